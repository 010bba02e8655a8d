// K4t -- lean band functions of the UNWEIGHTED truncated-linear (Felzenszwalb) aggregation: update_cost2_trunclinear /
// update_costW_trunclinear without image weights (mgm_core.cc:166-281 with minConvTruncatedLinear :152-163) inside the
// sweep loop (:489-580) -- the headline workload (BASELINE configs[2]: 2048x1536x256, TSGM=3).
//
// Same formulation, shared-memory layout (16-byte aligned rows, messages built in their ring slot, exact sequential
// min-convolution in place by lane pairs meeting in the middle), boundary lines and band scheduling as run_band /
// run_band_shear in aggregate.cu.  What changes is what aggregate_sgm.cu changed for the SGM potentials:
//   * compile-time label layout (NJ chunks per lane, 8 lanes per worker), incremental ring offsets and pixel pointers,
//     warp roles split at the top of the band;
//   * the hand-off between bands is OFF the step barriers: the two barriers of a step (after the gather, after the
//     chains; named barrier 1) hold the compute warps only and one thread bumps a step counter in shared memory.
//     The publisher warp follows that counter on its own: copy of the boundary vector(s), one fence + st.release for
//     everything copied since the last release, and a `copied` counter the boundary rows check before they overwrite a
//     ring slot.  The boundary consumer warp runs ahead: acquire of the previous band's progress, TMA loads into the
//     8-deep virtual-row ring as soon as published and free, mbarrier waits by the lanes that read the virtual row.
//     Measured on the generic kernel (option dbg, 2048x1536x256 TSGM=3): "top" of a step -- the wait for the boundary
//     pixel behind ld.acquire + TMA, and the publisher's two gpu-scope fences in front of the next barrier -- 3 800 to
//     5 500 cycles of a step whose gather and chains take 3 400 + 4 400.
// Arithmetic and its order are those of the generic path (bit-identical; tests/test_gpu_parity.py runs both).
#include "aggregate_dev.cuh"

#ifndef MGM_TRUNC_LB_T
#define MGM_TRUNC_LB_T MGM_AGG_MAX_THREADS   // __launch_bounds__ of the kernels: threads per CTA ...
#define MGM_TRUNC_LB_B 1                     // ... and CTAs per SM the register allocation allows
#endif

namespace mgm {

namespace {

__device__ __forceinline__ int lds_acquire(const int *p) {
   int v;
   asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
   return v;
}
__device__ __forceinline__ void sts_release(int *p, int v) {
   asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void compute_barrier(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

struct BandCtl {   // shared-memory control words of the running band
   int step;       // steps completed by the compute warps (sheared bands: v of the next step)
   int copied;     // steps whose boundary vectors the publisher has read out of the ring
};

static_assert(G == 8, "8 lanes per worker");

template <int NJ>
__device__ __forceinline__ void load_costs(float4 (&c)[NJ], const float4 *p) {
#pragma unroll
   for (int j = 0; j < NJ; ++j) c[j] = __ldcs(p + G * j);
}

// border pixel: the message is the matching cost (mgm_core.cc:538-541)
constexpr int LAY = MGM_TRUNC_LAY;   // shared-memory vector layout (ld16 / st16 in aggregate_dev.cuh)
constexpr int F2 = 2;                // float2 per 16-byte chunk

template <int NJ>
__device__ __forceinline__ float border_pixel(const float4 (&c)[NJ], float2 *cur, float4 *gout) {
   float m = MGM_INF;
#pragma unroll
   for (int j = 0; j < NJ; ++j) {
      m = hmin4(m, c[j]);
      st16<LAY>(cur, G * j, c[j]);
      __stcs(gout + G * j, c[j]);
   }
   return m;
}

// message of an interior pixel, built in its ring slot `cur` and streamed to the sweep's volume:
//   TSGM=2  c + (((a0 - m0) + a1) - m1) / 2      (update_cost2_trunclinear, mgm_core.cc:216)
//   else    c + (a0 + a1 + ...) / K              (update_costW_trunclinear :278; the producers subtracted their minima)
template <int K, int NJ>
__device__ __forceinline__ float gather_trunc(const float4 (&c)[NJ], const float2 *const (&S)[K], const float (&mk)[K],
                                              float2 *cur, float4 *gout) {
   float m = MGM_INF;
   constexpr int B = MGM_JB;
   static_assert(NJ % B == 0, "chunk count per lane is a multiple of the gather batch");
#pragma unroll
   for (int j0 = 0; j0 < NJ; j0 += B) {
      float4 a[K][B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
#pragma unroll
         for (int k = 0; k < K; ++k) a[k][jj] = ld16<LAY>(S[k], G * (j0 + jj));
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
         float4 o;
         if constexpr (K == 2) {
            o = add4(c[j0 + jj], mul4s(add4s(add4(add4s(a[0][jj], -mk[0]), a[1][jj]), -mk[1]), 0.5f));
         } else {
            float4 e = a[0][jj];
#pragma unroll
            for (int k = 1; k < K; ++k) e = add4(e, a[k][jj]);
            o = add4(c[j0 + jj], div4_by_k<K>(e));
         }
         m = hmin4(m, o);
         st16<LAY>(cur, G * (j0 + jj), o);
         __stcs(gout + G * (j0 + jj), o);
      }
   }
   return m;
}

__device__ __forceinline__ float group_min(float m, unsigned gmask) {
#pragma unroll
   for (int d = 1; d < G; d <<= 1) m = fminf(m, __shfl_xor_sync(gmask, m, d));
   return m;
}

// ---------------------------------------------------------------------------------------------------------
// Row-per-worker bands: axis sweeps 0-3 (lag 1; lag 2 with TSGM = 4) and, DIAG, sweeps 4-7 with TSGM = 4.
template <int K, int NJ, bool DIAG>
__device__ void run_band_trunc(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem, BandCtl *ctl) {
   constexpr int SIG = (DIAG || K == 4) ? 2 : 1;
   constexpr int R = SIG + 2;
   constexpr int CLS = DIAG ? CLS_DIAG : CLS_AXIS;
   constexpr int VS = 4 * G * NJ;
   constexpr int NQ = VS / 4;       // 16-byte chunks per vector
   constexpr bool NEEDM = (K == 2);
   constexpr uint32_t vbytes = (uint32_t)VS * 4u;

   const PassGeom g = pass_geometry(D.pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int T = P.T[CLS], TS = P.TS[CLS];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int ncomp = blockDim.x - 64;   // two service warps follow the row threads
   const int row0 = band * T;
   const int nrows = min(T, maxjj - row0);
   const bool has_prev = band > 0;
   const bool has_next = row0 + T < maxjj;
   const int nsteps = maxii + SIG * (nrows - 1);

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);      // [row][4] minima of the ring slots
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);     // [RV] minima of the virtual row
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);   // persistent mbarrier parities
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);

   if (tid == 0) { ctl->step = 0; ctl->copied = 0; }
   __syncthreads();

   if (tid >= ncomp + 32) {
      // ---------------- publisher warp: follows the step counter; boundary vector of the last row -> boundary line
      if (has_next) {
         float *bnd_out = D.bnd + (size_t)band * maxii * VS;
         float *bndm_out = D.bndm + (size_t)band * maxii;
         int *prog_out = D.progress + band;
         const float *last = thr + (size_t)(nrows - 1) * TS;
         const int lag = SIG * (nrows - 1);
         int slot = 0, seen = 0;
         for (int xl = 0; xl < maxii; ++xl) {
            const int need = xl + lag + 1;   // steps that must be complete
            if (seen < need) {
               if (lane == 0) { while ((seen = lds_acquire(&ctl->step)) < need) __nanosleep(40); }
               seen = __shfl_sync(0xffffffffu, seen, 0);
            }
            if (NEEDM && lane == 0) bndm_out[xl] = msr[(nrows - 1) * 4 + slot];
            warp_copy_vector(bnd_out + (size_t)xl * VS, last + slot * VS, VS, lane);
            __syncwarp();
            if (lane == 0) sts_release(&ctl->copied, need);   // the slot may be overwritten
            if (++slot == R) slot = 0;
            if (seen < need + 1 || xl + 1 == maxii) warp_publish(prog_out, xl + 1, lane);   // nothing more to copy right now
         }
      }
      __syncwarp();
   } else if (tid >= ncomp) {
      // ---------------- boundary consumer warp (lane 0): the previous band's last row -> virtual-row ring, running ahead
      if (has_prev && lane == 0) {
         const float *bnd_in = D.bnd + (size_t)(band - 1) * maxii * VS;
         const float *bndm_in = D.bndm + (size_t)(band - 1) * maxii;
         const int *prog_in = D.progress + band - 1;
         int avail = 0;
         for (int px = 0; px < maxii; ++px) {
            // slot of pixel px-RV: last read by row 0 in step px-RV+1
            while (lds_acquire(&ctl->step) < px - (RV - 2)) __nanosleep(20);
            while (avail < px + 1) {
               avail = ld_acquire(prog_in);
               if (avail < px + 1) __nanosleep(20);
            }
            fence_proxy_async();
            const int sl = px & (RV - 1);
            if (NEEDM) vms[sl] = __ldcg(bndm_in + px);   // ordered before the waiters' reads by the arrive / wait pair
            mbar_expect_tx(&vbar[sl], vbytes);
            tma_load_1d(virt + sl * VS, bnd_in + (size_t)px * VS, vbytes, &vbar[sl]);
         }
      }
      __syncwarp();
   } else {
      // ---------------- compute warps: 8 lanes per scan row, row r trails row r-1 by SIG pixels
      const int r = tid / G, gl = tid % G;
      const unsigned gmask = 0xffu << ((tid & 31) & ~(G - 1));
      const bool rowok = r < nrows;
      const int ys = row0 + r;
      float *rowf = thr + (size_t)r * TS;
      float2 *ownb = reinterpret_cast<float2 *>(rowf) + F2 * gl;            // chunk gl of my row's ring slot 0
      const bool upvirt = (r == 0);                                         // row -1 = the previous band's last row
      const float2 *upb = (upvirt ? reinterpret_cast<const float2 *>(virt) : reinterpret_cast<const float2 *>(rowf - TS)) + F2 * gl;
      const float *upm = upvirt ? vms : msr + (r - 1) * 4;
      const bool waiter = upvirt && has_prev;
      const bool bline = has_next && r == nrows - 1;
      const float p1 = P.P1, p2 = P.P2;
      uint32_t vph = waiter ? phase[vph_idx] : 0u;
      int vw = 0;   // next virtual pixel to wait for

      int xs = -SIG * r;
      int si = ((xs % R) + R) % R;   // ring slot of pixel xs
      const long long pix0 = g.base0 + (long long)ys * g.dys;
      const long long inc4 = g.dxs * NQ;
      const float4 *cp = reinterpret_cast<const float4 *>(D.cc) + (pix0 + (long long)(xs + 1) * g.dxs) * NQ + gl;   // pixel xs+1
      long long goff = (pix0 + (long long)xs * g.dxs) * NQ + gl;   // float4 offset of my chunk of pixel xs in a message volume
      const int nslabs = P.nslabs;                                 // > 1: row slabs of the volume live on peer GPUs (ldir_of_row)
      int yimg = g.y0 + xs * g.ydxs + ys * g.ydys;                 // image row of pixel xs

      // chain lanes: warps [0,ncw) run the upward halves of rows 32w+lane, warps [ncw,2ncw) the downward halves
      const int ncw = (T + 31) >> 5;
      const bool chain_warp = warp < 2 * ncw;
      const int cw = chain_warp ? warp % ncw : 0, cdir = chain_warp ? warp / ncw : 0;
      const int crow = cw * 32 + lane;
      int cxs = -SIG * crow;
      int csi = ((cxs % R) + R) % R;
      float *crowf = thr + (size_t)min(crow, T - 1) * TS;

      float4 c[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowok && xs == 0) load_costs<NJ>(c, cp - inc4);   // row 0 starts right away

      for (int s = 0; s < nsteps; ++s) {
         if (waiter) {   // virtual pixels xs-1, xs, xs+1 are read in this step (xs = s for row 0)
            const int need = min(maxii - 1, s + 1);
            while (vw <= need) {
               const int sl = vw & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
               ++vw;
            }
         }
         // ---------------- phase 1: gather
         if (rowok && (unsigned)xs < (unsigned)maxii) {
            // my slot held pixel xs-R: the publisher must have read it out (it does so two steps ahead of this)
            if (bline && xs >= R) { while (lds_acquire(&ctl->copied) < s - R + 1) {} }
            const bool border = (xs == 0) | (ys == 0) | (xs == maxii - 1);
            float2 *cur = ownb + si * (F2 * NQ);
            float4 *gp = reinterpret_cast<float4 *>(nslabs > 1 ? D.ldir[__umulhi((unsigned)yimg, P.slab_magic)] : D.ldir[0]) + goff;
            float m;
            if (border) m = border_pixel<NJ>(c, cur, gp);
            else {
               const int si_prev = (si == 0) ? R - 1 : si - 1;
               const int si_next = (si == R - 1) ? 0 : si + 1;
               const int ui = upvirt ? (xs & (RV - 1)) : si;
               const int ui_prev = upvirt ? ((xs - 1) & (RV - 1)) : si_prev;
               const int ui_next = upvirt ? ((xs + 1) & (RV - 1)) : si_next;
               (void)ui_next; (void)ui; (void)ui_prev;
               const float2 *S[K];
               float mk[K];
#pragma unroll
               for (int k = 0; k < K; ++k) {
                  const int pt = pred_type<DIAG>(k);
                  S[k] = (pt == PRED_SAME) ? ownb + si_prev * (F2 * NQ) : upb + ((pt == PRED_UP) ? ui : (pt == PRED_UPL) ? ui_prev : ui_next) * (F2 * NQ);
                  mk[k] = 0.f;
                  if (NEEDM) mk[k] = (pt == PRED_SAME) ? msr[r * 4 + si_prev] : upm[(pt == PRED_UP) ? ui : (pt == PRED_UPL) ? ui_prev : ui_next];
               }
               m = gather_trunc<K, NJ>(c, S, mk, cur, gp);
            }
            m = group_min(m, gmask);
            if (gl == 0) msr[r * 4 + si] = m;
         }
         // costs of the next pixel -> registers, in flight during phase 2
         if (rowok && (unsigned)(xs + 1) < (unsigned)maxii) load_costs<NJ>(c, cp);
         ++xs;
         cp += inc4;
         goff += inc4;
         yimg += g.ydxs;
         si = (si == R - 1) ? 0 : si + 1;
         compute_barrier(ncomp);

         // ---------------- phase 2: minConvTruncatedLinear of the finished messages, in place, one lane pair per row
         if (chain_warp) {
            const bool on = crow < nrows && (unsigned)cxs < (unsigned)maxii;
            const float cm = on ? msr[crow * 4 + csi] : 0.f;
            float2 *dst = reinterpret_cast<float2 *>(on ? crowf + csi * VS : thr);
            if (cdir == 0) minconv_half<0, false, LAY>(on, dst, dst, NQ, p1, cm + p2, NEEDM ? 0.0f : cm, 4 + cw);
            else minconv_half<1, false, LAY>(on, dst, dst, NQ, p1, cm + p2, NEEDM ? 0.0f : cm, 4 + cw);
            ++cxs;
            csi = (csi == R - 1) ? 0 : csi + 1;
         }
         compute_barrier(ncomp);
         if (tid == 0) sts_release(&ctl->step, s + 1);   // steps [0, s] are complete
      }
      if (waiter) {
         while (vw < maxii) {
            const int sl = vw & (RV - 1);
            mbar_wait(&vbar[sl], (vph >> sl) & 1u);
            vph ^= 1u << sl;
            ++vw;
         }
         if (tid == 0) phase[vph_idx] = vph;
      }
   }
   __syncthreads();
   band_finished(P, D, band);
}

// ---------------------------------------------------------------------------------------------------------
// Diagonal sweeps 4-7 with TSGM <= 3: sheared wavefront (run_band_shear, aggregate.cu): worker = anti-diagonal
// u = xs + ys, all workers at the same v = ys in a step, predecessors (u,v-1), (u-2,v-1), (u-1,v-1).
template <int K, int NJ>
__device__ void run_band_shear_trunc(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem, BandCtl *ctl) {
   static_assert(K <= 3, "the sheared wavefront needs predecessors in the row above only");
   constexpr int VS = 4 * G * NJ;
   constexpr int NQ = VS / 4;
   constexpr bool NEEDM = (K == 2);
   constexpr uint32_t vbytes = (uint32_t)VS * 4u;

   const PassGeom g = pass_geometry(D.pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int nu = maxii + maxjj - 1;   // anti-diagonals
   const int T = P.T[CLS_DIAG], TS = P.TS[CLS_DIAG];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int ncomp = blockDim.x - 64;
   const int u0 = band * T;
   const int nrows = min(T, nu - u0);
   const bool has_prev = band > 0;
   const bool has_next = u0 + T < nu;
   auto vlo = [&](int u) { return max(0, u - (maxii - 1)); };
   auto vhi = [&](int u) { return min(maxjj - 1, u); };
   const int sb = vlo(u0), se = vhi(u0 + nrows - 1);   // step window of the band (v = step)
   const int pf_lo = max(sb - 1, 0), pf_hi = has_prev ? min(se - 1, vhi(u0 - 1)) : -1;

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);     // [2][RV]
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);   // [2][RV][VS]: worker -1, worker -2
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);

   if (tid == 0) { ctl->step = sb; ctl->copied = sb; }   // v of the next step / of the next position to copy
   __syncthreads();

   if (tid >= ncomp + 32) {
      // ---------------- publisher warp: the band's last two workers -> boundary lines [line][maxjj][VS]
      if (has_next) {
         float *bnd_out = D.bnd + (size_t)band * 2 * maxjj * VS;
         float *bndm_out = D.bndm + (size_t)band * 2 * maxjj;
         int *prog_out = D.progress + band;
         int seen = sb;
         for (int v = sb; v <= se; ++v) {
            if (seen < v + 1) {
               if (lane == 0) { while ((seen = lds_acquire(&ctl->step)) < v + 1) __nanosleep(40); }
               seen = __shfl_sync(0xffffffffu, seen, 0);
            }
#pragma unroll
            for (int line = 0; line < 2; ++line) {
               const int br = nrows - 1 - line;
               if (v >= vlo(u0 + br) && v <= vhi(u0 + br)) {
                  if (NEEDM && lane == 0) bndm_out[(size_t)line * maxjj + v] = msr[br * 4 + (v & 1)];
                  warp_copy_vector(bnd_out + ((size_t)line * maxjj + v) * VS, thr + (size_t)br * TS + (v & 1) * VS, VS, lane);
               }
            }
            __syncwarp();
            if (lane == 0) sts_release(&ctl->copied, v + 1);
            if (seen < v + 2 || v == se) warp_publish(prog_out, v + 1, lane);   // positions <= v of both workers are in memory
         }
         warp_publish(prog_out, 0x7fffffff, lane);
      }
      __syncwarp();
   } else if (tid >= ncomp) {
      // ---------------- boundary consumer warp (lane 0), running ahead
      if (pf_hi >= pf_lo && lane == 0) {
         const float *bnd_in = D.bnd + (size_t)(band - 1) * 2 * maxjj * VS;
         const float *bndm_in = D.bndm + (size_t)(band - 1) * 2 * maxjj;
         const int *prog_in = D.progress + band - 1;
         int avail = 0;
         for (int p = pf_lo; p <= pf_hi; ++p) {
            // slot of position p-RV: last read in step p-RV+1
            while (lds_acquire(&ctl->step) < p - (RV - 2)) __nanosleep(20);
            while (avail < p + 1) {
               avail = ld_acquire(prog_in);
               if (avail < p + 1) __nanosleep(20);
            }
            fence_proxy_async();
            const int sl = p & (RV - 1);
            if (NEEDM) {
               vms[sl] = __ldcg(bndm_in + p);
               vms[RV + sl] = __ldcg(bndm_in + maxjj + p);
            }
            mbar_expect_tx(&vbar[sl], 2 * vbytes);
            tma_load_1d(virt + sl * VS, bnd_in + (size_t)p * VS, vbytes, &vbar[sl]);
            tma_load_1d(virt + (RV + sl) * VS, bnd_in + ((size_t)maxjj + p) * VS, vbytes, &vbar[sl]);
         }
      }
      __syncwarp();
   } else {
      // ---------------- compute warps: 8 lanes per anti-diagonal
      const int r = tid / G, gl = tid % G;
      const unsigned gmask = 0xffu << ((tid & 31) & ~(G - 1));
      const bool rowok = r < nrows;
      const int u = u0 + r;
      const int my_lo = vlo(u), my_hi = vhi(u);
      float *rowf = thr + (size_t)r * TS;
      float2 *ownb = reinterpret_cast<float2 *>(rowf) + F2 * gl;
      // workers r-2 and r-1: real rows, or the virtual workers -1 (line 0) and -2 (line 1) of the previous band
      const bool v1 = r < 2, v2 = r < 1;
      const float2 *p1b = (v1 ? reinterpret_cast<const float2 *>(virt + (r == 1 ? 0 : RV * VS)) : reinterpret_cast<const float2 *>(rowf - 2 * TS)) + F2 * gl;
      const float2 *p2b = (v2 ? reinterpret_cast<const float2 *>(virt) : reinterpret_cast<const float2 *>(rowf - TS)) + F2 * gl;
      const float *m1b = v1 ? vms + (r == 1 ? 0 : RV) : msr + (r - 2) * 4;
      const float *m2b = v2 ? vms : msr + (r - 1) * 4;
      const bool waiter = v1 && pf_hi >= pf_lo;
      const bool bline = has_next && rowok && (nrows - 1 - r) < 2;
      const float p1 = P.P1, p2 = P.P2;
      uint32_t vph = waiter ? phase[vph_idx] : 0u;
      int vw = pf_lo;

      const long long pix_u = g.base0 + (long long)u * g.dxs;   // pixel of (u, v): xs = u - v, ys = v
      const long long dv = g.dys - g.dxs;
      const long long inc4 = dv * NQ;
      const float4 *cp = reinterpret_cast<const float4 *>(D.cc) + (pix_u + (long long)(sb + 1) * dv) * NQ + gl;   // position v+1
      long long goff = (pix_u + (long long)sb * dv) * NQ + gl;
      const int nslabs = P.nslabs;
      int yimg = g.y0 + (u - sb) * g.ydxs + sb * g.ydys;   // image row of (xs = u - v, ys = v)

      const int ncw = (T + 31) >> 5;
      const bool chain_warp = warp < 2 * ncw;
      const int cw = chain_warp ? warp % ncw : 0, cdir = chain_warp ? warp / ncw : 0;
      const int crow = cw * 32 + lane;
      const int c_lo = vlo(u0 + crow), c_hi = vhi(u0 + crow);
      float *crowf = thr + (size_t)min(crow, T - 1) * TS;

      float4 c[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowok && sb >= my_lo && sb <= my_hi) load_costs<NJ>(c, cp - inc4);

      for (int v = sb; v <= se; ++v) {
         if (waiter) {   // position v-1 of the virtual workers is read in this step
            const int need = min(pf_hi, v - 1);
            while (vw <= need) {
               const int sl = vw & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
               ++vw;
            }
         }
         // ---------------- phase 1: gather
         if (rowok && v >= my_lo && v <= my_hi) {
            // my slot held position v-2: the publisher must have read it out
            if (bline && v - 2 >= sb) { while (lds_acquire(&ctl->copied) < v - 1) {} }
            const int xs = u - v;
            const bool border = (xs == 0) | (v == 0) | (xs == maxii - 1);
            float2 *cur = ownb + (v & 1) * (F2 * NQ);
            float4 *gp = reinterpret_cast<float4 *>(nslabs > 1 ? D.ldir[__umulhi((unsigned)yimg, P.slab_magic)] : D.ldir[0]) + goff;
            float m;
            if (border) m = border_pixel<NJ>(c, cur, gp);
            else {
               const int po = (v - 1) & 1;            // ring slot of position v-1 in a real row
               const int vo = (v - 1) & (RV - 1);     // ... in the virtual workers' rings
               const float2 *S3[3] = {ownb + po * (F2 * NQ), p1b + (v1 ? vo : po) * (F2 * NQ), p2b + (v2 ? vo : po) * (F2 * NQ)};
               const float2 *S[K];
               float mk[K];
#pragma unroll
               for (int k = 0; k < K; ++k) { S[k] = S3[k]; mk[k] = 0.f; }
               if (NEEDM) {
                  mk[0] = msr[r * 4 + po];
                  mk[1 % K] = m1b[v1 ? vo : po];
               }
               (void)m2b;
               m = gather_trunc<K, NJ>(c, S, mk, cur, gp);
            }
            m = group_min(m, gmask);
            if (gl == 0) msr[r * 4 + (v & 1)] = m;
         }
         if (rowok && v + 1 >= my_lo && v + 1 <= my_hi) load_costs<NJ>(c, cp);
         cp += inc4;
         goff += inc4;
         yimg += g.ydys - g.ydxs;
         compute_barrier(ncomp);

         // ---------------- phase 2: min-convolution chains, in place
         if (chain_warp) {
            const bool on = crow < nrows && v >= c_lo && v <= c_hi;
            const float cm = on ? msr[crow * 4 + (v & 1)] : 0.f;
            float2 *dst = reinterpret_cast<float2 *>(on ? crowf + (v & 1) * VS : thr);
            if (cdir == 0) minconv_half<0, false, LAY>(on, dst, dst, NQ, p1, cm + p2, NEEDM ? 0.0f : cm, 4 + cw);
            else minconv_half<1, false, LAY>(on, dst, dst, NQ, p1, cm + p2, NEEDM ? 0.0f : cm, 4 + cw);
         }
         compute_barrier(ncomp);
         if (tid == 0) sts_release(&ctl->step, v + 1);
      }
      if (waiter) {
         while (vw <= pf_hi) {
            const int sl = vw & (RV - 1);
            mbar_wait(&vbar[sl], (vph >> sl) & 1u);
            vph ^= 1u << sl;
            ++vw;
         }
         if (tid == 0) phase[vph_idx] = vph;
      }
   }
   __syncthreads();
   band_finished(P, D, band);
}

// The persistent kernel: same claim loop and finish tiles as mgm_aggregate_kernel (aggregate.cu).
template <int K, int NJ>
__global__ void __launch_bounds__(MGM_TRUNC_LB_T, MGM_TRUNC_LB_B) mgm_aggregate_trunc_kernel(const AggParams P) {
   extern __shared__ __align__(128) unsigned char smem[];
   __shared__ int2 s_ticket;
   __shared__ BandCtl s_ctl;
   __shared__ AggStage s_stage;
   __shared__ __align__(16) unsigned char s_tab_raw[MGM_MAX_NDIR * sizeof(SweepDesc)];
   const int t = threadIdx.x;
   const int ncomp = blockDim.x - 64;
   SweepDesc *s_tab = reinterpret_cast<SweepDesc *>(s_tab_raw);
   const bool small_tab = P.nsweeps <= MGM_MAX_NDIR;
   if (small_tab) {
      const uint4 *src = reinterpret_cast<const uint4 *>(P.sweeps);
      uint4 *dst = reinterpret_cast<uint4 *>(s_tab_raw);
      for (int i = t; i < P.nsweeps * (int)(sizeof(SweepDesc) / 16); i += blockDim.x) dst[i] = src[i];
   }
   const SweepDesc *tab = small_tab ? s_tab : P.sweeps;
   {
      uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
      uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);
      const int tmax = max(max(P.T[0], P.T[1]), P.T[2]);
      if (t == ncomp) { for (int i = 0; i < RV; ++i) mbar_init(&vbar[i], 1); }
      if (t == 0) phase[tmax] = 0;
      mbar_fence_init();
      __syncthreads();
   }
   int pending = -1;
   for (;;) {
      if (t < 32) {
         const int2 tk = claim_band(P, tab, pending, t);
         if (t == 0) s_ticket = tk;
         if (tk.x >= 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tab + tk.x);
            uint4 *dst = reinterpret_cast<uint4 *>(&s_stage.d);
            for (int i = t; i < (int)(sizeof(SweepDesc) / 16); i += 32) dst[i] = src[i];
         } else if (tk.x == -2 && P.npairs > 1) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.fins + tk.y / P.fin_ntiles);
            uint4 *dst = reinterpret_cast<uint4 *>(&s_stage.f);
            for (int i = t; i < (int)(sizeof(WtaParams) / 16); i += 32) dst[i] = src[i];
         }
      }
      __syncthreads();
      const int2 pb = s_ticket;
      if (pb.x == -1) break;
      if (pb.x == -2) {
         if (P.npairs == 1) run_finish_tile(P, P.fin0, pb.y, smem);
         else run_finish_tile(P, s_stage.f, pb.y % P.fin_ntiles, smem);
      } else {
         const SweepDesc &D = s_stage.d;
         if (D.pass < 4) run_band_trunc<K, NJ, false>(P, D, pb.y, smem, &s_ctl);
         else if constexpr (K <= 3) run_band_shear_trunc<K, NJ>(P, D, pb.y, smem, &s_ctl);
         else run_band_trunc<K, NJ, true>(P, D, pb.y, smem, &s_ctl);
      }
      __syncthreads();
   }
}

template <int K, int NJ>
cudaError_t launch_lean(const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   auto kern = mgm_aggregate_trunc_kernel<K, NJ>;
   cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.block, plan.smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   int grid = min(P.nbands, plan.num_sms * per_sm);
   if (grid < 1) grid = 1;
   if (plan.verbose)
      fprintf(stderr, "[mgmb200] aggregate (lean truncated linear K=%d chunks=%d): grid=%d block=%d smem=%zu CTAs/SM=%d sweeps=%d bands=%d rows=%d/%d\n",
              K, NJ, grid, plan.block, plan.smem, per_sm, P.nsweeps, P.nbands, plan.T[0], plan.T[1]);
   kern<<<grid, plan.block, plan.smem, st>>>(P);
   return cudaGetLastError();
}

template <int NJ>
cudaError_t launch_lean_k(int K, const AggParams &P, const AggPlan &plan, cudaStream_t st) {
#ifdef MGM_QUICK_K   // development builds: one TSGM value only
   if (K != MGM_QUICK_K) return cudaErrorNotSupported;
   return launch_lean<MGM_QUICK_K, NJ>(P, plan, st);
#else
   switch (K) {
   case 1: return launch_lean<1, NJ>(P, plan, st);
   case 2: return launch_lean<2, NJ>(P, plan, st);
   case 3: return launch_lean<3, NJ>(P, plan, st);
   default: return launch_lean<4, NJ>(P, plan, st);
   }
#endif
}

}  // namespace

// chunk counts per lane the lean kernels are built for (8 lanes per worker)
bool agg_trunc_lean_supported(int VS) {
   if (VS % 32) return false;
   const int nj = VS / 32;
   return nj == 2 || nj == 4 || nj == 6 || nj == 8;
}

cudaError_t agg_launch_trunc_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st) {
   switch (plan.VS / 32) {
   case 2: return launch_lean_k<2>(K, P, plan, st);
   case 4: return launch_lean_k<4>(K, P, plan, st);
   case 6: return launch_lean_k<6>(K, P, plan, st);
   case 8: return launch_lean_k<8>(K, P, plan, st);
   }
   return cudaErrorNotSupported;
}

}  // namespace mgm
