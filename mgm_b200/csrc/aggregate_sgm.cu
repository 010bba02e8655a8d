// K4s -- lean band functions of the UNWEIGHTED SGM-potential aggregation (update_cost2 / update_costW without image
// weights, mgm_core.cc:66-144, inside the sweep loop :489-580): the CLI's default potential and BASELINE configs 1, 3, 4.
//
// Same formulation, shared-memory layout, boundary lines and band scheduling as aggregate.cu (run_band /
// run_band_shear): these functions restate its register-resident SGM path (costs in registers, gather and min3
// transform fused, one barrier per step) for the case the generic functions pay most for their generality.
// Measured on the generic kernel, 1920x1080x128 TSGM=2 (profiles/r02_ncu_summary_cfg2.md, option dbg):
//   * 512 instructions per warp and step, about 320 of them index arithmetic, constant reloads and role tests;
//   * a step of 4 000 - 5 000 cycles of which the instructions explain 1 000: every step the publisher warp runs
//     copy -> __threadfence -> st.release (two gpu-scope fences, 3 600 cycles measured) and the boundary consumer an
//     ld.acquire (L1 invalidate) + TMA round trip, both in front of the barrier that every warp of the band waits at.
// Here:
//   * the chunk count per lane NJ and the lanes per worker GL are template parameters: VS = 4*GL*NJ is a compile-time
//     constant, every shared/global access of a step is a base register + immediate; ring-slot offsets, pixel pointers
//     and scan coordinates advance incrementally; the matching costs are double-buffered in registers (loads of
//     pixel xs+1 issued at the top of the step of pixel xs, step loop unrolled by two so the buffers swap without moves);
//   * the hand-off between bands is OFF the step barrier.  The step barrier (named barrier 1) holds the compute
//     warps only; after it one thread bumps a step counter in shared memory.
//       - boundary lines are written by the lanes that produce them (the last row / the last two anti-diagonals store
//         their transformed vector to global as well as to their ring slot): no copy through a publisher;
//       - the publisher warp polls the step counter and releases the progress counter of the band with ONE
//         st.release.gpu for everything finished since its last release (cumulative: the lanes' stores happen before the
//         barrier, the counter bump after it, the publisher's acquire of the counter after that);
//       - the boundary consumer warp runs ahead on its own: it acquires the previous band's progress, TMA-loads the
//         vectors into the 8-deep virtual-row ring as soon as they are published and their slot is free (step counter),
//         and the lanes that read the virtual row wait on the slot's mbarrier themselves.
// Arithmetic and its order are those of the generic path (bit-identical results; tests/test_gpu_parity.py runs every
// SGM case through both).  Not covered here (the generic kernel runs them): image-dependent weights,
// label counts whose chunk count per lane is odd, diagonal sweeps without the sheared wavefront.
#include "aggregate_dev.cuh"

namespace mgm {

// ---------------------------------------------------------------- shared-memory step counter (cta scope)
__device__ __forceinline__ int lds_acquire(const int *p) {
   int v;
   asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
   return v;
}
__device__ __forceinline__ void sts_release(int *p, int v) {
   asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void compute_barrier(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// message of one pixel: c[] holds the matching costs on entry and L_p = C_p + (sum_k A_k)/K on exit; returns this lane's
// minimum.  S[k] points at chunk gl of the k-th predecessor's transformed vector (8-byte halves, see ld16<0>).
template <int K, int NJ, int G>
__device__ __forceinline__ float gather_sgm(float4 (&c)[NJ], const float2 *const (&S)[K]) {
   float m = MGM_INF;
   constexpr int B = 2;
   static_assert(NJ % B == 0, "even chunk count per lane");
#pragma unroll
   for (int j0 = 0; j0 < NJ; j0 += B) {
      float4 a[K][B];
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
#pragma unroll
         for (int k = 0; k < K; ++k) {
            const float2 lo = S[k][2 * G * (j0 + jj)], hi = S[k][2 * G * (j0 + jj) + 1];
            a[k][jj] = make_float4(lo.x, lo.y, hi.x, hi.y);
         }
      }
#pragma unroll
      for (int jj = 0; jj < B; ++jj) {
         float4 e = a[0][jj];
#pragma unroll
         for (int k = 1; k < K; ++k) e = add4(e, a[k][jj]);
         // K == 2: the producers have halved their vectors (update_cost2); otherwise e / K (update_costW)
         const float4 o = add4(c[j0 + jj], (K == 2) ? e : div4_by_k<K>(e));
         c[j0 + jj] = o;
         m = hmin4(m, o);
      }
   }
   return m;
}

template <int NJ, int G>
__device__ __forceinline__ void load_costs(float4 (&c)[NJ], const float4 *p) {
#pragma unroll
   for (int j = 0; j < NJ; ++j) c[j] = __ldcs(p + G * j);
}

// Finish one pixel: store the message, reduce its minimum over the worker's lanes, write the transformed vector
// A(o) = min3(L(o), min(L(o-1),L(o+1))+P1, m+P2) - m  [x 1/2 for K=2]  (sgm_transform_regs, aggregate_dev.cuh) into the
// ring slot `cur` and, for a boundary worker, into its boundary line `gb` (chunk gl of the line, or nullptr).
template <int K, int NJ, int G>
__device__ __forceinline__ void finish_pixel(float4 (&v)[NJ], float m, float4 *gout, float2 *cur, float4 *gb, int gl,
                                             unsigned gmask, float p1, float p2) {
#pragma unroll
   for (int j = 0; j < NJ; ++j) __stcs(gout + G * j, v[j]);
#pragma unroll
   for (int d = 1; d < G; d <<= 1) m = fminf(m, __shfl_xor_sync(gmask, m, d));
   const float cap = m + p2;
   const float sc = (K == 2) ? 0.5f : 1.0f;
#pragma unroll
   for (int j = 0; j < NJ; ++j) {
      const int q = gl + G * j;
      // last label of chunk q-1, first label of chunk q+1: one rotation of the group each way, the SENDER picks the value
      const float snd_l = (gl == G - 1) ? v[j > 0 ? j - 1 : 0].w : v[j].w;
      const float snd_r = (gl == 0) ? v[j + 1 < NJ ? j + 1 : j].x : v[j].x;
      const float got_l = __shfl_sync(gmask, snd_l, (gl + G - 1) & (G - 1), G);
      const float got_r = __shfl_sync(gmask, snd_r, (gl + 1) & (G - 1), G);
      const float lft = (q == 0) ? MGM_INF : got_l;
      const float rgt = (q + 1 >= NJ * G) ? MGM_INF : got_r;
      float4 a;
      a.x = sgm_x(lft, v[j].x, v[j].y, p1, cap, m) * sc;
      a.y = sgm_x(v[j].x, v[j].y, v[j].z, p1, cap, m) * sc;
      a.z = sgm_x(v[j].y, v[j].z, v[j].w, p1, cap, m) * sc;
      a.w = sgm_x(v[j].z, v[j].w, rgt, p1, cap, m) * sc;
      st16<0>(cur, q, a);
      if (gb) gb[G * j] = a;
   }
}

// ---------------------------------------------------------------------------------------------------------
// Row-per-worker bands.  MODE 0: axis sweeps 0-3 (lag 1; lag 2 with TSGM = 4); MODE 1: sweeps 4-7 with TSGM = 4 (lag 2,
// predecessors in the order (+1,-1), (-1,-1), (0,-1), (-1,0)); MODE 2: sweeps 8-15 (lag 2, predecessor order by the
// parity of the scan coordinates, knight_pred_type in common.cuh).  See run_band (aggregate.cu).
template <int K, int GL, int NJ, int MODE>
__device__ void run_band_sgm(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem, int *s_step) {
   constexpr int G = GL;
   constexpr bool DIAG = (MODE == 1), KN = (MODE == 2);
   constexpr int SIG = (MODE != 0 || K == 4) ? 2 : 1;
   constexpr int R = SIG + 2;
   constexpr int CLS = KN ? CLS_KNIGHT : (DIAG ? CLS_DIAG : CLS_AXIS);
   const bool kn_diag = (D.pass & 7) >= 4;
   constexpr int VS = 4 * G * NJ;
   constexpr int V4 = VS / 4;       // float4 per vector
   constexpr int SLOT2 = VS / 2;    // float2 per ring slot
   constexpr uint32_t vbytes = (uint32_t)VS * 4u;

   const PassGeom g = pass_geometry(D.pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int T = P.T[CLS], TS = P.TS[CLS];
   const int tid = threadIdx.x, lane = tid & 31;
   const int ncomp = blockDim.x - 64;   // two service warps follow the row threads
   const int row0 = band * T;
   const int nrows = min(T, maxjj - row0);
   const bool has_prev = band > 0;
   const bool has_next = row0 + T < maxjj;
   const int nsteps = maxii + SIG * (nrows - 1);

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);   // persistent mbarrier parities
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);

   if (tid == 0) *s_step = 0;   // steps completed by the compute warps
   __syncthreads();

   if (tid >= ncomp + 32) {
      // ---------------- publisher warp (lane 0): the last row stores its boundary line itself; release what is finished
      if (has_next && lane == 0) {
         int *prog_out = D.progress + band;
         int pub = 0;
         while (pub < maxii) {
            const int done = min(maxii, lds_acquire(s_step) - SIG * (nrows - 1));   // finished pixels of the last row
            if (done > pub) { st_release(prog_out, done); pub = done; }
            else __nanosleep(40);
         }
      }
      __syncwarp();
   } else if (tid >= ncomp) {
      // ---------------- boundary consumer warp (lane 0): the previous band's last row -> virtual-row ring, running ahead
      if (has_prev && lane == 0) {
         const float *bnd_in = D.bnd + (size_t)(band - 1) * maxii * VS;
         const int *prog_in = D.progress + band - 1;
         int avail = 0;
         for (int px = 0; px < maxii; ++px) {
            // slot of pixel px-RV: last read by row 0 in step px-RV+1
            while (lds_acquire(s_step) < px - (RV - 2)) __nanosleep(20);
            while (avail < px + 1) {
               avail = ld_acquire(prog_in);
               if (avail < px + 1) __nanosleep(20);
            }
            fence_proxy_async();
            const int sl = px & (RV - 1);
            mbar_expect_tx(&vbar[sl], vbytes);
            tma_load_1d(virt + sl * VS, bnd_in + (size_t)px * VS, vbytes, &vbar[sl]);
         }
      }
      __syncwarp();
   } else {
      // ---------------- compute warps: G lanes per scan row, row r trails row r-1 by SIG pixels
      const int r = tid / G, gl = tid % G;
      const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
      const bool rowok = r < nrows;
      const int ys = row0 + r;
      float2 *ownb = reinterpret_cast<float2 *>(thr + (size_t)r * TS);     // my row's ring (TS is even: 8-byte aligned rows)
      const bool upvirt = (r == 0);                                        // row -1 = the previous band's last row
      const float2 *upb = upvirt ? reinterpret_cast<const float2 *>(virt) : ownb - (TS >> 1);
      const bool waiter = upvirt && has_prev;                              // my reads of the virtual row wait on its mbarriers
      const bool bline = has_next && r == nrows - 1;                       // my row is the band's boundary line
      const float p1 = P.P1, p2 = P.P2;
      const int cc_pf = P.cc_pf;
      uint32_t vph = waiter ? phase[vph_idx] : 0u;
      int vw = 0;   // next virtual pixel to wait for

      int xs = -SIG * r;
      int so = (((xs % R) + R) % R) * SLOT2;   // ring offset (float2) of pixel xs; the same in every real row
      const long long pix0 = g.base0 + (long long)ys * g.dys;
      const long long inc4 = g.dxs * V4;
      const float4 *cp = reinterpret_cast<const float4 *>(D.cc) + (pix0 + (long long)(xs + 1) * g.dxs) * V4 + gl;   // pixel xs+1
      long long goff = (pix0 + (long long)xs * g.dxs) * V4 + gl;   // float4 offset of my chunk of pixel xs in a message volume
      const int nslabs = P.nslabs;                                 // > 1: row slabs of the volume live on peer GPUs (ldir_of_row)
      int yimg = g.y0 + xs * g.ydxs + ys * g.ydys;                 // image row of pixel xs
      float4 *gb = bline ? reinterpret_cast<float4 *>(D.bnd + (size_t)band * maxii * VS) + (long long)xs * V4 + gl : nullptr;

      float4 c0[NJ], c1[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) c0[j] = c1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowok && xs == 0) load_costs<NJ, G>(c0, cp - inc4);   // row 0 starts right away

      int s = 0;
      auto step = [&](float4 (&cc)[NJ], float4 (&cn)[NJ]) {
         // costs of the pixel of the NEXT step: in flight during the whole step
         if (rowok && (unsigned)(xs + 1) < (unsigned)maxii) load_costs<NJ, G>(cn, cp);
         if (cc_pf > 0) {   // and further ahead into L2
            const int pp = xs + 1 + cc_pf;
            if (rowok && (unsigned)pp < (unsigned)maxii) {
               const float *line = reinterpret_cast<const float *>(cp - gl + (long long)cc_pf * inc4);
               for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + l * 32));
            }
         }
         if (waiter) {   // virtual pixels xs-1, xs, xs+1 are read in this step (xs = s for row 0)
            const int need = min(maxii - 1, s + 1);
            while (vw <= need) {
               const int sl = vw & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
               ++vw;
            }
         }
         if (rowok && (unsigned)xs < (unsigned)maxii) {
            const bool border = (xs == 0) | (ys == 0) | (xs == maxii - 1);   // mgm_core.cc:538-541
            float m = MGM_INF;
            if (border) {
#pragma unroll
               for (int j = 0; j < NJ; ++j) m = hmin4(m, cc[j]);
            } else {
               const int so_prev = (so == 0) ? (R - 1) * SLOT2 : so - SLOT2;
               const int so_next = (so == (R - 1) * SLOT2) ? 0 : so + SLOT2;
               // the same pixels in the row above: same ring offsets, or the 8-deep virtual-row ring
               const int uo = upvirt ? (xs & (RV - 1)) * SLOT2 : so;
               const int uo_prev = upvirt ? ((xs - 1) & (RV - 1)) * SLOT2 : so_prev;
               const int uo_next = upvirt ? ((xs + 1) & (RV - 1)) * SLOT2 : so_next;
               const float2 *same = ownb + so_prev + 2 * gl, *up = upb + uo + 2 * gl, *upl = upb + uo_prev + 2 * gl,
                            *upr = upb + uo_next + 2 * gl;
               (void)upr; (void)upl; (void)up; (void)same;
               const float2 *S[K];
#pragma unroll
               for (int k = 0; k < K; ++k) {
                  const int pt = KN ? knight_pred_type(kn_diag, k, xs, ys) : pred_type<DIAG>(k);
                  S[k] = (pt == PRED_SAME) ? same : (pt == PRED_UP) ? up : (pt == PRED_UPL) ? upl : upr;
               }
               m = gather_sgm<K, NJ, G>(cc, S);
            }
            float4 *gp = reinterpret_cast<float4 *>(nslabs > 1 ? D.ldir[__umulhi((unsigned)yimg, P.slab_magic)] : D.ldir[0]) + goff;
            finish_pixel<K, NJ, G>(cc, m, gp, ownb + so, gb, gl, gmask, p1, p2);
         }
         ++xs;
         ++s;
         cp += inc4;
         goff += inc4;
         yimg += g.ydxs;
         if (bline) gb += V4;
         so = (so == (R - 1) * SLOT2) ? 0 : so + SLOT2;
         compute_barrier(ncomp);
         if (tid == 0) sts_release(s_step, s);   // steps [0, s) are complete: boundary stores issued, virtual pixels < s-1 read
      };
      for (int i = 0; i < nsteps; i += 2) {
         step(c0, c1);
         if (i + 1 < nsteps) step(c1, c0);
      }
      if (waiter) {
         while (vw < maxii) {   // (nsteps >= maxii: nothing left; kept for symmetry with the sheared bands)
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
// Diagonal sweeps 4-7 with TSGM <= 3: sheared wavefront, worker = anti-diagonal u = xs + ys, all workers of a band at
// the same v = ys in a step, predecessors (u,v-1), (u-2,v-1), (u-1,v-1).  See run_band_shear (aggregate.cu).
template <int K, int GL, int NJ>
__device__ void run_band_shear_sgm(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem, int *s_step) {
   static_assert(K <= 3, "the sheared wavefront needs predecessors in the row above only");
   constexpr int G = GL;
   constexpr int VS = 4 * G * NJ;
   constexpr int V4 = VS / 4;
   constexpr int SLOT2 = VS / 2;
   constexpr uint32_t vbytes = (uint32_t)VS * 4u;

   const PassGeom g = pass_geometry(D.pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int nu = maxii + maxjj - 1;   // anti-diagonals
   const int T = P.T[CLS_DIAG], TS = P.TS[CLS_DIAG];
   const int tid = threadIdx.x, lane = tid & 31;
   const int ncomp = blockDim.x - 64;
   const int u0 = band * T;
   const int nrows = min(T, nu - u0);
   const bool has_prev = band > 0;
   const bool has_next = u0 + T < nu;
   auto vlo = [&](int u) { return max(0, u - (maxii - 1)); };
   auto vhi = [&](int u) { return min(maxjj - 1, u); };
   const int sb = vlo(u0), se = vhi(u0 + nrows - 1);   // step window of the band (v = step)
   // boundary positions this band ever reads from the previous band's last two workers
   const int pf_lo = max(sb - 1, 0), pf_hi = has_prev ? min(se - 1, vhi(u0 - 1)) : -1;

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);   // [2][RV][VS]: worker -1, worker -2
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);

   if (tid == 0) *s_step = sb;   // v of the next step: positions < *s_step are complete
   __syncthreads();

   if (tid >= ncomp + 32) {
      // ---------------- publisher warp (lane 0): positions <= v of both boundary workers are in global memory once step v
      // is complete (the workers store their lines themselves)
      if (has_next && lane == 0) {
         int *prog_out = D.progress + band;
         int pub = sb;
         while (pub <= se) {
            const int done = lds_acquire(s_step);
            if (done > pub) { st_release(prog_out, done); pub = done; }
            else __nanosleep(40);
         }
         st_release(prog_out, 0x7fffffff);
      }
      __syncwarp();
   } else if (tid >= ncomp) {
      // ---------------- boundary consumer warp (lane 0), running ahead
      if (pf_hi >= pf_lo && lane == 0) {
         const float *bnd_in = D.bnd + (size_t)(band - 1) * 2 * maxjj * VS;
         const int *prog_in = D.progress + band - 1;
         int avail = 0;
         for (int p = pf_lo; p <= pf_hi; ++p) {
            // slot of position p-RV: last read in step p-RV+1
            while (lds_acquire(s_step) < p - (RV - 2)) __nanosleep(20);
            while (avail < p + 1) {
               avail = ld_acquire(prog_in);
               if (avail < p + 1) __nanosleep(20);
            }
            fence_proxy_async();
            const int sl = p & (RV - 1);
            mbar_expect_tx(&vbar[sl], 2 * vbytes);
            tma_load_1d(virt + sl * VS, bnd_in + (size_t)p * VS, vbytes, &vbar[sl]);
            tma_load_1d(virt + (RV + sl) * VS, bnd_in + ((size_t)maxjj + p) * VS, vbytes, &vbar[sl]);
         }
      }
      __syncwarp();
   } else {
      // ---------------- compute warps: G lanes per anti-diagonal
      const int r = tid / G, gl = tid % G;
      const unsigned gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << ((tid & 31) & ~(G - 1));
      const bool rowok = r < nrows;
      const int u = u0 + r;
      const int my_lo = vlo(u), my_hi = vhi(u);
      float2 *ownb = reinterpret_cast<float2 *>(thr + (size_t)r * TS);
      // workers r-2 and r-1: real rows, or the virtual workers -1 (line 0) and -2 (line 1) of the previous band
      const bool v1 = r < 2, v2 = r < 1;
      const float2 *p1b = v1 ? reinterpret_cast<const float2 *>(virt + (r == 1 ? 0 : RV * VS)) : ownb - TS;
      const float2 *p2b = v2 ? reinterpret_cast<const float2 *>(virt) : ownb - (TS >> 1);
      const bool waiter = v1 && pf_hi >= pf_lo;   // workers 0 and 1 read the virtual workers
      // boundary lines [line][maxjj][VS]: line 0 = last worker of the band, line 1 = the one before
      const int bl = nrows - 1 - r;
      const bool bline = has_next && rowok && bl < 2;
      const float p1 = P.P1, p2 = P.P2;
      const int cc_pf = P.cc_pf;
      uint32_t vph = waiter ? phase[vph_idx] : 0u;
      int vw = pf_lo;   // next boundary position to wait for

      int v = sb;
      const long long pix_u = g.base0 + (long long)u * g.dxs;   // pixel of (u, v): xs = u - v, ys = v
      const long long dv = g.dys - g.dxs;
      const long long inc4 = dv * V4;
      const float4 *cp = reinterpret_cast<const float4 *>(D.cc) + (pix_u + (long long)(v + 1) * dv) * V4 + gl;   // position v+1
      long long goff = (pix_u + (long long)v * dv) * V4 + gl;
      const int nslabs = P.nslabs;
      int yimg = g.y0 + (u - v) * g.ydxs + v * g.ydys;   // image row of (xs = u - v, ys = v)
      float4 *gb = bline ? reinterpret_cast<float4 *>(D.bnd + ((size_t)band * 2 + bl) * maxjj * VS) + (long long)v * V4 + gl : nullptr;

      float4 c0[NJ], c1[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) c0[j] = c1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowok && v >= my_lo && v <= my_hi) load_costs<NJ, G>(c0, cp - inc4);

      auto step = [&](float4 (&cc)[NJ], float4 (&cn)[NJ]) {
         if (rowok && v + 1 >= my_lo && v + 1 <= my_hi) load_costs<NJ, G>(cn, cp);
         if (cc_pf > 0) {
            const int vp = v + 1 + cc_pf;
            if (rowok && vp >= my_lo && vp <= my_hi) {
               const float *line = reinterpret_cast<const float *>(cp - gl + (long long)cc_pf * inc4);
               for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + l * 32));
            }
         }
         if (waiter) {   // position v-1 of the virtual workers is read in this step
            const int need = min(pf_hi, v - 1);
            while (vw <= need) {
               const int sl = vw & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
               ++vw;
            }
         }
         if (rowok && v >= my_lo && v <= my_hi) {
            const int xs = u - v;
            const bool border = (xs == 0) | (v == 0) | (xs == maxii - 1);
            float m = MGM_INF;
            if (border) {
#pragma unroll
               for (int j = 0; j < NJ; ++j) m = hmin4(m, cc[j]);
            } else {
               const int po = ((v - 1) & 1) * SLOT2;            // ring slot of position v-1 in a real row
               const int vo = ((v - 1) & (RV - 1)) * SLOT2;     // ... in the virtual workers' rings
               const float2 *S3[3] = {ownb + po + 2 * gl, p1b + (v1 ? vo : po) + 2 * gl, p2b + (v2 ? vo : po) + 2 * gl};
               const float2 *S[K];
#pragma unroll
               for (int k = 0; k < K; ++k) S[k] = S3[k];
               m = gather_sgm<K, NJ, G>(cc, S);
            }
            float4 *gp = reinterpret_cast<float4 *>(nslabs > 1 ? D.ldir[__umulhi((unsigned)yimg, P.slab_magic)] : D.ldir[0]) + goff;
            finish_pixel<K, NJ, G>(cc, m, gp, ownb + (v & 1) * SLOT2, gb, gl, gmask, p1, p2);
         }
         ++v;
         cp += inc4;
         goff += inc4;
         yimg += g.ydys - g.ydxs;
         if (bline) gb += V4;
         compute_barrier(ncomp);
         if (tid == 0) sts_release(s_step, v);
      };
      const int nst = se - sb + 1;
      for (int i = 0; i < nst; i += 2) {
         step(c0, c1);
         if (i + 1 < nst) step(c1, c0);
      }
      if (waiter) {
         while (vw <= pf_hi) {   // every loaded position is waited for: the parities stay in step with the barriers
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
template <int K, int GL, int NJ>
__global__ void __launch_bounds__(MGM_AGG_MAX_THREADS, 1) mgm_aggregate_sgm_kernel(const AggParams P) {
   extern __shared__ __align__(128) unsigned char smem[];
   __shared__ int2 s_ticket;
   __shared__ int s_step;
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
         if (D.pass >= 8) run_band_sgm<K, GL, NJ, 2>(P, D, pb.y, smem, &s_step);
         else if (D.pass < 4) run_band_sgm<K, GL, NJ, 0>(P, D, pb.y, smem, &s_step);
         else if constexpr (K <= 3) run_band_shear_sgm<K, GL, NJ>(P, D, pb.y, smem, &s_step);
         else run_band_sgm<K, GL, NJ, 1>(P, D, pb.y, smem, &s_step);
      }
      __syncthreads();
   }
}

// ---------------------------------------------------------------- host side
template <int K, int GL, int NJ>
static cudaError_t launch_lean(const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   auto kern = mgm_aggregate_sgm_kernel<K, GL, NJ>;
   cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.block, plan.smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   int grid = min(P.nbands, plan.num_sms * per_sm);
   if (grid < 1) grid = 1;
   if (plan.verbose)
      fprintf(stderr, "[mgmb200] aggregate (lean SGM K=%d lanes=%d chunks=%d): grid=%d block=%d smem=%zu CTAs/SM=%d sweeps=%d bands=%d rows=%d/%d\n",
              K, GL, NJ, grid, plan.block, plan.smem, per_sm, P.nsweeps, P.nbands, plan.T[0], plan.T[1]);
   kern<<<grid, plan.block, plan.smem, st>>>(P);
   return cudaGetLastError();
}

template <int GL, int NJ>
static cudaError_t launch_lean_k(int K, const AggParams &P, const AggPlan &plan, cudaStream_t st) {
#ifdef MGM_QUICK_K   // development builds: one TSGM value only
   if (K != MGM_QUICK_K) return cudaErrorNotSupported;
   return launch_lean<MGM_QUICK_K, GL, NJ>(P, plan, st);
#else
   switch (K) {
   case 1: return launch_lean<1, GL, NJ>(P, plan, st);
   case 2: return launch_lean<2, GL, NJ>(P, plan, st);
   case 3: return launch_lean<3, GL, NJ>(P, plan, st);
   default: return launch_lean<4, GL, NJ>(P, plan, st);
   }
#endif
}

// chunk counts per lane the lean kernels are built for
bool agg_sgm_lean_supported(int VS, int lanes) {
   if (VS % (4 * lanes)) return false;
   const int nj = VS / (4 * lanes);
   return lanes == 8 ? (nj == 2 || nj == 4 || nj == 6 || nj == 8) : (lanes == 4 && (nj == 4 || nj == 8));
}

cudaError_t agg_launch_sgm_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st) {
   const int nj = plan.VS / (4 * plan.lanes);
   if (plan.lanes == 8) {
      switch (nj) {
      case 2: return launch_lean_k<8, 2>(K, P, plan, st);
      case 4: return launch_lean_k<8, 4>(K, P, plan, st);
      case 6: return launch_lean_k<8, 6>(K, P, plan, st);
      case 8: return launch_lean_k<8, 8>(K, P, plan, st);
      }
   } else if (plan.lanes == 4) {
      switch (nj) {
      case 4: return launch_lean_k<4, 4>(K, P, plan, st);
      case 8: return launch_lean_k<4, 8>(K, P, plan, st);
      }
   }
   return cudaErrorNotSupported;
}

}  // namespace mgm
