// K4sw -- lean band function of the SGM-potential aggregation WITH per-edge weights: update_costW with image-dependent
// weights (mgm_core.cc:95-144; weights from compute_mgm_weights, mgm_weights.h:63-85) inside the sweep loop (:489-580).
//
// Same structure as aggregate_sgm.cu (compile-time label layout, warp roles split, cost loads double-buffered in
// registers, band hand-off off the step barrier: the last row stores its boundary vector itself, the publisher only
// releases the progress counter, the boundary consumer runs ahead).  What the weights change: the neighbour term
//    min3(L_q(o), min(L_q(o-1), L_q(o+1)) + P1*w, m_q + P2*w) - m_q
// depends on the edge (p,q), so it cannot be formed once by the producer of L_q: the ring holds the RAW messages and
// their minima, every consumer transforms its K predecessors with its own weights (read AT the pixel, plane of the
// neighbour's offset: pass_weight_plane, mgm_core.cc:481-484,550-554), and every sweep runs row-per-worker (the sheared
// wavefront of the unweighted diagonal sweeps relies on the producer-side transform).  Arithmetic and its order are
// those of the generic kernel's weighted branch (aggregate.cu run_band): bit-identical, tests run both.
// Sweeps 8-15 and the per-pixel windows of the truncated-linear potentials stay on the generic kernel.
#include "aggregate_dev.cuh"

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

template <int NJ>
__device__ __forceinline__ void load_costs(float4 (&c)[NJ], const float4 *p) {
#pragma unroll
   for (int j = 0; j < NJ; ++j) c[j] = __ldcs(p + G * j);
}

// MODE 0: sweeps 0-3 (predecessors (-1,0), (0,-1), (-1,-1), (+1,-1); lag 1, lag 2 with TSGM = 4);
// MODE 1: sweeps 4-7 row-per-worker (predecessors (+1,-1), (-1,-1), (0,-1), (-1,0); lag 2)
template <int K, int NJ, int MODE>
__device__ void run_band_sgmw(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem, int *s_step) {
   constexpr bool DIAG = (MODE == 1);
   constexpr int SIG = (DIAG || K == 4) ? 2 : 1;
   constexpr int R = SIG + 2;
   constexpr int CLS = DIAG ? CLS_DIAG : CLS_AXIS;
   constexpr int VS = 4 * G * NJ;
   constexpr int V4 = VS / 4;       // float4 per vector
   constexpr int SLOT2 = VS / 2;    // float2 per ring slot
   constexpr uint32_t vbytes = (uint32_t)VS * 4u;

   const int pass = D.pass;
   const PassGeom g = pass_geometry(pass, P.nx, P.ny);
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
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);      // [row][4] minima of the ring slots
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);     // [RV] minima of the virtual row
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
         const float *bndm_in = D.bndm + (size_t)(band - 1) * maxii;
         const int *prog_in = D.progress + band - 1;
         int avail = 0;
         for (int px = 0; px < maxii; ++px) {
            while (lds_acquire(s_step) < px - (RV - 2)) __nanosleep(20);   // slot of pixel px-RV: last read in step px-RV+1
            while (avail < px + 1) {
               avail = ld_acquire(prog_in);
               if (avail < px + 1) __nanosleep(20);
            }
            fence_proxy_async();
            const int sl = px & (RV - 1);
            vms[sl] = __ldcg(bndm_in + px);   // ordered before the waiters' reads by the arrive / wait pair
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
      float2 *ownb = reinterpret_cast<float2 *>(rowf);                    // my row's ring of RAW messages (8-byte aligned rows)
      const bool upvirt = (r == 0);                                        // row -1 = the previous band's last row
      const float2 *upb = upvirt ? reinterpret_cast<const float2 *>(virt) : reinterpret_cast<const float2 *>(rowf - TS);
      const float *upm = upvirt ? vms : msr + (r - 1) * 4;
      const bool waiter = upvirt && has_prev;
      const bool bline = has_next && r == nrows - 1;
      const float P1 = P.P1, P2 = P.P2;
      const int cc_pf = P.cc_pf;
      uint32_t vph = waiter ? phase[vph_idx] : 0u;
      int vw = 0;   // next virtual pixel to wait for

      int xs = -SIG * r;
      int si = ((xs % R) + R) % R;   // ring slot of pixel xs
      const long long pix0 = g.base0 + (long long)ys * g.dys;
      const long long inc4 = g.dxs * V4;
      const float4 *cp = reinterpret_cast<const float4 *>(D.cc) + (pix0 + (long long)(xs + 1) * g.dxs) * V4 + gl;   // pixel xs+1
      long long goff = (pix0 + (long long)xs * g.dxs) * V4 + gl;
      const int nslabs = P.nslabs;
      int yimg = g.y0 + xs * g.ydxs + ys * g.ydys;
      float4 *gb = bline ? reinterpret_cast<float4 *>(D.bnd + (size_t)band * maxii * VS) + (long long)xs * V4 + gl : nullptr;
      float *gbm = bline ? D.bndm + (size_t)band * maxii + xs : nullptr;
      // weights of the K edges of a pixel: plane of neighbour k, read at the pixel
      const size_t wplane = (size_t)P.nx * P.ny;
      const float *wp[K];
#pragma unroll
      for (int k = 0; k < K; ++k) wp[k] = D.w + (size_t)pass_weight_plane(pass, k) * wplane + (pix0 + (long long)(xs + 1) * g.dxs);   // pixel xs+1

      float4 c0[NJ], c1[NJ];
      float w0[K], w1[K];
#pragma unroll
      for (int j = 0; j < NJ; ++j) c0[j] = c1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < K; ++k) w0[k] = w1[k] = 1.f;
      if (rowok && xs == 0) {   // row 0 starts right away
         load_costs<NJ>(c0, cp - inc4);
#pragma unroll
         for (int k = 0; k < K; ++k) w0[k] = __ldg(wp[k] - g.dxs);
      }

      int s = 0;
      auto step = [&](float4 (&cc)[NJ], float4 (&cn)[NJ], float (&wc)[K], float (&wn)[K]) {
         // costs and weights of the pixel of the NEXT step: in flight during the whole step
         if (rowok && (unsigned)(xs + 1) < (unsigned)maxii) {
            load_costs<NJ>(cn, cp);
#pragma unroll
            for (int k = 0; k < K; ++k) wn[k] = __ldg(wp[k]);
         }
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
               const int si_prev = (si == 0) ? R - 1 : si - 1;
               const int si_next = (si == R - 1) ? 0 : si + 1;
               const int ui = upvirt ? (xs & (RV - 1)) : si;
               const int ui_prev = upvirt ? ((xs - 1) & (RV - 1)) : si_prev;
               const int ui_next = upvirt ? ((xs + 1) & (RV - 1)) : si_next;
               (void)ui_next; (void)ui; (void)ui_prev;
               const float2 *S[K];
               float mk[K], pw[K], cap[K];
#pragma unroll
               for (int k = 0; k < K; ++k) {
                  const int pt = pred_type<DIAG>(k);
                  const int us = (pt == PRED_UP) ? ui : (pt == PRED_UPL) ? ui_prev : ui_next;
                  S[k] = (pt == PRED_SAME) ? ownb + si_prev * SLOT2 : upb + us * SLOT2;
                  mk[k] = (pt == PRED_SAME) ? msr[r * 4 + si_prev] : upm[us];
                  pw[k] = P1 * wc[k];
                  cap[k] = mk[k] + P2 * wc[k];
               }
#pragma unroll
               for (int j = 0; j < NJ; ++j) {
                  const int q = gl + G * j;
                  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int k = 0; k < K; ++k) {
                     const float *sf = reinterpret_cast<const float *>(S[k]);
                     const float4 v = ld16<0>(S[k], q);
                     const float lft = (q > 0) ? sf[4 * q - 1] : MGM_INF;
                     const float rgt = (q + 1 < NJ * G) ? sf[4 * q + 4] : MGM_INF;
                     e.x += sgm_x(lft, v.x, v.y, pw[k], cap[k], mk[k]);
                     e.y += sgm_x(v.x, v.y, v.z, pw[k], cap[k], mk[k]);
                     e.z += sgm_x(v.y, v.z, v.w, pw[k], cap[k], mk[k]);
                     e.w += sgm_x(v.z, v.w, rgt, pw[k], cap[k], mk[k]);
                  }
                  const float4 o = add4(cc[j], div4_by_k<K>(e));
                  cc[j] = o;
                  m = hmin4(m, o);
               }
            }
            // the message: to the sweep's volume, raw into my ring slot, and (last row) into the boundary line
            float4 *gp = reinterpret_cast<float4 *>(nslabs > 1 ? D.ldir[__umulhi((unsigned)yimg, P.slab_magic)] : D.ldir[0]) + goff;
            float2 *cur = ownb + si * SLOT2;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
               __stcs(gp + G * j, cc[j]);
               st16<0>(cur, gl + G * j, cc[j]);
               if (gb) gb[G * j] = cc[j];
            }
#pragma unroll
            for (int d = 1; d < G; d <<= 1) m = fminf(m, __shfl_xor_sync(gmask, m, d));
            if (gl == 0) {
               msr[r * 4 + si] = m;
               if (gbm) *gbm = m;
            }
         }
         ++xs;
         ++s;
         cp += inc4;
         goff += inc4;
         yimg += g.ydxs;
#pragma unroll
         for (int k = 0; k < K; ++k) wp[k] += g.dxs;
         if (bline) { gb += V4; gbm += 1; }
         si = (si == R - 1) ? 0 : si + 1;
         compute_barrier(ncomp);
         if (tid == 0) sts_release(s_step, s);   // steps [0, s) are complete
      };
      for (int i = 0; i < nsteps; i += 2) {
         step(c0, c1, w0, w1);
         if (i + 1 < nsteps) step(c1, c0, w1, w0);
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

// The persistent kernel: same claim loop and finish tiles as mgm_aggregate_kernel (aggregate.cu).
template <int K, int NJ>
__global__ void __launch_bounds__(MGM_AGG_MAX_THREADS, 1) mgm_aggregate_sgmw_kernel(const AggParams P) {
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
         if (D.pass < 4) run_band_sgmw<K, NJ, 0>(P, D, pb.y, smem, &s_step);
         else run_band_sgmw<K, NJ, 1>(P, D, pb.y, smem, &s_step);
      }
      __syncthreads();
   }
}

template <int K, int NJ>
cudaError_t launch_lean(const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   auto kern = mgm_aggregate_sgmw_kernel<K, NJ>;
   cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.block, plan.smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   int grid = min(P.nbands, plan.num_sms * per_sm);
   if (grid < 1) grid = 1;
   if (plan.verbose)
      fprintf(stderr, "[mgmb200] aggregate (lean SGM with per-edge weights K=%d chunks=%d): grid=%d block=%d smem=%zu sweeps=%d bands=%d rows=%d/%d\n",
              K, NJ, grid, plan.block, plan.smem, P.nsweeps, P.nbands, plan.T[0], plan.T[1]);
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

cudaError_t agg_launch_sgmw_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st) {
   switch (plan.VS / 32) {
   case 2: return launch_lean_k<2>(K, P, plan, st);
   case 4: return launch_lean_k<4>(K, P, plan, st);
   case 6: return launch_lean_k<6>(K, P, plan, st);
   case 8: return launch_lean_k<8>(K, P, plan, st);
   }
   return cudaErrorNotSupported;
}

}  // namespace mgm
