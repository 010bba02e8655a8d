// K4 -- MGM aggregation sweeps (replaces mgm_core.cc:489-580 with the message
// updates of :66-281) as ONE persistent sm_100a kernel running all requested
// sweeps concurrently.
//
// Formulation (checked against the reference by
// oracle/mgm_oracle.c): in the scan space of a sweep every pixel (xs,ys) reads the
// predecessors {(-1,0),(0,-1),(-1,-1),(+1,-1)}; sweeps 0-3 take them in that order,
// sweeps 4-7 in the order {(+1,-1),(-1,-1),(0,-1),(-1,0)}, TSGM=K keeps the first K.
// Pixels with xs==0, ys==0 or xs==maxii-1 keep their matching cost (mgm_core.cc:538-541).
//
// Mapping to the machine:
//   * one THREAD per scan row, sequential in xs; rows of a band of T consecutive
//     rows run in lock step, row t trailing row t-1 by SIGMA pixels (1, or 2 when
//     the (+1,-1) predecessor is used), one __syncthreads per pixel step;
//   * all per-pixel label vectors live in shared memory: a ring of SIGMA+2 message
//     slots per row, and two cost buffers per row that are filled by TMA bulk loads
//     (cp.async.bulk + mbarrier) one pixel ahead and drained by TMA bulk stores of
//     the finished message into the per-sweep volume;
//   * the unweighted paths store in the ring the neighbour-side transform of the
//     message (SGM min3 / truncated-linear min-convolution, minus its minimum),
//     computed ONCE by the producer row instead of once per consumer;
//   * bands are chained through a global boundary line (TMA store -> release flag
//     -> acquire -> TMA load into a virtual "row -1" ring), and claimed from an
//     atomic ticket in dependency order so any grid size is deadlock free.
// Precondition of this fast path (enforced by the host): every cost
// vector holds a finite value, no NaN / -INF, P1,P2 >= 0 and weights >= 0; then the
// hardware min is bit-identical to the reference's compare-select forms.
#include "aggregate.cuh"

namespace mgm {

static constexpr int RV = 8;   // virtual-row ring (pixels of the previous band's last row)
static constexpr int PF = 3;   // boundary prefetch distance in pixels

enum PredType { PRED_SAME = 0, PRED_UP = 1, PRED_UPL = 2, PRED_UPR = 3 };

template <bool DIAG>
__device__ __forceinline__ constexpr int pred_type(int k) {
   return DIAG ? (k == 0 ? PRED_UPR : k == 1 ? PRED_UPL : k == 2 ? PRED_UP : PRED_SAME)
               : (k == 0 ? PRED_SAME : k == 1 ? PRED_UP : k == 2 ? PRED_UPL : PRED_UPR);
}

__device__ __forceinline__ float hmin4(float m, const float4 &v) {
   return fminf(fminf(fminf(m, v.x), fminf(v.y, v.z)), v.w);
}

// SGM neighbour transform of one label: min3(L(o), min(L(o-1),L(o+1))+P1, m+P2) - m   (mgm_core.cc:113-116)
__device__ __forceinline__ float sgm_x(float l, float c, float r, float p1, float cap, float m) {
   return fminf(fminf(c, fminf(l, r) + p1), cap) - m;
}

template <int POT, int K, bool WEIGHTED, bool DIAG>
__device__ void run_band(const AggParams &P, const int pass, const int band, unsigned char *smem) {
   constexpr int SIG = (DIAG || K == 4) ? 2 : 1;
   constexpr int R = SIG + 2;
   constexpr bool NEEDM = WEIGHTED || (POT == POT_TRUNC && K == 2);
   constexpr int CLS = DIAG ? 1 : 0;

   const PassGeom g = pass_geometry(pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int T = P.T[CLS];
   const int TS = P.TS[CLS];
   const int VS = P.VS;
   const int nq = VS >> 2;
   const uint32_t vbytes = (uint32_t)VS * 4u;
   const int t = threadIdx.x;
   const int ncomp = blockDim.x - 32;
   const bool is_prod = (t == ncomp);
   const int row0 = band * T;
   const int nrows = min(T, maxjj - row0);
   const int ys = row0 + t;
   const bool rowok = (t < nrows);
   const bool has_prev = band > 0;
   const bool has_next = row0 + T < maxjj;
   const bool last_row = rowok && (t == nrows - 1) && has_next;
   const int nsteps = maxii + SIG * (nrows - 1);

   uint64_t *cbar = reinterpret_cast<uint64_t *>(smem + P.off_cbar);
   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);   // persistent mbarrier parities

   float *mybase = thr + (size_t)t * TS;                    // ring slots [R][VS]
   float *mycb = mybase + R * VS;                            // cost buffers [2][VS]
   float *myscr = mycb + 2 * VS;                             // scratch (weighted truncated-linear only)
   const float *upbase = mybase - TS;                        // row t-1 (t>0)

   const float *ccv = P.cc;
   float *ldir = P.ldir[pass];
   float *bnd_out = P.bnd[pass] + (size_t)band * maxii * VS;            // written by this band's last row
   float *bndm_out = P.bndm[pass] + (size_t)band * maxii;
   const float *bnd_in = has_prev ? P.bnd[pass] + (size_t)(band - 1) * maxii * VS : nullptr;
   const float *bndm_in = has_prev ? P.bndm[pass] + (size_t)(band - 1) * maxii : nullptr;
   int *prog_out = P.progress[pass] + band;
   const int *prog_in = has_prev ? P.progress[pass] + band - 1 : nullptr;

   const long long pix0 = g.base0 + (long long)ys * g.dys;   // pixel of (0, ys)

   uint32_t cph = 0, vph = 0;
   if (t < ncomp) cph = phase[t];
   if (t == 0) vph = phase[ncomp];
   int next_px = 0;   // producer: next boundary pixel to fetch

   // prologue: first cost vector of every row, first boundary pixels
   if (rowok) {
      fence_proxy_async_smem();
      mbar_expect_tx(&cbar[2 * t], vbytes);
      tma_load_1d(mycb, ccv + (size_t)pix0 * VS, vbytes, &cbar[2 * t]);
   }

   for (int s = 0; s < nsteps; ++s) {
      if (is_prod && has_prev) {
         const int lim = min(maxii - 1, s + 1 + PF);
         while (next_px <= lim) {
            while (ld_acquire(prog_in) < next_px + 1) __nanosleep(40);
            fence_proxy_async();
            const int slot = next_px & (RV - 1);
            if (NEEDM) vms[slot] = __ldcg(bndm_in + next_px);
            mbar_expect_tx(&vbar[slot], vbytes);
            tma_load_1d(virt + slot * VS, bnd_in + (size_t)next_px * VS, vbytes, &vbar[slot]);
            ++next_px;
         }
      }
      const int xs = s - SIG * t;
      if (rowok && xs >= 0 && xs < maxii) {
         const long long pix = pix0 + (long long)xs * g.dxs;
         // (a) my earlier bulk stores no longer read shared memory; publish the boundary progress
         if (last_row) {
            tma_wait_all<0>();
            if (xs > 0) {
               fence_proxy_async();
               __threadfence();
               st_release(prog_out, xs);   // pixels [0,xs) of the boundary row are in global memory
            }
         } else {
            tma_wait_read<0>();
         }
         // (b) prefetch the next cost vector of this row
         if (xs + 1 < maxii) {
            const int nb = (xs + 1) & 1;
            fence_proxy_async_smem();
            mbar_expect_tx(&cbar[2 * t + nb], vbytes);
            tma_load_1d(mycb + nb * VS, ccv + (size_t)(pix + g.dxs) * VS, vbytes, &cbar[2 * t + nb]);
         }
         // (c) this pixel's cost vector
         const int cbi = xs & 1;
         mbar_wait(&cbar[2 * t + cbi], (cph >> cbi) & 1u);
         cph ^= (1u << cbi);
         float4 *Cb = reinterpret_cast<float4 *>(mycb + cbi * VS);
         float4 *cur = reinterpret_cast<float4 *>(mybase + (xs % R) * VS);
         const bool border = (xs == 0) || (ys == 0) || (xs == maxii - 1);

         // virtual row: make sure the boundary pixels this step reads have landed
         if (t == 0 && has_prev) {
            if (xs == 1) {
               for (int px = 0; px <= min(2, maxii - 1); ++px) {
                  mbar_wait(&vbar[px & (RV - 1)], (vph >> (px & (RV - 1))) & 1u);
                  vph ^= 1u << (px & (RV - 1));
               }
            } else if (xs >= 2 && xs + 1 < maxii) {
               const int sl = (xs + 1) & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
            }
         }

         float m = MGM_INF;
         if (border) {
            for (int q = 0; q < nq; ++q) {
               float4 c = Cb[q];
               m = hmin4(m, c);
               cur[q] = c;
            }
         } else {
            // predecessor slots
            const float4 *S[K];
            float mk[K], wk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
               const int pt = pred_type<DIAG>(k);
               const int px = (pt == PRED_UP) ? xs : (pt == PRED_UPR ? xs + 1 : xs - 1);
               const bool own = (pt == PRED_SAME);
               const float *sp;
               if (own) sp = mybase + (px % R) * VS;
               else if (t == 0) sp = virt + (px & (RV - 1)) * VS;
               else sp = upbase + (px % R) * VS;
               S[k] = reinterpret_cast<const float4 *>(sp);
               mk[k] = 0.f; wk[k] = 1.f;
               if (NEEDM) mk[k] = own ? msr[t * 4 + (px % R)] : (t == 0 ? vms[px & (RV - 1)] : msr[(t - 1) * 4 + (px % R)]);
               if (WEIGHTED) wk[k] = __ldg(P.w + (size_t)pass_weight_plane(pass, k) * P.nx * P.ny + pix);
            }

            if constexpr (!WEIGHTED) {
               // ring slots already hold the producer-side transform
               for (int q = 0; q < nq; ++q) {
                  float4 c = Cb[q];
                  float4 a[K];
#pragma unroll
                  for (int k = 0; k < K; ++k) a[k] = S[k][q];
                  float4 o;
                  if constexpr (POT == POT_TRUNC && K == 2) {   // update_cost2_trunclinear association, mgm_core.cc:216
                     o.x = c.x + (((a[0].x - mk[0]) + a[1 % K].x) - mk[1 % K]) * 0.5f;
                     o.y = c.y + (((a[0].y - mk[0]) + a[1 % K].y) - mk[1 % K]) * 0.5f;
                     o.z = c.z + (((a[0].z - mk[0]) + a[1 % K].z) - mk[1 % K]) * 0.5f;
                     o.w = c.w + (((a[0].w - mk[0]) + a[1 % K].w) - mk[1 % K]) * 0.5f;
                  } else if constexpr (POT == POT_SGM && K == 2) {   // update_cost2: halves were taken by the producer
                     o.x = c.x + (a[0].x + a[1 % K].x);
                     o.y = c.y + (a[0].y + a[1 % K].y);
                     o.z = c.z + (a[0].z + a[1 % K].z);
                     o.w = c.w + (a[0].w + a[1 % K].w);
                  } else {
                     float4 e = a[0];
#pragma unroll
                     for (int k = 1; k < K; ++k) { e.x += a[k].x; e.y += a[k].y; e.z += a[k].z; e.w += a[k].w; }
                     o.x = c.x + div_by_k<K>(e.x);
                     o.y = c.y + div_by_k<K>(e.y);
                     o.z = c.z + div_by_k<K>(e.z);
                     o.w = c.w + div_by_k<K>(e.w);
                  }
                  m = hmin4(m, o);
                  Cb[q] = o;
                  cur[q] = o;
               }
            } else if constexpr (POT == POT_SGM) {
               // update_costW with per-edge weights, mgm_core.cc:95-144
               float pw[K], cap[K], lft[K];
               float4 cv[K];
#pragma unroll
               for (int k = 0; k < K; ++k) {
                  pw[k] = P.P1 * wk[k];
                  cap[k] = mk[k] + P.P2 * wk[k];
                  lft[k] = MGM_INF;
                  cv[k] = S[k][0];
               }
               for (int q = 0; q < nq; ++q) {
                  float4 c = Cb[q];
                  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int k = 0; k < K; ++k) {
                     float4 v = cv[k];
                     float4 nx4 = (q + 1 < nq) ? S[k][q + 1] : make_float4(MGM_INF, MGM_INF, MGM_INF, MGM_INF);
                     e.x += sgm_x(lft[k], v.x, v.y, pw[k], cap[k], mk[k]);
                     e.y += sgm_x(v.x, v.y, v.z, pw[k], cap[k], mk[k]);
                     e.z += sgm_x(v.y, v.z, v.w, pw[k], cap[k], mk[k]);
                     e.w += sgm_x(v.z, v.w, nx4.x, pw[k], cap[k], mk[k]);
                     lft[k] = v.w;
                     cv[k] = nx4;
                  }
                  float4 o;
                  o.x = c.x + div_by_k<K>(e.x);
                  o.y = c.y + div_by_k<K>(e.y);
                  o.z = c.z + div_by_k<K>(e.z);
                  o.w = c.w + div_by_k<K>(e.w);
                  m = hmin4(m, o);
                  Cb[q] = o;
                  cur[q] = o;
               }
            } else {
               // update_costW_trunclinear with per-edge slopes/caps, mgm_core.cc:229-281
               float *scr = myscr;
               float *E = reinterpret_cast<float *>(cur);
#pragma unroll
               for (int k = 0; k < K; ++k) {
                  const float pw = P.P1 * wk[k];
                  const float p2 = P.P2 * wk[k];
                  const float capv = mk[k] + p2;
                  const bool trunc = p2 < MGM_INF;
                  const float *src = reinterpret_cast<const float *>(S[k]);
                  float run = MGM_INF;
                  for (int o = 0; o < VS; ++o) {   // forward pass (:154-155)
                     run = fminf(run + pw, src[o]);
                     scr[o] = run;
                  }
                  run = MGM_INF;
                  for (int o = VS - 1; o >= 0; --o) {   // backward pass (:157-158) + truncation (:160-162)
                     run = fminf(run + pw, scr[o]);
                     float v = trunc ? fminf(run, capv) : run;
                     v = v - mk[k];
                     E[o] = (k == 0) ? v : E[o] + v;
                  }
               }
               for (int q = 0; q < nq; ++q) {
                  float4 c = Cb[q];
                  float4 e = cur[q];
                  float4 o;
                  o.x = c.x + div_by_k<K>(e.x);
                  o.y = c.y + div_by_k<K>(e.y);
                  o.z = c.z + div_by_k<K>(e.z);
                  o.w = c.w + div_by_k<K>(e.w);
                  m = hmin4(m, o);
                  Cb[q] = o;
                  cur[q] = o;
               }
            }
         }

         // (d) the finished message goes to this sweep's volume (smem -> HBM, TMA)
         fence_proxy_async_smem();
         tma_store_1d(ldir + (size_t)pix * VS, Cb, vbytes);
         tma_commit();

         // (e) producer-side transform of the message for its successors (unweighted paths)
         if (NEEDM) msr[t * 4 + (xs % R)] = m;
         if constexpr (!WEIGHTED) {
            if constexpr (POT == POT_SGM) {
               const float p1 = P.P1;
               const float cap = m + P.P2;
               const float sc = (K == 2) ? 0.5f : 1.0f;
               float lft = MGM_INF;
               float4 v = cur[0];
               for (int q = 0; q < nq; ++q) {
                  float4 nx4 = (q + 1 < nq) ? cur[q + 1] : make_float4(MGM_INF, MGM_INF, MGM_INF, MGM_INF);
                  float4 a;
                  a.x = sgm_x(lft, v.x, v.y, p1, cap, m) * sc;
                  a.y = sgm_x(v.x, v.y, v.z, p1, cap, m) * sc;
                  a.z = sgm_x(v.y, v.z, v.w, p1, cap, m) * sc;
                  a.w = sgm_x(v.z, v.w, nx4.x, p1, cap, m) * sc;
                  lft = v.w;
                  cur[q] = a;
                  v = nx4;
               }
            } else {
               // minConvTruncatedLinear (mgm_core.cc:152-163) in place, sequential like the reference
               const float p1 = P.P1;
               const float capv = m + P.P2;
               const bool trunc = P.P2 < MGM_INF;
               const float sub = (K == 2) ? 0.0f : m;   // A10 subtracts on the consumer side
               float run = MGM_INF;
               for (int q = 0; q < nq; ++q) {
                  float4 v = cur[q];
                  v.x = run = fminf(run + p1, v.x);
                  v.y = run = fminf(run + p1, v.y);
                  v.z = run = fminf(run + p1, v.z);
                  v.w = run = fminf(run + p1, v.w);
                  cur[q] = v;
               }
               run = MGM_INF;
               for (int q = nq - 1; q >= 0; --q) {
                  float4 v = cur[q];
                  run = fminf(run + p1, v.w); v.w = (trunc ? fminf(run, capv) : run) - sub;
                  run = fminf(run + p1, v.z); v.z = (trunc ? fminf(run, capv) : run) - sub;
                  run = fminf(run + p1, v.y); v.y = (trunc ? fminf(run, capv) : run) - sub;
                  run = fminf(run + p1, v.x); v.x = (trunc ? fminf(run, capv) : run) - sub;
                  cur[q] = v;
               }
            }
         }

         // (f) hand the boundary row to the next band
         if (last_row) {
            if (NEEDM) bndm_out[xs] = m;
            fence_proxy_async_smem();
            tma_store_1d(bnd_out + (size_t)xs * VS, cur, vbytes);
            tma_commit();
         }
      }
      __syncthreads();
   }

   // epilogue: drain stores, publish the full boundary row, save mbarrier parities
   tma_wait_all<0>();
   if (last_row) {
      fence_proxy_async();
      __threadfence();
      st_release(prog_out, maxii);
   }
   if (t < ncomp) phase[t] = cph;
   if (t == 0) phase[ncomp] = vph;
   __syncthreads();
}

template <int POT, int K, bool WEIGHTED>
__global__ void __launch_bounds__(MGM_AGG_MAX_THREADS, 1) mgm_aggregate_kernel(const AggParams P) {
   extern __shared__ __align__(128) unsigned char smem[];
   __shared__ int s_ticket;
   const int t = threadIdx.x;
   const int ncomp = blockDim.x - 32;

   // one-time barrier setup: two cost barriers per row thread, RV boundary barriers
   {
      uint64_t *cbar = reinterpret_cast<uint64_t *>(smem + P.off_cbar);
      uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
      uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);
      if (t < ncomp) { mbar_init(&cbar[2 * t], 1); mbar_init(&cbar[2 * t + 1], 1); phase[t] = 0; }
      if (t == ncomp) { for (int i = 0; i < RV; ++i) mbar_init(&vbar[i], 1); phase[ncomp] = 0; }
      mbar_fence_init();
      __syncthreads();
   }

   for (;;) {
      if (t == 0) s_ticket = (int)atomicAdd(P.ticket_counter, 1u);
      __syncthreads();
      const int tk = s_ticket;
      __syncthreads();
      if (tk >= P.ntickets) break;
      const int2 pb = P.tickets[tk];
      if (pb.x < 4) run_band<POT, K, WEIGHTED, false>(P, pb.x, pb.y, smem);
      else run_band<POT, K, WEIGHTED, true>(P, pb.x, pb.y, smem);
   }
}

// ---------------------------------------------------------------- host side
static int ring_slots(int cls, int K) { return ((cls == 1 || K == 4) ? 2 : 1) + 2; }

template <int POT, int K, bool WEIGHTED>
static cudaError_t launch_t(const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   auto kern = mgm_aggregate_kernel<POT, K, WEIGHTED>;
   cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.block, plan.smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   int grid = min(P.ntickets, plan.num_sms * per_sm);
   if (grid < 1) grid = 1;
   kern<<<grid, plan.block, plan.smem, st>>>(P);
   return cudaGetLastError();
}

template <int POT, bool WEIGHTED>
static cudaError_t launch_k(int K, const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   switch (K) {
   case 1: return launch_t<POT, 1, WEIGHTED>(P, plan, st);
   case 2: return launch_t<POT, 2, WEIGHTED>(P, plan, st);
   case 3: return launch_t<POT, 3, WEIGHTED>(P, plan, st);
   default: return launch_t<POT, 4, WEIGHTED>(P, plan, st);
   }
}

void agg_plan(AggPlan *plan, int nx, int ny, int L, int K, int pot, bool weighted, int max_smem, int num_sms,
              int t_override) {
   const int VS = (L + 3) & ~3;
   plan->VS = VS;
   const int xtra = (weighted && pot == POT_TRUNC) ? 1 : 0;
   // shared memory budget -> rows per band, per sweep class
   int tmax = MGM_AGG_MAX_THREADS - 32;
   for (int iter = 0; iter < 2; ++iter) {
      int T[2];
      for (int cls = 0; cls < 2; ++cls) {
         int nbuf = ring_slots(cls, K) + 2 + xtra;
         int TS = nbuf * VS;
         if (((TS >> 2) & 1) == 0) TS += 4;   // odd number of 16-byte units: conflict-free LDS.128 across rows
         plan->TS[cls] = TS;
         size_t fixed = 1024 + (size_t)RV * VS * 4 + (size_t)tmax * (16 + 16 + 4) + RV * 16;
         long avail = (long)max_smem - (long)fixed;
         int Tc = (int)(avail / ((long)TS * 4));
         if (Tc > tmax) Tc = tmax;
         if (Tc < 1) Tc = 0;
         T[cls] = Tc;
      }
      plan->T[0] = T[0];
      plan->T[1] = T[1];
      if (t_override > 0) { plan->T[0] = min(plan->T[0], t_override); plan->T[1] = min(plan->T[1], t_override); }
      int tm = max(plan->T[0], plan->T[1]);
      int ncomp = (tm + 31) & ~31;
      if (ncomp == tmax || iter == 1) { plan->block = ncomp + 32; break; }
      tmax = ncomp;   // recompute the fixed part with the real thread count
   }
   const int ncomp = plan->block - 32;
   size_t off = 0;
   plan->off_phase = off; off += (size_t)(ncomp + 1) * 4; off = (off + 15) & ~(size_t)15;
   plan->off_cbar = off; off += (size_t)ncomp * 16;
   plan->off_vbar = off; off += RV * 8;
   plan->off_ms = off; off += (size_t)ncomp * 16;
   plan->off_vms = off; off += RV * 4; off = (off + 127) & ~(size_t)127;
   plan->off_virt = off; off += (size_t)RV * VS * 4; off = (off + 127) & ~(size_t)127;
   plan->off_thr = off;
   size_t per_thr = (size_t)max(plan->TS[0] * plan->T[0], plan->TS[1] * plan->T[1]) * 4;
   plan->smem = off + per_thr;
   plan->num_sms = num_sms;
}

cudaError_t agg_launch(const AggParams &P, const AggPlan &plan, int pot, int K, bool weighted, cudaStream_t st) {
   if (pot == POT_SGM)
      return weighted ? launch_k<POT_SGM, true>(K, P, plan, st) : launch_k<POT_SGM, false>(K, P, plan, st);
   return weighted ? launch_k<POT_TRUNC, true>(K, P, plan, st) : launch_k<POT_TRUNC, false>(K, P, plan, st);
}

}  // namespace mgm
