// Compare-select aggregation sweeps for cost volumes OUTSIDE the fast path's envelope: vectors without any finite
// entry, NaN or -INF costs, negative / non-finite edge weights or penalties.  The reference has no such envelope: its
// minima are compare-and-select macros (mgm_core.cc:47-60) and its vector minimum a '<' scan (dvec.cc:81-88), so
// non-finite values propagate by plain IEEE rules (INF - INF = NaN at all-INF neighbours, NaN costs that never win a
// '<' test, ...).  This kernel restates the four update functions (mgm_core.cc:66-281) with exactly those forms --
// the hardware FMNMX used by mgm_aggregate_kernel differs from them as soon as a NaN is involved -- and is selected by
// the host-pointer entry points (mgmb200_mgm, mgmb200_mgm_labelmajor) when their input scan finds such values.
// It is a correctness path, not a fast one: one warp per scan row, lanes over labels, rows chained through
// release/acquire progress counters, predecessors read back from the sweep's message volume through L2.
#include "aggregate.cuh"

namespace mgm {

struct GenericParams {
   const float *cc;             // [ny][nx][VS]
   const float *w;              // 8 planes or nullptr
   float *ldir[MGM_MAX_NDIR];   // message volumes [ny][nx][VS]
   int *progress;               // [nsel][maxrows] finished pixels of each scan row
   int *ticket;                 // row claim counter
   float *scratch;              // [warps][4][L] min-convolution work rows
   int sel[MGM_MAX_NDIR];       // selected sweeps
   int nsel, maxrows;
   int nx, ny, L, VS, K, variant;
   float P1, P2;
};

__device__ __forceinline__ int generic_pred_type(int pass, int k, int xs, int ys) {
   if (pass >= 8) return knight_pred_type((pass & 7) >= 4, k, xs, ys);
   return (pass >= 4) ? 3 - k : k;   // sweeps 0-3: SAME, UP, UPL, UPR; sweeps 4-7: UPR, UPL, UP, SAME
}

// dvec.cc:81-88 get_minvalue: '<' scan from +INF (a NaN never wins); lanes scan their labels, then combine
__device__ __forceinline__ float generic_vec_min(const float *a, int L, int lane) {
   float m = MGM_INF;
   for (int o = lane; o < L; o += 32) { const float x = __ldcg(a + o); if (x < m) m = x; }
#pragma unroll
   for (int d = 16; d > 0; d >>= 1) { const float x = __shfl_xor_sync(0xffffffffu, m, d); if (x < m) m = x; }
   return m;
}
__device__ __forceinline__ float generic_at(const float *a, int L, int o) { return (o >= 0 && o < L) ? __ldcg(a + o) : MGM_INF; }

__global__ void __launch_bounds__(128) mgm_aggregate_generic_kernel(const GenericParams P) {
   const int lane = threadIdx.x & 31;
   const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   float *scratch = P.scratch + (size_t)gwarp * 4 * P.L;
   const size_t np = (size_t)P.nx * P.ny;
   const int L = P.L, K = P.K;
   for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(P.ticket, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      const int ys = t / P.nsel, si = t % P.nsel;   // rows of a sweep are claimed in increasing order: the row above is
      const int pass = P.sel[si];                   // running or finished, whatever the grid size
      const PassGeom g = pass_geometry(pass, P.nx, P.ny);
      if (ys >= P.maxrows) return;
      if (ys >= g.maxjj) continue;
      float *ldir = P.ldir[pass];
      int *prog = P.progress + (size_t)si * P.maxrows;
      int seen = 0;   // pixels of the row above known to be finished
      for (int xs = 0; xs < g.maxii; ++xs) {
         const long long pix = g.base0 + (long long)xs * g.dxs + (long long)ys * g.dys;
         const float *Cp = P.cc + (size_t)pix * P.VS;
         float *Lp = ldir + (size_t)pix * P.VS;
         const bool border = (xs == 0) || (ys == 0) || (xs == g.maxii - 1);   // all four neighbours inside (mgm_core.cc:538-541)
         if (border) {
            for (int o = lane; o < P.VS; o += 32) __stcg(Lp + o, o < L ? Cp[o] : MGM_INF);
         } else {
            // (xs+1, ys-1) must be finished: the row above has completed xs+2 pixels
            if (seen < xs + 2) {
               if (lane == 0) { while ((seen = ld_acquire(prog + ys - 1)) < xs + 2) __nanosleep(40); }
               seen = __shfl_sync(0xffffffffu, seen, 0);
            }
            const float *nb[4];
            float wk[4], m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
               const int pt = generic_pred_type(pass, k, xs, ys);
               const int pxs = (pt == PRED_UP) ? xs : (pt == PRED_UPR ? xs + 1 : xs - 1);
               const int pys = (pt == PRED_SAME) ? ys : ys - 1;
               nb[k] = ldir + (size_t)(g.base0 + (long long)pxs * g.dxs + (long long)pys * g.dys) * P.VS;
               wk[k] = P.w ? P.w[(size_t)pass_weight_plane_of_type(pass & 7, pt) * np + pix] : 1.0f;   // plane read AT p (:550-554)
               m[k] = (k < K) ? generic_vec_min(nb[k], L, lane) : 0.f;
            }
            if (P.variant == 0) {   // update_cost2 mgm_core.cc:66-90
               for (int o = lane; o < L; o += 32) {
                  float e = 0.f;
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                     const float a = __ldcg(nb[k] + o);
                     const float b = sel_min(generic_at(nb[k], L, o - 1), generic_at(nb[k], L, o + 1)) + P.P1;
                     const float c = m[k] + P.P2;
                     e += (min3_gt(a, b, c) - m[k]) / 2;
                  }
                  __stcg(Lp + o, Cp[o] + e);
               }
            } else if (P.variant == 1) {   // update_costW :95-144
               for (int o = lane; o < L; o += 32) {
                  float e = 0.f;
                  for (int k = 0; k < K; ++k) {
                     const float a = __ldcg(nb[k] + o);
                     const float b = sel_min(generic_at(nb[k], L, o - 1), generic_at(nb[k], L, o + 1)) + P.P1 * wk[k];
                     const float c = m[k] + P.P2 * wk[k];
                     e += min3_gt(a, b, c) - m[k];
                  }
                  __stcg(Lp + o, Cp[o] + e / K);
               }
            } else {
               // minConvTruncatedLinear :152-163 of each neighbour: lane k runs the sequential scans of neighbour k
               const int nk = (P.variant == 2) ? 2 : K;
               for (int k = 0; k < nk; ++k)
                  for (int o = lane; o < L; o += 32) scratch[(size_t)k * L + o] = __ldcg(nb[k] + o);
               __syncwarp();
               if (lane < nk) {
                  float *M = scratch + (size_t)lane * L;
                  const float p1 = (P.variant == 2) ? P.P1 : P.P1 * wk[lane], p2 = (P.variant == 2) ? P.P2 : P.P2 * wk[lane];
                  for (int o = 1; o < L; ++o) M[o] = sel_min(M[o - 1] + p1, M[o]);
                  for (int o = L - 2; o >= 0; --o) M[o] = sel_min(M[o + 1] + p1, M[o]);
                  if (p2 < MGM_INF) for (int o = 0; o < L; ++o) M[o] = sel_min(M[o], m[lane] + p2);
               }
               __syncwarp();
               if (P.variant == 2) {   // update_cost2_trunclinear :197-219
                  for (int o = lane; o < L; o += 32)
                     __stcg(Lp + o, Cp[o] + (scratch[o] - m[0] + scratch[(size_t)L + o] - m[1]) / 2);
               } else {                // update_costW_trunclinear :229-281
                  for (int o = lane; o < L; o += 32) {
                     float e = scratch[o] - m[0];
                     for (int k = 1; k < K; ++k) e += scratch[(size_t)k * L + o] - m[k];
                     __stcg(Lp + o, Cp[o] + e / K);
                  }
               }
               __syncwarp();
            }
            for (int o = L + lane; o < P.VS; o += 32) __stcg(Lp + o, MGM_INF);
         }
         __syncwarp();
         if (lane == 0) {
            __threadfence();
            st_release(prog + ys, xs + 1);
         }
      }
   }
}

cudaError_t agg_generic_launch(const float *cc, const float *w, int nx, int ny, int L, int VS, float P1, float P2, int NDIR,
                               int K, int variant, unsigned mask, float *const *ldir, int *d_counters, float *d_scratch,
                               int nwarps, cudaStream_t st) {
   GenericParams P;
   P.cc = cc; P.w = w;
   P.nsel = 0;
   for (int p = 0; p < MGM_MAX_NDIR; ++p) {
      P.ldir[p] = ldir[p];
      if (p < NDIR && (mask & (1u << p))) P.sel[P.nsel++] = p;
   }
   if (P.nsel == 0) return cudaSuccess;
   P.maxrows = nx > ny ? nx : ny;
   P.ticket = d_counters;
   P.progress = d_counters + 4;
   P.scratch = d_scratch;
   P.nx = nx; P.ny = ny; P.L = L; P.VS = VS; P.K = K; P.variant = variant; P.P1 = P1; P.P2 = P2;
   const int block = 128;
   mgm_aggregate_generic_kernel<<<(nwarps * 32 + block - 1) / block, block, 0, st>>>(P);
   return cudaGetLastError();
}

}  // namespace mgm
