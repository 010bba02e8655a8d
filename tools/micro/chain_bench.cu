// Not a test: micro-benchmark of the sequential min-convolution chain (the latency-bound phase of
// mgm_aggregate_kernel) in several exact formulations, plus raw dependent-issue latencies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o chain_bench chain_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>

#define INF CUDART_INF_F
static constexpr int VS = 256, NQ = VS / 4, TS = 5 * VS + 4;

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- form 0: the two-label form used by the kernel today
__device__ __forceinline__ void chain4(float &run, float &a0, float &a1, float &a2, float &a3, const float c) {
   { const float u1 = run + c, u2 = u1 + c, a0p = a0 + c; a0 = fminf(u1, a0); a1 = fminf(fminf(u2, a0p), a1); }
   { const float u1 = a1 + c, u2 = u1 + c, a2p = a2 + c; a2 = fminf(u1, a2); a3 = fminf(fminf(u2, a2p), a3); }
   run = a3;
}
// ---- form 1: local chain (independent of the carry) + pure add chain for the carry, 4 labels
__device__ __forceinline__ void chain4c(float &run, float &a0, float &a1, float &a2, float &a3, const float c) {
   const float b1 = fminf(a0 + c, a1), b2 = fminf(b1 + c, a2), b3 = fminf(b2 + c, a3);
   const float u1 = run + c, u2 = u1 + c, u3 = u2 + c, u4 = u3 + c;
   a0 = fminf(u1, a0); a1 = fminf(u2, b1); a2 = fminf(u3, b2); a3 = fminf(u4, b3);
   run = a3;
}
// ---- form 2: same over 8 labels (two chunks)
__device__ __forceinline__ void chain8c(float &run, float *a, const float c) {
   float b[8];
   b[0] = a[0];
#pragma unroll
   for (int i = 1; i < 8; ++i) b[i] = fminf(b[i - 1] + c, a[i]);
   float u = run;
#pragma unroll
   for (int i = 0; i < 8; ++i) { u = u + c; a[i] = fminf(u, b[i]); }
   run = a[7];
}

template <int FORM, int DIR>
__device__ __forceinline__ void step4(float &run, float4 &v, float c) {
   if (FORM == 0) { if (DIR) chain4(run, v.w, v.z, v.y, v.x, c); else chain4(run, v.x, v.y, v.z, v.w, c); }
   else { if (DIR) chain4c(run, v.w, v.z, v.y, v.x, c); else chain4c(run, v.x, v.y, v.z, v.w, c); }
}

// two-label chain, loads issued D chunks ahead of their use (register ring)
template <int D, int DIR>
__device__ __forceinline__ void minconv_half_pf(bool on, const float4 *src, float4 *dst, int nq, float c, float cap,
                                                float sub, int bar_id) {
   const int h = nq >> 1;   // multiple of 4
   const int dq = DIR ? -1 : 1;
   int q = DIR ? (nq - 1) : 0;
   float run = INF;
   float4 sb[D];
   if (on) {
#pragma unroll
      for (int d = 0; d < D; ++d) sb[d] = src[q + d * dq];
      for (int i = 0; i < h; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            float4 v = sb[d];
            sb[d] = src[q + D * dq];   // chunk i+d+D <= h-1+D <= nq-1
            if (DIR) chain4(run, v.w, v.z, v.y, v.x, c); else chain4(run, v.x, v.y, v.z, v.w, c);
            dst[q] = v;
            q += dq;
         }
      }
   }
   pair_barrier(bar_id);
   if (on) {
      float4 ob[D];
      const int qend = DIR ? 0 : (nq - 1);
#pragma unroll
      for (int d = 0; d < D; ++d) ob[d] = dst[q + d * dq];
      for (int i = h; i < nq; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            float4 v = sb[d];
            const float4 o = ob[d];
            int qn = q + D * dq;
            qn = DIR ? max(qn, qend) : min(qn, qend);   // clamp: the last D prefetches are redundant re-loads
            sb[d] = src[qn];
            ob[d] = dst[qn];
            if (DIR) chain4(run, v.w, v.z, v.y, v.x, c); else chain4(run, v.x, v.y, v.z, v.w, c);
            v.x = fminf(fminf(v.x, o.x), cap) - sub;
            v.y = fminf(fminf(v.y, o.y), cap) - sub;
            v.z = fminf(fminf(v.z, o.z), cap) - sub;
            v.w = fminf(fminf(v.w, o.w), cap) - sub;
            dst[q] = v;
            q += dq;
         }
      }
   }
}

template <int D, int DIR>
__device__ __forceinline__ void minconv_half_pf6(bool on, const float4 *src, float4 *dst, int nq, float c, float cap,
                                                 float sub, int bar_id) {
   const int h = nq >> 1;   // multiple of 4
   const int dq = DIR ? -1 : 1;
   int q = DIR ? (nq - 1) : 0;
   const int qend = DIR ? 0 : (nq - 1);
   float run = INF;
   float4 sb[D];
   if (on) {
#pragma unroll
      for (int d = 0; d < D; ++d) sb[d] = src[q + d * dq];
      for (int i = 0; i < h; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            float4 v = sb[d];
            if (i + d + D < h) sb[d] = src[q + D * dq];
            if (DIR) chain4(run, v.w, v.z, v.y, v.x, c); else chain4(run, v.x, v.y, v.z, v.w, c);
            dst[q] = v;
            q += dq;
         }
      }
   }
   pair_barrier(bar_id);
   if (on) {
#pragma unroll
      for (int d = 0; d < D; ++d) sb[d] = dst[q + d * dq];
      for (int i = h; i < nq; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            float4 v = sb[d];
            int qn = q + D * dq;
            qn = DIR ? max(qn, qend) : min(qn, qend);
            sb[d] = dst[qn];   // never a chunk this lane has already finalised within the last D steps? see note
            if (DIR) chain4(run, v.w, v.z, v.y, v.x, c); else chain4(run, v.x, v.y, v.z, v.w, c);
            v.x = fminf(v.x, cap) - sub; v.y = fminf(v.y, cap) - sub;
            v.z = fminf(v.z, cap) - sub; v.w = fminf(v.w, cap) - sub;
            dst[q] = v;
            q += dq;
         }
      }
   }
}

// form 9: four labels per carried step.  The local chain b (independent of the carry) of chunk i+1 is computed
// while the carry runs through chunk i:  F_k = min(g^(k+1)(r), b_k),  b_k = min(b_(k-1)+c, a_k).
struct Loc4 { float b0, b1, b2, b3; };
template <int DIR>
__device__ __forceinline__ Loc4 local4(const float4 &v, float c) {
   Loc4 l;
   const float a0 = DIR ? v.w : v.x, a1 = DIR ? v.z : v.y, a2 = DIR ? v.y : v.z, a3 = DIR ? v.x : v.w;
   l.b0 = a0; l.b1 = fminf(l.b0 + c, a1); l.b2 = fminf(l.b1 + c, a2); l.b3 = fminf(l.b2 + c, a3);
   return l;
}
template <int DIR>
__device__ __forceinline__ float4 carry4(float &run, const Loc4 &l, float c) {
   const float u1 = run + c, u2 = u1 + c, u3 = u2 + c, u4 = u3 + c;
   const float f0 = fminf(u1, l.b0), f1 = fminf(u2, l.b1), f2 = fminf(u3, l.b2), f3 = fminf(u4, l.b3);
   run = f3;
   return DIR ? make_float4(f3, f2, f1, f0) : make_float4(f0, f1, f2, f3);
}
template <int DIR>
__device__ __forceinline__ void minconv_half_c4(bool on, const float4 *src, float4 *dst, int nq, float c, float cap,
                                                float sub, int bar_id) {
   const int h = nq >> 1;
   const int dq = DIR ? -1 : 1;
   int q = DIR ? (nq - 1) : 0;
   float run = INF;
   Loc4 ln;
   if (on) {
      ln = local4<DIR>(src[q], c);
      float4 vn = src[q + dq];
      for (int i = 0; i < h; ++i, q += dq) {
         const Loc4 l = ln;
         ln = local4<DIR>(vn, c);
         vn = src[(i + 2 < nq) ? q + 2 * dq : q];
         dst[q] = carry4<DIR>(run, l, c);
      }
   }
   pair_barrier(bar_id);
   if (on) {
      ln = local4<DIR>(dst[q], c);
      float4 vn = dst[(h + 1 < nq) ? q + dq : q];
      for (int i = h; i < nq; ++i, q += dq) {
         const Loc4 l = ln;
         ln = local4<DIR>(vn, c);
         vn = dst[(i + 2 < nq) ? q + 2 * dq : q];
         float4 v = carry4<DIR>(run, l, c);
         v.x = fminf(v.x, cap) - sub; v.y = fminf(v.y, cap) - sub;
         v.z = fminf(v.z, cap) - sub; v.w = fminf(v.w, cap) - sub;
         dst[q] = v;
      }
   }
}

template <int FORM, int DIR>
__device__ __forceinline__ void minconv_half(bool on, const float4 *src, float4 *dst, int nq, float c, float cap,
                                             float sub, int bar_id) {
   const int h = nq >> 1;
   const int dq = DIR ? -1 : 1;
   int q = DIR ? (nq - 1) : 0;
   float run = INF;
   if (FORM == 5) {
      float4 v = on ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 acc = v;
      const int h2 = nq >> 1;
      for (int i = 0; i < h2; ++i) {
         float4 w = v;
         step4<0, DIR>(run, w, c);
         acc.x += w.x; acc.y = fminf(acc.y, w.y);
         v.x += 1.0f; v.y += 3.0f; v.z += 0.5f; v.w += 2.0f;
      }
      pair_barrier(bar_id);
      for (int i = h2; i < nq; ++i) {
         float4 w = v;
         step4<0, DIR>(run, w, c);
         w.x = fminf(fminf(w.x, acc.x), cap) - sub;
         w.y = fminf(fminf(w.y, acc.y), cap) - sub;
         w.z = fminf(fminf(w.z, acc.z), cap) - sub;
         w.w = fminf(fminf(w.w, acc.w), cap) - sub;
         acc.z += w.x + w.y; acc.w = fminf(acc.w, w.z + w.w);
         v.x += 1.0f; v.y += 3.0f; v.z += 0.5f; v.w += 2.0f;
      }
      if (on) dst[q] = acc;
      return;
   }
   if (FORM == 6) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on) {
         v = src[q];
         for (int i = 0; i < h; ++i, q += dq) {
            const float4 vn = src[q + dq];
            step4<0, DIR>(run, v, c);
            dst[q] = v;
            v = vn;
         }
      }
      pair_barrier(bar_id);
      if (on) {
         v = dst[q];
         for (int i = h; i < nq; ++i, q += dq) {
            const int qn = (i + 1 < nq) ? q + dq : q;
            const float4 vn = dst[qn];
            step4<0, DIR>(run, v, c);
            v.x = fminf(v.x, cap) - sub; v.y = fminf(v.y, cap) - sub;
            v.z = fminf(v.z, cap) - sub; v.w = fminf(v.w, cap) - sub;
            dst[q] = v;
            v = vn;
         }
      }
      return;
   }
   if (FORM == 9) { minconv_half_c4<DIR>(on, src, dst, nq, c, cap, sub, bar_id); return; }
   if (FORM == 7) { minconv_half_pf6<2, DIR>(on, src, dst, nq, c, cap, sub, bar_id); return; }
   if (FORM == 8) { minconv_half_pf6<4, DIR>(on, src, dst, nq, c, cap, sub, bar_id); return; }
   if (FORM == 3) { minconv_half_pf<2, DIR>(on, src, dst, nq, c, cap, sub, bar_id); return; }
   if (FORM == 4) { minconv_half_pf<4, DIR>(on, src, dst, nq, c, cap, sub, bar_id); return; }
   if (FORM < 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on) {
         v = src[q];
         for (int i = 0; i < h; ++i, q += dq) {
            const float4 vn = src[q + dq];
            step4<FORM, DIR>(run, v, c);
            dst[q] = v;
            v = vn;
         }
      }
      pair_barrier(bar_id);
      if (on) {
         float4 o = dst[q];
         auto finish = [&](float4 &vv, const float4 &oo) {
            step4<FORM, DIR>(run, vv, c);
            vv.x = fminf(fminf(vv.x, oo.x), cap) - sub;
            vv.y = fminf(fminf(vv.y, oo.y), cap) - sub;
            vv.z = fminf(fminf(vv.z, oo.z), cap) - sub;
            vv.w = fminf(fminf(vv.w, oo.w), cap) - sub;
         };
         for (int i = h; i + 1 < nq; ++i, q += dq) {
            const float4 vn = src[q + dq];
            const float4 on4 = dst[q + dq];
            finish(v, o);
            dst[q] = v;
            v = vn; o = on4;
         }
         finish(v, o);
         dst[q] = v;
      }
   } else {
      // 8 labels per dependent step; two chunks in flight, loads issued one pair ahead
      auto ld2 = [&](const float4 *p, int qq, float *a) {
         const float4 x = p[qq], y = p[qq + dq];
         if (DIR) { a[0] = x.w; a[1] = x.z; a[2] = x.y; a[3] = x.x; a[4] = y.w; a[5] = y.z; a[6] = y.y; a[7] = y.x; }
         else { a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w; }
      };
      auto st2 = [&](float4 *p, int qq, const float *a) {
         if (DIR) { p[qq] = make_float4(a[3], a[2], a[1], a[0]); p[qq + dq] = make_float4(a[7], a[6], a[5], a[4]); }
         else { p[qq] = make_float4(a[0], a[1], a[2], a[3]); p[qq + dq] = make_float4(a[4], a[5], a[6], a[7]); }
      };
      float a[8], an[8];
      if (on) {
         ld2(src, q, a);
         for (int i = 0; i < h; i += 2, q += 2 * dq) {
            ld2(src, q + 2 * dq, an);   // exists: h < nq
            chain8c(run, a, c);
            st2(dst, q, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = an[k];
         }
      }
      pair_barrier(bar_id);
      if (on) {
         float o[8], on8[8];
         ld2(dst, q, o);
         for (int i = h; i < nq; i += 2, q += 2 * dq) {
            const bool more = (i + 2 < nq);
            if (more) { ld2(src, q + 2 * dq, an); ld2(dst, q + 2 * dq, on8); }
            chain8c(run, a, c);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fminf(fminf(a[k], o[k]), cap) - sub;
            st2(dst, q, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) { a[k] = an[k]; o[k] = on8[k]; }
         }
      }
   }
}

// nrows vectors per CTA, chain warps: [0,ncw) upwards, [ncw,2ncw) downwards, as in the kernel
template <int FORM>
__global__ void __launch_bounds__(512, 1) chain_kernel(const float *in, float *out, int nrows, int iters, float c, float cap,
                                                        long long *cycles, int noise, int hi) {
   extern __shared__ __align__(128) float sm[];
   const int tid = threadIdx.x, lane_id = tid & 31;
   const int nwarps = blockDim.x >> 5;
   const int warp_id = hi ? (nwarps - 1 - (tid >> 5)) : (tid >> 5);   // role index: 0..3 = chain warps
   for (int i = tid; i < nrows * VS; i += blockDim.x) {
      const int r = i / VS, o = i % VS;
      sm[r * TS + 3 * VS + o] = in[(size_t)blockIdx.x * nrows * VS + i];   // "cost buffer" = chain source
   }
   if (tid == 0) *(int *)&sm[nrows * TS + 8] = 0;
   __syncthreads();
   const int ncw = (nrows + 31) >> 5;
   long long t0 = 0, t1 = 0;
   if (warp_id < 2 * ncw) {
      const int cw = warp_id % ncw, cdir = warp_id / ncw;
      const int crow = cw * 32 + lane_id;
      const bool on = crow < nrows;
      const float *src = on ? sm + crow * TS + 3 * VS : sm;
      float *dst = on ? sm + crow * TS : sm;
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
         if (cdir == 0) minconv_half<FORM, 0>(on, (const float4 *)src, (float4 *)dst, NQ, c, cap, 0.f, 1 + cw);
         else minconv_half<FORM, 1>(on, (const float4 *)src, (float4 *)dst, NQ, c, cap, 0.f, 1 + cw);
         pair_barrier(1 + cw);
      }
      t1 = clock64();
      if (lane_id == 0) atomicAdd((int *)&sm[nrows * TS + 8], 1);
   } else if (noise) {
      // emulate the gather traffic of other rows: LDS.128/STS.128 streams over a private region
      volatile int *done = (volatile int *)&sm[nrows * TS + 8];
      float4 *q0 = (float4 *)(sm + nrows * TS + 64);
      float4 *p = q0 + tid;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      while (*done < 2 * ncw) {
#pragma unroll
         for (int k = 0; k < 8; ++k) { float4 x = q0[(tid + k * 37) % 448]; acc.x += x.x; acc.y += x.y; }
         p[0] = acc;
         if (noise > 1) { const long long tw = clock64(); while (clock64() - tw < noise) {} }
      }
   }
   __syncthreads();
   if (warp_id == 0 && lane_id == 0) cycles[blockIdx.x] = (t1 - t0) / iters;
   for (int i = tid; i < nrows * VS; i += blockDim.x) {
      const int r = i / VS, o = i % VS;
      out[(size_t)blockIdx.x * nrows * VS + i] = sm[r * TS + o];
   }
}

// raw dependent-issue latencies
template <int KIND>
__global__ void lat_kernel(float *out, float c, int n, long long *cycles) {
   float x = out[threadIdx.x], y = out[threadIdx.x + 32], z = out[threadIdx.x + 64];
   const long long t0 = clock64();
#pragma unroll 1
   for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
         if (KIND == 0) x = x + c;                                   // FADD -> FADD
         if (KIND == 1) x = fminf(x + c, y);                         // FADD -> FMNMX -> FADD
         if (KIND == 2) x = fminf(fminf(x + c, y), z);               // FADD -> FMNMX3
         if (KIND == 3) { x = fminf(fminf((x + c) + c, y), z); }     // FADD FADD FMNMX3
         if (KIND == 4) x = fminf(x, y) , y = y + 1.0f;              // FMNMX -> FMNMX (y off chain)
         if (KIND == 5) x = __fmaf_rn(x, 1.0f, c);                   // FFMA -> FFMA
      }
   }
   const long long t1 = clock64();
   out[threadIdx.x] = x + y + z;
   if (threadIdx.x == 0) cycles[0] = (t1 - t0);
}

int main(int argc, char **argv) {
   const int nrows = argc > 1 ? atoi(argv[1]) : 43;
   const int iters = 200;
   const int grid = argc > 2 ? atoi(argv[2]) : 148;
   const size_t n = (size_t)grid * nrows * VS;
   std::vector<float> h(n);
   srand(1);
   for (size_t i = 0; i < n; ++i) h[i] = (float)(rand() % 4096) / 7.0f;
   float *din, *dout[10];
   long long *dcyc;
   cudaMalloc(&din, n * 4);
   cudaMemcpy(din, h.data(), n * 4, cudaMemcpyHostToDevice);
   for (int f = 0; f < 10; ++f) cudaMalloc(&dout[f], n * 4);
   cudaMalloc(&dcyc, grid * 8);
   const size_t smem = (size_t)nrows * TS * 4 + 1024 + 448 * 16 + 16;
   std::vector<std::vector<float>> res(10, std::vector<float>(n));
   for (int threads : {128, 448, -448}) {
      const int noise = threads < 0 ? (argc > 3 ? atoi(argv[3]) : 1) : 0;
      if (noise) threads = -threads;
      for (int f = 0; f < 10; ++f) {
         auto k = f == 9 ? chain_kernel<9> : f == 8 ? chain_kernel<8> : f == 7 ? chain_kernel<7> : f == 6 ? chain_kernel<6> : f == 5 ? chain_kernel<5> : f == 0 ? chain_kernel<0> : f == 1 ? chain_kernel<1> : f == 2 ? chain_kernel<2> : f == 3 ? chain_kernel<3> : chain_kernel<4>;
         cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
         for (int rep = 0; rep < 2; ++rep) k<<<grid, threads, smem>>>(din, dout[f], nrows, iters, 2.0f, 1e9f, dcyc, noise, argc > 4 ? atoi(argv[4]) : 0);
         cudaError_t e = cudaDeviceSynchronize();
         if (e != cudaSuccess) { printf("form %d: %s\n", f, cudaGetErrorString(e)); return 1; }
         std::vector<long long> cyc(grid);
         cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
         cudaMemcpy(res[f].data(), dout[f], n * 4, cudaMemcpyDeviceToHost);
         long long mn = cyc[0], mx = cyc[0];
         for (int i = 1; i < grid; ++i) { mn = cyc[i] < mn ? cyc[i] : mn; mx = cyc[i] > mx ? cyc[i] : mx; }
         printf("rows=%d threads=%d noise=%d form=%d: %lld..%lld cycles per 256-label min-convolution (%.2f cyc/label/lane) same_as_form0=%d\n",
                nrows, threads, noise, f, mn, mx, (double)mn / 256.0, (int)(memcmp(res[f].data(), res[0].data(), n * 4) == 0));
      }
   }
   const char *names[] = {"FADD->FADD", "FADD->FMNMX->", "FADD->FMNMX3->", "FADD,FADD,FMNMX3->", "FMNMX->FMNMX", "FFMA->FFMA"};
   for (int kind = 0; kind < 6; ++kind) {
      const int nn = 1000;
      switch (kind) {
      case 0: lat_kernel<0><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      case 1: lat_kernel<1><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      case 2: lat_kernel<2><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      case 3: lat_kernel<3><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      case 4: lat_kernel<4><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      case 5: lat_kernel<5><<<1, 32>>>(dout[0], 2.0f, nn, dcyc); break;
      }
      cudaDeviceSynchronize();
      long long c0;
      cudaMemcpy(&c0, dcyc, 8, cudaMemcpyDeviceToHost);
      printf("latency %-22s %.2f cycles per link\n", names[kind], (double)c0 / (nn * 16.0));
   }
   return 0;
}
