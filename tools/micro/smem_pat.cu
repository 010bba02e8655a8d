// Not a test: LSU cost of the aggregation kernel's shared-memory access patterns (layout: 56 workers, row stride
// 772 floats = 193 x 16 B, 3 ring slots of 256 floats).
#include <cstdio>
#include <cuda_runtime.h>
#ifndef TS
#define TS 772
#endif
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ void sts128(unsigned a, float4 v) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
   extern __shared__ __align__(128) float sm[];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   for (int i = tid; i < 56 * TS; i += blockDim.x) sm[i] = i;
   __syncthreads();
   const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
   float acc = 0.f;
   // gather mapping: 8 lanes per row
   const int r = tid >> 3, gl = tid & 7;            // 448 threads -> rows 0..55
   // chain mapping: lane = row (warps 0,1: rows 0..31, 32..55)
   const int crow = (warp & 1) * 32 + lane;
   const long long t0 = clock64();
   for (int it = 0; it < iters; ++it) {
      const int slot = it % 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
         if (MODE == 0 && r < 56) {           // gather LDS.128: chunk gl + 8j of row r-1's slot
            const int pr = r > 0 ? r - 1 : 0;
            float4 v = lds128(base + 4u * (pr * TS + slot * 256) + 16u * (gl + 8 * j)); acc += v.x + v.w;
         }
         if (MODE == 1 && r < 56) {           // gather LDS.64 x2, halves alternate with the row parity
            const int pr = r > 0 ? r - 1 : 0;
            const unsigned a = base + 4u * (pr * TS + slot * 256) + 16u * (gl + 8 * j);
            const unsigned h = (r & 1) ? 8u : 0u;
            float2 p = lds64(a + h), q = lds64(a + (h ^ 8u)); acc += p.x + q.y;
         }
         if (MODE == 2 && r < 56) {           // gather LDS.64 x2 without alternation
            const int pr = r > 0 ? r - 1 : 0;
            const unsigned a = base + 4u * (pr * TS + slot * 256) + 16u * (gl + 8 * j);
            float2 p = lds64(a), q = lds64(a + 8u); acc += p.x + q.y;
         }
         if (MODE == 3 && crow < 56) {        // chain LDS.128: lane = row, chunk it-dependent
            float4 v = lds128(base + 4u * (crow * TS + slot * 256) + 16u * ((j + it) & 63)); acc += v.x + v.w;
         }
         if (MODE == 4 && crow < 56) {        // chain LDS.64 x2, halves alternate with bit 3 of the lane
            const unsigned a = base + 4u * (crow * TS + slot * 256) + 16u * ((j + it) & 63);
            const unsigned h = (lane & 8) ? 8u : 0u;
            float2 p = lds64(a + h), q = lds64(a + (h ^ 8u)); acc += p.x + q.y;
         }
         if (MODE == 5 && crow < 56) {        // chain STS.128
            sts128(base + 4u * (crow * TS + slot * 256) + 16u * ((j + it) & 63), make_float4(acc, 1.f, 2.f, 3.f));
         }
         if (MODE == 7 && crow < 56) {        // chain 2xLDS.64 same order
            const unsigned a = base + 4u * (crow * TS + slot * 256) + 16u * ((j + it) & 63);
            float2 p = lds64(a), q = lds64(a + 8u); acc += p.x + q.y;
         }
         if (MODE == 8 && crow < 56) {        // chain 2xSTS.64
            const unsigned a = base + 4u * (crow * TS + slot * 256) + 16u * ((j + it) & 63);
            sts64(a, make_float2(acc, 1.f)); sts64(a + 8u, make_float2(2.f, 3.f));
         }
         if (MODE == 9 && r < 56) {           // gather 2xSTS.64 own slot
            const unsigned a = base + 4u * (r * TS + slot * 256) + 16u * (gl + 8 * j);
            sts64(a, make_float2(acc, 1.f)); sts64(a + 8u, make_float2(2.f, 3.f));
         }
         if (MODE == 10 && crow < 56) {       // chain, planar halves: lanes (rows) contiguous
            const unsigned e = (unsigned)(slot * 128 + 2 * ((j + it) & 63));
            float2 p = lds64(base + 8u * (e * 57 + crow)), q = lds64(base + 8u * ((e + 1) * 57 + crow)); acc += p.x + q.y;
         }
         if (MODE == 11 && crow < 56) {       // chain STS, planar halves
            const unsigned e = (unsigned)(slot * 128 + 2 * ((j + it) & 63));
            sts64(base + 8u * (e * 57 + crow), make_float2(acc, 1.f)); sts64(base + 8u * ((e + 1) * 57 + crow), make_float2(2.f, 3.f));
         }
         if (MODE == 12 && r < 56) {          // gather LDS, planar halves: lane (r, gl) reads chunk gl+8j of row r-1
            const int pr = r > 0 ? r - 1 : 0;
            const unsigned e = (unsigned)(slot * 128 + 2 * (gl + 8 * j));
            float2 p = lds64(base + 8u * (e * 57 + pr)), q = lds64(base + 8u * ((e + 1) * 57 + pr)); acc += p.x + q.y;
         }
         if (MODE == 13 && r < 56) {          // gather STS, planar halves
            const unsigned e = (unsigned)(slot * 128 + 2 * (gl + 8 * j));
            sts64(base + 8u * (e * 57 + r), make_float2(acc, 1.f)); sts64(base + 8u * ((e + 1) * 57 + r), make_float2(2.f, 3.f));
         }
         if (MODE == 6 && r < 56) {           // gather STS.128 own slot
            sts128(base + 4u * (r * TS + slot * 256) + 16u * (gl + 8 * j), make_float4(acc, 1.f, 2.f, 3.f));
         }
      }
   }
   const long long t1 = clock64();
   out[blockIdx.x * blockDim.x + tid] = acc;
   if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
   float *out; long long *cyc;
   cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
   const int iters = 1000;
   const char *names[] = {"", "", "", "", "", "", "", "chain 2xLDS.64 same order", "chain 2xSTS.64", "gather 2xSTS.64", "PLANAR chain 2xLDS.64", "PLANAR chain 2xSTS.64", "PLANAR gather 2xLDS.64", "PLANAR gather 2xSTS.64"};
   const char *names0[] = {"gather LDS.128 (3 preds = x3)", "gather 2xLDS.64 alternating", "gather 2xLDS.64 same half", "chain LDS.128 (4 warps)", "chain 2xLDS.64 alternating", "chain STS.128", "gather STS.128"};
   const size_t smem = 3 * 128 * 57 * 8 + 56 * 16 + 256 > 56 * TS * 4 + 256 ? 3 * 128 * 57 * 8 + 56 * 16 + 256 : 56 * TS * 4 + 256;
   for (int mode = 0; mode < 14; ++mode) {
      if ((TS % 4) != 0 && (mode == 0 || mode == 3 || mode == 5 || mode == 6)) continue;   // 16-byte accesses need 16-byte aligned rows
      auto kk = mode == 13 ? k<13> : mode == 12 ? k<12> : mode == 11 ? k<11> : mode == 10 ? k<10> : mode == 9 ? k<9> : mode == 8 ? k<8> : mode == 7 ? k<7> : mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : mode == 3 ? k<3> : mode == 4 ? k<4> : mode == 5 ? k<5> : k<6>;
      cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const bool chain = (mode >= 3 && mode <= 5) || mode == 7 || mode == 8 || mode == 10 || mode == 11;
      const int threads = chain ? 128 : 448;
      kk<<<148, threads, smem>>>(out, iters, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      // bytes moved per iteration: gather modes: 56 rows x 8 chunks x 8 lanes x 16 B = 57344; chain modes: 56 rows x 16 B x 8 (x2 directions = 4 warps -> 112 lanes)
      const double bytes = chain ? 112.0 * 16 * 8 : 57344.0;
      printf("TS=%d %-32s %.1f cycles per iteration, %.1f B/clk/SM\n", TS, mode < 7 ? names0[mode] : names[mode], (double)c / iters, bytes * iters / c);
   }
   return 0;
}
