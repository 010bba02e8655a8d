// Not a test: shared-memory / LSU throughput per SM for the access shapes the aggregation kernel uses.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
   extern __shared__ __align__(16) float4 sm[];
   const int tid = threadIdx.x;
   for (int i = tid; i < 8192; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
   __syncthreads();
   float4 acc = make_float4(0, 0, 0, 0);
   const long long t0 = clock64();
   // MODE 0: LDS.128 conflict-free (lane -> consecutive 16 B)     MODE 1: STS.128      MODE 2: LDS.128 + STS.128 alternating
   // MODE 3: LDS.128 where only 8 lanes per quarter... all lanes active but each warp reads 32 different "rows" (stride 321 units)
   // MODE 4: LDS.32 consecutive
   int idx = tid & 1023;
   const int ridx = ((tid & 31) * 321 + (tid >> 5) * 7) & 8191;
   const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
   for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
         const unsigned a0 = base + 16u * ((idx + u * 1024 + it * 32) & 8191);
         const unsigned a3 = base + 16u * ((ridx + u * 8 + it) & 8191);
         float4 v;
         if (MODE == 0 || MODE == 2) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a0)); acc.x += v.x; acc.y += v.w; }
         if (MODE == 1 || MODE == 2) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(base + 16u * ((idx + u * 1024 + it * 32 + 512) & 8191)), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory"); }
         if (MODE == 3) { asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a3)); acc.x += v.x; acc.y += v.w; }
         if (MODE == 5) { float2 w2; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(w2.x), "=f"(w2.y) : "r"(base + 8u * ((tid + u * 1024 + it * 32) & 16383))); acc.x += w2.x; acc.y += w2.y; }
         if (MODE == 6) {   // gather pattern: 4 rows x 8 lanes, row stride 193 x 16 B, each lane one half (8 B) of its 16-B chunk; odd rows take the other half
            const int lane = tid & 31, row = lane >> 3, gl = lane & 7;
            const unsigned a6 = base + (unsigned)(((tid >> 5) * 4 + row) * 193 * 16 + (gl + 8 * u) * 16 + (((row ^ it) & 1) ? 8 : 0));
            float2 w2; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(w2.x), "=f"(w2.y) : "r"(base + ((a6 - base) & 131071u & ~7u))); acc.x += w2.x; acc.y += w2.y; }
         if (MODE == 7) {   // same without alternating halves
            const int lane = tid & 31, row = lane >> 3, gl = lane & 7;
            const unsigned a6 = (unsigned)(((tid >> 5) * 4 + row) * 193 * 16 + (gl + 8 * u) * 16 + ((it & 1) ? 8 : 0));
            float2 w2; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(w2.x), "=f"(w2.y) : "r"(base + (a6 & 131071u & ~7u))); acc.x += w2.x; acc.y += w2.y; }
         if (MODE == 4) { float w; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(base + 4u * ((tid + u * 1024 + it * 32) & 32767))); acc.x += w; }
      }
   }
   const long long t1 = clock64();
   out[blockIdx.x * blockDim.x + tid] = acc.x + acc.y;
   if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
   float *out; long long *cyc;
   cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
   const int iters = 2000;
   const char *names[] = {"LDS.128 linear", "STS.128 linear", "LDS.128+STS.128", "LDS.128 rows(stride 321)", "LDS.32 linear", "LDS.64 linear", "LDS.64 rows alt halves", "LDS.64 rows same half"};
   for (int threads : {128, 512}) {
      for (int mode = 0; mode < 8; ++mode) {
         auto kk = mode == 7 ? k<7> : mode == 6 ? k<6> : mode == 5 ? k<5> : mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : mode == 3 ? k<3> : k<4>;
         cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
         kk<<<148, threads, 131072>>>(out, iters, cyc);
         cudaError_t e = cudaDeviceSynchronize();
         if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
         long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
         const double ops = (double)iters * 8 * (threads / 32) * (mode == 2 ? 2 : 1);
         const double bytes = ops * 32 * (mode == 4 ? 4 : mode >= 5 ? 8 : 16);
         printf("threads=%4d %-26s %.2f cycles per warp instruction, %.1f B/clk/SM\n", threads, names[mode], c / ops, bytes / c);
      }
   }
   return 0;
}
