// K5 -- ordered sweep sum + over-count fix + winner-take-all + sub-pixel refinement.
// Replaces mgm_core.cc:582-609 (S += Lr in sweep order, S -= (NDIR-1)*C, first
// finite minimum wins) fused with mgm_refine.h:40-70 / refine.h (vfit, parabola,
// cubic, parabolaOCV on S[o-1..o+2]).
//
// One warp per pixel, lanes over labels with 16-byte accesses.  The per-sweep
// volumes may live on peer GPUs (multi-GPU direction sharding): the kernel only
// sees pointers, the additions are always performed in sweep order 0..NDIR-1 so
// the result does not depend on where a sweep was computed.
#include "wta_device.cuh"

namespace mgm {

template <int LP>
__device__ __forceinline__ void wta_loop(const WtaParams &P, float *s_all) {
   constexpr int NP = 32 / LP;   // pixels per warp
   const int warps_per_cta = blockDim.x >> 5;
   const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LP;
   float *sS = s_all + ((size_t)wid * NP + sub) * P.VS;
   for (long long p0 = ((long long)blockIdx.x * warps_per_cta + wid) * NP + P.pix_begin; p0 < P.pix_end;
        p0 += (long long)gridDim.x * warps_per_cta * NP)
      wta_pixel<false, LP>(P, p0 + sub, sS, lane, p0 + sub < P.pix_end);
}

__global__ void __launch_bounds__(256) mgm_wta_kernel(const WtaParams P) {
   extern __shared__ float s_all[];
   const int lp = wta_lanes_per_pixel(P.VS);
   if (lp == 32) wta_loop<32>(P, s_all);
   else if (lp == 16) wta_loop<16>(P, s_all);
   else wta_loop<8>(P, s_all);
}

// Stand-alone sub-pixel refinement of given labels on a dense volume S [npix][L]
// (subpixel_refinement_sgm mgm_refine.h:40-70, for callers that keep the reference's
// two-call sequence mgm() -> refine()).
__global__ void mgm_refine_kernel(const float *__restrict__ S, long long npix, int L, int dmin, int method,
                                  const float *__restrict__ smin, const float *__restrict__ smax,
                                  float *__restrict__ out, float *__restrict__ outcost) {
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npix) return;
   float minP = out[i], minL = outcost[i];
   if (minP != minP) return;   // no finite label (undefined in the reference)
   const int o = (int)minP;
   const int lo = smin ? (int)smin[i] : dmin, hi = smax ? (int)smax[i] : dmin + L - 1;
   if (o - 1 >= lo && o + 2 <= hi) {
      const float *s = S + (size_t)i * L + (o - dmin);
      const float v[4] = {s[-1], s[0], s[1], s[2]};
      float dx = 0.f;
      if (method == 1) fit_vshape(v, &minL, &dx);
      else if (method == 2) fit_parabola(v, &minL, &dx, false);
      else if (method == 3) fit_cubic(v, &minL, &dx);
      else fit_parabola(v, &minL, &dx, true);
      minP = (float)o + dx;
   }
   out[i] = minP;
   outcost[i] = minL;
}

cudaError_t refine_launch(const float *d_S, long long npix, int L, int dmin, int method, const float *d_smin,
                          const float *d_smax, float *d_out, float *d_outcost, cudaStream_t st) {
   mgm_refine_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, st>>>(d_S, npix, L, dmin, method, d_smin, d_smax, d_out,
                                                                      d_outcost);
   return cudaGetLastError();
}

struct SumSrc { const float *p[16]; };
__global__ void __launch_bounds__(256) mgm_sum_volumes_kernel(const SumSrc S, int n, float4 *__restrict__ out, long long n4) {
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < n; ++k) {
         const float4 l = __ldcs(reinterpret_cast<const float4 *>(S.p[k]) + i);
         s.x += l.x; s.y += l.y; s.z += l.z; s.w += l.w;
      }
      out[i] = s;
   }
}
cudaError_t sum_volumes_launch(const float *const *src, int n, float *out, long long nelem, int num_sms, cudaStream_t st) {
   if (n < 0 || n > 16 || (nelem & 3)) return cudaErrorInvalidValue;
   SumSrc S;
   for (int k = 0; k < 16; ++k) S.p[k] = k < n ? src[k] : nullptr;
   mgm_sum_volumes_kernel<<<num_sms * 8, 256, 0, st>>>(S, n, reinterpret_cast<float4 *>(out), nelem / 4);
   return cudaGetLastError();
}

cudaError_t wta_launch(const WtaParams &P, int num_sms, cudaStream_t st) {
   const int block = 256;
   const size_t smem = (size_t)(block / 32) * (32 / wta_lanes_per_pixel(P.VS)) * P.VS * sizeof(float);
   cudaError_t e = cudaFuncSetAttribute(mgm_wta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mgm_wta_kernel, block, smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   const long long npix = P.pix_end - P.pix_begin;
   const long long ppb = (long long)(block / 32) * (32 / wta_lanes_per_pixel(P.VS));   // pixels per block and pass
   long long want = (npix + ppb - 1) / ppb;
   long long grid = (long long)num_sms * per_sm;
   if (grid > want) grid = want;
   if (grid < 1) grid = 1;
   mgm_wta_kernel<<<(unsigned)grid, block, smem, st>>>(P);
   return cudaGetLastError();
}

}  // namespace mgm
