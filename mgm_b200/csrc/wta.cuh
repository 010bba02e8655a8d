// Parameters of the fused sum + over-count fix + WTA + sub-pixel kernel (wta.cu).
#pragma once
#include "common.cuh"

namespace mgm {

struct WtaParams {
   const float *ldir[16];   // per-sweep message volumes [npix][VS] in sweep order (local or peer memory)
   const float *cc;         // matching costs [npix][VS]
   float *out;              // [npix] disparity (integer WTA label + sub-pixel offset)
   float *outcost;          // [npix]
   float *S_out;            // optional dense [npix][L] corrected aggregated volume (+INF outside [smin,smax])
   // per-pixel ranges (optional), float images truncated to int like Dvec::init:
   const float *smin, *smax;     // range of the output volume S = the dminI/dmaxI arguments of mgm() (mgm_core.cc:408)
   const float *ccmin, *ccmax;   // range of the cost vectors: S is only incremented there (mgm_core.cc:582-587)
   long long pix_begin, pix_end;
   int ndir, L, VS, dmin;
   int fix;                 // SGM_FIX_OVERCOUNT
   int fix_count;           // sweeps the over-count fix refers to; 0 = ndir (differs when ldir[0] is a pre-summed volume)
   int refine;              // 0 none, 1 vfit, 2 parabola, 3 cubic, 4 parabolaOCV (mgm_refine.h:14-27)
};

// lanes of a warp that share one pixel in the finish stage (32, or 16 / 8 for 64..128 / 32 padded labels)
#ifndef MGM_WTA_LP32_MIN
#define MGM_WTA_LP32_MIN 160   // padded labels from which a whole warp takes one pixel (measured: two pixels per warp at 128
                              // labels make the fused finish tiles of 1920x1080x128 7 % cheaper, the stand-alone kernel is unchanged)
#endif
__host__ __device__ inline int wta_lanes_per_pixel(int VS) { return VS >= MGM_WTA_LP32_MIN ? 32 : (VS >= 64 ? 16 : 8); }
cudaError_t wta_launch(const WtaParams &P, int num_sms, cudaStream_t st);
// out = ((0 + src[0]) + src[1]) + ...  element-wise over n volumes (partial sums of the all-reduce exchange)
cudaError_t sum_volumes_launch(const float *const *src, int n, float *out, long long nelem, int num_sms, cudaStream_t st);
cudaError_t refine_launch(const float *d_S, long long npix, int L, int dmin, int method, const float *d_smin,
                          const float *d_smax, float *d_out, float *d_outcost, cudaStream_t st);

}  // namespace mgm
