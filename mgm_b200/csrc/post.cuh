// O(W*H) stages around the hot path: left-right test, NaN-aware median,
// disparity-range update and back-projection, so that the default CLI flow (mgm.cc:396-443) stays on the device.
#pragma once
#include "common.cuh"

namespace mgm {

// leftright_test, mgm.cc:68-91: out-of-place (the reference passes copies of the two maps)
cudaError_t leftright_launch(const float *d_dx, int nx, int ny, const float *d_rdx, int rnx, float threshold,
                             float *d_out, cudaStream_t st);
// median_filter, img_tools.h:203-238: window clipped to the image, NaNs skipped, element size/2 of the sorted window
constexpr int MGM_MEDIAN_MAX_RADIUS = 7;
cudaError_t median_launch(const float *d_u, int nx, int ny, int nch, int radius, float *d_out, cudaStream_t st);
// image_minmax, img_tools.h:183-200: d_mm[0] = finite min (+INF if none), d_mm[1] = finite max (-INF if none)
cudaError_t minmax_launch(const float *d_u, long long n, float *d_mm, int num_sms, cudaStream_t st);
// update_dmin_dmax, mgm.cc:120-158 (d_mm from minmax_launch); out-of-place on the two range images
cudaError_t update_range_launch(const float *d_off, int nx, int ny, const float *d_mm, int slack, int radius,
                                const float *d_lo_in, const float *d_hi_in, float *d_lo, float *d_hi, cudaStream_t st);
// remove_nonfinite_values_Img (mgm.cc:335 and :391-392) with the replacement read from d_value[0]
cudaError_t replace_nonfinite_launch(float *d_u, long long n, const float *d_value, cudaStream_t st);
// back-projected image, mgm.cc:432-443
cudaError_t backproject_launch(const float *d_off, const float *d_u, const float *d_v, int nx, int ny, int nch, int vnx,
                               int vny, float *d_syn, cudaStream_t st);

}  // namespace mgm
