// Launchers of the weights / census / cost-volume kernels (costvolume.cu).
#pragma once
#include "common.cuh"

namespace mgm {

// index tables of mgm_costvolume.h:170-207 and mgm_refine.h:14-27
enum DistKind { DIST_AD = 0, DIST_SD = 1, DIST_CENSUS = 2, DIST_NCC = 3, DIST_BTAD = 4, DIST_BTSD = 5 };
enum PrefilterKind { PF_NONE = 0, PF_CENSUS = 1, PF_SOBELX = 2, PF_GBLUR = 3 };

// number of 32-bit words of the census signature (census_tools.cc:79-83)
inline int census_nwords(int nch, int win) {
   const int r = win / 2, side = 2 * r + 1;
   const int nbytes = nch * (side * side - 1) / 8;
   return (nbytes + 3) / 4;
}

cudaError_t weights_launch(const float *d_u, int nx, int ny, int nch, float aP, float aThresh, float *d_w,
                           int *d_flag, cudaStream_t st);
cudaError_t census_launch(const float *d_u, int nx, int ny, int nch, int win, uint32_t *d_out, cudaStream_t st);
cudaError_t sobelx_launch(const float *d_u, int nx, int ny, int nch, float *d_out, cudaStream_t st);
cudaError_t gblur_launch(const float *d_u, int nx, int ny, int nch, float sigma, float *d_tmp, float *d_out,
                         cudaStream_t st);
cudaError_t costvolume_launch(int dist, const float *d_u, const float *d_v, const uint32_t *d_cu,
                              const uint32_t *d_cv, int nx, int ny, int vnx, int vny, int nch, int win, int dmin,
                              int L, int VS, float truncDist, const float *d_rlo, const float *d_rhi, float *d_cc,
                              int num_sms, cudaStream_t st, float *d_scratch = nullptr);
// floats of scratch the NCC fast path needs (window statistics of both images); without scratch the direct form runs
inline size_t costvolume_ncc_scratch_floats(int nx, int ny, int vnx, int vny, int nch) {
   return ((size_t)nx * ny + (size_t)vnx * vny) * (2 * (size_t)nch + 1);
}
cudaError_t mask_volume_launch(float *d_cc, long long npix, int L, int VS, int dmin, const float *d_rlo,
                               const float *d_rhi, cudaStream_t st);
cudaError_t pad_volume_launch(const float *d_src, float *d_dst, long long npix, int L, int VS, int label_major,
                              cudaStream_t st);
cudaError_t unpad_volume_launch(const float *d_src, float *d_dst, long long npix, int L, int VS, cudaStream_t st);
cudaError_t validate_volume_launch(const float *d_cc, long long npix, int L, int VS, int *d_flags, int num_sms,
                                   cudaStream_t st);
cudaError_t scan_weights_launch(const float *d_w, long long n, int *d_flags, cudaStream_t st);

}  // namespace mgm
