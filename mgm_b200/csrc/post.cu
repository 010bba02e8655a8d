// Post-processing kernels of the CLI flow.  All of them are O(W*H) 2-D work on one or two
// float maps: one thread per pixel, coalesced rows, no reuse beyond a small window (median: shared-memory tile).
#include "post.cuh"

namespace mgm {

// ------------------------------------------------------------------ leftright_test, mgm.cc:68-91
// dx is tested against the map of the other view: a pixel survives when it lands inside the other image and the
// disparity found there brings it back within `threshold`.  A NaN on the other side does not invalidate (the
// comparison fabs(NaN) > t is false), a NaN or out-of-range value on this side does (mgm.cc:79-88).
__global__ void mgm_leftright_kernel(const float *__restrict__ dx, int nx, int ny, const float *__restrict__ rdx, int rnx,
                                     float threshold, float *__restrict__ out) {
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= nx) return;
   const size_t i = x + (size_t)y * nx;
   const float d = dx[i];
   const float r = roundf((float)x + d);   // int Lx = round(x + dx[i])
   float o = CUDART_NAN_F;
   if (r >= 0.f && r < (float)rnx) {
      const int Lx = (int)r;
      const float Rx = (float)Lx + rdx[Lx + (size_t)y * rnx];
      if (!(fabsf(Rx - (float)x) > threshold)) o = d;
   }
   out[i] = o;
}

cudaError_t leftright_launch(const float *d_dx, int nx, int ny, const float *d_rdx, int rnx, float threshold,
                             float *d_out, cudaStream_t st) {
   dim3 grid((nx + 255) / 256, ny);
   mgm_leftright_kernel<<<grid, 256, 0, st>>>(d_dx, nx, ny, d_rdx, rnx, threshold, d_out);
   return cudaGetLastError();
}

// ------------------------------------------------------------------ median_filter, img_tools.h:203-238
// The reference collects the non-NaN values of the (2r+1)^2 window clipped to the image and takes element
// size/2 of their sorted order (nth_element).  Here the window lives in a shared tile whose out-of-image cells are
// NaN, and the k-th smallest is found by rank counting: v is the answer iff #(w < v) <= k < #(w < v) + #(w == v).
__global__ void mgm_median_kernel(const float *__restrict__ u, int nx, int ny, int radius, float *__restrict__ out) {
   extern __shared__ float tile[];
   const int tw = blockDim.x + 2 * radius, th = blockDim.y + 2 * radius;
   const int x0 = blockIdx.x * blockDim.x - radius, y0 = blockIdx.y * blockDim.y - radius;
   const size_t plane = (size_t)blockIdx.z * nx * ny;
   for (int t = threadIdx.y * blockDim.x + threadIdx.x; t < tw * th; t += blockDim.x * blockDim.y) {
      const int xx = x0 + t % tw, yy = y0 + t / tw;
      tile[t] = (xx >= 0 && yy >= 0 && xx < nx && yy < ny) ? u[plane + xx + (size_t)yy * nx] : CUDART_NAN_F;
   }
   __syncthreads();
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
   if (x >= nx || y >= ny) return;
   const float *win = tile + threadIdx.y * tw + threadIdx.x;   // top-left cell of this pixel's window
   const int side = 2 * radius + 1;
   int n = 0;
   for (int j = 0; j < side; j++)
      for (int i = 0; i < side; i++) n += (win[j * tw + i] == win[j * tw + i]);
   float res = win[radius * tw + radius];   // empty window: the pixel keeps its value (img_tools.h:230)
   const int k = n / 2;
   bool found = (n == 0);
   for (int j = 0; j < side && !found; j++)
      for (int i = 0; i < side && !found; i++) {
         const float v = win[j * tw + i];
         if (!(v == v)) continue;
         int less = 0, eq = 0;
         for (int jj = 0; jj < side; jj++)
            for (int ii = 0; ii < side; ii++) {
               const float w = win[jj * tw + ii];
               less += (w < v);
               eq += (w == v);
            }
         if (less <= k && k < less + eq) { res = v; found = true; }
      }
   out[plane + x + (size_t)y * nx] = res;
}

cudaError_t median_launch(const float *d_u, int nx, int ny, int nch, int radius, float *d_out, cudaStream_t st) {
   if (radius < 0 || radius > MGM_MEDIAN_MAX_RADIUS) return cudaErrorInvalidValue;
   dim3 block(32, 8), grid((nx + 31) / 32, (ny + 7) / 8, nch);
   const size_t smem = (size_t)(32 + 2 * radius) * (8 + 2 * radius) * sizeof(float);
   mgm_median_kernel<<<grid, block, smem, st>>>(d_u, nx, ny, radius, d_out);
   return cudaGetLastError();
}

// ------------------------------------------------------------------ image_minmax, img_tools.h:183-200
// finite min/max through atomicMax on an order-preserving integer code; 0 means "no finite value seen"
__device__ __forceinline__ unsigned order_code(float f) {
   const unsigned b = __float_as_uint(f);
   return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_decode(unsigned c) {
   return __uint_as_float((c & 0x80000000u) ? (c & 0x7fffffffu) : ~c);
}
__global__ void mgm_minmax_kernel(const float *__restrict__ u, long long n, unsigned *__restrict__ code) {
   unsigned cmin = 0, cmax = 0;   // cmin holds ~code so that both reductions are maxima
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float v = u[i];
      if (fabsf(v) < MGM_INF) {
         const unsigned c = order_code(v);
         cmax = max(cmax, c);
         cmin = max(cmin, ~c);
      }
   }
   for (int o = 16; o; o >>= 1) {
      cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      cmin = max(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
   }
   if ((threadIdx.x & 31) == 0) {
      if (cmin) atomicMax(code + 0, cmin);
      if (cmax) atomicMax(code + 1, cmax);
   }
}
__global__ void mgm_minmax_decode_kernel(const unsigned *__restrict__ code, float *__restrict__ mm) {
   mm[0] = code[0] ? order_decode(~code[0]) : MGM_INF;
   mm[1] = code[1] ? order_decode(code[1]) : -MGM_INF;
}

cudaError_t minmax_launch(const float *d_u, long long n, float *d_mm, int num_sms, cudaStream_t st) {
   unsigned *code = reinterpret_cast<unsigned *>(d_mm + 2);   // d_mm holds 4 words: min, max, two codes
   cudaError_t e = cudaMemsetAsync(code, 0, 2 * sizeof(unsigned), st);
   if (e != cudaSuccess) return e;
   long long blocks = (n + 255) / 256;
   if (blocks > 4LL * num_sms) blocks = 4LL * num_sms;
   if (blocks < 1) blocks = 1;
   mgm_minmax_kernel<<<(int)blocks, 256, 0, st>>>(d_u, n, code);
   mgm_minmax_decode_kernel<<<1, 1, 0, st>>>(code, d_mm);
   return cudaGetLastError();
}

// ------------------------------------------------------------------ update_dmin_dmax, mgm.cc:120-158
// new range of a pixel = [min - slack, max + slack] over its (2r+1)^2 Neumann neighbourhood of the current
// disparities, non-finite neighbours counting as the whole finite range of the map (mm[0], mm[1]).
__global__ void mgm_update_range_kernel(const float *__restrict__ off, int nx, int ny, const float *__restrict__ mm,
                                        int slack, int radius, const float *__restrict__ lo_in,
                                        const float *__restrict__ hi_in, float *__restrict__ lo, float *__restrict__ hi) {
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= nx) return;
   const float gmin = mm[0], gmax = mm[1], sl = (float)slack;
   float dmin = MGM_INF, dmax = -MGM_INF;
   for (int dj = -radius; dj <= radius; dj++) {
      const int yy = min(max(y + dj, 0), ny - 1);
      for (int di = -radius; di <= radius; di++) {
         const int xx = min(max(x + di, 0), nx - 1);
         const float v = off[xx + (size_t)yy * nx];
         const bool fin = fabsf(v) < MGM_INF;
         dmin = fminf(dmin, (fin ? v : gmin) - sl);
         dmax = fmaxf(dmax, (fin ? v : gmax) + sl);
      }
   }
   const size_t i = x + (size_t)y * nx;
   const bool ok = fabsf(dmin) < MGM_INF;   // mgm.cc:149 tests dmin only
   lo[i] = ok ? dmin : lo_in[i];
   hi[i] = ok ? dmax : hi_in[i];
}

cudaError_t update_range_launch(const float *d_off, int nx, int ny, const float *d_mm, int slack, int radius,
                                const float *d_lo_in, const float *d_hi_in, float *d_lo, float *d_hi, cudaStream_t st) {
   dim3 grid((nx + 255) / 256, ny);
   mgm_update_range_kernel<<<grid, 256, 0, st>>>(d_off, nx, ny, d_mm, slack < 0 ? -slack : slack, radius, d_lo_in, d_hi_in,
                                                d_lo, d_hi);
   return cudaGetLastError();
}

__global__ void mgm_replace_nonfinite_kernel(float *__restrict__ u, long long n, const float *__restrict__ value) {
   const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
   if (i < n && !(fabsf(u[i]) < MGM_INF)) u[i] = value[0];
}
cudaError_t replace_nonfinite_launch(float *d_u, long long n, const float *d_value, cudaStream_t st) {
   mgm_replace_nonfinite_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_u, n, d_value);
   return cudaGetLastError();
}

// ------------------------------------------------------------------ back-projection, mgm.cc:432-443
// syn(x,y,c) = v(x + d, y, c) when (x + d, y) falls inside v, else u(x,y,c).  The reference indexes v.data with a
// FLOAT expression (x+q.x+(y+q.y)*v.nx + c*v.npix, truncated): the same float arithmetic is done here, including
// its loss of integer precision beyond 2^24 elements.
__global__ void mgm_backproject_kernel(const float *__restrict__ off, const float *__restrict__ u,
                                       const float *__restrict__ v, int nx, int ny, int nch, int vnx, int vny,
                                       float *__restrict__ syn) {
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= nx) return;
   const size_t i = x + (size_t)y * nx, np = (size_t)nx * ny;
   const size_t vnp = (size_t)vnx * vny, vtotal = vnp * nch;
   const float qx = off[i], qy = 0.f;
   const float px = (float)x + qx, py = (float)y + qy;
   const bool inside = px >= 0.f && py >= 0.f && px < (float)vnx && py < (float)vny;
   for (int c = 0; c < nch; c++) {
      float val = u[i + c * np];
      if (inside) {
         const float fidx = (px + py * (float)vnx) + (float)(int)(c * vnp);
         size_t idx = (size_t)fidx;
         if (idx >= vtotal) idx = vtotal - 1;
         val = v[idx];
      }
      syn[i + c * np] = val;
   }
}
cudaError_t backproject_launch(const float *d_off, const float *d_u, const float *d_v, int nx, int ny, int nch, int vnx,
                               int vny, float *d_syn, cudaStream_t st) {
   dim3 grid((nx + 255) / 256, ny);
   mgm_backproject_kernel<<<grid, 256, 0, st>>>(d_off, d_u, d_v, nx, ny, nch, vnx, vny, d_syn);
   return cudaGetLastError();
}

}  // namespace mgm
