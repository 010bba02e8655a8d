// K1-K3 -- edge weights, census transform and the W x H x L matching-cost volume.
//   K3 weights      : compute_mgm_weights            mgm_weights.h:63-85
//   K1 census       : census_transform               census_tools.cc:127-153
//   K2 cost volume  : allocate_and_fill_sgm_costvolume mgm_costvolume.h:337-424
//                     with the cost functions of mgm_costvolume.h:23-165
// Device layout of a volume: [y][x][VS] floats, VS = L rounded up to a multiple of
// 32; the padding labels hold +INF, which is exactly what Dvec::operator[] returns
// outside [min,max] (dvec.cc:129), so the aggregation needs no range checks.
#include "costvolume.cuh"

namespace mgm {

// ------------------------------------------------------------------ K3 weights
__global__ void mgm_weights_kernel(const float *__restrict__ u, int nx, int ny, int nch, float aP, float aThresh,
                                   float *__restrict__ w, int *__restrict__ not_all_ones) {
   // grid: x over columns, y = image row; one thread writes the eight planes of its pixel (no index divisions,
   // blocks with enough work: both dominated earlier versions of this kernel)
   const long long np = (long long)nx * ny;
   const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
   if (x >= nx) return;
   const long long p = x + (long long)y * nx;
   const float thr2 = aThresh * aThresh;
   bool other = false;
#pragma unroll
   for (int k = 0; k < 8; ++k) {
      // plane order W E S N NW NE SE SW (mgm_weights.h:69)
      const int ox = (k == 0 || k == 4 || k == 7) ? -1 : ((k == 1 || k == 5 || k == 6) ? 1 : 0);
      const int oy = (k == 3 || k == 4 || k == 5) ? -1 : ((k == 2 || k == 6 || k == 7) ? 1 : 0);
      const int qx = x + ox, qy = y + oy;
      float wv = 1.0f;
      if (qx >= 0 && qy >= 0 && qx < nx && qy < ny) {
         float d = 0.f;
         for (int c = 0; c < nch; ++c) {
            const float diff = u[p + c * np] - u[qx + (long long)qy * nx + c * np];
            d += diff * diff;
         }
         d = __fdiv_rn(d, (float)nch);
         if (fabsf(d) < thr2) wv = aP;
      }
      w[p + (long long)k * np] = wv;
      other |= (wv != 1.0f);
   }
   // the scan of mgm_core.cc:420-422: one atomic per warp that saw a weight other than 1
   const unsigned am = __activemask();
   if (__any_sync(am, other) && not_all_ones && (int)(threadIdx.x & 31) == __ffs(am) - 1) atomicOr(not_all_ones, 1);
}

// ------------------------------------------------------------------ K1 census transform
// Bit b of the reference's MSB-first bit string sits in byte b/8; bytes are memcpy'd into
// little-endian floats, i.e. byte j of a word occupies bits [8j, 8j+8).
__global__ void mgm_census_kernel(const float *__restrict__ u, int nx, int ny, int nch, int r, int nwords,
                                  uint32_t *__restrict__ out) {
   const long long np = (long long)nx * ny;
   const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (p >= np) return;
   const int x = (int)(p % nx), y = (int)(p / nx);
   uint32_t word = 0;
   int bit = 0, widx = 0;
   for (int c = 0; c < nch; ++c) {
      const float a = u[p + c * np];
      for (int dy = -r; dy <= r; ++dy)
         for (int dx = -r; dx <= r; ++dx) {
            if (!dx && !dy) continue;
            const int qx = x + dx, qy = y + dy;
            bool b = false;
            if (qx >= 0 && qx < nx && qy >= 0 && qy < ny) b = a < u[qx + (long long)qy * nx + c * np];
            if (b) word |= 1u << (8 * ((bit >> 3) & 3) + 7 - (bit & 7));
            ++bit;
            if ((bit & 31) == 0) { out[p + (long long)widx * np] = word; word = 0; ++widx; }
         }
   }
   if (widx < nwords) out[p + (long long)widx * np] = word;
}

// ------------------------------------------------------------------ N3 sobelx prefilter
__global__ void mgm_sobelx_kernel(const float *__restrict__ u, int nx, int ny, int nch, float *__restrict__ out) {
   const long long np = (long long)nx * ny;
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= np * nch) return;
   const int c = (int)(i / np);
   const long long p = i - (long long)c * np;
   const int x = (int)(p % nx), y = (int)(p / nx);
   const float kx[3] = {-1.f, 0.f, 1.f};
   const float ky[3] = {1.f, 2.f, 1.f};
   float v = 0.f;
   for (int jj = 0; jj < 3; ++jj)
      for (int ii = 0; ii < 3; ++ii) {   // img_tools.h:105-124, Neumann borders, accumulate in tap order
         int xx = x + ii - 1, yy = y + jj - 1;
         xx = xx < 0 ? 0 : (xx >= nx ? nx - 1 : xx);
         yy = yy < 0 ? 0 : (yy >= ny ? ny - 1 : yy);
         v += u[xx + (long long)yy * nx + c * np] * (kx[ii] * ky[jj]);
      }
   out[i] = v;
}

// ------------------------------------------------------------------ N3 gblur prefilter
// apply_filter (img_tools.h:105-127) with a 1-D kernel of up to 39 taps along x or y, Neumann borders, taps
// accumulated in order.  gblur_truncated (:169-180) = the horizontal pass followed by the vertical one; the taps
// are computed on the host exactly as fill_gaussian_kernel does (:151-167).
struct FilterTaps { float k[39]; int n; };
__global__ void mgm_filter1d_kernel(const float *__restrict__ u, int nx, int ny, int nch, const FilterTaps taps,
                                    int vertical, float *__restrict__ out) {
   const long long np = (long long)nx * ny;
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= np * nch) return;
   const int c = (int)(i / np);
   const long long p = i - (long long)c * np;
   const int x = (int)(p % nx), y = (int)(p / nx);
   const int h = taps.n / 2;
   const float *uc = u + c * np;
   float v = 0.f;
   for (int t = 0; t < taps.n; ++t) {
      int xx = vertical ? x : x + t - h, yy = vertical ? y + t - h : y;
      xx = xx < 0 ? 0 : (xx >= nx ? nx - 1 : xx);
      yy = yy < 0 ? 0 : (yy >= ny ? ny - 1 : yy);
      v += uc[xx + (long long)yy * nx] * taps.k[t];
   }
   out[i] = v;
}

// ------------------------------------------------------------------ per-cell costs
struct CostArgs {
   const float *u, *v;          // (prefiltered) images, planar
   const uint32_t *cu, *cv;     // census words, planar [nwords][ny][nx]
   int nx, ny, vnx, vny, nch;   // nch: channels of the images the cost reads (census: nwords)
   int win;                     // CENSUS_NCC_WIN
};

__device__ __forceinline__ float bt_channel(const float *__restrict__ ur, const float *__restrict__ vr, int nx,
                                            int vnx, int px, int qx) {   // BTAD mgm_costvolume.h:82-110
   const float IL = ur[px];
   float ILp = IL, ILm = IL;
   if (px < nx - 1) ILp = (IL + ur[px + 1]) * 0.5f;   // (a+b)/2.0 in double then to float == exact halving
   if (px >= 1) ILm = (IL + ur[px - 1]) * 0.5f;
   const float IR = vr[qx];
   float IRp = IR, IRm = IR;
   if (qx < vnx - 1) IRp = (IR + vr[qx + 1]) * 0.5f;
   if (qx >= 1) IRm = (IR + vr[qx - 1]) * 0.5f;
   // the reference's min3 / max3 / __min are compare-and-select macros (:85-86, mgm_core.cc:48): restated as such, so
   // that NaN samples (census bit strings read as floats, "-p census -t btad") select the same operand
   auto min3 = [](float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((c < b) ? c : b); };
   auto max3 = [](float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((c > b) ? c : b); };
   const float IminR = min3(IRm, IRp, IR), ImaxR = max3(IRm, IRp, IR);
   const float IminL = min3(ILm, ILp, IL), ImaxL = max3(ILm, ILp, IL);
   const float dLR = max3(0.f, IL - ImaxR, IminR - IL);
   const float dRL = max3(0.f, IR - ImaxL, IminL - IR);
   return fabsf(sel_min(dLR, dRL));
}

template <int DIST>
__device__ __forceinline__ float cell_cost(const CostArgs &A, int px, int py, int qx, int qy) {
   const long long np = (long long)A.nx * A.ny, vnp = (long long)A.vnx * A.vny;
   const long long pi = px + (long long)py * A.nx, qi = qx + (long long)qy * A.vnx;
   if (DIST == DIST_AD || DIST == DIST_SD) {   // mgm_costvolume.h:23-44
      float acc = 0.f;
      for (int c = 0; c < A.nch; ++c) {
         float x = __ldg(A.u + pi + c * np) - __ldg(A.v + qi + c * vnp);
         x = sel_max(x, -x);
         acc += (DIST == DIST_SD) ? x * x : x;
      }
      return acc;
   } else if (DIST == DIST_CENSUS) {   // mgm_costvolume.h:65-78
      float r = 0.f;
      for (int t = 0; t < A.nch; ++t) r += (float)__popc(__ldg(A.cu + pi + t * np) ^ __ldg(A.cv + qi + t * vnp));
      if (A.nch == 1) return r;
      return (float)((double)r / (double)A.nch);
   } else if (DIST == DIST_BTAD || DIST == DIST_BTSD) {   // mgm_costvolume.h:114-133
      float acc = 0.f;
      for (int c = 0; c < A.nch; ++c) {
         const float x = bt_channel(A.u + (long long)py * A.nx + c * np, A.v + (long long)qy * A.vnx + c * vnp,
                                    A.nx, A.vnx, px, qx);
         acc += (DIST == DIST_BTSD) ? x * x : x;
      }
      return acc;
   } else {   // clipped NCC, mgm_costvolume.h:137-165
      const int h = A.win / 2;
      float NCC = 0.f;
      for (int c = 0; c < A.nch; ++c) {
         float mu1 = 0.f, mu2 = 0.f, s1 = 0.f, s2 = 0.f, prod = 0.f;
         int n = 0;
         for (int i = -h; i <= h; ++i)
            for (int j = -h; j <= h; ++j) {   // dx outer, dy inner
               const int ax = px + i, ay = py + j, bx = qx + i, by = qy + j;
               if (ax < 0 || ay < 0 || ax >= A.nx || ay >= A.ny) return MGM_INF;
               if (bx < 0 || by < 0 || bx >= A.vnx || by >= A.vny) return MGM_INF;
               const float v1 = __ldg(A.u + ax + (long long)ay * A.nx + c * np);
               const float v2 = __ldg(A.v + bx + (long long)by * A.vnx + c * vnp);
               if (v1 != v1 || v2 != v2) return MGM_INF;
               mu1 += v1; mu2 += v2;
               s1 += v1 * v1; s2 += v2 * v2; prod += v1 * v2;
               ++n;
            }
         const float fn = (float)n;
         mu1 = __fdiv_rn(mu1, fn); mu2 = __fdiv_rn(mu2, fn);
         s1 = __fdiv_rn(s1, fn); s2 = __fdiv_rn(s2, fn); prod = __fdiv_rn(prod, fn);
         const float num = prod - mu1 * mu2;
         const float var = (s1 - mu1 * mu1) * (s2 - mu2 * mu2);
         const double den = (0.0000001 > (double)var) ? 0.0000001 : (double)var;
         NCC = (float)((double)NCC + (double)num / sqrt(den));
      }
      const float fnch = (float)A.nch;
      float t = (NCC < fnch) ? NCC : fnch;
      t = (0.f > t) ? 0.f : t;
      return (fnch - t) * 64.f;
   }
}

// ------------------------------------------------------------------ K2 cost volume: one warp per pixel
// rlo/rhi (optional): per-pixel disparity ranges as float images, truncated to int like Dvec::init
// (mgm_costvolume.h:323); labels outside a pixel's range do not exist in the reference: +INF here.
// NCT: compile-time number of channels / census words (1..4) for the fast path, 0 = generic cell_cost
template <int DIST, int NCT>
__global__ void __launch_bounds__(256) mgm_costvolume_kernel(const CostArgs A, int dmin, int L, int VS, float cap,
                                                             const float *__restrict__ rlo,
                                                             const float *__restrict__ rhi,
                                                             float *__restrict__ cc) {
   const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const long long np = (long long)A.nx * A.ny;
   const int nq = VS >> 2;
   // Fast path (AD, SD, census with up to four channels / census words): everything that depends on the pixel
   // only -- the left pixel's values, the row pointers of the right image -- is hoisted out of the label loop and
   // the loop indexes with 32-bit offsets; same operations in the same order as cell_cost.
   constexpr bool fast = (NCT > 0);
   const long long vnp = (long long)A.vnx * A.vny;
   const bool small = np < 0x7fffffffLL;   // 32-bit pixel arithmetic (the 64-bit division costs as much as a pixel)
   for (long long p = (long long)blockIdx.x * warps + wid; p < np; p += (long long)gridDim.x * warps) {
      int x, y;
      if (small) { y = (int)((unsigned)p / (unsigned)A.nx); x = (int)((unsigned)p - (unsigned)y * (unsigned)A.nx); }
      else { x = (int)(p % A.nx); y = (int)(p / A.nx); }
      float4 *dst = reinterpret_cast<float4 *>(cc + (size_t)p * VS);
      bool anyfinite = false;
      int klo = 0, khi = L - 1;   // the pixel's own range, as label indices of the dense envelope
      if (rlo) {
         klo = max((int)rlo[p] - dmin, 0);
         khi = min((int)rhi[p] - dmin, L - 1);
      }
      if constexpr (fast) {
         constexpr int NC = NCT > 0 ? NCT : 1;
         uint32_t lw[NC];            // left census words, or the bits of the left channel values
         const uint32_t *rrow[NC];   // right image row, per channel / word
         const bool vy_ok = y < A.vny;
#pragma unroll
         for (int t = 0; t < NC; ++t) {
            if (DIST == DIST_CENSUS) {
               lw[t] = __ldg(A.cu + p + t * np);
               rrow[t] = A.cv + (long long)y * A.vnx + t * vnp;
            } else {
               lw[t] = __float_as_uint(__ldg(A.u + p + t * np));
               rrow[t] = reinterpret_cast<const uint32_t *>(A.v + (long long)y * A.vnx + t * vnp);
            }
         }
         const int qx0 = x + dmin;
         // labels [va,vb] have a match inside the right image; a 16-byte chunk that lies entirely in there (all but
         // two per pixel) takes the straight-line path without per-label range tests
         const int va = vy_ok ? max(klo, -qx0) : 1, vb = vy_ok ? min(khi, A.vnx - 1 - qx0) : 0;
         for (int q = lane; q < nq; q += 32) {
            float e4[4];
            if (q * 4 >= va && q * 4 + 3 <= vb) {
               const int qb = qx0 + q * 4;
#pragma unroll
               for (int j = 0; j < 4; ++j) {
                  float acc = 0.f;
#pragma unroll
                  for (int t = 0; t < NC; ++t) {
                     const uint32_t rv = __ldg(rrow[t] + qb + j);
                     if (DIST == DIST_CENSUS) {
                        const float pc = __uint_as_float(0x4B000000u | (unsigned)__popc(lw[t] ^ rv)) - 8388608.0f;
                        acc = (t == 0) ? pc : acc + pc;
                     } else {
                        float d = __uint_as_float(lw[t]) - __uint_as_float(rv);
                        d = sel_max(d, -d);
                        acc += (DIST == DIST_SD) ? d * d : d;
                     }
                  }
                  float e = (DIST == DIST_CENSUS && NC != 1) ? (float)((double)acc / (double)NC) : acc;
                  e = sel_min(e, cap);
                  anyfinite |= (fabsf(e) < MGM_INF);
                  e4[j] = e;
               }
               dst[q] = make_float4(e4[0], e4[1], e4[2], e4[3]);
               continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
               const int k = q * 4 + j;
               float e = MGM_INF;   // padding labels, labels outside the pixel's range
               if (k >= klo && k <= khi) {
                  const int qx = qx0 + k;
                  e = cap;
                  if (vy_ok && (unsigned)qx < (unsigned)A.vnx) {
                     float acc = 0.f;
#pragma unroll
                     for (int t = 0; t < NC; ++t) {
                        const uint32_t rv = __ldg(rrow[t] + qx);
                        if (DIST == DIST_CENSUS) {
                           // (float)popc without the quarter-rate I2F: 2^23 + n is exact for n <= 32
                           const float pc = __uint_as_float(0x4B000000u | (unsigned)__popc(lw[t] ^ rv)) - 8388608.0f;
                           acc = (t == 0) ? pc : acc + pc;   // 0.f + pc == pc (pc >= +0)
                        }
                        else {
                           float d = __uint_as_float(lw[t]) - __uint_as_float(rv);
                           d = sel_max(d, -d);
                           acc += (DIST == DIST_SD) ? d * d : d;
                        }
                     }
                     e = (DIST == DIST_CENSUS && NC != 1) ? (float)((double)acc / (double)NC) : acc;
                  }
                  e = sel_min(e, cap);
                  anyfinite |= (fabsf(e) < MGM_INF);
               }
               e4[j] = e;
            }
            dst[q] = make_float4(e4[0], e4[1], e4[2], e4[3]);
         }
      } else
      for (int q = lane; q < nq; q += 32) {
         float e4[4];
#pragma unroll
         for (int j = 0; j < 4; ++j) {
            const int k = q * 4 + j;
            float e = MGM_INF;   // padding labels, labels outside the pixel's range
            if (k >= klo && k <= khi) {
               const int qx = x + dmin + k;
               e = cap;   // truncDist * nch' when the match falls outside v (mgm_costvolume.h:398-400)
               if (qx >= 0 && qx < A.vnx && y < A.vny) e = cell_cost<DIST>(A, x, y, qx, y);
               e = sel_min(e, cap);   // :403 (NaN cost -> cap)
               anyfinite |= (fabsf(e) < MGM_INF);
            }
            e4[j] = e;
         }
         dst[q] = make_float4(e4[0], e4[1], e4[2], e4[3]);
      }
      if (!__any_sync(0xffffffffu, anyfinite)) {   // no valid hypothesis: all costs become 0 (:414-421)
         for (int q = lane; q < nq; q += 32) {
            const int k = q * 4;
            auto z = [&](int kk) { return (kk >= klo && kk <= khi) ? 0.f : MGM_INF; };
            dst[q] = make_float4(z(k), z(k + 1), z(k + 2), z(k + 3));
         }
      }
   }
}

// ------------------------------------------------------------------ clipped NCC, fast path
// computeC_clippedNCC (mgm_costvolume.h:137-165) spends most of its additions on quantities that do not depend on the
// label: the window sums of v1, v1^2 (left pixel) and v2, v2^2 (right pixel).  They are computed ONCE per pixel of each
// image, with the reference's own loop order (dx outer, dy inner) and float operations, so every value is the one
// the reference forms: mean = (sum v)/n and dev = (sum v^2)/n - mean*mean per channel, valid = the whole window is
// inside the image and holds no NaN in any channel (otherwise the cost is +INF whatever the other image holds).
// Only sum v1*v2 is left per (pixel, label): 1/3 of the loads and additions of the direct form.
__global__ void mgm_ncc_stats_kernel(const float *__restrict__ img, int nx, int ny, int nch, int h,
                                     float *__restrict__ mean, float *__restrict__ dev, float *__restrict__ valid) {
   const long long np = (long long)nx * ny;
   const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (p >= np) return;
   const int x = (int)(p % nx), y = (int)(p / nx);
   bool ok = true;
   for (int c = 0; c < nch && ok; ++c) {
      float mu = 0.f, s = 0.f;
      int n = 0;
      for (int i = -h; i <= h && ok; ++i)
         for (int j = -h; j <= h; ++j) {
            const int ax = x + i, ay = y + j;
            if (ax < 0 || ay < 0 || ax >= nx || ay >= ny) { ok = false; break; }
            const float v = __ldg(img + ax + (long long)ay * nx + c * np);
            if (v != v) { ok = false; break; }
            mu += v; s += v * v;
            ++n;
         }
      if (!ok) break;
      const float fn = (float)n;
      mu = __fdiv_rn(mu, fn);
      s = __fdiv_rn(s, fn);
      mean[p + c * np] = mu;
      dev[p + c * np] = s - mu * mu;
   }
   valid[p] = ok ? 1.f : 0.f;
}

#define MGM_NCC_MAXWIN 512   // floats of the left window a warp stages (channels x window samples)
__global__ void __launch_bounds__(256) mgm_costvolume_ncc_kernel(const CostArgs A, int dmin, int L, int VS, float cap,
                                                                 const float *__restrict__ rlo, const float *__restrict__ rhi,
                                                                 const float *__restrict__ um, const float *__restrict__ ud,
                                                                 const float *__restrict__ uok, const float *__restrict__ vm,
                                                                 const float *__restrict__ vd, const float *__restrict__ vok,
                                                                 float *__restrict__ cc) {
   __shared__ float s_win[8][MGM_NCC_MAXWIN];
   const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const long long np = (long long)A.nx * A.ny, vnp = (long long)A.vnx * A.vny;
   const int h = A.win / 2, w = 2 * h + 1, w2 = w * w;
   const float fn = (float)w2, fnch = (float)A.nch;
   float *uw = s_win[wid];
   for (long long p = (long long)blockIdx.x * warps + wid; p < np; p += (long long)gridDim.x * warps) {
      const int x = (int)(p % A.nx), y = (int)(p / A.nx);
      float *dst = cc + (size_t)p * VS;
      int klo = 0, khi = L - 1;
      if (rlo) {
         klo = max((int)rlo[p] - dmin, 0);
         khi = min((int)rhi[p] - dmin, L - 1);
      }
      const bool pok = __ldg(uok + p) != 0.f;
      __syncwarp();
      if (pok)
         for (int t = lane; t < A.nch * w2; t += 32) {
            const int c = t / w2, r = t - c * w2, i = r / w - h, j = r % w - h;   // dx outer, dy inner
            uw[t] = __ldg(A.u + (x + i) + (long long)(y + j) * A.nx + c * np);
         }
      __syncwarp();
      bool anyfinite = false;
      for (int k = lane; k < VS; k += 32) {
         float e = MGM_INF;   // padding labels, labels outside the pixel's range
         if (k >= klo && k <= khi) {
            const int qx = x + dmin + k;
            e = cap;   // the match falls outside v (mgm_costvolume.h:398-400)
            if (qx >= 0 && qx < A.vnx && y < A.vny) {
               const long long q = qx + (long long)y * A.vnx;
               e = MGM_INF;   // a window sample outside an image or NaN (:150-154)
               if (pok && __ldg(vok + q) != 0.f) {
                  float NCC = 0.f;
                  for (int c = 0; c < A.nch; ++c) {
                     const float *vb = A.v + q + c * vnp;
                     const float *uc = uw + c * w2;
                     float prod = 0.f;
                     int t = 0;
                     for (int i = -h; i <= h; ++i)
                        for (int j = -h; j <= h; ++j, ++t) prod += uc[t] * __ldg(vb + i + (long long)j * A.vnx);
                     prod = __fdiv_rn(prod, fn);
                     const float mu1 = __ldg(um + p + c * np), mu2 = __ldg(vm + q + c * vnp);
                     const float num = prod - mu1 * mu2;
                     const float var = __ldg(ud + p + c * np) * __ldg(vd + q + c * vnp);
                     const double den = (0.0000001 > (double)var) ? 0.0000001 : (double)var;
                     NCC = (float)((double)NCC + (double)num / sqrt(den));
                  }
                  float t2 = (NCC < fnch) ? NCC : fnch;
                  t2 = (0.f > t2) ? 0.f : t2;
                  e = (fnch - t2) * 64.f;
               }
            }
            e = sel_min(e, cap);   // :403
            anyfinite |= (fabsf(e) < MGM_INF);
         }
         dst[k] = e;
      }
      if (!__any_sync(0xffffffffu, anyfinite))   // no valid hypothesis: all costs become 0 (:414-421)
         for (int k = lane; k < VS; k += 32) dst[k] = (k >= klo && k <= khi) ? 0.f : MGM_INF;
   }
}

// Single-channel NCC with windows up to 7x7 and up to MGM_NCC1_MAXL labels: one THREAD per pixel, labels in a loop.
// The window of the right image slides by one column per label, so it lives in registers as a ring of columns
// (one new column of 2H+1 shared-memory loads per label instead of (2H+1)^2 loads); the left window and the
// pixel's statistics stay in registers for all labels; the right image rows and statistics of the block's pixels x
// labels are staged in shared memory once; costs go to global through a shared-memory transpose (32 labels at a
// time) so that every store is a full 128-byte line.  Same operations in the same order as the kernel above
// (products summed dx outer / dy inner, mgm_costvolume.h:137-165): bit-identical, 4096x4096x64 5x5: 28.7 -> see DESIGN.
#define MGM_NCC1_TP 128     // pixels (threads) per block
#define MGM_NCC1_MAXL 512   // labels the staged rows are sized for
template <int H>
__global__ void __launch_bounds__(MGM_NCC1_TP) mgm_costvolume_ncc1_kernel(const CostArgs A, int dmin, int L, int VS, float cap,
                                                                          const float *__restrict__ rlo, const float *__restrict__ rhi,
                                                                          const float *__restrict__ um, const float *__restrict__ ud,
                                                                          const float *__restrict__ uok, const float *__restrict__ vm,
                                                                          const float *__restrict__ vd, const float *__restrict__ vok,
                                                                          float *__restrict__ cc) {
   constexpr int W = 2 * H + 1, TP = MGM_NCC1_TP;
   constexpr int NCMAX = TP + MGM_NCC1_MAXL - 1 + 2 * H;
   __shared__ float s_v[W][NCMAX];                        // right image rows y-H..y+H, columns x0+dmin-H ...
   __shared__ float s_vm[TP + MGM_NCC1_MAXL], s_vd[TP + MGM_NCC1_MAXL], s_vs[TP + MGM_NCC1_MAXL];   // mean, dev, state of the match pixel
   __shared__ float s_out[TP][33];                        // 32 labels of every pixel, transposed on the way out
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   const int x0 = blockIdx.x * TP, y = blockIdx.y;
   const int x = x0 + tid;
   const bool inimg = x < A.nx;
   const long long p = (long long)y * A.nx + (inimg ? x : A.nx - 1);
   const int nc = TP + L - 1 + 2 * H;       // staged columns: vx = x0 + dmin - H + c
   for (int idx = tid; idx < W * nc; idx += TP) {
      const int j = idx / nc, c = idx - j * nc;
      const int vx = x0 + dmin - H + c, vy = y + j - H;
      s_v[j][c] = (vx >= 0 && vx < A.vnx && vy >= 0 && vy < A.vny) ? __ldg(A.v + vx + (long long)vy * A.vnx) : 0.f;
   }
   for (int c = tid; c < TP + L - 1; c += TP) {   // match pixel qx = x0 + dmin + c
      const int qx = x0 + dmin + c;
      float st = -1.f, m = 0.f, d = 0.f;          // -1: the match falls outside v (cost = cap, mgm_costvolume.h:398-400)
      if (qx >= 0 && qx < A.vnx && y < A.vny) {
         const long long q = qx + (long long)y * A.vnx;
         st = (__ldg(vok + q) != 0.f) ? 1.f : 0.f;   // 0: a window sample outside the image or NaN (+INF, :150-154)
         m = __ldg(vm + q); d = __ldg(vd + q);
      }
      s_vs[c] = st; s_vm[c] = m; s_vd[c] = d;
   }
   const bool pok = inimg && __ldg(uok + p) != 0.f;
   float uw[W][W];   // left window, [dx][dy]
#pragma unroll
   for (int i = 0; i < W; ++i)
#pragma unroll
      for (int j = 0; j < W; ++j) uw[i][j] = pok ? __ldg(A.u + (x + i - H) + (long long)(y + j - H) * A.nx) : 0.f;
   const float mu1 = pok ? __ldg(um + p) : 0.f, dv1 = pok ? __ldg(ud + p) : 0.f;
   int klo = 0, khi = L - 1;
   if (rlo && inimg) {
      klo = max((int)rlo[p] - dmin, 0);
      khi = min((int)rhi[p] - dmin, L - 1);
   }
   const float fn = (float)(W * W);
   __syncthreads();

   float col[W][W];   // ring of window columns of the right image: column position c sits in slot c % W
#pragma unroll
   for (int i = 0; i < W - 1; ++i)
#pragma unroll
      for (int j = 0; j < W; ++j) col[i][j] = s_v[j][tid + i];
   bool anyfinite = false;
   float *dst0 = cc + ((long long)y * A.nx + x0) * VS;   // the block's first pixel
   for (int kb = 0; kb < VS; kb += W) {
#pragma unroll
      for (int t = 0; t < W; ++t) {
         const int k = kb + t;
         if (k < VS) {   // uniform
            // new column: position k + W - 1 -> slot (t + W - 1) % W
            if (k < L) {
#pragma unroll
               for (int j = 0; j < W; ++j) col[(t + W - 1) % W][j] = s_v[j][tid + k + W - 1];
            }
            float e = MGM_INF;   // padding labels, labels outside the pixel's range
            if (k >= klo && k <= khi) {
               const float st = s_vs[tid + k];
               e = cap;
               if (st >= 0.f) {
                  e = MGM_INF;
                  if (pok && st > 0.f) {
                     float prod = 0.f;
#pragma unroll
                     for (int i = 0; i < W; ++i)
#pragma unroll
                        for (int j = 0; j < W; ++j) prod += uw[i][j] * col[(t + i) % W][j];
                     prod = __fdiv_rn(prod, fn);
                     const float num = prod - mu1 * s_vm[tid + k];
                     const float var = dv1 * s_vd[tid + k];
                     const double den = (0.0000001 > (double)var) ? 0.0000001 : (double)var;
                     const float NCC = (float)((double)0.f + (double)num / sqrt(den));
                     float t2 = (NCC < 1.f) ? NCC : 1.f;
                     t2 = (0.f > t2) ? 0.f : t2;
                     e = (1.f - t2) * 64.f;
                  }
               }
               e = sel_min(e, cap);   // :403
               anyfinite |= (fabsf(e) < MGM_INF);
            }
            s_out[tid][k & 31] = e;
            if ((k & 31) == 31 || k == VS - 1) {   // flush 32 labels of the block's pixels, one 128-byte line per pixel
               __syncthreads();
               const int k0 = k & ~31, nk = (k & 31) + 1;
               for (int r = 0; r < 32; ++r) {
                  const int pr = wid * 32 + r;
                  if (x0 + pr < A.nx && lane < nk) dst0[(long long)pr * VS + k0 + lane] = s_out[pr][lane];
               }
               __syncthreads();
            }
         }
      }
   }
   if (inimg && !anyfinite) {   // no valid hypothesis: all costs become 0 (:414-421); the staged stores are ordered before
      float *dst = cc + (size_t)p * VS;
      for (int k = 0; k < VS; ++k) dst[k] = (k >= klo && k <= khi) ? 0.f : MGM_INF;
   }
}

// ------------------------------------------------------------------ layout helpers
// dense [npix][L] (the flat Dvec layout of mgm_costvolume.h:276-299) <-> padded [npix][VS]
__global__ void mgm_pad_volume_kernel(const float *__restrict__ src, float *__restrict__ dst, long long npix, int L,
                                      int VS, long long sp, long long sl) {
   // src element (p,o) at src[p*sp + o*sl]: (sp,sl) = (L,1) pixel-major, (1,npix) label-major planes
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npix * VS) return;
   const long long p = i / VS;
   const int o = (int)(i - p * VS);
   dst[i] = (o < L) ? src[p * sp + (long long)o * sl] : MGM_INF;
}
__global__ void mgm_unpad_volume_kernel(const float *__restrict__ src, float *__restrict__ dst, long long npix,
                                        int L, int VS) {
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npix * L) return;
   const long long p = i / L;
   const int o = (int)(i - p * L);
   dst[i] = src[p * VS + o];
}

// entries of a padded volume outside the per-pixel range [rlo,rhi] become +INF (they do not exist in a Dvec)
__global__ void mgm_mask_volume_kernel(float *__restrict__ cc, long long npix, int L, int VS, int dmin,
                                       const float *__restrict__ rlo, const float *__restrict__ rhi) {
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= npix * VS) return;
   const long long p = i / VS;
   const int k = (int)(i - p * VS);
   if (k < L && (k < (int)rlo[p] - dmin || k > (int)rhi[p] - dmin)) cc[i] = MGM_INF;
}

// flags: bit0 a vector without any finite entry, bit1 NaN, bit2 -INF (fast-path preconditions)
__global__ void mgm_validate_volume_kernel(const float *__restrict__ cc, long long npix, int L, int VS,
                                           int *__restrict__ flags) {
   const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
   for (long long p = (long long)blockIdx.x * warps + wid; p < npix; p += (long long)gridDim.x * warps) {
      bool fin = false, bad_nan = false, bad_ninf = false;
      for (int o = lane; o < L; o += 32) {
         const float v = cc[(size_t)p * VS + o];
         fin |= fabsf(v) < MGM_INF;
         bad_nan |= (v != v);
         bad_ninf |= (v == -MGM_INF);
      }
      const bool f = __any_sync(0xffffffffu, fin);
      const bool n = __any_sync(0xffffffffu, bad_nan);
      const bool i = __any_sync(0xffffffffu, bad_ninf);
      if (lane == 0) {
         const int fl = (f ? 0 : 1) | (n ? 2 : 0) | (i ? 4 : 0);
         if (fl) atomicOr(flags, fl);
      }
   }
}

// flags: bit0 some weight != 1, bit1 a weight that is negative / NaN / INF
__global__ void mgm_scan_weights_kernel(const float *__restrict__ w, long long n, int *__restrict__ flags) {
   const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   int fl = 0;
   if (i < n) {
      const float v = w[i];
      if (v != 1.0f) fl |= 1;
      if (!(v >= 0.f) || !(fabsf(v) < MGM_INF)) fl |= 2;
   }
   fl = __reduce_or_sync(0xffffffffu, fl);
   if ((threadIdx.x & 31) == 0 && fl) atomicOr(flags, fl);
}

// ------------------------------------------------------------------ host launchers
static inline unsigned blocks_for(long long n, int b) { return (unsigned)((n + b - 1) / b); }

cudaError_t weights_launch(const float *d_u, int nx, int ny, int nch, float aP, float aThresh, float *d_w,
                           int *d_flag, cudaStream_t st) {
   if (ny > 65535) return cudaErrorInvalidValue;   // one grid row per image row
   mgm_weights_kernel<<<dim3((nx + 255) / 256, ny), 256, 0, st>>>(d_u, nx, ny, nch, aP, aThresh, d_w, d_flag);
   return cudaGetLastError();
}

cudaError_t census_launch(const float *d_u, int nx, int ny, int nch, int win, uint32_t *d_out, cudaStream_t st) {
   const int r = win / 2;
   mgm_census_kernel<<<blocks_for((long long)nx * ny, 128), 128, 0, st>>>(d_u, nx, ny, nch, r,
                                                                          census_nwords(nch, win), d_out);
   return cudaGetLastError();
}

cudaError_t sobelx_launch(const float *d_u, int nx, int ny, int nch, float *d_out, cudaStream_t st) {
   mgm_sobelx_kernel<<<blocks_for((long long)nx * ny * nch, 256), 256, 0, st>>>(d_u, nx, ny, nch, d_out);
   return cudaGetLastError();
}

cudaError_t gblur_launch(const float *d_u, int nx, int ny, int nch, float sigma, float *d_tmp, float *d_out,
                         cudaStream_t st) {
   // gaussian_kernel_width / fill_gaussian_kernel, img_tools.h:143-167 (float arithmetic, exp(float))
   FilterTaps taps;
   const float radius = 3 * fabsf(sigma);
   int w = (int)ceilf(1 + 2 * radius);
   if (w < 1) w = 1;
   if (w > 39) w = 39;
   const int cw = (w - 1) / 2;
   float m = 0;
   for (int i = 0; i < w; i++) {
      const float x = (float)hypot((double)(i - cw), 0.0);
      const float v = expf(-x * x / (2 * sigma * sigma));
      taps.k[i] = v;
      m += v;
   }
   for (int i = 0; i < w; i++) taps.k[i] /= m;
   for (int i = w; i < 39; i++) taps.k[i] = 0.f;
   taps.n = w;
   const long long n = (long long)nx * ny * nch;
   mgm_filter1d_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_u, nx, ny, nch, taps, 0, d_tmp);
   mgm_filter1d_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_tmp, nx, ny, nch, taps, 1, d_out);
   return cudaGetLastError();
}

static inline unsigned blocks_for(long long n, int b);
cudaError_t costvolume_launch(int dist, const float *d_u, const float *d_v, const uint32_t *d_cu,
                              const uint32_t *d_cv, int nx, int ny, int vnx, int vny, int nch, int win, int dmin,
                              int L, int VS, float truncDist, const float *d_rlo, const float *d_rhi, float *d_cc,
                              int num_sms, cudaStream_t st, float *d_scratch) {
   CostArgs A;
   A.u = d_u; A.v = d_v; A.cu = d_cu; A.cv = d_cv;
   A.nx = nx; A.ny = ny; A.vnx = vnx; A.vny = vny; A.nch = nch; A.win = win;
   const float cap = truncDist * (float)nch;   // nch of the image the cost function reads (:398)
   const long long np = (long long)nx * ny;
   long long want = (np + 7) / 8;
   unsigned grid = (unsigned)((want < (long long)num_sms * 16) ? want : (long long)num_sms * 16);
   if (grid < 1) grid = 1;
#define MGM_CV_LAUNCH(D, N) mgm_costvolume_kernel<D, N><<<grid, 256, 0, st>>>(A, dmin, L, VS, cap, d_rlo, d_rhi, d_cc)
#define MGM_CV_FAST(D)                                  \
   switch (nch) {                                       \
   case 1: MGM_CV_LAUNCH(D, 1); break;                  \
   case 2: MGM_CV_LAUNCH(D, 2); break;                  \
   case 3: MGM_CV_LAUNCH(D, 3); break;                  \
   case 4: MGM_CV_LAUNCH(D, 4); break;                  \
   default: MGM_CV_LAUNCH(D, 0); break;                 \
   }
   if (dist == DIST_NCC && d_scratch && nch * (2 * (win / 2) + 1) * (2 * (win / 2) + 1) <= MGM_NCC_MAXWIN) {
      // label-independent window statistics once per pixel of each image, then one product sum per cell
      const long long vnp = (long long)vnx * vny;
      float *um = d_scratch, *ud = um + np * nch, *uok = ud + np * nch;
      float *vm = uok + np, *vd = vm + vnp * nch, *vok = vd + vnp * nch;
      mgm_ncc_stats_kernel<<<blocks_for(np, 128), 128, 0, st>>>(d_u, nx, ny, nch, win / 2, um, ud, uok);
      mgm_ncc_stats_kernel<<<blocks_for(vnp, 128), 128, 0, st>>>(d_v, vnx, vny, nch, win / 2, vm, vd, vok);
      if (nch == 1 && L <= MGM_NCC1_MAXL && ny <= 65535 && (win / 2 >= 1 && win / 2 <= 3)) {
         // one thread per pixel, the right window as a register ring of columns (mgm_costvolume_ncc1_kernel)
         const dim3 g1((unsigned)((nx + MGM_NCC1_TP - 1) / MGM_NCC1_TP), (unsigned)ny);
         if (win / 2 == 1) mgm_costvolume_ncc1_kernel<1><<<g1, MGM_NCC1_TP, 0, st>>>(A, dmin, L, VS, cap, d_rlo, d_rhi, um, ud, uok, vm, vd, vok, d_cc);
         else if (win / 2 == 2) mgm_costvolume_ncc1_kernel<2><<<g1, MGM_NCC1_TP, 0, st>>>(A, dmin, L, VS, cap, d_rlo, d_rhi, um, ud, uok, vm, vd, vok, d_cc);
         else mgm_costvolume_ncc1_kernel<3><<<g1, MGM_NCC1_TP, 0, st>>>(A, dmin, L, VS, cap, d_rlo, d_rhi, um, ud, uok, vm, vd, vok, d_cc);
         return cudaGetLastError();
      }
      mgm_costvolume_ncc_kernel<<<grid, 256, 0, st>>>(A, dmin, L, VS, cap, d_rlo, d_rhi, um, ud, uok, vm, vd, vok, d_cc);
      return cudaGetLastError();
   }
   switch (dist) {
   case DIST_AD: MGM_CV_FAST(DIST_AD); break;
   case DIST_SD: MGM_CV_FAST(DIST_SD); break;
   case DIST_CENSUS: MGM_CV_FAST(DIST_CENSUS); break;
   case DIST_NCC: MGM_CV_LAUNCH(DIST_NCC, 0); break;
   case DIST_BTAD: MGM_CV_LAUNCH(DIST_BTAD, 0); break;
   default: MGM_CV_LAUNCH(DIST_BTSD, 0); break;
   }
#undef MGM_CV_FAST
#undef MGM_CV_LAUNCH
   return cudaGetLastError();
}

cudaError_t pad_volume_launch(const float *d_src, float *d_dst, long long npix, int L, int VS, int label_major,
                              cudaStream_t st) {
   const long long sp = label_major ? 1 : L, sl = label_major ? npix : 1;
   mgm_pad_volume_kernel<<<blocks_for(npix * VS, 256), 256, 0, st>>>(d_src, d_dst, npix, L, VS, sp, sl);
   return cudaGetLastError();
}
cudaError_t unpad_volume_launch(const float *d_src, float *d_dst, long long npix, int L, int VS, cudaStream_t st) {
   mgm_unpad_volume_kernel<<<blocks_for(npix * L, 256), 256, 0, st>>>(d_src, d_dst, npix, L, VS);
   return cudaGetLastError();
}
cudaError_t mask_volume_launch(float *d_cc, long long npix, int L, int VS, int dmin, const float *d_rlo,
                               const float *d_rhi, cudaStream_t st) {
   mgm_mask_volume_kernel<<<blocks_for(npix * VS, 256), 256, 0, st>>>(d_cc, npix, L, VS, dmin, d_rlo, d_rhi);
   return cudaGetLastError();
}
cudaError_t validate_volume_launch(const float *d_cc, long long npix, int L, int VS, int *d_flags, int num_sms,
                                   cudaStream_t st) {
   long long want = (npix + 7) / 8;
   unsigned grid = (unsigned)((want < (long long)num_sms * 16) ? want : (long long)num_sms * 16);
   if (grid < 1) grid = 1;
   mgm_validate_volume_kernel<<<grid, 256, 0, st>>>(d_cc, npix, L, VS, d_flags);
   return cudaGetLastError();
}
cudaError_t scan_weights_launch(const float *d_w, long long n, int *d_flags, cudaStream_t st) {
   mgm_scan_weights_kernel<<<blocks_for(n, 256), 256, 0, st>>>(d_w, n, d_flags);
   return cudaGetLastError();
}

}  // namespace mgm
