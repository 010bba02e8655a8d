// Device code of the finish stage (ordered sweep sum + over-count fix + winner-take-all + sub-pixel refinement),
// shared by mgm_wta_kernel (wta.cu) and the fused finish tiles of mgm_aggregate_kernel (aggregate.cu).
#pragma once
#include "wta.cuh"

namespace mgm {

// ---- refine.h restated; v = {S(o-1), S(o), S(o+1), S(o+2)} -------------------------------
__device__ inline void fit_vshape(const float *v, float *vmin, float *xmin) {   // refine.h:70-92
   if ((v[1] > v[0]) && (v[1] > v[2])) { *vmin = v[1]; *xmin = 0.f; return; }
   float slope = v[2] - v[1];
   if ((v[2] - v[1]) < (v[0] - v[1])) slope = v[0] - v[1];
   const float x = __fdiv_rn(v[0] - v[2], 2.f * slope);
   *xmin = x;
   *vmin = v[2] + (x - 1.f) * slope;
}
__device__ inline void fit_parabola(const float *v, float *vmin, float *xmin, bool ocv) {   // refine.h:6-68
   if (v[1] > v[0] && v[1] > v[2]) { *xmin = 0.f; *vmin = v[1]; return; }
   const float c = v[1];
   float b = (v[2] - v[0]) * 0.5f;
   float a = (v[2] - 2.f * v[1] + v[0]) * 0.5f;
   float x;
   if (ocv) {
      a *= 2.f; b *= 2.f;
      a = (a > 1.0f) ? a : 1.0f;
      x = __fdiv_rn(-b + a, 2.f * a);
   } else {
      x = __fdiv_rn(-b, 2.f * a);
   }
   if (x > 1.f) x = 1.f;
   if (x < -1.f) x = -1.f;
   *vmin = (a * x + b) * x + c;
   *xmin = x;
}
__device__ inline float cubic_at(const float *p, const float x) {   // refine.h:94-98 (double arithmetic, float x)
   const double xd = (double)x;
   const double p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3];
   const float d12 = p[1] - p[2];            // float subtraction inside 3.0*(p[1]-p[2])
   const float d20 = p[2] - p[0];            // p[2]-p[0] is a float subtraction too
   double inner = 3.0 * (double)d12 + p3 - p0;
   double mid = 2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3 + xd * inner;
   double outer = (double)d20 + xd * mid;
   return (float)(p1 + 0.5 * xd * outer);
}
__device__ inline void fit_cubic(const float *p, float *vmin, float *xmin) {   // refine.h:102-145
   float pm, xm;
   if (p[1] < p[2]) { pm = p[1]; xm = 0.f; } else { pm = p[2]; xm = 1.f; }
   const float d12 = p[1] - p[2];
   const float d20 = p[2] - p[0];
   const double a = 1.5 * (3.0 * (double)d12 + (double)p[3] - (double)p[0]);
   const double b = 2.0 * (double)p[0] - 5.0 * (double)p[1] + 4.0 * (double)p[2] - (double)p[3];
   const double c = 0.5 * (double)d20;
   const double discr = b * b - 4.0 * a * c;
   if (discr >= 0) {
      const double sq = sqrt(discr);
      const double z1 = (-b + sq) / (2.0 * a);
      const double z2 = (-b - sq) / (2.0 * a);
      if (z1 > 0.0 && z1 < 1.0) {
         float t = cubic_at(p, (float)z1);
         if (t < pm) { pm = t; xm = (float)z1; }
      }
      if (z2 > 0.0 && z2 < 1.0) {
         float t = cubic_at(p, (float)z2);
         if (t < pm) { pm = t; xm = (float)z2; }
      }
   }
   *vmin = pm; *xmin = xm;
}

__device__ __forceinline__ bool finitef(float v) { return fabsf(v) < MGM_INF; }   // false for NaN and +-INF


// COHERENT: the sweep volumes were written by other CTAs of the SAME launch (fused finish): read them through L2
// (ld.global.cg) after the acquire of the bands' completion flags; otherwise streaming loads.
template <bool COHERENT>
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
   return COHERENT ? __ldcg(p) : __ldcs(p);
}

// One pixel by LP lanes of a warp (LP = 32: one pixel per warp; 16 or 8 for short label vectors, so that 64 or 32
// padded labels still use every lane: the warp then finishes 2 or 4 pixels at once): lanes over labels with 16-byte
// accesses; sS = VS floats of shared memory owned by the pixel's lane group; `valid` = the group has a pixel (all
// lanes take part in the shuffles either way).
template <bool COHERENT, int LP = 32>
__device__ __forceinline__ void wta_pixel(const WtaParams &P, const long long pix, float *sS, const int lane_in_warp,
                                          const bool valid = true) {
   const int lane = lane_in_warp & (LP - 1);
   const int nq = valid ? (P.VS >> 2) : 0;
   const float fixmul = (float)((P.fix_count ? P.fix_count : P.ndir) - 1);
   {
      float best = MGM_INF;
      int besto = -1;
      const size_t base = (size_t)pix * P.VS;
      // per-pixel ranges as label indices (defaults: the whole envelope)
      int slo = 0, shi = P.L - 1, clo = 0, chi = P.L - 1;
      if (P.smin && valid) { slo = (int)P.smin[pix] - P.dmin; shi = (int)P.smax[pix] - P.dmin; }
      if (P.ccmin && valid) { clo = (int)P.ccmin[pix] - P.dmin; chi = (int)P.ccmax[pix] - P.dmin; }
      const bool ranged = (P.smin != nullptr) || (P.ccmin != nullptr);
      for (int q = lane; q < nq; q += LP) {
         const size_t off = base + (size_t)q * 4;
         float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
         for (int p = 0; p < P.ndir; ++p) {   // S = ((0 + L0) + L1) + ...   mgm_core.cc:582-587
            const float4 l = ld_stream<COHERENT>(reinterpret_cast<const float4 *>(P.ldir[p] + off));
            s.x += l.x; s.y += l.y; s.z += l.z; s.w += l.w;
         }
         if (ranged) {   // S is only incremented inside the cost vector's range: it stays 0 elsewhere
            const int o0 = q * 4;
            if (o0 + 0 < clo || o0 + 0 > chi) s.x = 0.f;
            if (o0 + 1 < clo || o0 + 1 > chi) s.y = 0.f;
            if (o0 + 2 < clo || o0 + 2 > chi) s.z = 0.f;
            if (o0 + 3 < clo || o0 + 3 > chi) s.w = 0.f;
         }
         if (P.fix) {   // mgm_core.cc:598-599
            const float4 c = __ldcs(reinterpret_cast<const float4 *>(P.cc + off));
            s.x = s.x - fixmul * c.x; s.y = s.y - fixmul * c.y;
            s.z = s.z - fixmul * c.z; s.w = s.w - fixmul * c.w;
         }
         if (ranged) {   // labels outside S's range do not exist: they read as +INF and never win
            const int o0 = q * 4;
            if (o0 + 0 < slo || o0 + 0 > shi) s.x = MGM_INF;
            if (o0 + 1 < slo || o0 + 1 > shi) s.y = MGM_INF;
            if (o0 + 2 < slo || o0 + 2 > shi) s.z = MGM_INF;
            if (o0 + 3 < slo || o0 + 3 > shi) s.w = MGM_INF;
         }
         *reinterpret_cast<float4 *>(sS + q * 4) = s;
         if (P.S_out) {
            // S_out is dense [pix][L]: element-wise stores (L need not be a multiple of 4)
            float *dst = P.S_out + (size_t)pix * P.L + (size_t)q * 4;
            const int o0 = q * 4;
            if (o0 + 0 < P.L) dst[0] = s.x;
            if (o0 + 1 < P.L) dst[1] = s.y;
            if (o0 + 2 < P.L) dst[2] = s.z;
            if (o0 + 3 < P.L) dst[3] = s.w;
         }
         const int o0 = q * 4;   // labels beyond L hold INF-INF=NaN or INF: never finite
         if (o0 + 0 < P.L && finitef(s.x) && best > s.x) { best = s.x; besto = o0; }
         if (o0 + 1 < P.L && finitef(s.y) && best > s.y) { best = s.y; besto = o0 + 1; }
         if (o0 + 2 < P.L && finitef(s.z) && best > s.z) { best = s.z; besto = o0 + 2; }
         if (o0 + 3 < P.L && finitef(s.w) && best > s.w) { best = s.w; besto = o0 + 3; }
      }
      // first minimum wins: smaller value, then smaller label
#pragma unroll
      for (int d = LP / 2; d > 0; d >>= 1) {
         const float ob = __shfl_xor_sync(0xffffffffu, best, d);
         const int oo = __shfl_xor_sync(0xffffffffu, besto, d);
         const bool take = (oo >= 0) && (besto < 0 || ob < best || (ob == best && oo < besto));
         if (take) { best = ob; besto = oo; }
      }
      __syncwarp();
      if (lane == 0 && valid) {
         float minP, minL = best;
         if (besto < 0) {
            minP = __int_as_float(0x7fc00000);   // reference leaves it uninitialised (mgm_core.cc:594)
         } else {
            const int o = besto + P.dmin;
            minP = (float)o;
            if (P.refine != 0) {
               const int oi = (int)minP;   // mgm_refine.h:57
               if (oi - 1 >= P.dmin + slo && oi + 2 <= P.dmin + shi) {   // S[i].min / S[i].max, mgm_refine.h:58
                  const float *v = sS + (besto - 1);
                  float dx = 0.f;
                  if (P.refine == 1) fit_vshape(v, &minL, &dx);
                  else if (P.refine == 2) fit_parabola(v, &minL, &dx, false);
                  else if (P.refine == 3) fit_cubic(v, &minL, &dx);
                  else fit_parabola(v, &minL, &dx, true);
                  minP = (float)oi + dx;
               }
            }
         }
         P.out[pix] = minP;
         P.outcost[pix] = minL;
      }
      __syncwarp();
   }
}

}  // namespace mgm
