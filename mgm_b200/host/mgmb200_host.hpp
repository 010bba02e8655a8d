// Host-side C++ mirror of the reference's operator interface for the hot path, on top of the C ABI
// (include/mgmb200.h).  Same names, argument meaning and value semantics as the reference so that a
// caller written against mgm.cc:372-385 / matlab/readme.txt:67-98 compiles unchanged:
//
//   struct Img                                  img.h:9-59 (planar float, data[x + y*nx + c*nx*ny])
//   struct costvolume_t                         mgm_costvolume.h:311-315 (dense here: uniform [dmin,dmax])
//   Img compute_mgm_weights(Img&, aP, aThresh)  mgm_weights.h:63
//   costvolume_t allocate_and_fill_sgm_costvolume(Img&, Img&, Img&, Img&, char*, char*, float)
//                                               mgm_costvolume.h:337
//   costvolume_t mgm(costvolume_t, const Img&, const Img&, const Img&, Img*, Img*, P1, P2, NDIR, MGM,
//                    USE_FELZENSZWALB_POTENTIALS = 0, SGM_FIX_OVERCOUNT = 1)        mgm_core.cc:408
//   void subpixel_refinement_sgm(costvolume_t&, std::vector<float>&, std::vector<float>&, char*)
//                                               mgm_refine.h:40
// All arithmetic happens on the GPU; there is no host fallback (errors throw std::runtime_error).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mgmb200.h"

namespace mgmb200 {

// SMART_PARAMETER (smartparameter.h:26-50): environment read once, parsed with "%lf", cached
inline double smart_parameter(const char *name, double dflt) {
   const char *sv = getenv(name);
   double y;
   if (sv && sscanf(sv, "%lf", &y) == 1) return y;
   return dflt;
}
#define MGMB200_SMART_PARAMETER(n, v)                      \
   static inline double n(void) {                          \
      static bool known = false;                           \
      static double value = v;                             \
      if (!known) { value = mgmb200::smart_parameter(#n, v); known = true; } \
      return value;                                        \
   }

struct Img {
   std::vector<float> data;
   int nx = 0, ny = 0, nch = 0, npix = 0;
   Img() {}
   Img(int nx_, int ny_, int nch_ = 1) : data((size_t)nx_ * ny_ * nch_, 0.f), nx(nx_), ny(ny_), nch(nch_), npix(nx_ * ny_) {}
   Img(const float *copydata, int nx_, int ny_, int nch_ = 1) : Img(nx_, ny_, nch_) {
      memcpy(data.data(), copydata, data.size() * sizeof(float));
   }
   float operator[](int i) const { return data[i]; }
   float &operator[](int i) { return data[i]; }
   float &val(int i, int j, int c) { return data[i + j * nx + (size_t)c * nx * ny]; }
   float val(int i, int j, int c) const { return data[i + j * nx + (size_t)c * nx * ny]; }
};

// dense stand-in for std::vector<Dvec>: values[(x + y*nx)*L + (o - dmin)]
struct costvolume_t {
   std::vector<float> values;
   int nx = 0, ny = 0, dmin = 0, dmax = -1;
   int nlabels() const { return dmax - dmin + 1; }
   // Dvec::operator[] (dvec.cc:129): +INF outside the range
   float at(int pix, int o) const {
      return (o >= dmin && o <= dmax) ? values[(size_t)pix * nlabels() + (o - dmin)] : INFINITY;
   }
};

inline mgmb200_ctx *context() {
   static mgmb200_ctx *ctx = nullptr;
   if (!ctx) {
      int dev = -1;
      if (const char *e = getenv("MGMB200_DEVICE")) dev = atoi(e);
      if (mgmb200_create(dev, &ctx) != 0) throw std::runtime_error(std::string("mgmb200: ") + mgmb200_last_error());
   }
   return ctx;
}
inline void check(int rc) {
   if (rc != 0) throw std::runtime_error(std::string("mgmb200: ") + mgmb200_last_error());
}

// the reference takes per-pixel range images; the GPU path supports the uniform case
inline void uniform_range(const Img &dminI, const Img &dmaxI, int *dmin, int *dmax) {
   if (dminI.data.empty() || dmaxI.data.size() != dminI.data.size()) throw std::runtime_error("mgmb200: empty range images");
   const int lo = (int)dminI[0], hi = (int)dmaxI[0];   // allocate_costvolume truncates float->int (mgm_costvolume.h:323)
   for (size_t i = 0; i < dminI.data.size(); i++)
      if ((int)dminI.data[i] != lo || (int)dmaxI.data[i] != hi)
         throw std::runtime_error("mgmb200: per-pixel disparity ranges (-m/-M) are not supported yet");
   *dmin = lo; *dmax = hi;
}

MGMB200_SMART_PARAMETER(CENSUS_NCC_WIN, 3)

inline Img compute_mgm_weights(Img &u, float aP, float aThresh) {
   Img w(u.nx, u.ny, 8);
   check(mgmb200_compute_mgm_weights(context(), u.data.data(), u.nx, u.ny, u.nch, aP, aThresh, w.data.data()));
   return w;
}

inline costvolume_t allocate_and_fill_sgm_costvolume(Img &in_u, Img &in_v, Img &dminI, Img &dmaxI, char *prefilter,
                                                     char *distance, float truncDist) {
   costvolume_t CC;
   uniform_range(dminI, dmaxI, &CC.dmin, &CC.dmax);
   CC.nx = in_u.nx; CC.ny = in_u.ny;
   CC.values.resize((size_t)CC.nx * CC.ny * CC.nlabels());
   check(mgmb200_costvolume(context(), in_u.data.data(), in_v.data.data(), in_u.nx, in_u.ny, in_u.nch, in_v.nx, in_v.ny,
                            CC.dmin, CC.dmax, prefilter, distance, truncDist, (int)CENSUS_NCC_WIN(), CC.values.data()));
   return CC;
}

inline costvolume_t mgm(costvolume_t CC, const Img &in_w, const Img &dminI, const Img &dmaxI, Img *out, Img *outcost,
                        const float P1, const float P2, const int NDIR, const int MGM,
                        const int USE_FELZENSZWALB_POTENTIALS = 0, int SGM_FIX_OVERCOUNT = 1) {
   int dmin, dmax;
   uniform_range(dminI, dmaxI, &dmin, &dmax);
   if (dmin != CC.dmin || dmax != CC.dmax) throw std::runtime_error("mgmb200: cost volume / range mismatch");
   costvolume_t S;
   S.nx = CC.nx; S.ny = CC.ny; S.dmin = dmin; S.dmax = dmax;
   S.values.resize(CC.values.size());
   if (out->npix != CC.nx * CC.ny) *out = Img(CC.nx, CC.ny);
   if (outcost->npix != CC.nx * CC.ny) *outcost = Img(CC.nx, CC.ny);
   const float *w = in_w.data.empty() ? nullptr : in_w.data.data();
   check(mgmb200_mgm(context(), CC.values.data(), w, CC.nx, CC.ny, dmin, dmax, P1, P2, NDIR, MGM,
                     USE_FELZENSZWALB_POTENTIALS, SGM_FIX_OVERCOUNT, out->data.data(), outcost->data.data(),
                     S.values.data()));
   // side effect kept for drop-in parity of the console output: the sweep digits of mgm_core.cc:491
   for (int pass = 0; pass < NDIR; pass++) printf("%d", pass);
   fflush(stdout);
   return S;
}

inline void subpixel_refinement_sgm(costvolume_t &S, std::vector<float> &out, std::vector<float> &outcost,
                                    char *refinement) {
   check(mgmb200_subpixel_refinement_sgm(context(), S.values.data(), S.nx, S.ny, S.dmin, S.dmax, out.data(),
                                         outcost.data(), refinement));
}

}  // namespace mgmb200
