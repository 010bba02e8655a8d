// Host-side C++ mirror of the reference's operator interface for the hot path, on top of the C ABI
// (include/mgmb200.h).  Same names, argument meaning and value semantics as the reference so that a
// caller written against mgm.cc:372-385 / matlab/readme.txt:67-98 compiles unchanged:
//
//   struct Img                                  img.h:9-59 (planar float, data[x + y*nx + c*nx*ny])
//   struct costvolume_t                         mgm_costvolume.h:311-315 (dense here over an envelope [dmin,dmax];
//                                               per-pixel ranges are kept beside it, +INF outside them)
//   Img compute_mgm_weights(Img&, aP, aThresh)  mgm_weights.h:63
//   costvolume_t allocate_and_fill_sgm_costvolume(Img&, Img&, Img&, Img&, char*, char*, float)
//                                               mgm_costvolume.h:337
//   costvolume_t mgm(costvolume_t, const Img&, const Img&, const Img&, Img*, Img*, P1, P2, NDIR, MGM,
//                    USE_FELZENSZWALB_POTENTIALS = 0, SGM_FIX_OVERCOUNT = 1)        mgm_core.cc:408
//   void subpixel_refinement_sgm(costvolume_t&, std::vector<float>&, std::vector<float>&, char*)
//                                               mgm_refine.h:40
//   void leftright_test(Img&, Img&, float)      mgm.cc:68      Img median_filter(Img const&, int)  img_tools.h:203
//   std::pair<float,float> update_dmin_dmax(Img, Img*, Img*, int, int)              mgm.cc:120
// All arithmetic happens on the GPU; there is no host fallback (errors throw std::runtime_error).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>
#include <utility>
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

// dense stand-in for std::vector<Dvec>: values[(x + y*nx)*L + (o - dmin)] over the envelope [dmin,dmax];
// lo/hi (empty = every pixel spans the envelope) are the per-pixel Dvec ranges, truncated to int
struct costvolume_t {
   std::vector<float> values;
   int nx = 0, ny = 0, dmin = 0, dmax = -1;
   std::vector<float> lo, hi;
   int nlabels() const { return dmax - dmin + 1; }
   bool uniform() const { return lo.empty(); }
   int min_of(int pix) const { return lo.empty() ? dmin : (int)lo[pix]; }
   int max_of(int pix) const { return hi.empty() ? dmax : (int)hi[pix]; }
   // Dvec::operator[] (dvec.cc:129): +INF outside the range
   float at(int pix, int o) const {
      return (o >= min_of(pix) && o <= max_of(pix)) ? values[(size_t)pix * nlabels() + (o - dmin)] : INFINITY;
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

// range images -> envelope [*dmin,*dmax]; returns true when every pixel spans the whole envelope.
// allocate_costvolume truncates float -> int (mgm_costvolume.h:323).
inline bool range_envelope(const Img &dminI, const Img &dmaxI, int *dmin, int *dmax) {
   if (dminI.data.empty() || dmaxI.data.size() != dminI.data.size()) throw std::runtime_error("mgmb200: empty range images");
   int lo = (int)dminI[0], hi = (int)dmaxI[0];
   bool uniform = true;
   for (size_t i = 0; i < dminI.data.size(); i++) {
      const int a = (int)dminI.data[i], b = (int)dmaxI.data[i];
      if (a != (int)dminI[0] || b != (int)dmaxI[0]) uniform = false;
      lo = a < lo ? a : lo;
      hi = b > hi ? b : hi;
   }
   *dmin = lo; *dmax = hi;
   return uniform;
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
   const bool uniform = range_envelope(dminI, dmaxI, &CC.dmin, &CC.dmax);
   CC.nx = in_u.nx; CC.ny = in_u.ny;
   CC.values.resize((size_t)CC.nx * CC.ny * CC.nlabels());
   if (uniform) {
      check(mgmb200_costvolume(context(), in_u.data.data(), in_v.data.data(), in_u.nx, in_u.ny, in_u.nch, in_v.nx,
                               in_v.ny, CC.dmin, CC.dmax, prefilter, distance, truncDist, (int)CENSUS_NCC_WIN(),
                               CC.values.data()));
   } else {
      CC.lo = dminI.data; CC.hi = dmaxI.data;
      check(mgmb200_costvolume_ranges(context(), in_u.data.data(), in_v.data.data(), in_u.nx, in_u.ny, in_u.nch, in_v.nx,
                                      in_v.ny, CC.lo.data(), CC.hi.data(), CC.dmin, CC.dmax, prefilter, distance,
                                      truncDist, (int)CENSUS_NCC_WIN(), CC.values.data()));
   }
   return CC;
}

inline costvolume_t mgm(costvolume_t CC, const Img &in_w, const Img &dminI, const Img &dmaxI, Img *out, Img *outcost,
                        const float P1, const float P2, const int NDIR, const int MGM,
                        const int USE_FELZENSZWALB_POTENTIALS = 0, int SGM_FIX_OVERCOUNT = 1) {
   int dmin, dmax;
   const bool s_uniform = range_envelope(dminI, dmaxI, &dmin, &dmax);
   const size_t np = (size_t)CC.nx * CC.ny;
   if (dminI.data.size() != np) throw std::runtime_error("mgmb200: cost volume / range image size mismatch");
   costvolume_t S;
   S.nx = CC.nx; S.ny = CC.ny;
   if (out->npix != CC.nx * CC.ny) *out = Img(CC.nx, CC.ny);
   if (outcost->npix != CC.nx * CC.ny) *outcost = Img(CC.nx, CC.ny);
   const float *w = in_w.data.empty() ? nullptr : in_w.data.data();
   if (s_uniform && CC.uniform() && dmin == CC.dmin && dmax == CC.dmax) {
      S.dmin = dmin; S.dmax = dmax;
      S.values.resize(CC.values.size());
      check(mgmb200_mgm(context(), CC.values.data(), w, CC.nx, CC.ny, dmin, dmax, P1, P2, NDIR, MGM,
                        USE_FELZENSZWALB_POTENTIALS, SGM_FIX_OVERCOUNT, out->data.data(), outcost->data.data(),
                        S.values.data()));
   } else {
      // per-pixel ranges: the returned volume has the ranges [dminI,dmaxI], the cost vectors keep theirs; both live
      // in one dense envelope that covers the two
      const int emin = dmin < CC.dmin ? dmin : CC.dmin, emax = dmax > CC.dmax ? dmax : CC.dmax;
      const int Le = emax - emin + 1, Lc = CC.nlabels();
      std::vector<float> cclo(np), cchi(np), dense;
      for (size_t i = 0; i < np; i++) { cclo[i] = (float)CC.min_of((int)i); cchi[i] = (float)CC.max_of((int)i); }
      const float *ccp = CC.values.data();
      if (emin != CC.dmin || emax != CC.dmax) {   // re-base the cost volume into the wider envelope
         dense.assign(np * Le, INFINITY);
         for (size_t i = 0; i < np; i++)
            memcpy(&dense[i * Le + (CC.dmin - emin)], &CC.values[i * Lc], sizeof(float) * Lc);
         ccp = dense.data();
      }
      S.dmin = emin; S.dmax = emax; S.lo = dminI.data; S.hi = dmaxI.data;
      S.values.resize(np * Le);
      check(mgmb200_mgm_ranges(context(), ccp, cclo.data(), cchi.data(), w, CC.nx, CC.ny, emin, emax, S.lo.data(),
                               S.hi.data(), P1, P2, NDIR, MGM, USE_FELZENSZWALB_POTENTIALS, SGM_FIX_OVERCOUNT,
                               out->data.data(), outcost->data.data(), S.values.data()));
   }
   // side effect kept for drop-in parity of the console output: the sweep digits of mgm_core.cc:491
   for (int pass = 0; pass < NDIR; pass++) printf("%d", pass);
   fflush(stdout);
   return S;
}

// mgm_naive_parallelism (mgm_core.cc:632): one thread per sweep, S accumulated in completion order.  Here the sweeps
// always run concurrently and are summed in sweep order, so it is the same call.
inline costvolume_t mgm_naive_parallelism(costvolume_t CC, const Img &in_w, const Img &dminI, const Img &dmaxI, Img *out,
                                          Img *outcost, const float P1, const float P2, const int NDIR, const int MGM,
                                          const int USE_FELZENSZWALB_POTENTIALS = 0, int SGM_FIX_OVERCOUNT = 1) {
   return mgm(CC, in_w, dminI, dmaxI, out, outcost, P1, P2, NDIR, MGM, USE_FELZENSZWALB_POTENTIALS, SGM_FIX_OVERCOUNT);
}

inline void subpixel_refinement_sgm(costvolume_t &S, std::vector<float> &out, std::vector<float> &outcost,
                                    char *refinement) {
   if (S.uniform())
      check(mgmb200_subpixel_refinement_sgm(context(), S.values.data(), S.nx, S.ny, S.dmin, S.dmax, out.data(),
                                            outcost.data(), refinement));
   else
      check(mgmb200_subpixel_refinement_sgm_ranges(context(), S.values.data(), S.lo.data(), S.hi.data(), S.nx, S.ny,
                                                   S.dmin, S.dmax, out.data(), outcost.data(), refinement));
}

// ---- the O(W*H) stages of the command-line flow, same signatures as the reference's
// leftright_test, mgm.cc:68-91
inline void leftright_test(Img &dx, Img &Rdx, float threshold = 1) {
   check(mgmb200_leftright_test(context(), dx.data.data(), dx.nx, dx.ny, Rdx.data.data(), Rdx.nx, Rdx.ny, threshold));
}
// median_filter, img_tools.h:203-238
inline Img median_filter(Img const &u, int radius) {
   Img M(u.nx, u.ny, u.nch);
   check(mgmb200_median_filter(context(), u.data.data(), u.nx, u.ny, u.nch, radius, M.data.data()));
   return M;
}
// update_dmin_dmax, mgm.cc:120-158
inline std::pair<float, float> update_dmin_dmax(const Img &outoff, Img *dminI, Img *dmaxI, int slack = 3, int radius = 2) {
   float gmin = 0, gmax = 0;
   check(mgmb200_update_dmin_dmax(context(), outoff.data.data(), outoff.nx, outoff.ny, dminI->data.data(),
                                  dmaxI->data.data(), slack, radius, &gmin, &gmax));
   return std::pair<float, float>(gmin, gmax);
}
// the back-projected image of mgm.cc:432-443
inline Img backproject(const Img &outoff, const Img &u, const Img &v) {
   Img syn(u.nx, u.ny, u.nch);
   check(mgmb200_backproject(context(), outoff.data.data(), u.data.data(), v.data.data(), u.nx, u.ny, u.nch, v.nx, v.ny,
                             syn.data.data()));
   return syn;
}

}  // namespace mgmb200
