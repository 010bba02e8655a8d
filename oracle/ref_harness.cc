// TEST INFRASTRUCTURE ONLY -- never linked or called by the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference sources, compiled
// from where they lie under /root/reference into oracle/_ref/ (git-ignored).
// It exposes the reference's hot-path functions over flat arrays so that the
// C restatement in oracle/mgm_oracle.c (and through it the CUDA path) can be
// pinned against outputs of the reference itself:
//   compute_mgm_weights                mgm_weights.h:63
//   allocate_and_fill_sgm_costvolume   mgm_costvolume.h:337
//   mgm                                mgm_core.cc:408
//   subpixel_refinement_sgm            mgm_refine.h:40
// No reference source is copied into this repository; this file only
// #includes the headers by name (the include path is given by the Makefile).
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <unistd.h>
#include <fcntl.h>
#include <vector>
#include <chrono>

// The reference reads its tunables through SMART_PARAMETER (smartparameter.h:26),
// which caches getenv() on first use.  The harness must be able to change them
// between calls inside one process, so it supplies its own definition of the
// macro (the reference headers use the macro without including its header).
#define SMART_PARAMETER(n, v)             \
   static double harness_value_##n = v;   \
   static double n(void) { return harness_value_##n; }

SMART_PARAMETER(TSGM_DEBUG, 0)

#include "img.h"
#include "point.h"
#include "img_tools.h"
#include "mgm_costvolume.h"   // defines SMART_PARAMETER(CENSUS_NCC_WIN,3)
#include "mgm_core.cc"
#include "mgm_weights.h"
#include "mgm_refine.h"

// img_tools.h defines (non-inline) wrappers that call into iio; the harness
// never uses them, these stubs only satisfy the dynamic linker.
extern "C" float *iio_read_image_float_split(const char *, int *, int *, int *) { abort(); }
extern "C" void iio_save_image_float_split(char *, float *, int, int, int) { abort(); }

namespace {
struct QuietStdout {   // mgm() prints the pass index digits (mgm_core.cc:491)
   int saved;
   QuietStdout() {
      fflush(stdout);
      saved = dup(1);
      int nul = open("/dev/null", O_WRONLY);
      dup2(nul, 1);
      close(nul);
   }
   ~QuietStdout() {
      fflush(stdout);
      dup2(saved, 1);
      close(saved);
   }
};

Img make_img(const float *data, int nx, int ny, int nch) {
   Img im(nx, ny, nch);
   if (data) memcpy(&im.data[0], data, sizeof(float) * (size_t)nx * ny * nch);
   return im;
}

// flat volume layout used across the harness: [pixel][o - dmin], L = dmax-dmin+1
costvolume_t volume_from_flat(const float *cc, Img &dminI, Img &dmaxI, int L) {
   costvolume_t CC = allocate_costvolume(dminI, dmaxI);
   int npix = dminI.nx * dminI.ny;
   for (int i = 0; i < npix; i++) {
      int lo = (int)dminI[i];
      for (int k = 0; k < L; k++) CC[i].set_nolock(lo + k, cc[(size_t)i * L + k]);
   }
   return CC;
}

void volume_to_flat(costvolume_t &CC, float *out, int npix, int dmin, int L) {
   for (int i = 0; i < npix; i++)
      for (int k = 0; k < L; k++) out[(size_t)i * L + k] = CC[i][dmin + k];
}
}  // namespace

extern "C" {

int ref_abi_version(void) { return 1; }

void ref_set_census_ncc_win(int win) { harness_value_CENSUS_NCC_WIN = win; }

// compute_mgm_weights (mgm_weights.h:63); w_out: 8 planes of nx*ny
void ref_weights(const float *u, int nx, int ny, int nch, float aP, float aThresh, float *w_out) {
   Img U = make_img(u, nx, ny, nch);
   Img w = compute_mgm_weights(U, aP, aThresh);
   memcpy(w_out, &w.data[0], sizeof(float) * (size_t)nx * ny * 8);
}

// allocate_and_fill_sgm_costvolume (mgm_costvolume.h:337), uniform range.
// cc_out: [ny*nx][L]
void ref_costvolume(const float *u, const float *v, int nx, int ny, int nch, int vnx, int vny, int dmin,
                    int dmax, const char *prefilter, const char *distance, float truncDist,
                    int census_ncc_win, float *cc_out) {
   harness_value_CENSUS_NCC_WIN = census_ncc_win;
   Img U = make_img(u, nx, ny, nch), V = make_img(v, vnx, vny, nch);
   Img dminI(nx, ny), dmaxI(nx, ny);
   for (int i = 0; i < nx * ny; i++) { dminI[i] = dmin; dmaxI[i] = dmax; }
   costvolume_t CC = allocate_and_fill_sgm_costvolume(U, V, dminI, dmaxI, (char *)prefilter,
                                                      (char *)distance, truncDist);
   volume_to_flat(CC, cc_out, nx * ny, dmin, dmax - dmin + 1);
}

// mgm (mgm_core.cc:408) or mgm_naive_parallelism (:632) on a flat cost volume.
// cc: [ny*nx][L]; w: 8 planes; out/outcost: nx*ny; S_out (optional): [ny*nx][L]
// Returns the wall time of the mgm() call itself in seconds.
double ref_mgm(const float *cc, const float *w, int nx, int ny, int dmin, int dmax, float P1, float P2,
               int NDIR, int MGM, int felz, int fix, int naive, float *out, float *outcost,
               float *S_out) {
   int L = dmax - dmin + 1;
   Img dminI(nx, ny), dmaxI(nx, ny);
   for (int i = 0; i < nx * ny; i++) { dminI[i] = dmin; dmaxI[i] = dmax; }
   Img W = make_img(w, nx, ny, 8);
   if (!w) for (size_t i = 0; i < W.data.size(); i++) W[i] = 1.0f;
   costvolume_t CC = volume_from_flat(cc, dminI, dmaxI, L);
   Img o(nx, ny), oc(nx, ny);
   QuietStdout quiet;
   auto t0 = std::chrono::steady_clock::now();
   costvolume_t S = naive ? mgm_naive_parallelism(CC, W, dminI, dmaxI, &o, &oc, P1, P2, NDIR, MGM, felz, fix)
                          : mgm(CC, W, dminI, dmaxI, &o, &oc, P1, P2, NDIR, MGM, felz, fix);
   auto t1 = std::chrono::steady_clock::now();
   memcpy(out, &o.data[0], sizeof(float) * (size_t)nx * ny);
   memcpy(outcost, &oc.data[0], sizeof(float) * (size_t)nx * ny);
   if (S_out) volume_to_flat(S, S_out, nx * ny, dmin, L);
   return std::chrono::duration<double>(t1 - t0).count();
}

// subpixel_refinement_sgm (mgm_refine.h:40); S: [npix][L]; out/outcost updated in place
void ref_refine(const float *S, int nx, int ny, int dmin, int dmax, float *out, float *outcost,
                const char *refinement) {
   int L = dmax - dmin + 1;
   Img dminI(nx, ny), dmaxI(nx, ny);
   for (int i = 0; i < nx * ny; i++) { dminI[i] = dmin; dmaxI[i] = dmax; }
   costvolume_t SS = volume_from_flat(S, dminI, dmaxI, L);
   std::vector<float> o(out, out + (size_t)nx * ny), oc(outcost, outcost + (size_t)nx * ny);
   subpixel_refinement_sgm(SS, o, oc, (char *)refinement);
   memcpy(out, &o[0], sizeof(float) * o.size());
   memcpy(outcost, &oc[0], sizeof(float) * oc.size());
}

// Whole hot path the way mgm.cc:372-385 strings it together (one direction of
// the LR pair, TSGM_ITER=1): weights -> cost volume -> mgm -> refine.
// P1/P2 are the CLI values; they are scaled by nch here as mgm.cc:356-357 does.
// times[4] (optional): seconds for weights, costvolume, mgm, refine.
void ref_pipeline(const float *u, const float *v, int nx, int ny, int nch, int dmin, int dmax, float P1,
                  float P2, int NDIR, int MGM, int felz, int fix, float aP, float aThresh,
                  const char *prefilter, const char *distance, float truncDist, int census_ncc_win,
                  const char *refinement, float *out, float *outcost, double *times) {
   harness_value_CENSUS_NCC_WIN = census_ncc_win;
   Img U = make_img(u, nx, ny, nch), V = make_img(v, nx, ny, nch);
   Img dminI(nx, ny), dmaxI(nx, ny);
   for (int i = 0; i < nx * ny; i++) { dminI[i] = dmin; dmaxI[i] = dmax; }
   P1 *= nch; P2 *= nch;
   Img o(nx, ny), oc(nx, ny);
   QuietStdout quiet;
   auto t0 = std::chrono::steady_clock::now();
   Img W = compute_mgm_weights(U, aP, aThresh);
   auto t1 = std::chrono::steady_clock::now();
   costvolume_t CC = allocate_and_fill_sgm_costvolume(U, V, dminI, dmaxI, (char *)prefilter,
                                                      (char *)distance, truncDist);
   auto t2 = std::chrono::steady_clock::now();
   costvolume_t S = mgm(CC, W, dminI, dmaxI, &o, &oc, P1, P2, NDIR, MGM, felz, fix);
   auto t3 = std::chrono::steady_clock::now();
   subpixel_refinement_sgm(S, o.data, oc.data, (char *)refinement);
   auto t4 = std::chrono::steady_clock::now();
   memcpy(out, &o.data[0], sizeof(float) * (size_t)nx * ny);
   memcpy(outcost, &oc.data[0], sizeof(float) * (size_t)nx * ny);
   if (times) {
      times[0] = std::chrono::duration<double>(t1 - t0).count();
      times[1] = std::chrono::duration<double>(t2 - t1).count();
      times[2] = std::chrono::duration<double>(t3 - t2).count();
      times[3] = std::chrono::duration<double>(t4 - t3).count();
   }
}

// ---------------------------------------------------------------- per-pixel disparity ranges
// Volumes cross the harness as DENSE arrays over an envelope [emin, emax] (Le labels per pixel); a pixel's Dvec
// only holds [min_i, max_i]: entries outside are ignored on input and read back as +INF (Dvec::operator[]).
// Range images are floats, truncated by Dvec::init like in the reference (mgm_costvolume.h:323).
static costvolume_t volume_from_dense(const float *cc, Img &lo, Img &hi, int emin, int Le) {
   costvolume_t CC = allocate_costvolume(lo, hi);
   int npix = lo.nx * lo.ny;
   for (int i = 0; i < npix; i++)
      for (int o = CC[i].min; o <= CC[i].max; o++)
         if (o >= emin && o < emin + Le) CC[i].set_nolock(o, cc[(size_t)i * Le + (o - emin)]);
   return CC;
}

void ref_costvolume_ranges(const float *u, const float *v, int nx, int ny, int nch, int vnx, int vny,
                           const float *dminI, const float *dmaxI, int emin, int emax, const char *prefilter,
                           const char *distance, float truncDist, int census_ncc_win, float *cc_out) {
   harness_value_CENSUS_NCC_WIN = census_ncc_win;
   Img U = make_img(u, nx, ny, nch), V = make_img(v, vnx, vny, nch);
   Img lo = make_img(dminI, nx, ny, 1), hi = make_img(dmaxI, nx, ny, 1);
   costvolume_t CC = allocate_and_fill_sgm_costvolume(U, V, lo, hi, (char *)prefilter, (char *)distance, truncDist);
   volume_to_flat(CC, cc_out, nx * ny, emin, emax - emin + 1);
}

// mgm() with a cost volume whose vectors have the ranges [ccmin, ccmax] and an output volume with the ranges
// [smin, smax] (the dminI/dmaxI arguments of mgm_core.cc:408-413; different from the former when TSGM_ITER > 1)
double ref_mgm_ranges(const float *cc, const float *ccmin, const float *ccmax, const float *w, int nx, int ny,
                      int emin, int emax, const float *smin, const float *smax, float P1, float P2, int NDIR,
                      int MGM, int felz, int fix, float *out, float *outcost, float *S_out) {
   int Le = emax - emin + 1;
   Img clo = make_img(ccmin, nx, ny, 1), chi = make_img(ccmax, nx, ny, 1);
   Img slo = make_img(smin, nx, ny, 1), shi = make_img(smax, nx, ny, 1);
   Img W = make_img(w, nx, ny, 8);
   if (!w) for (size_t i = 0; i < W.data.size(); i++) W[i] = 1.0f;
   costvolume_t CC = volume_from_dense(cc, clo, chi, emin, Le);
   Img o(nx, ny), oc(nx, ny);
   QuietStdout quiet;
   auto t0 = std::chrono::steady_clock::now();
   costvolume_t S = mgm(CC, W, slo, shi, &o, &oc, P1, P2, NDIR, MGM, felz, fix);
   auto t1 = std::chrono::steady_clock::now();
   memcpy(out, &o.data[0], sizeof(float) * (size_t)nx * ny);
   memcpy(outcost, &oc.data[0], sizeof(float) * (size_t)nx * ny);
   if (S_out) volume_to_flat(S, S_out, nx * ny, emin, Le);
   return std::chrono::duration<double>(t1 - t0).count();
}

void ref_refine_ranges(const float *S, const float *smin, const float *smax, int nx, int ny, int emin, int emax,
                       float *out, float *outcost, const char *refinement) {
   Img slo = make_img(smin, nx, ny, 1), shi = make_img(smax, nx, ny, 1);
   costvolume_t SS = volume_from_dense(S, slo, shi, emin, emax - emin + 1);
   std::vector<float> o(out, out + (size_t)nx * ny), oc(outcost, outcost + (size_t)nx * ny);
   subpixel_refinement_sgm(SS, o, oc, (char *)refinement);
   memcpy(out, &o[0], sizeof(float) * o.size());
   memcpy(outcost, &oc[0], sizeof(float) * oc.size());
}

}  // extern "C"
