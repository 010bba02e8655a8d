// `mgm` command line on the B200 hot path: same options, environment variables, console output and file
// outputs as the reference CLI (mgm.cc:266-450), with the hot path (weights, cost volume, aggregation, WTA,
// sub-pixel: mgm.cc:372-385 and :405-414) and the O(W*H) stages around it (left-right test mgm.cc:68-91, median
// img_tools.h:203-238, range update mgm.cc:120-158, back-projection mgm.cc:432-443) executed by libmgmb200 on the
// GPU.  The default flow is one mgmb200_stereo_lr call: the images go up once and the maps come back once.
// -m/-M range images and TSGM_ITER > 1 take mgmb200_stereo_ranges per direction (mgm.cc:372-395 on the device: cost
// volume once, then mgm + refinement + range update per iteration); MGMB200_STEPWISE=1 runs the same sequence call
// by call through the mirror functions of mgmb200_host.hpp (volumes cross the host between the calls).  -O 9..16 run the
// sweeps 8-15 DEFINED by this build (include/mgmb200.h; the reference reads past its 8-entry table there).  Not
// supported (exit code 2 or 3 with a message, never a silent fallback): -O > 16, MEDIAN > 7.
#include <algorithm>
#include <cmath>

#include "imgio.hpp"

using namespace mgmb200;

// The command-line SURFACE is kept verbatim from the reference, as north_star asks: pick_option below is taken from
// mgm.cc:165-179 ("-name value" anywhere on the line, removed in place; with d == NULL a flag) and the option block in
// main from mgm.cc:303-325 -- same names, defaults and order, because the order of extraction is observable (an
// option value that looks like an option name).  About 40 lines of glue; everything behind them is this build's.
static char *pick_option(int *c, char ***v, const char *o, const char *d) {
   int argc = *c;
   char **argv = *v;
   const int id = d ? 1 : 0;
   for (int i = 0; i < argc - id; i++)
      if (argv[i][0] == '-' && 0 == strcmp(argv[i] + 1, o)) {
         char *r = argv[i + id] + 1 - id;
         *c -= id + 1;
         for (int j = i; j < argc - id; j++) (*v)[j] = (*v)[j + id + 1];
         return r;
      }
   return const_cast<char *>(d);
}

MGMB200_SMART_PARAMETER(TSGM, 4)
MGMB200_SMART_PARAMETER(TSGM_FIX_OVERCOUNT, 1)
MGMB200_SMART_PARAMETER(USE_TRUNCATED_LINEAR_POTENTIALS, 0)
MGMB200_SMART_PARAMETER(WITH_MGM2, 0)
MGMB200_SMART_PARAMETER(TSGM_ITER, 1)
MGMB200_SMART_PARAMETER(TESTLRRL, 1)
MGMB200_SMART_PARAMETER(TESTLRRL_TAU, 1.0)
MGMB200_SMART_PARAMETER(MEDIAN, 0)

static void remove_nonfinite_values_Img(Img &u, float newval) {
   for (float &f : u.data) if (!std::isfinite(f)) f = newval;
}

static const char *help_usage = "usage:\n\tmgm [-options] u v out [cost [backflow]]";

int main(int argc, char *argv[]) {
   if (argc < 2 || !strcmp(argv[1], "-h")) return 0 * puts(help_usage);
   if (!strcmp(argv[1], "-?")) return 0 * puts("Compute stereo disparities by the MGM algorithm.");
   if (!strcmp(argv[1], "--version")) return 0 * puts("mgm 2.0 (mgm_b200)");
   if (!strcmp(argv[1], "--help")) {
      puts("mgm [options] in_u in_v out_disp [out_cost [out_backflow]]\n"
           " -r {-30} -R {30} -O {4} -P1 {8} -P2 {32} -p {none|census|sobelx|gblur} -t {ad|sd|census|ncc|btad|btsd}\n"
           " -truncDist {inf} -s {none|vfit|parabola|cubic|parabolaOCV} -aP1 {1} -aP2 {1} -aThresh {5}\n"
           " -m FILE -M FILE -l FILE\n"
           "Images: PNM, PFM, NPY, PNG (non-interlaced), uncompressed TIFF in; .npy, .tif/.tiff (float32), .pfm out.\n"
           "-O 9..16 add the knight-move sweeps defined by this build (the reference is undefined beyond 8).\n"
           "u and v must have the same size and channels; at most 65535 rows, 4096 disparities, MEDIAN <= 7.\n"
           "Environment: CENSUS_NCC_WIN=3 TESTLRRL=1 TESTLRRL_TAU=1.0 MEDIAN=0 TSGM=4 TSGM_ITER=1\n"
           "             TSGM_FIX_OVERCOUNT=1 USE_TRUNCATED_LINEAR_POTENTIALS=0 MGMB200_DEVICE=<gpu>\n"
           "Same surface as gfacciol/mgm; the hot path runs on the GPU (libmgmb200).");
      return 0;
   }
   if (argc < 4) {
      fprintf(stderr, "too few parameters\n   %s\n", help_usage);
      return 1;
   }
   try {
      // option extraction: names, defaults and order of mgm.cc:303-318
      char *in_min_disp_file = pick_option(&argc, &argv, "m", "");
      char *in_max_disp_file = pick_option(&argc, &argv, "M", "");
      int dmin = atoi(pick_option(&argc, &argv, "r", "-30"));
      int dmax = atoi(pick_option(&argc, &argv, "R", "30"));
      int NDIR = atoi(pick_option(&argc, &argv, "O", "4"));
      float P1 = atof(pick_option(&argc, &argv, "P1", "8"));
      float P2 = atof(pick_option(&argc, &argv, "P2", "32"));
      float aP1 = atof(pick_option(&argc, &argv, "aP1", "1"));
      float aP2 = atof(pick_option(&argc, &argv, "aP2", "1"));
      float aThresh = atof(pick_option(&argc, &argv, "aThresh", "5"));
      char *distance = pick_option(&argc, &argv, "t", "ad");
      char *prefilter = pick_option(&argc, &argv, "p", "none");
      char *refine = pick_option(&argc, &argv, "s", "none");
      float truncDist = atof(pick_option(&argc, &argv, "truncDist", "inf"));
      char *nolr_disp_file = pick_option(&argc, &argv, "l", "");
      (void)aP1;   // parsed but unused by the reference as well (mgm.cc:372)

      int i = 1;
      char *filename_u = (argc > i) ? argv[i] : NULL; i++;
      char *filename_v = (argc > i) ? argv[i] : NULL; i++;
      char *filename_out = (argc > i) ? argv[i] : NULL; i++;
      char *filename_cost = (argc > i) ? argv[i] : NULL; i++;
      char *filename_back = (argc > i) ? argv[i] : NULL; i++;
      if (!filename_u || !filename_v || !filename_out) { fprintf(stderr, "too few parameters\n"); return 1; }

      printf("%d %d\n", dmin, dmax);

      const bool have_ranges = strcmp(in_min_disp_file, "") != 0;   // mgm.cc:342 tests the -m file only
      const bool stepwise = have_ranges || (int)TSGM_ITER() != 1;
      if (WITH_MGM2()) fprintf(stderr, "mgm_b200: WITH_MGM2 ignored (the sweeps always run concurrently, summed in order)\n");

      Img u = io::read_image(filename_u);
      Img v = io::read_image(filename_v);
      remove_nonfinite_values_Img(u, 0);
      remove_nonfinite_values_Img(v, 0);
      if (u.nx != v.nx || u.ny != v.ny || u.nch != v.nch) {
         fprintf(stderr, "mgm_b200: u and v must have the same size and channels\n");
         return 2;
      }

      mgmb200_stereo_params p;
      mgmb200_stereo_params_default(&p);
      p.P1 = P1; p.P2 = P2;   // scaled by nch inside mgmb200_stereo exactly like mgm.cc:356-357
      p.NDIR = NDIR;
      p.MGM = (int)TSGM();   // double -> int truncation like mgm.cc:380
      p.use_felzenszwalb_potentials = (int)USE_TRUNCATED_LINEAR_POTENTIALS();
      p.sgm_fix_overcount = (int)TSGM_FIX_OVERCOUNT();
      p.aP = aP2; p.aThresh = aThresh;
      p.prefilter = prefilter; p.distance = distance; p.truncDist = truncDist;
      p.census_ncc_win = (int)CENSUS_NCC_WIN();
      p.refinement = refine;

      Img outoff(u.nx, u.ny), outcost(u.nx, u.ny), outoffR(v.nx, v.ny), outcostR(v.nx, v.ny);
      // range images, mgm.cc:338-353 (L->R) and :366-368 (R->L: always the uniform mirrored range)
      Img dminI(u.nx, u.ny), dmaxI(u.nx, u.ny), dminRI(v.nx, v.ny), dmaxRI(v.nx, v.ny);
      for (int k = 0; k < u.npix; k++) { dminI[k] = dmin; dmaxI[k] = dmax; }
      for (int k = 0; k < v.npix; k++) { dminRI[k] = -dmax; dmaxRI[k] = -dmin; }
      if (have_ranges) {
         dminI = io::read_image(in_min_disp_file);
         dmaxI = io::read_image(in_max_disp_file);
         if (dminI.npix != u.npix || dmaxI.npix != u.npix) { fprintf(stderr, "mgm_b200: range images must have the size of u\n"); return 2; }
         remove_nonfinite_values_Img(dminI, dmin);
         remove_nonfinite_values_Img(dmaxI, dmax);
         for (int k = 0; k < u.npix; k++)
            if (dmaxI[k] < dminI[k] + 1) dmaxI[k] = ceil(dminI[k] + 1);
      }
      const bool call_by_call = getenv("MGMB200_STEPWISE") != NULL;   // keep every volume on the host between the calls
      auto run = [&](Img &a, Img &b, Img &off, Img &cost, Img &loI, Img &hiI) {
         if (!call_by_call) {
            // the same sequence resident on the device (cost volume built once, ranges updated in place)
            check(mgmb200_stereo_ranges(context(), a.data.data(), b.data.data(), a.nx, a.ny, a.nch, &p, loI.data.data(),
                                        hiI.data.data(), (int)TSGM_ITER(), off.data.data(), cost.data.data()));
            for (int it = 0; it < (int)TSGM_ITER(); it++) {
               for (int pass = 0; pass < NDIR; pass++) printf("%d", pass);   // mgm_core.cc:491
               printf("\n");                                                  // mgm_print_energy.h:110
            }
            return;
         }
         // mgm.cc:372-395 call by call: the cost volume keeps its ranges, every iteration aggregates, refines and
         // updates the range images the next winner-take-all is restricted to
         Img w = compute_mgm_weights(a, aP2, aThresh);
         costvolume_t CC = allocate_and_fill_sgm_costvolume(a, b, loI, hiI, prefilter, distance, truncDist);
         const float P1s = P1 * a.nch, P2s = P2 * a.nch;   // mgm.cc:356-357
         for (int it = 0; it < (int)TSGM_ITER(); it++) {
            costvolume_t S = mgm(CC, w, loI, hiI, &off, &cost, P1s, P2s, NDIR, (int)TSGM(),
                                 (int)USE_TRUNCATED_LINEAR_POTENTIALS(), (int)TSGM_FIX_OVERCOUNT());
            printf("\n");
            subpixel_refinement_sgm(S, off.data, cost.data, refine);
            const std::pair<float, float> gminmax = update_dmin_dmax(off, &loI, &hiI);
            remove_nonfinite_values_Img(loI, gminmax.first);
            remove_nonfinite_values_Img(hiI, gminmax.second);
         }
      };
      Img syn(u.nx, u.ny, u.nch);
      if (!stepwise) {
         // mgm.cc:372-443 in one device-resident call
         mgmb200_post_params q;
         q.testlrrl = TESTLRRL() ? 1 : 0;
         q.testlrrl_tau = (float)TESTLRRL_TAU();
         q.median = (int)MEDIAN();
         p.dmin = dmin; p.dmax = dmax;
         Img nolr(u.nx, u.ny);
         check(mgmb200_stereo_lr(context(), u.data.data(), v.data.data(), u.nx, u.ny, u.nch, &p, &q, outoff.data.data(),
                                 outcost.data.data(), NULL, NULL, nolr.data.data(), syn.data.data()));
         for (int r = 0; r < (q.testlrrl ? 2 : 1); r++) {
            for (int pass = 0; pass < NDIR; pass++) printf("%d", pass);   // mgm_core.cc:491
            printf("\n");                                                  // print_solution_energy, mgm_print_energy.h:110
         }
         if (0 != strcmp(nolr_disp_file, "")) io::write_image(nolr_disp_file, nolr);
      } else {
         run(u, v, outoff, outcost, dminI, dmaxI);
         if (MEDIAN()) outoff = median_filter(outoff, (int)MEDIAN());
         if (0 != strcmp(nolr_disp_file, "")) io::write_image(nolr_disp_file, outoff);
         if (TESTLRRL()) {
            run(v, u, outoffR, outcostR, dminRI, dmaxRI);
            if (MEDIAN()) outoffR = median_filter(outoffR, (int)MEDIAN());
            Img tmpL(outoff), tmpR(outoffR);
            leftright_test(outoffR, tmpL, (float)TESTLRRL_TAU());
            leftright_test(outoff, tmpR, (float)TESTLRRL_TAU());
         }
         syn = backproject(outoff, u, v);
      }

      io::write_image(filename_out, outoff);
      if (filename_cost) io::write_image(filename_cost, outcost);
      if (filename_back) io::write_image(filename_back, syn);
      return 0;
   } catch (const std::exception &e) {
      fprintf(stderr, "%s\n", e.what());
      return 3;
   }
}
