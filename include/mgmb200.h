/* mgmb200.h -- C ABI of the B200-native MGM stereo hot path.
 *
 * Drop-in boundary for the hot path of gfacciol/mgm (cost volume -> multi-direction
 * MGM aggregation -> WTA + sub-pixel).  Every entry point states the reference
 * interface it replaces (file:line under the reference tree).  Plain pointers and
 * sizes only; the library owns nothing but the opaque context (stream, device
 * scratch).  All functions return 0 on success or a negative MGMB200_E* code;
 * mgmb200_last_error() gives the message of the last failure on the calling
 * thread.  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with MGMB200_ECUDA.
 *
 * Host layouts (identical to the reference):
 *   image    planar float  data[x + y*nx + c*nx*ny]                       img.h:35
 *   weights  Img(nx,ny,8), planes W,E,S,N,NW,NE,SE,SW                      mgm_weights.h:69
 *   volume   pixel-major, label-fastest  vol[(x + y*nx)*L + (o - dmin)]    dvec.cc:129,
 *            i.e. the flat layout of mgm_costvolume.h:276-299; L = dmax-dmin+1
 * Device layout of a volume ("padded"): [ny*nx][VS] floats with
 *   VS = mgmb200_padded_labels(L); labels >= L hold +INF.
 */
#ifndef MGMB200_H_
#define MGMB200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGMB200_VERSION 100

enum {
   MGMB200_OK = 0,
   MGMB200_EINVAL = -1,       /* bad argument */
   MGMB200_ECUDA = -2,        /* CUDA runtime failure (incl. no device) */
   MGMB200_EUNSUPPORTED = -3, /* outside the supported envelope */
   MGMB200_ENOMEM = -4
};

/* mgm_costvolume.h:170-190  distance table: ad sd census ncc btad btsd (unknown -> 0) */
int mgmb200_distance_index(const char *name);
/* mgm_costvolume.h:194-207  prefilter table: none census sobelx gblur (unknown -> 0) */
int mgmb200_prefilter_index(const char *name);
/* mgm_refine.h:14-35        refinement table: none vfit parabola cubic parabolaOCV (unknown -> 0) */
int mgmb200_refinement_index(const char *name);

typedef struct mgmb200_ctx mgmb200_ctx;

int mgmb200_version(void);
const char *mgmb200_last_error(void);

/* One context per GPU (one process per GPU in multi-GPU runs).  device < 0: current device. */
int mgmb200_create(int device, mgmb200_ctx **ctx);
void mgmb200_destroy(mgmb200_ctx *ctx);
/* Use a caller-provided CUDA stream (cudaStream_t as void*); NULL restores the context's own. */
int mgmb200_set_stream(mgmb200_ctx *ctx, void *cuda_stream);
/* Tuning / test knob: rows per band of the aggregation wavefront (0 = derive from shared memory). */
int mgmb200_set_rows_per_band(mgmb200_ctx *ctx, int rows);
/* Tuning / debugging knobs of the aggregation launch (no reference counterpart).  They are read ONCE from the
 * environment by mgmb200_create (MGMB200_<NAME>) and changed afterwards only through this call -- nothing on the
 * per-call path reads the environment.  Names: rows_per_band, rows_axis, rows_diag, groups, no_creg, no_fused_sgm, reg_chains,
 * no_lean_sgm / no_lean_trunc (run the unweighted potentials through the generic kernel instead of the lean ones), full_block,
 * lanes4, lanes8, no_shear, static_order, no_fused_finish, fused_finish, fin_tile ("WxH"), cc_pf, lr_sequential, batch (pairs per launch of the batch
 * entry points), verbose; "reset" re-reads the
 * environment.  Every layout they select is parity-tested (tests/test_gpu_parity.py). */
int mgmb200_set_option(mgmb200_ctx *ctx, const char *name, const char *value);
/* Block until everything queued on the context's stream is finished. */
int mgmb200_synchronize(mgmb200_ctx *ctx);

int mgmb200_padded_labels(int L);
size_t mgmb200_volume_bytes(int nx, int ny, int L);

/* ------------------------------------------------------------------------------------------
 * Host-pointer entry points: the reference's operator interface, one call per reference
 * function.  Host<->device copies happen inside the call.
 * ---------------------------------------------------------------------------------------- */

/* struct Img compute_mgm_weights(struct Img &u, float aP, float aThresh)        mgm_weights.h:63
 * w_out: 8*nx*ny floats. */
int mgmb200_compute_mgm_weights(mgmb200_ctx *ctx, const float *u, int nx, int ny, int nch, float aP,
                                float aThresh, float *w_out);

/* struct costvolume_t allocate_and_fill_sgm_costvolume(Img &in_u, Img &in_v, Img &dminI, Img &dmaxI,
 *        char *prefilter, char *distance, float truncDist)                      mgm_costvolume.h:337
 * Uniform disparity range [dmin,dmax]; census_ncc_win is the CENSUS_NCC_WIN environment
 * parameter (mgm_costvolume.h:61).  cc_out: nx*ny*L floats (host volume layout). */
int mgmb200_costvolume(mgmb200_ctx *ctx, const float *u, const float *v, int nx, int ny, int nch, int vnx,
                       int vny, int dmin, int dmax, const char *prefilter, const char *distance,
                       float truncDist, int census_ncc_win, float *cc_out);

/* struct costvolume_t mgm(struct costvolume_t CC, const Img &in_w, const Img &dminI, const Img &dmaxI,
 *        Img *out, Img *outcost, float P1, float P2, int NDIR, int MGM,
 *        int USE_FELZENSZWALB_POTENTIALS, int SGM_FIX_OVERCOUNT)                 mgm_core.cc:408
 * cc: nx*ny*L floats; w: 8*nx*ny floats or NULL (all ones); out/outcost: nx*ny floats;
 * S_out: NULL or nx*ny*L floats receiving the returned (over-count corrected) volume.
 * NDIR 1..8 are the reference's sweeps; 9..16 add the sweeps 8-15 DEFINED by this library (the reference advertises
 * -O 16, mgm.cc:223, but reads past its 8-entry table, mgm_core.cc:463-473): sweep 8+b scans like sweep b and takes the
 * same four neighbours in an order given by the parity of the scan coordinates, so that the neighbour chains follow the
 * knight-move (22.5 degree) directions (mgm_b200/csrc/common.cuh knight_pred_type).
 * Inputs are scanned: volumes with NaN / -INF costs or vectors without a finite entry, negative or non-finite weights
 * or penalties run on a slower compare-select kernel that propagates them like the reference's macros
 * (mgm_core.cc:47-60, dvec.cc:81-88); a pixel without any finite label gets a NaN disparity (uninitialised in the
 * reference, mgm_core.cc:594). */
int mgmb200_mgm(mgmb200_ctx *ctx, const float *cc, const float *w, int nx, int ny, int dmin, int dmax,
                float P1, float P2, int NDIR, int MGM, int use_felzenszwalb_potentials,
                int sgm_fix_overcount, float *out, float *outcost, float *S_out);

/* Same, cost volume given as label-major planes costs[i + o*nx*ny] -- the input.bin protocol of
 * the standalone aggregator (matlab/mgm_o.cc:544-595, MGM_wrapper.m:88-94). Labels 0..nlab-1. */
int mgmb200_mgm_labelmajor(mgmb200_ctx *ctx, const float *costs, const float *w, int ncol, int nrow, int nlab,
                           float P1, float P2, int NDIR, int MGM, int use_felzenszwalb_potentials,
                           float *labels_out, float *outcost);

/* void subpixel_refinement_sgm(struct costvolume_t &S, std::vector<float> &out,
 *        std::vector<float> &outcost, char *refinement)                          mgm_refine.h:40
 * S: nx*ny*L floats; out/outcost updated in place. */
int mgmb200_subpixel_refinement_sgm(mgmb200_ctx *ctx, const float *S, int nx, int ny, int dmin, int dmax,
                                    float *out, float *outcost, const char *refinement);

/* ---- per-pixel disparity ranges: the dminI / dmaxI images of the three functions above (-m / -M range files
 * mgm.cc:342-353, range updates between TSGM_ITER iterations mgm.cc:377-395).  A Dvec only holds its own
 * [min,max] (dvec.cc:49-131); here volumes stay DENSE over an envelope [emin,emax] (L = emax-emin+1 labels per
 * pixel) and a label outside a pixel's range holds +INF, which is what Dvec::operator[] returns for it.
 * Range images are floats and are truncated to int like Dvec::init does (mgm_costvolume.h:323).          */

/* allocate_and_fill_sgm_costvolume(in_u, in_v, dminI, dmaxI, ...)                 mgm_costvolume.h:337-424
 * cc_out: nx*ny*L floats, +INF outside [dminI,dmaxI]. */
int mgmb200_costvolume_ranges(mgmb200_ctx *ctx, const float *u, const float *v, int nx, int ny, int nch, int vnx,
                              int vny, const float *dminI, const float *dmaxI, int emin, int emax,
                              const char *prefilter, const char *distance, float truncDist, int census_ncc_win,
                              float *cc_out);

/* mgm(CC, in_w, dminI, dmaxI, ...)                                                   mgm_core.cc:408-613
 * cc: nx*ny*L floats whose vectors have the ranges [ccmin,ccmax] (entries outside are ignored); dminI/dmaxI: the
 * ranges of the returned volume S, over which the winner is taken (they differ from the former from the second
 * TSGM_ITER iteration on, mgm.cc:377-395).  S_out: NULL or nx*ny*L floats, +INF outside [dminI,dmaxI].
 * Truncated-linear potentials with non-uniform cost ranges: MGM=2 without image-dependent weights folds the
 * out-of-range labels back in (mgm_core.cc:166-219), every other variant convolves inside the receiving pixel's
 * range (mgm_core.cc:229-281); both are reproduced. */
int mgmb200_mgm_ranges(mgmb200_ctx *ctx, const float *cc, const float *ccmin, const float *ccmax, const float *w,
                       int nx, int ny, int emin, int emax, const float *dminI, const float *dmaxI, float P1,
                       float P2, int NDIR, int MGM, int use_felzenszwalb_potentials, int sgm_fix_overcount,
                       float *out, float *outcost, float *S_out);

/* subpixel_refinement_sgm(S, out, outcost, refinement) on a volume with the ranges [dminI,dmaxI]
 * (the test against S[i].min / S[i].max of mgm_refine.h:58). */
int mgmb200_subpixel_refinement_sgm_ranges(mgmb200_ctx *ctx, const float *S, const float *dminI, const float *dmaxI,
                                           int nx, int ny, int emin, int emax, float *out, float *outcost,
                                           const char *refinement);

/* The whole hot path as mgm.cc:356-385 strings it together for one direction of the LR pair:
 * weights(u) -> cost volume(u,v) -> mgm -> sub-pixel, nothing but the images going in and the two
 * maps coming out.  P1/P2 are the command-line values: they are multiplied by nch here exactly
 * like mgm.cc:356-357.  aP is the CLI's -aP2 (mgm.cc:372). */
typedef struct mgmb200_stereo_params {
   int dmin, dmax;                /* -r -R */
   float P1, P2;                  /* -P1 -P2 */
   int NDIR;                      /* -O */
   int MGM;                       /* env TSGM */
   int use_felzenszwalb_potentials; /* env USE_TRUNCATED_LINEAR_POTENTIALS */
   int sgm_fix_overcount;         /* env TSGM_FIX_OVERCOUNT */
   float aP, aThresh;             /* -aP2 -aThresh */
   const char *prefilter;         /* -p */
   const char *distance;          /* -t */
   float truncDist;               /* -truncDist */
   int census_ncc_win;            /* env CENSUS_NCC_WIN */
   const char *refinement;        /* -s */
} mgmb200_stereo_params;
void mgmb200_stereo_params_default(mgmb200_stereo_params *p);   /* defaults of mgm.cc:186-196,303-318 */
/* Boundary limitations of the fused mgmb200_stereo* calls (all fail with a message, none silently):
 *  - u and v must have the same size (the reference lets them differ: outoffR then has v's size, mgm.cc:404-424;
 *    the stage-by-stage calls mgmb200_costvolume / _dev take vnx, vny and do support it);
 *  - at most 65535 image rows, 4096 labels, 16 sweeps, median radius 7.
 * Limitations of mgmb200_mgm_ranges (per-pixel ranges): costs must be finite where present, weights finite and >= 0,
 * penalties >= 0 -- the compare-select kernel that reproduces the reference on non-finite inputs (mgmb200_mgm,
 * mgmb200_mgm_labelmajor) covers uniform ranges only. */
int mgmb200_stereo(mgmb200_ctx *ctx, const float *u, const float *v, int nx, int ny, int nch,
                   const mgmb200_stereo_params *p, float *out, float *outcost);
/* The same for npairs pairs of one shape (a driver looping over frames or tiles): images up, weights + cost volume per
 * pair, the aggregation of several pairs per launch (mgmb200_aggregate_batch_dev), maps down.  u, v, out, outcost are
 * arrays of npairs host pointers.  Results are those of npairs mgmb200_stereo calls. */
int mgmb200_stereo_batch(mgmb200_ctx *ctx, int npairs, const float *const *u, const float *const *v, int nx, int ny,
                         int nch, const mgmb200_stereo_params *params, float *const *out, float *const *outcost);

/* ---- the O(W*H) stages around the hot path in the CLI flow (mgm.cc:396-443), as small kernels so that the two
 * directions, the median, the left-right tests and the back-projection stay on the device. */

/* leftright_test(dx, Rdx, threshold)                                                        mgm.cc:68-91
 * dx (nx*ny, tested in place) against the map of the other view Rdx (rnx*rny, rny >= ny). */
int mgmb200_leftright_test(mgmb200_ctx *ctx, float *dx, int nx, int ny, const float *Rdx, int rnx, int rny,
                           float threshold);
/* median_filter(u, radius)                                                            img_tools.h:203-238
 * NaN-aware median of the (2*radius+1)^2 window clipped to the image; radius 0..7. */
int mgmb200_median_filter(mgmb200_ctx *ctx, const float *u, int nx, int ny, int nch, int radius, float *out);
/* update_dmin_dmax(outoff, &dminI, &dmaxI, slack=3, radius=2)                             mgm.cc:120-158
 * dminI/dmaxI updated in place; *gmin,*gmax (may be NULL) = the returned finite min/max of outoff. */
int mgmb200_update_dmin_dmax(mgmb200_ctx *ctx, const float *outoff, int nx, int ny, float *dminI, float *dmaxI,
                             int slack, int radius, float *gmin, float *gmax);
/* the back-projected image of mgm.cc:432-443: v sampled at x + outoff where that is inside v, else u */
int mgmb200_backproject(mgmb200_ctx *ctx, const float *outoff, const float *u, const float *v, int nx, int ny,
                        int nch, int vnx, int vny, float *syn);

/* The default command-line flow mgm.cc:372-443 (uniform range, TSGM_ITER=1) in one call: L->R run, median,
 * R->L run with the mirrored range, median, both left-right tests, back-projection.  out/outcost are required;
 * outR, outcostR (need testlrrl), out_nolr (the -l map: after the median, before the test) and backproj
 * (nx*ny*nch) may be NULL. */
typedef struct mgmb200_post_params {
   int testlrrl;        /* env TESTLRRL */
   float testlrrl_tau;  /* env TESTLRRL_TAU */
   int median;          /* env MEDIAN: radius, 0 = off */
} mgmb200_post_params;
void mgmb200_post_params_default(mgmb200_post_params *q);   /* mgm.cc:194-196 */
int mgmb200_stereo_lr(mgmb200_ctx *ctx, const float *u, const float *v, int nx, int ny, int nch,
                      const mgmb200_stereo_params *p, const mgmb200_post_params *q, float *out, float *outcost,
                      float *outR, float *outcostR, float *out_nolr, float *backproj);

/* One direction of mgm.cc:372-395 with per-pixel range images (-m / -M, or the uniform range) and TSGM_ITER
 * iterations, device resident: cost volume over [dminI,dmaxI] once, then per iteration mgm() + sub-pixel refinement
 * over the current ranges and update_dmin_dmax + remove_nonfinite_values_Img on them.  dminI/dmaxI (nx*ny, finite,
 * min <= max after truncation) are updated in place; p->dmin / p->dmax are not used. */
int mgmb200_stereo_ranges(mgmb200_ctx *ctx, const float *u, const float *v, int nx, int ny, int nch,
                          const mgmb200_stereo_params *p, float *dminI, float *dmaxI, int tsgm_iter, float *out,
                          float *outcost);

/* ------------------------------------------------------------------------------------------
 * Device-pointer entry points (inputs and outputs resident in HBM, asynchronous on the
 * context's stream).  d_cc is a padded volume.
 * ---------------------------------------------------------------------------------------- */
int mgmb200_weights_dev(mgmb200_ctx *ctx, const float *d_u, int nx, int ny, int nch, float aP, float aThresh,
                        float *d_w);
int mgmb200_costvolume_dev(mgmb200_ctx *ctx, const float *d_u, const float *d_v, int nx, int ny, int nch,
                           int vnx, int vny, int dmin, int dmax, int prefilter_index, int distance_index,
                           float truncDist, int census_ncc_win, float *d_cc);
/* weights_mode: 0 = d_w ignored (all ones), 1 = use d_w as image dependent weights,
 *               2 = decide like mgm_core.cc:420-422 (scans d_w, synchronises the stream).
 * refinement_index as mgmb200_refinement_index.  d_S: NULL or dense nx*ny*L floats.
 */
int mgmb200_aggregate_dev(mgmb200_ctx *ctx, const float *d_cc, const float *d_w, int weights_mode, int nx,
                          int ny, int dmin, int dmax, float P1, float P2, int NDIR, int MGM,
                          int use_felzenszwalb_potentials, int sgm_fix_overcount, int refinement_index,
                          float *d_out, float *d_outcost, float *d_S);

/* A batch of npairs stereo pairs of one shape and one parameter set (BASELINE.json configs[3]; the reference
 * processes one pair per process run, mgm.cc:266-450): the sweeps of several pairs share one launch.  d_w: NULL, or
 * npairs weight-plane pointers (then the per-edge kernels run, as weights_mode 1).  Equivalent to npairs
 * mgmb200_aggregate_dev calls. */
int mgmb200_aggregate_batch_dev(mgmb200_ctx *ctx, int npairs, const float *const *d_cc, const float *const *d_w, int nx,
                                int ny, int dmin, int dmax, float P1, float P2, int NDIR, int MGM,
                                int use_felzenszwalb_potentials, int sgm_fix_overcount, int refinement_index,
                                float *const *d_out, float *const *d_outcost);

/* Multi-GPU direction sharding (the reference's own precedent is mgm_naive_parallelism,
 * mgm_core.cc:632-831: one thread per sweep, S += Lr at the end).  One process per GPU; sweep p is aggregated by one
 * rank (sweep_mask), and the sum S = sum_p L_p (mgm_core.cc:582-587) is the only exchange.  Two exchanges are provided:
 *
 * (1) ORDERED, bit-exact: the image rows are cut into nslabs slabs of slab_rows rows, slab r is finished by rank r.
 *     Every rank allocates all NDIR message volumes (mgmb200_sweeps_alloc), exports them (mgmb200_ipc_export) and maps
 *     the other ranks' (mgmb200_ipc_open).  mgmb200_aggregate_sweeps_slabs_dev then stores the messages of slab r of
 *     sweep p into d_slab_volumes[p*nslabs + r] -- rank r's volume p, a peer mapping for r != own rank: the stores
 *     cross NVLink from inside the aggregation kernel while the sweep runs.  After one barrier between the ranks
 *     every rank holds all sweeps of its slab locally and mgmb200_finish_rows_dev adds them IN SWEEP ORDER.
 * (2) ALL-REDUCE, as north_star names it: mgmb200_aggregate_sweeps_dev into local volumes, mgmb200_sum_sweeps_dev adds
 *     the rank's sweeps into one partial volume, the caller all-reduces the partial volumes (NCCL), and
 *     mgmb200_finish_sum_dev applies fix + WTA + refinement.  The summation order is not the reference's.
 */
int mgmb200_aggregate_sweeps_dev(mgmb200_ctx *ctx, const float *d_cc, const float *d_w, int weights_mode,
                                 int nx, int ny, int dmin, int dmax, float P1, float P2, int NDIR, int MGM,
                                 int use_felzenszwalb_potentials, unsigned sweep_mask);
int mgmb200_sweeps_alloc(mgmb200_ctx *ctx, int nx, int ny, int dmin, int dmax, int NDIR);
/* frees the message volumes (also the exported ones: peers must have closed their mappings) */
int mgmb200_sweeps_release(mgmb200_ctx *ctx);
int mgmb200_aggregate_sweeps_slabs_dev(mgmb200_ctx *ctx, const float *d_cc, const float *d_w, int weights_mode, int nx,
                                       int ny, int dmin, int dmax, float P1, float P2, int NDIR, int MGM,
                                       int use_felzenszwalb_potentials, unsigned sweep_mask, int nslabs, int slab_rows,
                                       float *const *d_slab_volumes);
/* device pointer of the volume this context holds for sweep p (after mgmb200_sweeps_alloc or an aggregation) */
int mgmb200_sweep_volume(mgmb200_ctx *ctx, int sweep, float **d_ptr, size_t *bytes);
/* 64-byte CUDA IPC handle of a device allocation / mapping of a peer handle.  An exported message volume is never
 * reallocated: a later call that needs a larger one fails until mgmb200_sweeps_release. */
int mgmb200_ipc_export(mgmb200_ctx *ctx, const void *d_ptr, unsigned char handle_out[64]);
int mgmb200_ipc_open(mgmb200_ctx *ctx, const unsigned char handle[64], void **d_ptr);
int mgmb200_ipc_close(mgmb200_ctx *ctx, void *d_ptr);
/* ordered sum + fix + WTA + sub-pixel over rows [row_begin,row_end) with explicit per-sweep
 * volume pointers (local or peer); d_out/d_outcost are full nx*ny maps, only the slab is written */
int mgmb200_finish_rows_dev(mgmb200_ctx *ctx, const float *const *d_sweeps, const float *d_cc, int nx, int ny,
                            int dmin, int dmax, int NDIR, int sgm_fix_overcount, int refinement_index,
                            int row_begin, int row_end, float *d_out, float *d_outcost);
/* d_sum (padded volume) = sum of this context's sweeps in sweep_mask, in increasing sweep order */
int mgmb200_sum_sweeps_dev(mgmb200_ctx *ctx, int nx, int ny, int dmin, int dmax, unsigned sweep_mask, float *d_sum);
/* over-count fix for NDIR sweeps + WTA + sub-pixel on a pre-summed padded volume */
int mgmb200_finish_sum_dev(mgmb200_ctx *ctx, const float *d_sum, const float *d_cc, int nx, int ny, int dmin, int dmax,
                           int NDIR, int sgm_fix_overcount, int refinement_index, int row_begin, int row_end,
                           float *d_out, float *d_outcost);

/* device-pointer forms of the post-processing stages; all out of place except update_dmin_dmax, whose
 * d_minmax is 4 floats of scratch that receive the finite min and max of d_outoff in [0] and [1] */
int mgmb200_leftright_test_dev(mgmb200_ctx *ctx, const float *d_dx, int nx, int ny, const float *d_Rdx, int rnx,
                               int rny, float threshold, float *d_out);
int mgmb200_median_filter_dev(mgmb200_ctx *ctx, const float *d_u, int nx, int ny, int nch, int radius,
                              float *d_out);
int mgmb200_update_dmin_dmax_dev(mgmb200_ctx *ctx, const float *d_outoff, int nx, int ny, float *d_dminI,
                                 float *d_dmaxI, int slack, int radius, float *d_minmax);
int mgmb200_backproject_dev(mgmb200_ctx *ctx, const float *d_outoff, const float *d_u, const float *d_v, int nx,
                            int ny, int nch, int vnx, int vny, float *d_syn);

/* device memory helpers so that non-CUDA hosts (ctypes, cgo, JNI) can drive the dev entry points */
int mgmb200_malloc(mgmb200_ctx *ctx, size_t bytes, void **d_ptr);
int mgmb200_free(mgmb200_ctx *ctx, void *d_ptr);
int mgmb200_memcpy_h2d(mgmb200_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int mgmb200_memcpy_d2h(mgmb200_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
/* dense host-layout volume <-> padded device volume (both device pointers) */
int mgmb200_pad_volume_dev(mgmb200_ctx *ctx, const float *d_dense, float *d_padded, int nx, int ny, int L,
                           int label_major);
int mgmb200_unpad_volume_dev(mgmb200_ctx *ctx, const float *d_padded, float *d_dense, int nx, int ny, int L);

/* counters of the last aggregate call, for benchmarks: kernels launched, rows per band used */
int mgmb200_last_launch_info(mgmb200_ctx *ctx, int *kernel_launches, int *rows_axis, int *rows_diag,
                             int *threads_per_cta, size_t *smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MGMB200_H_ */
