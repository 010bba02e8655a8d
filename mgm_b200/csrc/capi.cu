// C ABI of the B200 MGM hot path (include/mgmb200.h): context, buffer cache and the
// orchestration of the kernels in costvolume.cu / aggregate.cu / wta.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <ctype.h>
#include <algorithm>
#include <vector>

#include "../../include/mgmb200.h"
#include "aggregate.cuh"
#include "costvolume.cuh"
#include "wta.cuh"
#include "post.cuh"

using namespace mgm;

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(g_err, sizeof(g_err), fmt, ap);
   va_end(ap);
   return code;
}
#define CU(call)                                                                              \
   do {                                                                                       \
      cudaError_t e_ = (call);                                                                \
      if (e_ != cudaSuccess)                                                                  \
         return fail(MGMB200_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
   } while (0)
#define RET(call)              \
   do {                        \
      int r_ = (call);         \
      if (r_ != 0) return r_;  \
   } while (0)

struct DevBuf {
   void *p = nullptr;
   size_t cap = 0;
   bool exported = false;   // a CUDA IPC handle of this allocation is in other processes' hands: never reallocate it
   int reserve(size_t bytes) {
      if (bytes <= cap) return 0;
      if (exported)
         return fail(MGMB200_EINVAL, "a message volume exported to other processes (mgmb200_ipc_export) would have to grow "
                     "from %zu to %zu bytes: call mgmb200_sweeps_release and exchange the handles again", cap, bytes);
      if (p) cudaFree(p);
      p = nullptr; cap = 0;
      cudaError_t e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess) {
         cudaGetLastError();
         return fail(MGMB200_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
      }
      cap = bytes;
      return 0;
   }
   void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
   template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Small host->device tables (sweep descriptors, finish parameters, tile order) go through a ring of pinned staging
// slots: a plain cudaMemcpyAsync from pageable memory would synchronise the stream before every aggregation launch.
struct PinnedRing {
   static constexpr int NSLOT = 4;
   void *slot[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
   size_t cap[NSLOT] = {0, 0, 0, 0};
   cudaEvent_t ev[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
   int next = 0;
   // copies `bytes` from h to device memory d on stream st without blocking on earlier work of the stream
   cudaError_t push(void *d, const void *h, size_t bytes, cudaStream_t st) {
      const int i = next;
      next = (next + 1) % NSLOT;
      cudaError_t e;
      if (!ev[i]) { e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
      else { e = cudaEventSynchronize(ev[i]); if (e != cudaSuccess) return e; }   // the slot's previous copy has been consumed
      if (cap[i] < bytes) {
         if (slot[i]) cudaFreeHost(slot[i]);
         slot[i] = nullptr; cap[i] = 0;
         const size_t want = std::max(bytes, (size_t)65536);
         e = cudaMallocHost(&slot[i], want);
         if (e != cudaSuccess) return e;
         cap[i] = want;
      }
      memcpy(slot[i], h, bytes);
      e = cudaMemcpyAsync(d, slot[i], bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      return cudaEventRecord(ev[i], st);
   }
   void release() {
      for (int i = 0; i < NSLOT; i++) {
         if (ev[i]) cudaEventDestroy(ev[i]);
         if (slot[i]) cudaFreeHost(slot[i]);
         ev[i] = nullptr; slot[i] = nullptr; cap[i] = 0;
      }
   }
};

struct mgmb200_ctx {
   int device = 0, num_sms = 0, max_smem = 0;
   cudaStream_t own_stream = nullptr, stream = nullptr;
   int rows_override = 0;
   AggTuning tune;          // MGMB200_* knobs: read once by mgmb200_create, changed only by mgmb200_set_option
   // scratch (grow-only, reused across calls)
   DevBuf u, v, fu, fv, ftmp, cu, cv, w, cc, dense, out, outcost, flags, progress, bnd, bndm, rg[4], ncc;
   // per-pixel ranges of the call in flight (device pointers or nullptr): S range, cost-vector range
   const float *r_smin = nullptr, *r_smax = nullptr, *r_ccmin = nullptr, *r_ccmax = nullptr;
   bool r_window = false;   // truncated-linear update inside the receiving pixel's cost range (consumer-side kernels)
   int r_emin = 0;
   std::vector<DevBuf> sweepv;   // per-sweep message volumes, slot pair*MGM_MAX_NDIR + sweep
   DevBuf desc, fins;            // device tables of the launch in flight (SweepDesc, WtaParams per pair)
   std::vector<DevBuf> bcc, bw, bout;   // per-pair cost volumes, weights and maps of mgmb200_stereo_batch
   PinnedRing staging;
   DevBuf post[10];   // maps of the post-processing stages (N1/N2)
   // fused finish: tile order (expected readiness) cached per geometry
   DevBuf tiles;
   std::vector<int> tile_order;
   int tile_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
   bool fin_done = false;   // the last run_sweeps finished the maps inside the aggregation launch
   DevBuf dbgbuf;
   size_t sweep_bytes = 0;
   // last launch info
   int n_launches = 0, rows_axis = 0, rows_diag = 0, block = 0;
   size_t smem = 0;
};

// ------------------------------------------------------------------------------------------ tables
static int lookup(const char *name, const char *const *names, int n) {
   int r = 0;   // unknown names silently map to entry 0, like the reference
   if (!name) return 0;
   for (int i = 0; i < n; i++) if (!strcmp(name, names[i])) r = i;
   return r;
}
extern "C" int mgmb200_distance_index(const char *name) {
   static const char *const t[] = {"ad", "sd", "census", "ncc", "btad", "btsd"};
   return lookup(name, t, 6);
}
extern "C" int mgmb200_prefilter_index(const char *name) {
   static const char *const t[] = {"none", "census", "sobelx", "gblur"};
   return lookup(name, t, 4);
}
extern "C" int mgmb200_refinement_index(const char *name) {
   static const char *const t[] = {"none", "vfit", "parabola", "cubic", "parabolaOCV"};
   return lookup(name, t, 5);
}

extern "C" int mgmb200_version(void) { return MGMB200_VERSION; }
extern "C" const char *mgmb200_last_error(void) { return g_err; }
extern "C" int mgmb200_padded_labels(int L) { return (L + 31) & ~31; }
extern "C" size_t mgmb200_volume_bytes(int nx, int ny, int L) {
   return (size_t)nx * ny * mgmb200_padded_labels(L) * sizeof(float);
}

// ------------------------------------------------------------------------------------------ context
// Tuning / debugging knobs (names = the MGMB200_* environment variables without the prefix, lower case).  A NULL or
// empty value switches a flag on, like an environment variable that is merely set.
static bool apply_option(mgmb200_ctx *c, const char *name, const char *value) {
   AggTuning &t = c->tune;
   const int iv = (value && *value) ? atoi(value) : 1;
   if (!strcmp(name, "rows_per_band")) { c->rows_override = iv < 0 ? 0 : iv; return true; }
   if (!strcmp(name, "rows_axis")) { t.rows_axis = iv; return true; }
   if (!strcmp(name, "rows_diag")) { t.rows_diag = iv; return true; }
   if (!strcmp(name, "groups")) { t.groups = iv; return true; }
   if (!strcmp(name, "no_creg")) { t.no_creg = iv; return true; }
   if (!strcmp(name, "no_fused_sgm")) { t.no_fused_sgm = iv; return true; }
   if (!strcmp(name, "no_lean_sgm")) { t.no_lean_sgm = iv; return true; }
   if (!strcmp(name, "no_lean_trunc")) { t.no_lean_trunc = iv; return true; }
   if (!strcmp(name, "full_block")) { t.full_block = iv; return true; }
   if (!strcmp(name, "reg_chains")) { t.reg_chains = iv; return true; }
   if (!strcmp(name, "lanes4")) { t.lanes = iv ? 4 : 0; return true; }
   if (!strcmp(name, "lanes8")) { t.lanes = iv ? 8 : 0; return true; }
   if (!strcmp(name, "no_shear")) { t.no_shear = iv; return true; }
   if (!strcmp(name, "static_order")) { t.static_order = iv; return true; }
   if (!strcmp(name, "no_fused_finish")) { t.no_fused_finish = iv; return true; }
   if (!strcmp(name, "fused_finish")) { t.fused_finish = iv; return true; }
   if (!strcmp(name, "cc_pf")) { t.cc_pf = iv; return true; }
   if (!strcmp(name, "batch")) { t.batch = iv < 1 ? 1 : iv; return true; }
   if (!strcmp(name, "lr_sequential")) { t.lr_sequential = iv; return true; }
   if (!strcmp(name, "verbose")) { t.verbose = iv; return true; }
   if (!strcmp(name, "dbg")) { t.dbg = iv; return true; }
   if (!strcmp(name, "fin_tile")) {
      int a = 0, b = 0;
      if (!value || sscanf(value, "%dx%d", &a, &b) != 2 || a < 1 || b < 1) return false;
      t.fin_tw = a; t.fin_th = b;
      return true;
   }
   return false;
}
static void tuning_from_env(mgmb200_ctx *c) {
   c->tune = AggTuning();
   c->rows_override = 0;
   static const char *const names[] = {"rows_per_band", "rows_axis", "rows_diag", "groups", "no_creg", "no_fused_sgm", "no_lean_sgm", "no_lean_trunc", "full_block", "reg_chains", "lanes4",
                                       "lanes8", "no_shear", "static_order", "no_fused_finish", "fused_finish", "cc_pf", "batch", "lr_sequential", "verbose", "dbg", "fin_tile"};
   for (const char *n : names) {
      char env[64] = "MGMB200_";
      size_t k = strlen(env);
      for (const char *q = n; *q && k + 1 < sizeof(env); ++q) env[k++] = (char)toupper((unsigned char)*q);
      env[k] = 0;
      if (const char *e = getenv(env)) apply_option(c, n, e);
   }
}

extern "C" int mgmb200_create(int device, mgmb200_ctx **out) {
   if (!out) return fail(MGMB200_EINVAL, "ctx output pointer is NULL");
   *out = nullptr;
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      return fail(MGMB200_ECUDA, "no CUDA device available (%s); mgmb200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
   }
   if (device < 0) CU(cudaGetDevice(&device));
   if (device >= ndev) return fail(MGMB200_EINVAL, "device %d out of range (%d devices)", device, ndev);
   CU(cudaSetDevice(device));
   mgmb200_ctx *c = new mgmb200_ctx();
   c->device = device;
   cudaDeviceProp prop;
   CU(cudaGetDeviceProperties(&prop, device));
   c->num_sms = prop.multiProcessorCount;
   c->max_smem = (int)prop.sharedMemPerBlockOptin;
   if (prop.major < 9) {
      delete c;
      return fail(MGMB200_EUNSUPPORTED, "compute capability %d.%d: the kernels need TMA bulk copies (sm_100a build)",
                  prop.major, prop.minor);
   }
   CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
   c->stream = c->own_stream;
   tuning_from_env(c);
   *out = c;
   return 0;
}

extern "C" int mgmb200_set_option(mgmb200_ctx *c, const char *name, const char *value) {
   if (!c || !name) return fail(MGMB200_EINVAL, "NULL argument");
   if (!strcmp(name, "reset")) { tuning_from_env(c); return 0; }
   if (!apply_option(c, name, value)) return fail(MGMB200_EINVAL, "unknown option '%s' or bad value '%s'", name, value ? value : "(null)");
   return 0;
}

extern "C" void mgmb200_destroy(mgmb200_ctx *c) {
   if (!c) return;
   cudaSetDevice(c->device);
   cudaStreamSynchronize(c->stream);
   DevBuf *bufs[] = {&c->u, &c->v, &c->fu, &c->fv, &c->cu, &c->cv, &c->w, &c->cc, &c->dense, &c->out,
                     &c->outcost, &c->flags, &c->ftmp, &c->progress, &c->bnd, &c->bndm, &c->ncc, &c->rg[0], &c->rg[1],
                     &c->rg[2], &c->rg[3]};
   for (DevBuf *b : bufs) b->release();
   for (DevBuf &b : c->sweepv) b.release();
   for (DevBuf &b : c->bcc) b.release();
   for (DevBuf &b : c->bw) b.release();
   for (DevBuf &b : c->bout) b.release();
   c->desc.release(); c->fins.release(); c->staging.release();
   for (DevBuf &b : c->post) b.release();
   c->tiles.release();
   if (c->own_stream) cudaStreamDestroy(c->own_stream);
   delete c;
}

extern "C" int mgmb200_set_stream(mgmb200_ctx *c, void *s) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   c->stream = s ? (cudaStream_t)s : c->own_stream;
   return 0;
}
extern "C" int mgmb200_set_rows_per_band(mgmb200_ctx *c, int rows) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   c->rows_override = rows < 0 ? 0 : rows;
   return 0;
}
extern "C" int mgmb200_synchronize(mgmb200_ctx *c) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   CU(cudaSetDevice(c->device));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" int mgmb200_malloc(mgmb200_ctx *c, size_t bytes, void **p) {
   if (!c || !p) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   cudaError_t e = cudaMalloc(p, bytes);
   if (e != cudaSuccess) { cudaGetLastError(); return fail(MGMB200_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
   return 0;
}
extern "C" int mgmb200_free(mgmb200_ctx *c, void *p) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   CU(cudaSetDevice(c->device));
   CU(cudaFree(p));
   return 0;
}
extern "C" int mgmb200_memcpy_h2d(mgmb200_ctx *c, void *d, const void *h, size_t bytes) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   CU(cudaSetDevice(c->device));
   CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}
extern "C" int mgmb200_memcpy_d2h(mgmb200_ctx *c, void *h, const void *d, size_t bytes) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   CU(cudaSetDevice(c->device));
   CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

// ------------------------------------------------------------------------------------------ argument checks
static int check_dims(int nx, int ny, int dmin, int dmax) {
   if (nx < 1 || ny < 1) return fail(MGMB200_EINVAL, "image size %dx%d", nx, ny);
   if (dmax <= dmin) return fail(MGMB200_EINVAL, "empty disparity range [%d,%d] (the reference asserts min<max, dvec.cc:57)", dmin, dmax);
   long long L = (long long)dmax - dmin + 1;
   if (L > 4096) return fail(MGMB200_EUNSUPPORTED, "%lld labels: at most 4096 are supported", L);
   if (ny > 65535) return fail(MGMB200_EUNSUPPORTED, "%d image rows: at most 65535 are supported (one grid row per image row in the "
                               "O(W*H) kernels)", ny);
   return 0;
}

static int read_flags(mgmb200_ctx *c, int *h) {
   CU(cudaMemcpyAsync(h, c->flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}
static int clear_flags(mgmb200_ctx *c) {
   RET(c->flags.reserve(64));
   CU(cudaMemsetAsync(c->flags.p, 0, 64, c->stream));
   return 0;
}

// ------------------------------------------------------------------------------------------ device stages
extern "C" int mgmb200_weights_dev(mgmb200_ctx *c, const float *d_u, int nx, int ny, int nch, float aP,
                                   float aThresh, float *d_w) {
   if (!c || !d_u || !d_w) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d", nx, ny, nch);
   CU(cudaSetDevice(c->device));
   CU(weights_launch(d_u, nx, ny, nch, aP, aThresh, d_w, nullptr, c->stream));
   return 0;
}

static int costvolume_dev_impl(mgmb200_ctx *c, const float *d_u, const float *d_v, int nx, int ny, int nch, int vnx,
                               int vny, int dmin, int dmax, int pf, int dist, float truncDist, int win,
                               const float *d_rlo, const float *d_rhi, float *d_cc);
extern "C" int mgmb200_costvolume_dev(mgmb200_ctx *c, const float *d_u, const float *d_v, int nx, int ny, int nch,
                                      int vnx, int vny, int dmin, int dmax, int pf, int dist, float truncDist,
                                      int win, float *d_cc) {
   return costvolume_dev_impl(c, d_u, d_v, nx, ny, nch, vnx, vny, dmin, dmax, pf, dist, truncDist, win, nullptr, nullptr,
                              d_cc);
}
static int costvolume_dev_impl(mgmb200_ctx *c, const float *d_u, const float *d_v, int nx, int ny, int nch, int vnx,
                               int vny, int dmin, int dmax, int pf, int dist, float truncDist, int win,
                               const float *d_rlo, const float *d_rhi, float *d_cc) {
   if (!c || !d_u || !d_v || !d_cc) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   if (vnx < 1 || vny < 1 || nch < 1) return fail(MGMB200_EINVAL, "image v %dx%dx%d", vnx, vny, nch);
   if (dist < 0 || dist > 5 || pf < 0 || pf > 3) return fail(MGMB200_EINVAL, "distance/prefilter index");
   if (win < 1) return fail(MGMB200_EINVAL, "CENSUS_NCC_WIN=%d", win);
   CU(cudaSetDevice(c->device));
   const int L = dmax - dmin + 1, VS = mgmb200_padded_labels(L);
   const float *uu = d_u, *vv = d_v;
   const uint32_t *cu = nullptr, *cv = nullptr;
   int cnch = nch;
   if (dist == DIST_CENSUS || pf == PF_CENSUS) {
      cnch = census_nwords(nch, win);
      if (cnch < 1 || cnch > 32) return fail(MGMB200_EUNSUPPORTED, "census window %d with %d channels", win, nch);
      RET(c->cu.reserve((size_t)nx * ny * cnch * 4));
      RET(c->cv.reserve((size_t)vnx * vny * cnch * 4));
      CU(census_launch(d_u, nx, ny, nch, win, c->cu.as<uint32_t>(), c->stream));
      CU(census_launch(d_v, vnx, vny, nch, win, c->cv.as<uint32_t>(), c->stream));
      cu = c->cu.as<uint32_t>(); cv = c->cv.as<uint32_t>();
      if (dist != DIST_CENSUS) {
         // "-p census -t ad|sd|ncc|bt*": the reference picks the cost function before it forces census/census
         // (mgm_costvolume.h:355 vs :358-362), so the distance runs on the census bit strings held in float channels
         // (census_tools.cc:76-99) -- reproduced: the words are the same bits, read as floats, one channel per word
         uu = reinterpret_cast<const float *>(cu); vv = reinterpret_cast<const float *>(cv);
         cu = cv = nullptr;
      }
   } else if (pf == PF_SOBELX) {
      RET(c->fu.reserve((size_t)nx * ny * nch * 4));
      RET(c->fv.reserve((size_t)vnx * vny * nch * 4));
      CU(sobelx_launch(d_u, nx, ny, nch, c->fu.as<float>(), c->stream));
      CU(sobelx_launch(d_v, vnx, vny, nch, c->fv.as<float>(), c->stream));
      uu = c->fu.as<float>(); vv = c->fv.as<float>();
   } else if (pf == PF_GBLUR) {   // gblur_truncated(., 1.0), mgm_costvolume.h:380-384
      const size_t nmax = std::max((size_t)nx * ny, (size_t)vnx * vny) * nch * 4;
      RET(c->fu.reserve((size_t)nx * ny * nch * 4));
      RET(c->fv.reserve((size_t)vnx * vny * nch * 4));
      RET(c->ftmp.reserve(nmax));
      CU(gblur_launch(d_u, nx, ny, nch, 1.0f, c->ftmp.as<float>(), c->fu.as<float>(), c->stream));
      CU(gblur_launch(d_v, vnx, vny, nch, 1.0f, c->ftmp.as<float>(), c->fv.as<float>(), c->stream));
      uu = c->fu.as<float>(); vv = c->fv.as<float>();
   }
   float *scratch = nullptr;
   if (dist == DIST_NCC) {
      RET(c->ncc.reserve(costvolume_ncc_scratch_floats(nx, ny, vnx, vny, cnch) * sizeof(float)));
      scratch = c->ncc.as<float>();
   }
   CU(costvolume_launch(dist, uu, vv, cu, cv, nx, ny, vnx, vny, cnch, win, dmin, L, VS, truncDist, d_rlo, d_rhi, d_cc,
                        c->num_sms, c->stream, scratch));
   return 0;
}

extern "C" int mgmb200_pad_volume_dev(mgmb200_ctx *c, const float *d_dense, float *d_padded, int nx, int ny, int L,
                                      int label_major) {
   if (!c || !d_dense || !d_padded) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   CU(pad_volume_launch(d_dense, d_padded, (long long)nx * ny, L, mgmb200_padded_labels(L), label_major, c->stream));
   return 0;
}
extern "C" int mgmb200_unpad_volume_dev(mgmb200_ctx *c, const float *d_padded, float *d_dense, int nx, int ny,
                                        int L) {
   if (!c || !d_dense || !d_padded) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   CU(unpad_volume_launch(d_padded, d_dense, (long long)nx * ny, L, mgmb200_padded_labels(L), c->stream));
   return 0;
}

// One aggregation launch: the sweeps in `mask` of `npairs` stereo pairs of the same shape and parameters.
struct SweepRun {
   int npairs = 1;
   const float *const *cc = nullptr;     // [npairs] padded cost volumes
   const float *const *w = nullptr;      // [npairs] weight planes, or nullptr
   WtaParams *fin = nullptr;             // [npairs] finish parameters (sweep pointers are filled in here), or nullptr:
                                         // with it, all sweeps are requested and the finish stage (ordered sum, fix,
                                         // WTA, sub-pixel) may run inside the same launch on tiles whose bands are
                                         // complete (c->fin_done tells whether it did)
   unsigned mask = 0;
   int nslabs = 1, slab_rows = 0;        // sweep-sharded multi-GPU layout (npairs == 1): image rows
   float *const *slab_ptrs = nullptr;    // [r*slab_rows, (r+1)*slab_rows) of sweep p are stored to slab_ptrs[p*nslabs+r]
};

static int sweep_slot(int pair, int p) { return pair * MGM_MAX_NDIR + p; }

static int run_sweeps_multi(mgmb200_ctx *c, const SweepRun &R, int weights_mode, int nx, int ny, int L, float P1,
                            float P2, int NDIR, int K, int felz) {
   c->fin_done = false;
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR)
      return fail(MGMB200_EUNSUPPORTED, "NDIR=%d: 1..8 sweeps of the reference (mgm_core.cc:463-471) or up to 16 with the "
                  "sweeps 8-15 defined by this library (include/mgmb200.h)", NDIR);
   if (K < 1 || K > 4) return fail(MGMB200_EINVAL, "MGM/TSGM=%d not in 1..4", K);
   if (!(P1 >= 0.f) || !(P2 >= 0.f)) return fail(MGMB200_EUNSUPPORTED, "P1=%g P2=%g must be >= 0", P1, P2);
   if (!(P1 < INFINITY)) return fail(MGMB200_EUNSUPPORTED, "P1 must be finite");
   const int pot = felz > 0 ? POT_TRUNC : POT_SGM;
   if (pot == POT_SGM && !(P2 < INFINITY)) return fail(MGMB200_EUNSUPPORTED, "P2 must be finite with SGM potentials");
   if (R.npairs < 1) return 0;
   if (R.nslabs > 1 && (R.npairs != 1 || R.nslabs > MGM_MAX_SLABS || R.slab_rows < 2 || !R.slab_ptrs || R.fin ||
                        (long long)R.slab_rows * R.nslabs < ny || ny > 65535))
      return fail(MGMB200_EINVAL, "row slabs: one pair, at most %d slabs that cover the image, no fused finish", MGM_MAX_SLABS);
   const int VS = mgmb200_padded_labels(L);
   const size_t vol = (size_t)nx * ny * VS * sizeof(float);
   const unsigned mask = R.mask & ((NDIR >= 32) ? ~0u : ((1u << NDIR) - 1u));
   c->n_launches = 0;

   bool weighted = false;
   if (weights_mode == 1) weighted = true;
   if (weights_mode == 2 && R.w && R.w[0]) {
      if (R.npairs != 1) return fail(MGMB200_EINVAL, "weights_mode 2 (scan) is a single-pair mode");
      RET(clear_flags(c));
      CU(scan_weights_launch(R.w[0], (long long)nx * ny * 8, c->flags.as<int>(), c->stream));
      c->n_launches++;
      int fl = 0;
      RET(read_flags(c, &fl));
      weighted = fl & 1;
      if (fl & 2) return fail(MGMB200_EUNSUPPORTED, "edge weights must be finite and >= 0");
   }
   if (weighted && !(R.w && R.w[0])) return fail(MGMB200_EINVAL, "weighted aggregation without weights");
   // windowed truncated-linear updates need the consumer-side (per-edge) kernels, with or without weights
   const bool window = c->r_window && pot == POT_TRUNC;
   bool use_w = weighted;
   if (window) { weighted = true; }

   if ((int)c->sweepv.size() < R.npairs * MGM_MAX_NDIR) c->sweepv.resize((size_t)R.npairs * MGM_MAX_NDIR);
   if (R.nslabs == 1)
      for (int b = 0; b < R.npairs; b++)
         for (int p = 0; p < NDIR; p++)
            if (mask & (1u << p)) RET(c->sweepv[sweep_slot(b, p)].reserve(vol));
   c->sweep_bytes = vol;

   int nsw = 0;
   for (int p = 0; p < NDIR; p++) nsw += (mask >> p) & 1;
   AggPlan plan;
   const bool knight = (mask >> 8) != 0;
   // The finish stage runs as tile work inside the launch where that is a gain: one pair (its parameters are kernel
   // arguments) and label vectors of at least 128 floats.  Measured otherwise (profiles/r02_experiments.md): 32 x
   // 1242x375x192 TSGM=4 in launches of 8 pairs 95.7 ms fused vs 74.5 ms with separate finish launches, 4096x4096x64
   // -O 16 113.2 vs 97.4 ms; 2048x1536x256 20.0 vs 22.2 ms, 1920x1080x128 8.4 vs 9.4.
   // ... and, with SGM potentials, frames of a megapixel or more: their band steps are short, a small frame leaves the tiles
   // to the end of the launch where one CTA of 16 warps per SM is slower than the stand-alone kernel (measured 640x480x100
   // TSGM=2: 1.80 ms fused vs 1.46 ms, 1242x375x192: 5.4 vs 4.1 ms; 1920x1080x128: 5.68 vs 5.97 ms)
   const bool fuse_gain = c->tune.fused_finish > 0 || (c->tune.fused_finish < 0 && R.npairs == 1 && VS >= 128 &&
                                                       (pot == POT_TRUNC || (long long)nx * ny >= 1000000LL));
   const bool want_fuse = R.fin && !c->tune.no_fused_finish && fuse_gain;
   AggTuning tune = c->tune;
   if (want_fuse) tune.full_block = 1;   // bands with fewer rows than the block holds: the spare warps serve the finish tiles
   agg_plan(&plan, nx, ny, L, K, pot, weighted, c->max_smem, c->num_sms, c->rows_override, knight, tune);
   if (!c->rows_override && !c->tune.rows_axis && !c->tune.rows_diag && plan.T[0] > 40 && R.npairs == 1 && nsw <= 2) {
      // One or two sweeps on this GPU (sweep-sharded layouts): with fewer bands than SMs the launch is bound by the
      // dependency depth alone, and bands of 40 workers step faster than bands of 56 (measured: one axis sweep of
      // the headline shape 15.7 -> 14.2 ms, an axis + a diagonal sweep 16.2 -> 15.3 ms; small images with all
      // eight sweeps do not gain)
      long bands = 0;
      for (int p = 0; p < NDIR; p++) {
         if (!(mask & (1u << p))) continue;
         int nb = 0; size_t a = 0, bq = 0;
         agg_sweep_bands(plan, p, nx, ny, &nb, &a, &bq);
         bands += nb;
      }
      if (bands < c->num_sms) agg_plan(&plan, nx, ny, L, K, pot, weighted, c->max_smem, c->num_sms, 40, knight, tune);
   } else if (!c->rows_override && !c->tune.rows_axis && !c->tune.rows_diag && R.npairs == 1) {
      // One small pair (a KITTI-size frame yields fewer bands than 1.5 x the SMs): the launch is bound by the dependency
      // depth of its sweeps and the step of a band shortens with its rows: 40, then 28 rows per band (measured 1242x375x192
      // TSGM=4 SGM: 7.83 / 6.62 / 6.08 ms with 56 / 40 / 28 rows; truncated linear TSGM=3: 5.97 / 5.30 / 5.01 ms, 640x480x64:
      // 1.68 / 1.58 / 1.53 ms)
      for (int t : {40, 28}) {
         long bands = 0;
         for (int p = 0; p < NDIR; p++) {
            if (!(mask & (1u << p))) continue;
            int nb = 0; size_t a = 0, bq = 0;
            agg_sweep_bands(plan, p, nx, ny, &nb, &a, &bq);
            bands += nb;
         }
         if (2 * bands >= 3L * c->num_sms || plan.T[0] <= t) break;
         agg_plan(&plan, nx, ny, L, K, pot, weighted, c->max_smem, c->num_sms, t, knight, tune);
      }
   }
   // SGM potentials, 128 padded labels, 8 lanes per worker: a step is short enough that bands of 44 rows beat bands of 56
   // (measured 1920x1080x128 TSGM=2, full block kept for the finish tiles: 5.55 / 5.39 / 5.11 / 5.05 / 5.24 / 5.47 ms with
   // 56 / 52 / 48 / 44 / 40 / 36 rows; 256 labels and 64 labels with 4 lanes keep the largest bands; the two directions of a
   // left-right run in one launch do not gain: 11.6 vs 11.5 ms)
   if (!c->rows_override && !c->tune.rows_axis && !c->tune.rows_diag && pot == POT_SGM && !weighted && VS == 128 &&
       plan.lanes == 8 && plan.T[0] > 44 && nsw > 2 && R.npairs == 1)
      agg_plan(&plan, nx, ny, L, K, pot, weighted, c->max_smem, c->num_sms, 44, knight, tune);
   if (plan.T[0] < 1 || plan.T[1] < 1 || plan.T[2] < 1)
      return fail(MGMB200_EUNSUPPORTED, "%d labels do not fit the shared-memory wavefront (max_smem=%d)", L, c->max_smem);
   c->rows_axis = plan.T[0]; c->rows_diag = plan.T[1]; c->block = plan.block; c->smem = plan.smem;

   // the sweep table; the kernel claims bands dynamically, in order within a sweep (aggregate.cu claim_band)
   const int nsweeps = R.npairs * NDIR;
   std::vector<SweepDesc> desc((size_t)nsweeps);
   std::vector<size_t> bnd_off((size_t)nsweeps), bndm_off((size_t)nsweeps), prog_off((size_t)nsweeps);
   size_t bnd_total = 0, bndm_total = 0, prog_total = 0;
   int nbands = 0;
   memset(desc.data(), 0, sizeof(SweepDesc) * desc.size());
   for (int b = 0; b < R.npairs; b++)
      for (int p = 0; p < NDIR; p++) {
         SweepDesc &d = desc[(size_t)b * NDIR + p];
         d.pass = p; d.pair = b;
         d.cls = agg_sweep_class(plan, p);
         d.filler = (d.cls == CLS_DIAG && plan.shear) ? 1 : 0;
         if (!(mask & (1u << p))) continue;
         size_t bf = 0, bmf = 0;
         agg_sweep_bands(plan, p, nx, ny, &d.nb, &bf, &bmf);
         const size_t v = (size_t)b * NDIR + p;
         bnd_off[v] = bnd_total; bnd_total += bf;
         bndm_off[v] = bndm_total; bndm_total += bmf;
         prog_off[v] = prog_total; prog_total += d.nb;
         nbands += d.nb;
      }
   if (nbands == 0) return 0;

   // layout of the counters: [prog_total] progress | [nsweeps] claim counters | [4] finish-tile claim counter | [prog_total] done flags
   const size_t ncount = 2 * prog_total + nsweeps + 4;
   RET(c->progress.reserve(ncount * sizeof(int)));
   RET(c->bnd.reserve(bnd_total * sizeof(float)));
   RET(c->bndm.reserve(bndm_total * sizeof(float)));
   CU(cudaMemsetAsync(c->progress.p, 0, ncount * sizeof(int), c->stream));
   int *d_prog = c->progress.as<int>(), *d_next = d_prog + prog_total, *d_fin_next = d_next + nsweeps,
       *d_done = d_fin_next + 4;

   for (int b = 0; b < R.npairs; b++)
      for (int p = 0; p < NDIR; p++) {
         const size_t v = (size_t)b * NDIR + p;
         SweepDesc &d = desc[v];
         d.cc = R.cc[b];
         d.w = (use_w && R.w) ? R.w[b] : nullptr;
         if (window) { d.win_lo = c->r_ccmin; d.win_hi = c->r_ccmax; }
         if (R.nslabs > 1) { for (int r = 0; r < R.nslabs; r++) d.ldir[r] = R.slab_ptrs[(size_t)p * R.nslabs + r]; }
         else d.ldir[0] = c->sweepv[sweep_slot(b, p)].as<float>();
         d.bnd = c->bnd.as<float>() + bnd_off[v];
         d.bndm = c->bndm.as<float>() + bndm_off[v];
         d.progress = d_prog + prog_off[v];
         d.band_done = d_done + prog_off[v];
      }
   RET(c->desc.reserve(sizeof(SweepDesc) * desc.size()));
   CU(c->staging.push(c->desc.p, desc.data(), sizeof(SweepDesc) * desc.size(), c->stream));

   // fused finish: tiles of pixels in the order they are expected to become complete (the row-per-worker sweeps
   // decide: band b of such a sweep is done after about maxii + (b+1)*T steps), pairs interleaved
   const int tw = c->tune.fin_tw, th = c->tune.fin_th;
   const size_t rows_region = plan.smem - plan.off_thr;
   const bool fuse = want_fuse &&
                     (size_t)(plan.block / 32) * (32 / wta_lanes_per_pixel(VS)) * VS * 4 <= rows_region;
   const int tiles_x = (nx + tw - 1) / tw, tiles_y = (ny + th - 1) / th, ntiles = tiles_x * tiles_y;
   if (R.fin)
      for (int b = 0; b < R.npairs; b++)
         for (int p = 0; p < MGM_MAX_NDIR; p++) R.fin[b].ldir[p] = p < NDIR ? c->sweepv[sweep_slot(b, p)].as<float>() : nullptr;
   if (fuse) {
      const int key[8] = {nx, ny, plan.T[0], plan.T[2], tw, th, NDIR, R.npairs};
      if (memcmp(key, c->tile_key, sizeof(key)) != 0 || (int)c->tile_order.size() != ntiles * R.npairs) {
         std::vector<std::pair<long long, int>> when((size_t)ntiles);
         for (int t = 0; t < ntiles; t++) {
            const int x0 = (t % tiles_x) * tw, y0 = (t / tiles_x) * th;
            const int x1 = std::min(x0 + tw, nx) - 1, y1 = std::min(y0 + th, ny) - 1;
            long long ready = 0;
            for (int p = 0; p < NDIR; p++) {
               const int cls = agg_sweep_class(plan, p);
               if (cls == CLS_DIAG && plan.shear) continue;
               const PassGeom g = pass_geometry(p, nx, ny);
               const int q = p & 7;
               const int rm = (0x53 >> q) & 1, incx = (0xC5 >> q) & 1, incy = (0x99 >> q) & 1;
               const int ax1 = incx ? x1 : nx - 1 - x0, ay1 = incy ? y1 : ny - 1 - y0;
               const int ys1 = rm ? ay1 : ax1;
               const int T = plan.T[cls], sig = (cls != CLS_AXIS || K == 4) ? 2 : 1;
               ready = std::max(ready, (long long)g.maxii + (long long)(ys1 / T + 1) * T * sig);
            }
            when[t] = std::make_pair(ready, t);
         }
         std::stable_sort(when.begin(), when.end());
         c->tile_order.resize((size_t)ntiles * R.npairs);
         for (int t = 0; t < ntiles; t++)
            for (int b = 0; b < R.npairs; b++) c->tile_order[(size_t)t * R.npairs + b] = b * ntiles + when[t].second;
         memcpy(c->tile_key, key, sizeof(key));
         RET(c->tiles.reserve(c->tile_order.size() * sizeof(int)));
         CU(c->staging.push(c->tiles.p, c->tile_order.data(), c->tile_order.size() * sizeof(int), c->stream));
      }
      RET(c->fins.reserve(sizeof(WtaParams) * (size_t)R.npairs));
      CU(c->staging.push(c->fins.p, R.fin, sizeof(WtaParams) * (size_t)R.npairs, c->stream));
   }

   AggParams P;
   memset(&P, 0, sizeof(P));
   P.sweeps = c->desc.as<SweepDesc>(); P.nsweeps = nsweeps;
   P.next_band = d_next;
   P.nbands = nbands;
   P.static_order = c->tune.static_order;
   P.win_emin = c->r_emin;
   P.nx = nx; P.ny = ny; P.L = L; P.VS = VS; P.ndir = NDIR;
   P.nslabs = R.nslabs; P.slab_rows = R.slab_rows;
   P.slab_magic = R.nslabs > 1 ? (unsigned)((0x100000000ull + (unsigned)R.slab_rows - 1) / (unsigned)R.slab_rows) : 0u;
   for (int i = 0; i < 3; i++) { P.T[i] = plan.T[i]; P.TS[i] = plan.TS[i]; P.ng[i] = plan.ng[i]; }
   P.ncb = plan.ncb; P.shear = plan.shear; P.fused_sgm = plan.fused_sgm; P.regchain = plan.regchain;
   // measured (profiles/r02_experiments.md): 1920x1080x128 TSGM=2 8.37 -> 6.96 ms, 4096x4096x64 25.8 -> 22.8 ms
   P.cc_pf = c->tune.cc_pf >= 0 ? c->tune.cc_pf : (pot == POT_SGM ? 3 : 0);
   P.P1 = P1; P.P2 = P2;
   if (fuse) {
      P.fin_enabled = 1;
      P.fin_ntiles = ntiles; P.fin_total = ntiles * R.npairs; P.fin_tw = tw; P.fin_th = th; P.fin_tiles_x = tiles_x;
      P.fin_order = c->tiles.as<int>();
      P.fin_next = d_fin_next;
      P.fins = c->fins.as<WtaParams>(); P.npairs = R.npairs;
      P.fin0 = R.fin[0];
   }
   P.off_phase = (unsigned)plan.off_phase; P.off_cbar = (unsigned)plan.off_cbar; P.off_vbar = (unsigned)plan.off_vbar;
   P.off_ms = (unsigned)plan.off_ms; P.off_vms = (unsigned)plan.off_vms; P.off_virt = (unsigned)plan.off_virt;
   P.off_thr = (unsigned)plan.off_thr;
   if (c->tune.dbg) {
      RET(c->dbgbuf.reserve(8 * (8 + 600)));
      CU(cudaMemsetAsync(c->dbgbuf.p, 0, 8 * (8 + 600), c->stream));
      P.dbg = c->dbgbuf.as<unsigned long long>();
   }
   CU(agg_launch(P, plan, pot, K, weighted, c->stream));
   c->n_launches++;
   c->fin_done = fuse;
   if (c->tune.dbg) {   // profiling aid: synchronises
      static unsigned long long h[8 + 600];
      CU(cudaMemcpyAsync(h, c->dbgbuf.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      if (c->tune.dbg > 1)
         for (int b = 0; b < 60 && h[8 + b * 10 + 6]; b++) {
            const unsigned long long *q = h + 8 + b * 10;
            fprintf(stderr, "[mgmb200 band %2d of sweep 0] steps=%llu cycles/step: top=%.0f gather=%.0f sync1=%.0f transform=%.0f rest=%.0f "
                    "barrier=%.0f end=%.3f ms\n", b, q[6], (double)q[0] / q[6], (double)q[1] / q[6], (double)q[2] / q[6],
                    (double)q[3] / q[6], (double)q[4] / q[6], (double)q[5] / q[6], (double)(q[7] - h[8 + 7]) * 1e-6);
         }
      if (h[6])
         fprintf(stderr, "[mgmb200 phase timing] steps=%llu cycles/step: top=%.0f gather=%.0f sync1=%.0f transform=%.0f rest=%.0f barrier=%.0f x7=%.0f\n",
                 h[6], (double)h[0] / h[6], (double)h[1] / h[6], (double)h[2] / h[6], (double)h[3] / h[6], (double)h[4] / h[6], (double)h[5] / h[6], (double)h[7] / h[6]);
   }
   return 0;
}

// single pair (the reference's mgm() shape)
static int run_sweeps(mgmb200_ctx *c, const float *d_cc, const float *d_w, int weights_mode, int nx, int ny,
                      int L, float P1, float P2, int NDIR, int K, int felz, unsigned mask, WtaParams *fin = nullptr) {
   SweepRun R;
   const float *ccs[1] = {d_cc}, *ws[1] = {d_w};
   R.cc = ccs; R.w = ws; R.fin = fin; R.mask = mask;
   return run_sweeps_multi(c, R, weights_mode, nx, ny, L, P1, P2, NDIR, K, felz);
}

static WtaParams finish_params(mgmb200_ctx *c, const float *const *d_sweeps, const float *d_cc, int nx, int dmin, int L,
                               int NDIR, int fix, int refine, int row_begin, int row_end, float *d_out, float *d_outcost,
                               float *d_S) {
   WtaParams W;
   memset(&W, 0, sizeof(W));
   for (int p = 0; p < NDIR; p++) W.ldir[p] = d_sweeps[p];
   W.cc = d_cc; W.out = d_out; W.outcost = d_outcost; W.S_out = d_S;
   W.pix_begin = (long long)row_begin * nx; W.pix_end = (long long)row_end * nx;
   W.ndir = NDIR; W.L = L; W.VS = mgmb200_padded_labels(L); W.dmin = dmin;
   W.fix = (fix == 1); W.refine = refine;
   W.smin = c->r_smin; W.smax = c->r_smax; W.ccmin = c->r_ccmin; W.ccmax = c->r_ccmax;
   return W;
}

static int finish_rows(mgmb200_ctx *c, const float *const *d_sweeps, const float *d_cc, int nx, int ny, int dmin,
                       int L, int NDIR, int fix, int refine, int row_begin, int row_end, float *d_out,
                       float *d_outcost, float *d_S) {
   const WtaParams W = finish_params(c, d_sweeps, d_cc, nx, dmin, L, NDIR, fix, refine, row_begin, row_end, d_out, d_outcost, d_S);
   CU(wta_launch(W, c->num_sms, c->stream));
   c->n_launches++;
   return 0;
}

extern "C" int mgmb200_aggregate_sweeps_dev(mgmb200_ctx *c, const float *d_cc, const float *d_w, int weights_mode,
                                            int nx, int ny, int dmin, int dmax, float P1, float P2, int NDIR,
                                            int K, int felz, unsigned sweep_mask) {
   if (!c || !d_cc) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   CU(cudaSetDevice(c->device));
   return run_sweeps(c, d_cc, d_w, weights_mode, nx, ny, dmax - dmin + 1, P1, P2, NDIR, K, felz, sweep_mask);
}

extern "C" int mgmb200_sweep_volume(mgmb200_ctx *c, int sweep, float **d_ptr, size_t *bytes) {
   if (!c || sweep < 0 || sweep >= MGM_MAX_NDIR || !d_ptr) return fail(MGMB200_EINVAL, "bad argument");
   if ((int)c->sweepv.size() < MGM_MAX_NDIR) c->sweepv.resize(MGM_MAX_NDIR);
   *d_ptr = c->sweepv[sweep].as<float>();
   if (bytes) *bytes = c->sweepv[sweep].p ? c->sweepv[sweep].cap : 0;
   return 0;
}

// Sweep-sharded multi-GPU layout.  Every rank allocates all NDIR message volumes and exports them
// (mgmb200_ipc_export marks them: they are never reallocated while exported); the rank that aggregates sweep p
// stores the messages of image rows [r*slab_rows, (r+1)*slab_rows) straight into rank r's volume p -- peer stores
// over NVLink issued by the aggregation kernel itself, overlapped with the sweep -- so that after one barrier every
// rank finds all sweeps of its own slab in LOCAL memory and finishes it in sweep order (mgmb200_finish_rows_dev).
extern "C" int mgmb200_sweeps_alloc(mgmb200_ctx *c, int nx, int ny, int dmin, int dmax, int NDIR) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   RET(check_dims(nx, ny, dmin, dmax));
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return fail(MGMB200_EINVAL, "NDIR=%d", NDIR);
   CU(cudaSetDevice(c->device));
   if ((int)c->sweepv.size() < MGM_MAX_NDIR) c->sweepv.resize(MGM_MAX_NDIR);
   const size_t vol = mgmb200_volume_bytes(nx, ny, dmax - dmin + 1);
   for (int p = 0; p < NDIR; p++) RET(c->sweepv[p].reserve(vol));
   c->sweep_bytes = vol;
   return 0;
}
extern "C" int mgmb200_sweeps_release(mgmb200_ctx *c) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   CU(cudaSetDevice(c->device));
   CU(cudaStreamSynchronize(c->stream));
   for (DevBuf &b : c->sweepv) { b.exported = false; b.release(); }
   return 0;
}
extern "C" int mgmb200_aggregate_sweeps_slabs_dev(mgmb200_ctx *c, const float *d_cc, const float *d_w, int weights_mode,
                                                  int nx, int ny, int dmin, int dmax, float P1, float P2, int NDIR, int K,
                                                  int felz, unsigned sweep_mask, int nslabs, int slab_rows,
                                                  float *const *d_slab_volumes) {
   if (!c || !d_cc || !d_slab_volumes) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   CU(cudaSetDevice(c->device));
   SweepRun R;
   const float *ccs[1] = {d_cc}, *ws[1] = {d_w};
   R.cc = ccs; R.w = ws; R.mask = sweep_mask;
   R.nslabs = nslabs; R.slab_rows = slab_rows; R.slab_ptrs = d_slab_volumes;
   if (nslabs == 1) {   // degenerate: one slab = the plain layout with caller-provided volumes is not needed; use own volumes
      R.slab_ptrs = nullptr; R.slab_rows = 0;
   }
   return run_sweeps_multi(c, R, weights_mode, nx, ny, dmax - dmin + 1, P1, P2, NDIR, K, felz);
}

// The exchange north_star names: every rank adds ITS sweeps (in increasing sweep order) into one partial
// volume, the partial volumes are summed by an NCCL all-reduce (the caller's collective), and mgmb200_finish_sum_dev
// applies the over-count fix for all NDIR sweeps, WTA and refinement.  The floating-point summation order differs
// from the reference's (mgm_core.cc:582-587), so the bits may: bench.py counts the differences.
extern "C" int mgmb200_sum_sweeps_dev(mgmb200_ctx *c, int nx, int ny, int dmin, int dmax, unsigned sweep_mask,
                                      float *d_sum) {
   if (!c || !d_sum) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   CU(cudaSetDevice(c->device));
   const float *src[MGM_MAX_NDIR];
   int n = 0;
   for (int p = 0; p < MGM_MAX_NDIR; p++)
      if (sweep_mask & (1u << p)) {
         if (p >= (int)c->sweepv.size() || !c->sweepv[p].p) return fail(MGMB200_EINVAL, "sweep %d has not been aggregated", p);
         src[n++] = c->sweepv[p].as<float>();
      }
   CU(sum_volumes_launch(src, n, d_sum, (long long)nx * ny * mgmb200_padded_labels(dmax - dmin + 1), c->num_sms, c->stream));
   c->n_launches++;
   return 0;
}
extern "C" int mgmb200_finish_sum_dev(mgmb200_ctx *c, const float *d_sum, const float *d_cc, int nx, int ny, int dmin,
                                      int dmax, int NDIR, int fix, int refine, int row_begin, int row_end, float *d_out,
                                      float *d_outcost) {
   if (!c || !d_sum || !d_cc || !d_out || !d_outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return fail(MGMB200_EUNSUPPORTED, "NDIR=%d", NDIR);
   if (row_begin < 0 || row_end > ny || row_begin > row_end) return fail(MGMB200_EINVAL, "rows [%d,%d)", row_begin, row_end);
   if (refine < 0 || refine > 4) return fail(MGMB200_EINVAL, "refinement index %d", refine);
   CU(cudaSetDevice(c->device));
   if (row_begin == row_end) return 0;
   const float *one[1] = {d_sum};
   WtaParams W = finish_params(c, one, d_cc, nx, dmin, dmax - dmin + 1, 1, fix, refine, row_begin, row_end, d_out, d_outcost, nullptr);
   W.fix_count = NDIR;   // S - (NDIR-1)*C although a single (pre-summed) volume is read
   CU(wta_launch(W, c->num_sms, c->stream));
   c->n_launches++;
   return 0;
}

extern "C" int mgmb200_finish_rows_dev(mgmb200_ctx *c, const float *const *d_sweeps, const float *d_cc, int nx,
                                       int ny, int dmin, int dmax, int NDIR, int fix, int refine, int row_begin,
                                       int row_end, float *d_out, float *d_outcost) {
   if (!c || !d_sweeps || !d_cc || !d_out || !d_outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return fail(MGMB200_EUNSUPPORTED, "NDIR=%d", NDIR);
   if (row_begin < 0 || row_end > ny || row_begin > row_end) return fail(MGMB200_EINVAL, "rows [%d,%d)", row_begin, row_end);
   if (refine < 0 || refine > 4) return fail(MGMB200_EINVAL, "refinement index %d", refine);
   CU(cudaSetDevice(c->device));
   if (row_begin == row_end) return 0;
   return finish_rows(c, d_sweeps, d_cc, nx, ny, dmin, dmax - dmin + 1, NDIR, fix, refine, row_begin, row_end,
                      d_out, d_outcost, nullptr);
}

extern "C" int mgmb200_aggregate_dev(mgmb200_ctx *c, const float *d_cc, const float *d_w, int weights_mode, int nx,
                                     int ny, int dmin, int dmax, float P1, float P2, int NDIR, int K, int felz,
                                     int fix, int refine, float *d_out, float *d_outcost, float *d_S) {
   if (!c || !d_cc || !d_out || !d_outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   if (refine < 0 || refine > 4) return fail(MGMB200_EINVAL, "refinement index %d", refine);
   CU(cudaSetDevice(c->device));
   const int L = dmax - dmin + 1;
   const unsigned mask = (NDIR >= 1 && NDIR <= MGM_MAX_NDIR) ? ((NDIR == 32 ? 0u : (1u << NDIR)) - 1u) : 0u;
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return run_sweeps(c, d_cc, d_w, weights_mode, nx, ny, L, P1, P2, NDIR, K, felz, mask);   // fails with the message
   const float *none[MGM_MAX_NDIR] = {nullptr};
   WtaParams W = finish_params(c, none, d_cc, nx, dmin, L, NDIR, fix, refine, 0, ny, d_out, d_outcost, d_S);
   RET(run_sweeps(c, d_cc, d_w, weights_mode, nx, ny, L, P1, P2, NDIR, K, felz, mask, &W));   // fills W.ldir
   if (c->fin_done) return 0;   // finished tile by tile inside the aggregation launch
   CU(wta_launch(W, c->num_sms, c->stream));
   c->n_launches++;
   return 0;
}

// A batch of stereo pairs of one shape (BASELINE.json configs[3]: 32 KITTI-shape pairs): the sweeps of up to
// `batch` (mgmb200_set_option, default 8) pairs share one launch, so that the machine stays full although one small
// pair alone is bound by the dependency depth of its sweeps.  Results are those of npairs mgmb200_aggregate_dev calls.
// npairs pairs of one shape, each with its own label origin dmins[b] (the two directions of a left-right run have
// mirrored ranges): several pairs per aggregation launch
static int aggregate_pairs(mgmb200_ctx *c, int npairs, const float *const *d_cc, const float *const *d_w, int nx, int ny,
                           const int *dmins, int L, float P1, float P2, int NDIR, int K, int felz, int fix, int refine,
                           float *const *d_out, float *const *d_outcost) {
   const unsigned mask = (1u << NDIR) - 1u;
   const int chunk = std::max(1, c->tune.batch);
   int launches = 0;
   for (int b0 = 0; b0 < npairs; b0 += chunk) {
      const int nb = std::min(chunk, npairs - b0);
      std::vector<WtaParams> W((size_t)nb);
      const float *none[MGM_MAX_NDIR] = {nullptr};
      for (int b = 0; b < nb; b++) {
         if (!d_cc[b0 + b] || !d_out[b0 + b] || !d_outcost[b0 + b]) return fail(MGMB200_EINVAL, "NULL pointer for pair %d", b0 + b);
         W[b] = finish_params(c, none, d_cc[b0 + b], nx, dmins[b0 + b], L, NDIR, fix, refine, 0, ny, d_out[b0 + b], d_outcost[b0 + b], nullptr);
      }
      SweepRun R;
      R.npairs = nb; R.cc = d_cc + b0; R.w = d_w ? d_w + b0 : nullptr; R.fin = W.data(); R.mask = mask;
      RET(run_sweeps_multi(c, R, d_w ? 1 : 0, nx, ny, L, P1, P2, NDIR, K, felz));
      launches += c->n_launches;
      if (!c->fin_done)
         for (int b = 0; b < nb; b++) { CU(wta_launch(W[b], c->num_sms, c->stream)); launches++; }
   }
   c->n_launches = launches;
   return 0;
}

extern "C" int mgmb200_aggregate_batch_dev(mgmb200_ctx *c, int npairs, const float *const *d_cc, const float *const *d_w,
                                           int nx, int ny, int dmin, int dmax, float P1, float P2, int NDIR, int K,
                                           int felz, int fix, int refine, float *const *d_out, float *const *d_outcost) {
   if (!c || !d_cc || !d_out || !d_outcost || npairs < 0) return fail(MGMB200_EINVAL, "bad argument");
   RET(check_dims(nx, ny, dmin, dmax));
   if (refine < 0 || refine > 4) return fail(MGMB200_EINVAL, "refinement index %d", refine);
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return fail(MGMB200_EUNSUPPORTED, "NDIR=%d", NDIR);
   if (c->r_smin || c->r_ccmin) return fail(MGMB200_EINVAL, "per-pixel ranges are a single-pair mode");
   CU(cudaSetDevice(c->device));
   std::vector<int> dmins((size_t)std::max(npairs, 1), dmin);
   return aggregate_pairs(c, npairs, d_cc, d_w, nx, ny, dmins.data(), dmax - dmin + 1, P1, P2, NDIR, K, felz, fix, refine, d_out,
                          d_outcost);
}

// ------------------------------------------------------------------------------------------ IPC (multi-GPU)
extern "C" int mgmb200_ipc_export(mgmb200_ctx *c, const void *d_ptr, unsigned char handle_out[64]) {
   if (!c || !d_ptr || !handle_out) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   cudaIpcMemHandle_t h;
   CU(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
   for (DevBuf &b : c->sweepv) if (b.p == d_ptr) b.exported = true;
   static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
   memcpy(handle_out, &h, 64);
   return 0;
}
extern "C" int mgmb200_ipc_open(mgmb200_ctx *c, const unsigned char handle[64], void **d_ptr) {
   if (!c || !handle || !d_ptr) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   cudaIpcMemHandle_t h;
   memcpy(&h, handle, 64);
   CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
   return 0;
}
extern "C" int mgmb200_ipc_close(mgmb200_ctx *c, void *d_ptr) {
   if (!c || !d_ptr) return fail(MGMB200_EINVAL, "NULL argument");
   CU(cudaSetDevice(c->device));
   CU(cudaIpcCloseMemHandle(d_ptr));
   return 0;
}

extern "C" int mgmb200_last_launch_info(mgmb200_ctx *c, int *launches, int *rows_axis, int *rows_diag, int *threads,
                                        size_t *smem) {
   if (!c) return fail(MGMB200_EINVAL, "ctx is NULL");
   if (launches) *launches = c->n_launches;
   if (rows_axis) *rows_axis = c->rows_axis;
   if (rows_diag) *rows_diag = c->rows_diag;
   if (threads) *threads = c->block;
   if (smem) *smem = c->smem;
   return 0;
}

// ------------------------------------------------------------------------------------------ host-pointer API
static int upload(mgmb200_ctx *c, DevBuf &b, const void *h, size_t bytes) {
   RET(b.reserve(bytes));
   CU(cudaMemcpyAsync(b.p, h, bytes, cudaMemcpyHostToDevice, c->stream));
   return 0;
}
static int download(mgmb200_ctx *c, void *h, const void *d, size_t bytes) {
   CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
   return 0;
}

extern "C" int mgmb200_compute_mgm_weights(mgmb200_ctx *c, const float *u, int nx, int ny, int nch, float aP,
                                           float aThresh, float *w_out) {
   if (!c || !u || !w_out) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d", nx, ny, nch);
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny;
   RET(upload(c, c->u, u, np * nch * 4));
   RET(c->w.reserve(np * 8 * 4));
   RET(mgmb200_weights_dev(c, c->u.as<float>(), nx, ny, nch, aP, aThresh, c->w.as<float>()));
   RET(download(c, w_out, c->w.p, np * 8 * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" int mgmb200_costvolume(mgmb200_ctx *c, const float *u, const float *v, int nx, int ny, int nch, int vnx,
                                  int vny, int dmin, int dmax, const char *prefilter, const char *distance,
                                  float truncDist, int win, float *cc_out) {
   if (!c || !u || !v || !cc_out) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   CU(cudaSetDevice(c->device));
   int pf = mgmb200_prefilter_index(prefilter), di = mgmb200_distance_index(distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;   // consistency fix of mgm_costvolume.h:358-362
   const int L = dmax - dmin + 1, VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny;
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, (size_t)vnx * vny * nch * 4));
   RET(c->cc.reserve(np * VS * 4));
   RET(mgmb200_costvolume_dev(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, vnx, vny, dmin, dmax, pf, di,
                              truncDist, win, c->cc.as<float>()));
   if (VS == L) {
      RET(download(c, cc_out, c->cc.p, np * L * 4));
   } else {
      RET(c->dense.reserve(np * L * 4));
      CU(unpad_volume_launch(c->cc.as<float>(), c->dense.as<float>(), (long long)np, L, VS, c->stream));
      RET(download(c, cc_out, c->dense.p, np * L * 4));
   }
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

// uploads a host volume (pixel-major or label-major) into the padded c->cc and scans it for values outside the fast
// path's envelope (*flags: bit0 a vector without any finite entry, bit1 NaN, bit2 -INF); with flags == nullptr such a
// volume is rejected
static int upload_volume(mgmb200_ctx *c, const float *cc, int nx, int ny, int L, int label_major, int *flags = nullptr) {
   const int VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny;
   RET(c->cc.reserve(np * VS * 4));
   if (VS == L && !label_major) {
      CU(cudaMemcpyAsync(c->cc.p, cc, np * L * 4, cudaMemcpyHostToDevice, c->stream));
   } else {
      RET(upload(c, c->dense, cc, np * L * 4));
      CU(pad_volume_launch(c->dense.as<float>(), c->cc.as<float>(), (long long)np, L, VS, label_major, c->stream));
   }
   RET(clear_flags(c));
   CU(validate_volume_launch(c->cc.as<float>(), (long long)np, L, VS, c->flags.as<int>(), c->num_sms, c->stream));
   int fl = 0;
   RET(read_flags(c, &fl));
   if (flags) { *flags = fl; return 0; }
   if (fl)
      return fail(MGMB200_EUNSUPPORTED, "cost volume outside the supported envelope:%s%s%s",
                  (fl & 1) ? " a pixel without any finite cost" : "", (fl & 2) ? " NaN cost" : "",
                  (fl & 4) ? " -INF cost" : "");
   return 0;
}

namespace mgm {
cudaError_t agg_generic_launch(const float *cc, const float *w, int nx, int ny, int L, int VS, float P1, float P2, int NDIR,
                               int K, int variant, unsigned mask, float *const *ldir, int *d_counters, float *d_scratch,
                               int nwarps, cudaStream_t st);
}
// mgm() for inputs outside the fast path's envelope (non-finite costs, weights or penalties): the compare-select
// kernel of aggregate_generic.cu, then the ordinary finish kernel (whose sums, isfinite test and '>' argmin are the
// reference's own, mgm_core.cc:582-609).  Uniform ranges only.
static int aggregate_generic(mgmb200_ctx *c, const float *d_cc, const float *d_w, int nx, int ny, int dmin, int L, float P1,
                             float P2, int NDIR, int K, int felz, int fix, int refine, float *d_out, float *d_outcost,
                             float *d_S) {
   if (NDIR < 1 || NDIR > MGM_MAX_NDIR) return fail(MGMB200_EUNSUPPORTED, "NDIR=%d", NDIR);
   if (K < 1 || K > 4) return fail(MGMB200_EINVAL, "MGM/TSGM=%d not in 1..4", K);
   const int VS = mgmb200_padded_labels(L);
   const size_t vol = (size_t)nx * ny * VS * sizeof(float);
   bool weighted = false;
   if (d_w) {   // mgm_core.cc:420-422: any weight != 1 selects the W variants
      RET(clear_flags(c));
      CU(scan_weights_launch(d_w, (long long)nx * ny * 8, c->flags.as<int>(), c->stream));
      int fl = 0;
      RET(read_flags(c, &fl));
      weighted = fl & 1;
   }
   const int variant = weighted ? (felz > 0 ? 3 : 1) : (felz > 0 ? (K == 2 ? 2 : 3) : (K == 2 ? 0 : 1));   // :543-576
   if ((int)c->sweepv.size() < MGM_MAX_NDIR) c->sweepv.resize(MGM_MAX_NDIR);
   float *ldir[MGM_MAX_NDIR] = {nullptr};
   for (int p = 0; p < NDIR; p++) { RET(c->sweepv[p].reserve(vol)); ldir[p] = c->sweepv[p].as<float>(); }
   const int maxrows = std::max(nx, ny);
   const size_t ncount = 4 + (size_t)NDIR * maxrows;
   RET(c->progress.reserve(ncount * sizeof(int)));
   CU(cudaMemsetAsync(c->progress.p, 0, ncount * sizeof(int), c->stream));
   const int nwarps = (int)std::min<long long>((long long)NDIR * maxrows, (long long)c->num_sms * 32);
   RET(c->bnd.reserve((size_t)(nwarps + 4) * 4 * L * sizeof(float)));
   CU(agg_generic_launch(d_cc, weighted ? d_w : nullptr, nx, ny, L, VS, P1, P2, NDIR, K, variant, (1u << NDIR) - 1u, ldir,
                         c->progress.as<int>(), c->bnd.as<float>(), nwarps, c->stream));
   c->n_launches = 1;
   return finish_rows(c, ldir, d_cc, nx, ny, dmin, L, NDIR, fix, refine, 0, ny, d_out, d_outcost, d_S);
}

static int mgm_host(mgmb200_ctx *c, const float *cc, const float *w, int nx, int ny, int dmin, int dmax, float P1,
                    float P2, int NDIR, int K, int felz, int fix, int label_major, float *out, float *outcost,
                    float *S_out) {
   if (!c || !cc || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   CU(cudaSetDevice(c->device));
   const int L = dmax - dmin + 1;
   const size_t np = (size_t)nx * ny;
   int vflags = 0;
   RET(upload_volume(c, cc, nx, ny, L, label_major, &vflags));
   if (w) RET(upload(c, c->w, w, np * 8 * 4));
   RET(c->out.reserve(np * 4));
   RET(c->outcost.reserve(np * 4));
   float *dS = nullptr;
   if (S_out) { RET(c->dense.reserve(np * L * 4)); dS = c->dense.as<float>(); }
   // the fast kernels assume finite vectors, penalties and weights (hardware minima = the reference's compare-select
   // forms only then); anything else takes the compare-select kernel, which propagates non-finite values like the
   // reference
   bool generic = vflags != 0 || !(P1 >= 0.f) || !(P2 >= 0.f) || !(P1 < INFINITY) || (felz <= 0 && !(P2 < INFINITY));
   if (w && !generic) {
      RET(clear_flags(c));
      CU(scan_weights_launch(c->w.as<float>(), (long long)np * 8, c->flags.as<int>(), c->stream));
      int fl = 0;
      RET(read_flags(c, &fl));
      generic = (fl & 2) != 0;   // a negative, NaN or infinite weight
   }
   if (generic)
      RET(aggregate_generic(c, c->cc.as<float>(), w ? c->w.as<float>() : nullptr, nx, ny, dmin, L, P1, P2, NDIR, K, felz, fix, 0,
                            c->out.as<float>(), c->outcost.as<float>(), dS));
   else
      RET(mgmb200_aggregate_dev(c, c->cc.as<float>(), w ? c->w.as<float>() : nullptr, w ? 2 : 0, nx, ny, dmin, dmax, P1,
                                P2, NDIR, K, felz, fix, 0, c->out.as<float>(), c->outcost.as<float>(), dS));
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   if (S_out) RET(download(c, S_out, dS, np * L * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" int mgmb200_mgm(mgmb200_ctx *c, const float *cc, const float *w, int nx, int ny, int dmin, int dmax,
                           float P1, float P2, int NDIR, int MGM, int felz, int fix, float *out, float *outcost,
                           float *S_out) {
   return mgm_host(c, cc, w, nx, ny, dmin, dmax, P1, P2, NDIR, MGM, felz, fix, 0, out, outcost, S_out);
}

extern "C" int mgmb200_mgm_labelmajor(mgmb200_ctx *c, const float *costs, const float *w, int ncol, int nrow,
                                      int nlab, float P1, float P2, int NDIR, int MGM, int felz, float *labels_out,
                                      float *outcost) {
   if (!labels_out) return fail(MGMB200_EINVAL, "NULL argument");
   std::vector<float> tmp;
   if (!outcost) { tmp.resize((size_t)ncol * nrow); outcost = tmp.data(); }
   return mgm_host(c, costs, w, ncol, nrow, 0, nlab - 1, P1, P2, NDIR, MGM, felz, 1, 1, labels_out, outcost, nullptr);
}

extern "C" int mgmb200_subpixel_refinement_sgm(mgmb200_ctx *c, const float *S, int nx, int ny, int dmin, int dmax,
                                               float *out, float *outcost, const char *refinement) {
   if (!c || !S || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, dmin, dmax));
   const int m = mgmb200_refinement_index(refinement);
   if (m == 0) return 0;   // "none": out/outcost untouched (mgm_refine.h:47)
   // Separate entry point for callers that keep the reference call sequence mgm() -> refine();
   // the fused path (mgmb200_stereo / mgmb200_aggregate_dev) refines inside the WTA kernel.
   CU(cudaSetDevice(c->device));
   const int L = dmax - dmin + 1;
   const size_t np = (size_t)nx * ny;
   RET(upload(c, c->dense, S, np * L * 4));
   RET(upload(c, c->out, out, np * 4));
   RET(upload(c, c->outcost, outcost, np * 4));
   CU(refine_launch(c->dense.as<float>(), (long long)np, L, dmin, m, nullptr, nullptr, c->out.as<float>(),
                    c->outcost.as<float>(), c->stream));
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

// ------------------------------------------------------------------------------------------ per-pixel ranges
// (-m/-M range images, TSGM_ITER > 1.)  Volumes cross the boundary DENSE over an envelope [emin,emax];
// a label outside a pixel's range does not exist in the reference's Dvec: +INF in the dense array.
struct RangeScan { bool ok, ragged; };
static RangeScan scan_ranges(const float *lo, const float *hi, size_t np, int emin, int emax) {
   RangeScan r = {true, false};
   for (size_t i = 0; i < np; i++) {
      const int a = (int)lo[i], b = (int)hi[i];
      if (!(lo[i] == lo[i]) || !(hi[i] == hi[i]) || a > b || a < emin || b > emax) { r.ok = false; return r; }
      if (a != (int)lo[0] || b != (int)hi[0]) r.ragged = true;   // ragged = the pixels do not all have the same range
   }
   return r;
}

extern "C" int mgmb200_costvolume_ranges(mgmb200_ctx *c, const float *u, const float *v, int nx, int ny, int nch,
                                         int vnx, int vny, const float *dminI, const float *dmaxI, int emin, int emax,
                                         const char *prefilter, const char *distance, float truncDist, int win,
                                         float *cc_out) {
   if (!c || !u || !v || !dminI || !dmaxI || !cc_out) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, emin, emax));
   const size_t np = (size_t)nx * ny;
   if (!scan_ranges(dminI, dmaxI, np, emin, emax).ok)
      return fail(MGMB200_EINVAL, "a disparity range is empty, NaN or outside the envelope [%d,%d]", emin, emax);
   CU(cudaSetDevice(c->device));
   int pf = mgmb200_prefilter_index(prefilter), di = mgmb200_distance_index(distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;
   const int L = emax - emin + 1, VS = mgmb200_padded_labels(L);
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, (size_t)vnx * vny * nch * 4));
   RET(upload(c, c->rg[0], dminI, np * 4));
   RET(upload(c, c->rg[1], dmaxI, np * 4));
   RET(c->cc.reserve(np * VS * 4));
   RET(costvolume_dev_impl(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, vnx, vny, emin, emax, pf, di, truncDist,
                           win, c->rg[0].as<float>(), c->rg[1].as<float>(), c->cc.as<float>()));
   if (VS == L) {
      RET(download(c, cc_out, c->cc.p, np * L * 4));
   } else {
      RET(c->dense.reserve(np * L * 4));
      CU(unpad_volume_launch(c->cc.as<float>(), c->dense.as<float>(), (long long)np, L, VS, c->stream));
      RET(download(c, cc_out, c->dense.p, np * L * 4));
   }
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" int mgmb200_mgm_ranges(mgmb200_ctx *c, const float *cc, const float *ccmin, const float *ccmax,
                                  const float *w, int nx, int ny, int emin, int emax, const float *dminI,
                                  const float *dmaxI, float P1, float P2, int NDIR, int K, int felz, int fix,
                                  float *out, float *outcost, float *S_out) {
   if (!c || !cc || !ccmin || !ccmax || !dminI || !dmaxI || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, emin, emax));
   const int L = emax - emin + 1, VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny;
   const RangeScan rc = scan_ranges(ccmin, ccmax, np, emin, emax), rs = scan_ranges(dminI, dmaxI, np, emin, emax);
   if (!rc.ok || !rs.ok)
      return fail(MGMB200_EINVAL, "a disparity range is empty, NaN or outside the envelope [%d,%d]", emin, emax);
   bool weighted = false;
   if (w) for (size_t i = 0; i < np * 8 && !weighted; i++) weighted = (w[i] != 1.0f);
   // Truncated-linear potentials convolve inside the RECEIVING pixel's range (mgm_core.cc:229-281); only the
   // two-neighbour unweighted variant folds the rest back in (FixBounrady..., :166-186, :197-219) and thereby
   // equals the dense-envelope result.  The other variants with non-uniform ranges run on the consumer-side
   // (per-edge) kernels, which mask the neighbour's vector to the receiving pixel's range before convolving.
   const bool window = felz && rc.ragged && !(K == 2 && !weighted);
   CU(cudaSetDevice(c->device));
   RET(upload(c, c->rg[0], ccmin, np * 4));
   RET(upload(c, c->rg[1], ccmax, np * 4));
   RET(upload(c, c->rg[2], dminI, np * 4));
   RET(upload(c, c->rg[3], dmaxI, np * 4));
   // dense envelope -> padded volume, entries outside the cost ranges forced to +INF, preconditions checked
   RET(c->cc.reserve(np * VS * 4));
   RET(upload(c, c->dense, cc, np * L * 4));
   CU(pad_volume_launch(c->dense.as<float>(), c->cc.as<float>(), (long long)np, L, VS, 0, c->stream));
   CU(mask_volume_launch(c->cc.as<float>(), (long long)np, L, VS, emin, c->rg[0].as<float>(), c->rg[1].as<float>(),
                         c->stream));
   RET(clear_flags(c));
   CU(validate_volume_launch(c->cc.as<float>(), (long long)np, L, VS, c->flags.as<int>(), c->num_sms, c->stream));
   int fl = 0;
   RET(read_flags(c, &fl));
   if (fl)
      return fail(MGMB200_EUNSUPPORTED, "cost volume outside the supported envelope:%s%s%s",
                  (fl & 1) ? " a pixel without any finite cost" : "", (fl & 2) ? " NaN cost" : "",
                  (fl & 4) ? " -INF cost" : "");
   if (w) RET(upload(c, c->w, w, np * 8 * 4));
   RET(c->out.reserve(np * 4));
   RET(c->outcost.reserve(np * 4));
   float *dS = nullptr;
   if (S_out) { RET(c->dense.reserve(np * L * 4)); dS = c->dense.as<float>(); }
   c->r_ccmin = c->rg[0].as<float>(); c->r_ccmax = c->rg[1].as<float>();
   c->r_smin = c->rg[2].as<float>(); c->r_smax = c->rg[3].as<float>();
   c->r_window = window; c->r_emin = emin;
   const int rc2 = mgmb200_aggregate_dev(c, c->cc.as<float>(), w ? c->w.as<float>() : nullptr, w ? 2 : 0, nx, ny, emin,
                                         emax, P1, P2, NDIR, K, felz, fix, 0, c->out.as<float>(), c->outcost.as<float>(), dS);
   c->r_ccmin = c->r_ccmax = c->r_smin = c->r_smax = nullptr;
   c->r_window = false;
   if (rc2) return rc2;
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   if (S_out) RET(download(c, S_out, dS, np * L * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" int mgmb200_subpixel_refinement_sgm_ranges(mgmb200_ctx *c, const float *S, const float *dminI,
                                                      const float *dmaxI, int nx, int ny, int emin, int emax,
                                                      float *out, float *outcost, const char *refinement) {
   if (!c || !S || !dminI || !dmaxI || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, emin, emax));
   const int m = mgmb200_refinement_index(refinement);
   if (m == 0) return 0;
   const int L = emax - emin + 1;
   const size_t np = (size_t)nx * ny;
   if (!scan_ranges(dminI, dmaxI, np, emin, emax).ok)
      return fail(MGMB200_EINVAL, "a disparity range is empty, NaN or outside the envelope [%d,%d]", emin, emax);
   CU(cudaSetDevice(c->device));
   RET(upload(c, c->dense, S, np * L * 4));
   RET(upload(c, c->out, out, np * 4));
   RET(upload(c, c->outcost, outcost, np * 4));
   RET(upload(c, c->rg[2], dminI, np * 4));
   RET(upload(c, c->rg[3], dmaxI, np * 4));
   CU(refine_launch(c->dense.as<float>(), (long long)np, L, emin, m, c->rg[2].as<float>(), c->rg[3].as<float>(),
                    c->out.as<float>(), c->outcost.as<float>(), c->stream));
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" void mgmb200_stereo_params_default(mgmb200_stereo_params *p) {
   if (!p) return;
   p->dmin = -30; p->dmax = 30;
   p->P1 = 8.f; p->P2 = 32.f;
   p->NDIR = 4; p->MGM = 4;
   p->use_felzenszwalb_potentials = 0; p->sgm_fix_overcount = 1;
   p->aP = 1.f; p->aThresh = 5.f;
   p->prefilter = "none"; p->distance = "ad";
   p->truncDist = INFINITY; p->census_ncc_win = 3;
   p->refinement = "none";
}

// one direction of mgm.cc:372-385 on device-resident images: weights, cost volume, aggregation, WTA + refinement
// weights + cost volume of one direction into (wbuf, ccbuf); *weighted = the image has weights != 1 (mgm_core.cc:420-422)
static int stereo_build(mgmb200_ctx *c, const float *d_u, const float *d_v, int nx, int ny, int nch,
                        const mgmb200_stereo_params *p, int dmin, int dmax, int pf, int di, DevBuf &wbuf, DevBuf &ccbuf,
                        int *weighted) {
   const int L = dmax - dmin + 1, VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny;
   RET(wbuf.reserve(np * 8 * 4));
   RET(ccbuf.reserve(np * VS * 4));
   RET(clear_flags(c));
   // weights + the "all ones?" scan of mgm_core.cc:420-422 in one kernel
   CU(weights_launch(d_u, nx, ny, nch, p->aP, p->aThresh, wbuf.as<float>(), c->flags.as<int>(), c->stream));
   RET(mgmb200_costvolume_dev(c, d_u, d_v, nx, ny, nch, nx, ny, dmin, dmax, pf, di, p->truncDist, p->census_ncc_win,
                              ccbuf.as<float>()));
   int fl = 0;
   RET(read_flags(c, &fl));
   if ((fl & 1) && !(p->aP >= 0.f && p->aP < INFINITY)) return fail(MGMB200_EUNSUPPORTED, "aP must be finite and >= 0");
   *weighted = fl & 1;
   return 0;
}

static int stereo_dev(mgmb200_ctx *c, const float *d_u, const float *d_v, int nx, int ny, int nch,
                      const mgmb200_stereo_params *p, int dmin, int dmax, int pf, int di, float *d_out, float *d_outcost) {
   int fl = 0;
   RET(stereo_build(c, d_u, d_v, nx, ny, nch, p, dmin, dmax, pf, di, c->w, c->cc, &fl));
   const float P1 = p->P1 * nch, P2 = p->P2 * nch;   // mgm.cc:356-357
   return mgmb200_aggregate_dev(c, c->cc.as<float>(), c->w.as<float>(), (fl & 1) ? 1 : 0, nx, ny, dmin, dmax, P1, P2, p->NDIR,
                                p->MGM, p->use_felzenszwalb_potentials, p->sgm_fix_overcount,
                                mgmb200_refinement_index(p->refinement), d_out, d_outcost, nullptr);
}

extern "C" int mgmb200_stereo(mgmb200_ctx *c, const float *u, const float *v, int nx, int ny, int nch,
                              const mgmb200_stereo_params *p, float *out, float *outcost) {
   if (!c || !u || !v || !p || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, p->dmin, p->dmax));
   CU(cudaSetDevice(c->device));
   const int L = p->dmax - p->dmin + 1, VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny;
   int pf = mgmb200_prefilter_index(p->prefilter), di = mgmb200_distance_index(p->distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, np * nch * 4));
   RET(c->out.reserve(np * 4));
   RET(c->outcost.reserve(np * 4));
   RET(stereo_dev(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, p, p->dmin, p->dmax, pf, di, c->out.as<float>(),
                  c->outcost.as<float>()));
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

// ------------------------------------------------------------------------------------------ post-processing (N1/N2)
extern "C" int mgmb200_leftright_test_dev(mgmb200_ctx *c, const float *d_dx, int nx, int ny, const float *d_Rdx, int rnx,
                                          int rny, float threshold, float *d_out) {
   if (!c || !d_dx || !d_Rdx || !d_out) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || rnx < 1 || rny < ny) return fail(MGMB200_EINVAL, "maps %dx%d vs %dx%d: the other view needs at least as many rows", nx, ny, rnx, rny);
   if (d_out == d_dx || d_out == d_Rdx) return fail(MGMB200_EINVAL, "the test is out of place (mgm.cc:421-424 passes copies)");
   CU(cudaSetDevice(c->device));
   CU(leftright_launch(d_dx, nx, ny, d_Rdx, rnx, threshold, d_out, c->stream));
   return 0;
}
extern "C" int mgmb200_median_filter_dev(mgmb200_ctx *c, const float *d_u, int nx, int ny, int nch, int radius,
                                         float *d_out) {
   if (!c || !d_u || !d_out) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1 || d_out == d_u) return fail(MGMB200_EINVAL, "image %dx%dx%d (out of place)", nx, ny, nch);
   if (radius < 0 || radius > MGM_MEDIAN_MAX_RADIUS) return fail(MGMB200_EUNSUPPORTED, "median radius %d: 0..%d supported", radius, MGM_MEDIAN_MAX_RADIUS);
   CU(cudaSetDevice(c->device));
   CU(median_launch(d_u, nx, ny, nch, radius, d_out, c->stream));
   return 0;
}
// d_minmax: 4 floats of device scratch; [0], [1] receive the finite min/max of d_outoff
extern "C" int mgmb200_update_dmin_dmax_dev(mgmb200_ctx *c, const float *d_outoff, int nx, int ny, float *d_dminI,
                                            float *d_dmaxI, int slack, int radius, float *d_minmax) {
   if (!c || !d_outoff || !d_dminI || !d_dmaxI || !d_minmax) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || radius < 0) return fail(MGMB200_EINVAL, "image %dx%d radius %d", nx, ny, radius);
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny;
   RET(c->post[8].reserve(np * 4));
   RET(c->post[9].reserve(np * 4));
   CU(minmax_launch(d_outoff, (long long)np, d_minmax, c->num_sms, c->stream));
   CU(update_range_launch(d_outoff, nx, ny, d_minmax, slack, radius, d_dminI, d_dmaxI, c->post[8].as<float>(),
                          c->post[9].as<float>(), c->stream));
   CU(cudaMemcpyAsync(d_dminI, c->post[8].p, np * 4, cudaMemcpyDeviceToDevice, c->stream));
   CU(cudaMemcpyAsync(d_dmaxI, c->post[9].p, np * 4, cudaMemcpyDeviceToDevice, c->stream));
   return 0;
}
extern "C" int mgmb200_backproject_dev(mgmb200_ctx *c, const float *d_outoff, const float *d_u, const float *d_v, int nx,
                                       int ny, int nch, int vnx, int vny, float *d_syn) {
   if (!c || !d_outoff || !d_u || !d_v || !d_syn) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1 || vnx < 1 || vny < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d / %dx%d", nx, ny, nch, vnx, vny);
   CU(cudaSetDevice(c->device));
   CU(backproject_launch(d_outoff, d_u, d_v, nx, ny, nch, vnx, vny, d_syn, c->stream));
   return 0;
}

extern "C" int mgmb200_leftright_test(mgmb200_ctx *c, float *dx, int nx, int ny, const float *Rdx, int rnx, int rny,
                                      float threshold) {
   if (!c || !dx || !Rdx) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || rnx < 1 || rny < ny) return fail(MGMB200_EINVAL, "maps %dx%d vs %dx%d: the other view needs at least as many rows", nx, ny, rnx, rny);
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny, rnp = (size_t)rnx * rny;
   RET(upload(c, c->post[0], dx, np * 4));
   RET(upload(c, c->post[1], Rdx, rnp * 4));
   RET(c->post[2].reserve(np * 4));
   RET(mgmb200_leftright_test_dev(c, c->post[0].as<float>(), nx, ny, c->post[1].as<float>(), rnx, rny, threshold,
                                  c->post[2].as<float>()));
   RET(download(c, dx, c->post[2].p, np * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}
extern "C" int mgmb200_median_filter(mgmb200_ctx *c, const float *u, int nx, int ny, int nch, int radius, float *out) {
   if (!c || !u || !out) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d", nx, ny, nch);
   CU(cudaSetDevice(c->device));
   const size_t n = (size_t)nx * ny * nch;
   RET(upload(c, c->post[0], u, n * 4));
   RET(c->post[2].reserve(n * 4));
   RET(mgmb200_median_filter_dev(c, c->post[0].as<float>(), nx, ny, nch, radius, c->post[2].as<float>()));
   RET(download(c, out, c->post[2].p, n * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}
extern "C" int mgmb200_update_dmin_dmax(mgmb200_ctx *c, const float *outoff, int nx, int ny, float *dminI, float *dmaxI,
                                        int slack, int radius, float *gmin, float *gmax) {
   if (!c || !outoff || !dminI || !dmaxI) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1) return fail(MGMB200_EINVAL, "image %dx%d", nx, ny);
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny;
   RET(upload(c, c->post[0], outoff, np * 4));
   RET(upload(c, c->post[1], dminI, np * 4));
   RET(upload(c, c->post[2], dmaxI, np * 4));
   RET(c->post[7].reserve(64));
   RET(mgmb200_update_dmin_dmax_dev(c, c->post[0].as<float>(), nx, ny, c->post[1].as<float>(), c->post[2].as<float>(), slack,
                                    radius, c->post[7].as<float>()));
   float mm[2];
   RET(download(c, dminI, c->post[1].p, np * 4));
   RET(download(c, dmaxI, c->post[2].p, np * 4));
   RET(download(c, mm, c->post[7].p, 8));
   CU(cudaStreamSynchronize(c->stream));
   if (gmin) *gmin = mm[0];
   if (gmax) *gmax = mm[1];
   return 0;
}
extern "C" int mgmb200_backproject(mgmb200_ctx *c, const float *outoff, const float *u, const float *v, int nx, int ny,
                                   int nch, int vnx, int vny, float *syn) {
   if (!c || !outoff || !u || !v || !syn) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1 || vnx < 1 || vny < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d / %dx%d", nx, ny, nch, vnx, vny);
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny;
   RET(upload(c, c->post[0], outoff, np * 4));
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, (size_t)vnx * vny * nch * 4));
   RET(c->post[2].reserve(np * nch * 4));
   RET(mgmb200_backproject_dev(c, c->post[0].as<float>(), c->u.as<float>(), c->v.as<float>(), nx, ny, nch, vnx, vny,
                               c->post[2].as<float>()));
   RET(download(c, syn, c->post[2].p, np * nch * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

extern "C" void mgmb200_post_params_default(mgmb200_post_params *q) {
   if (!q) return;
   q->testlrrl = 1; q->testlrrl_tau = 1.f; q->median = 0;   // mgm.cc:194-196
}

// The default CLI flow of mgm.cc:372-443 for uniform ranges and TSGM_ITER=1, resident on the device: the images go up
// once, both directions run back to back, median + left-right tests + back-projection follow as small kernels, and
// only the requested maps come back.
extern "C" int mgmb200_stereo_lr(mgmb200_ctx *c, const float *u, const float *v, int nx, int ny, int nch,
                                 const mgmb200_stereo_params *p, const mgmb200_post_params *q, float *out,
                                 float *outcost, float *outR, float *outcostR, float *out_nolr, float *backproj) {
   if (!c || !u || !v || !p || !q || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   RET(check_dims(nx, ny, p->dmin, p->dmax));
   if (q->median < 0 || q->median > MGM_MEDIAN_MAX_RADIUS) return fail(MGMB200_EUNSUPPORTED, "median radius %d: 0..%d supported", q->median, MGM_MEDIAN_MAX_RADIUS);
   if (!q->testlrrl && (outR || outcostR)) return fail(MGMB200_EINVAL, "outR/outcostR need testlrrl (the second run, mgm.cc:404)");
   CU(cudaSetDevice(c->device));
   const size_t np = (size_t)nx * ny;
   int pf = mgmb200_prefilter_index(p->prefilter), di = mgmb200_distance_index(p->distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, np * nch * 4));
   for (int i = 0; i < 6; i++) RET(c->post[i].reserve(np * 4));
   float *offL = c->post[0].as<float>(), *costL = c->post[1].as<float>(), *offR = c->post[2].as<float>(),
         *costR = c->post[3].as<float>(), *t0 = c->post[4].as<float>(), *t1 = c->post[5].as<float>();
   // Both directions (mgm.cc:372-385 and :405-414) are independent until the left-right test: when they take the same
   // kernels (both images with or both without image-dependent weights) their 2 x NDIR sweeps share one launch -- the
   // tail of one direction is filled with bands of the other -- instead of two launches back to back.
   bool both_done = false;
   if (q->testlrrl && !c->tune.lr_sequential) {
      int wl = 0, wr = 0;
      if (c->bcc.empty()) { c->bcc.resize(1); c->bw.resize(1); }
      RET(stereo_build(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, p, p->dmin, p->dmax, pf, di, c->w, c->cc, &wl));
      RET(stereo_build(c, c->v.as<float>(), c->u.as<float>(), nx, ny, nch, p, -p->dmax, -p->dmin, pf, di, c->bw[0], c->bcc[0], &wr));
      const float P1 = p->P1 * nch, P2 = p->P2 * nch;   // mgm.cc:356-357
      const int refine = mgmb200_refinement_index(p->refinement);
      if (wl == wr) {
         const float *ccs[2] = {c->cc.as<float>(), c->bcc[0].as<float>()}, *ws[2] = {c->w.as<float>(), c->bw[0].as<float>()};
         float *outs[2] = {offL, offR}, *costs[2] = {costL, costR};
         const int dmins[2] = {p->dmin, -p->dmax};
         RET(aggregate_pairs(c, 2, ccs, wl ? ws : nullptr, nx, ny, dmins, p->dmax - p->dmin + 1, P1, P2, p->NDIR, p->MGM,
                             p->use_felzenszwalb_potentials, p->sgm_fix_overcount, refine, outs, costs));
      } else {
         RET(mgmb200_aggregate_dev(c, c->cc.as<float>(), c->w.as<float>(), wl, nx, ny, p->dmin, p->dmax, P1, P2, p->NDIR, p->MGM,
                                   p->use_felzenszwalb_potentials, p->sgm_fix_overcount, refine, offL, costL, nullptr));
         RET(mgmb200_aggregate_dev(c, c->bcc[0].as<float>(), c->bw[0].as<float>(), wr, nx, ny, -p->dmax, -p->dmin, P1, P2, p->NDIR,
                                   p->MGM, p->use_felzenszwalb_potentials, p->sgm_fix_overcount, refine, offR, costR, nullptr));
      }
      both_done = true;
   } else
   RET(stereo_dev(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, p, p->dmin, p->dmax, pf, di, offL, costL));
   if (q->median) {   // mgm.cc:396
      CU(median_launch(offL, nx, ny, 1, q->median, t0, c->stream));
      std::swap(offL, t0);
   }
   if (out_nolr) RET(download(c, out_nolr, offL, np * 4));   // mgm.cc:399-401
   if (q->testlrrl) {
      // mgm.cc:404-419: the other direction with the mirrored range
      if (!both_done)
         RET(stereo_dev(c, c->v.as<float>(), c->u.as<float>(), nx, ny, nch, p, -p->dmax, -p->dmin, pf, di, offR, costR));
      if (q->median) {
         CU(median_launch(offR, nx, ny, 1, q->median, t0, c->stream));
         std::swap(offR, t0);
      }
      // mgm.cc:420-424: both tests read the untested maps
      CU(leftright_launch(offR, nx, ny, offL, nx, q->testlrrl_tau, t0, c->stream));
      CU(leftright_launch(offL, nx, ny, offR, nx, q->testlrrl_tau, t1, c->stream));
      offR = t0;
      offL = t1;
      if (outR) RET(download(c, outR, offR, np * 4));
      if (outcostR) RET(download(c, outcostR, costR, np * 4));
   }
   RET(download(c, out, offL, np * 4));
   RET(download(c, outcost, costL, np * 4));
   if (backproj) {
      RET(c->post[6].reserve(np * nch * 4));
      CU(backproject_launch(offL, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, nx, ny, c->post[6].as<float>(), c->stream));
      RET(download(c, backproj, c->post[6].p, np * nch * 4));
   }
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}

// The hot path of mgm.cc:372-385 for a batch of pairs of one shape (BASELINE.json configs[3]: 32 KITTI-size frames):
// every pair's images go up, weights and cost volume are built, then the aggregation runs for several pairs per launch
// (mgmb200_aggregate_batch_dev) and the maps come back.  Same results as npairs mgmb200_stereo calls.
extern "C" int mgmb200_stereo_batch(mgmb200_ctx *c, int npairs, const float *const *u, const float *const *v, int nx, int ny,
                                    int nch, const mgmb200_stereo_params *p, float *const *out, float *const *outcost) {
   if (!c || !u || !v || !p || !out || !outcost || npairs < 0) return fail(MGMB200_EINVAL, "bad argument");
   RET(check_dims(nx, ny, p->dmin, p->dmax));
   CU(cudaSetDevice(c->device));
   const int L = p->dmax - p->dmin + 1, VS = mgmb200_padded_labels(L);
   const size_t np = (size_t)nx * ny, vol = np * VS * 4;
   int pf = mgmb200_prefilter_index(p->prefilter), di = mgmb200_distance_index(p->distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;
   const float P1 = p->P1 * nch, P2 = p->P2 * nch;   // mgm.cc:356-357
   const int refine = mgmb200_refinement_index(p->refinement);
   const int chunk = std::max(1, c->tune.batch);
   if ((int)c->bcc.size() < chunk) { c->bcc.resize(chunk); c->bw.resize(chunk); c->bout.resize(2 * (size_t)chunk); }
   int launches = 0;
   for (int b0 = 0; b0 < npairs; b0 += chunk) {
      const int nb = std::min(chunk, npairs - b0);
      int weighted_any = 0, weighted_all = 1;
      std::vector<const float *> ccs((size_t)nb), ws((size_t)nb);
      std::vector<float *> outs((size_t)nb), costs((size_t)nb);
      for (int b = 0; b < nb; b++) {
         if (!u[b0 + b] || !v[b0 + b] || !out[b0 + b] || !outcost[b0 + b]) return fail(MGMB200_EINVAL, "NULL pointer for pair %d", b0 + b);
         RET(upload(c, c->u, u[b0 + b], np * nch * 4));
         RET(upload(c, c->v, v[b0 + b], np * nch * 4));
         RET(c->bcc[b].reserve(vol));
         RET(c->bw[b].reserve(np * 8 * 4));
         RET(c->bout[2 * b].reserve(np * 4));
         RET(c->bout[2 * b + 1].reserve(np * 4));
         RET(clear_flags(c));
         CU(weights_launch(c->u.as<float>(), nx, ny, nch, p->aP, p->aThresh, c->bw[b].as<float>(), c->flags.as<int>(), c->stream));
         RET(mgmb200_costvolume_dev(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, nx, ny, p->dmin, p->dmax, pf, di,
                                    p->truncDist, p->census_ncc_win, c->bcc[b].as<float>()));
         int fl = 0;
         RET(read_flags(c, &fl));
         if ((fl & 1) && !(p->aP >= 0.f && p->aP < INFINITY)) return fail(MGMB200_EUNSUPPORTED, "aP must be finite and >= 0");
         weighted_any |= (fl & 1);
         weighted_all &= (fl & 1);
         ccs[b] = c->bcc[b].as<float>(); ws[b] = c->bw[b].as<float>();
         outs[b] = c->bout[2 * b].as<float>(); costs[b] = c->bout[2 * b + 1].as<float>();
      }
      if (weighted_any == weighted_all) {   // all pairs take the same kernels (mgm_core.cc:420-423 decides per pair)
         RET(mgmb200_aggregate_batch_dev(c, nb, ccs.data(), weighted_any ? ws.data() : nullptr, nx, ny, p->dmin, p->dmax, P1, P2,
                                         p->NDIR, p->MGM, p->use_felzenszwalb_potentials, p->sgm_fix_overcount, refine,
                                         outs.data(), costs.data()));
         launches += c->n_launches;
      } else {
         for (int b = 0; b < nb; b++) {
            RET(mgmb200_aggregate_dev(c, ccs[b], ws[b], 2, nx, ny, p->dmin, p->dmax, P1, P2, p->NDIR, p->MGM,
                                      p->use_felzenszwalb_potentials, p->sgm_fix_overcount, refine, outs[b], costs[b], nullptr));
            launches += c->n_launches;
         }
      }
      for (int b = 0; b < nb; b++) {
         RET(download(c, out[b0 + b], outs[b], np * 4));
         RET(download(c, outcost[b0 + b], costs[b], np * 4));
      }
   }
   CU(cudaStreamSynchronize(c->stream));
   c->n_launches = launches;
   return 0;
}

// mgm.cc:372-395 for one direction with range images and TSGM_ITER iterations, resident on the device: the cost
// volume is built once over an envelope wide enough for every iteration (update_dmin_dmax widens a range by its slack
// of 3 per iteration), then each iteration aggregates + finishes over the current ranges and updates them from the
// refined disparities.  dminI/dmaxI come back as left by the last update (mgm.cc:390-392).
extern "C" int mgmb200_stereo_ranges(mgmb200_ctx *c, const float *u, const float *v, int nx, int ny, int nch,
                                     const mgmb200_stereo_params *p, float *dminI, float *dmaxI, int tsgm_iter,
                                     float *out, float *outcost) {
   if (!c || !u || !v || !p || !dminI || !dmaxI || !out || !outcost) return fail(MGMB200_EINVAL, "NULL argument");
   if (nx < 1 || ny < 1 || nch < 1) return fail(MGMB200_EINVAL, "image %dx%dx%d", nx, ny, nch);
   if (tsgm_iter < 1) tsgm_iter = 0;   // the reference's loop simply does not run (outputs stay untouched)
   const size_t np = (size_t)nx * ny;
   int rmin = 0x7fffffff, rmax = -0x7fffffff;
   for (size_t i = 0; i < np; i++) {
      if (!(dminI[i] == dminI[i]) || !(dmaxI[i] == dmaxI[i]) || !(fabsf(dminI[i]) < 1e9f) || !(fabsf(dmaxI[i]) < 1e9f))
         return fail(MGMB200_EINVAL, "non-finite disparity range at pixel %zu", i);
      const int a = (int)dminI[i], b = (int)dmaxI[i];
      if (a > b) return fail(MGMB200_EINVAL, "empty disparity range [%d,%d] at pixel %zu", a, b, i);
      rmin = std::min(rmin, a); rmax = std::max(rmax, b);
   }
   const RangeScan rc = scan_ranges(dminI, dmaxI, np, rmin, rmax);
   const int slack = 3, radius = 2;   // defaults of update_dmin_dmax, mgm.cc:120
   const int grow = tsgm_iter > 1 ? slack * (tsgm_iter - 1) : 0;
   const int emin = rmin - grow, emax = rmax + grow;
   RET(check_dims(nx, ny, emin, emax));
   if (tsgm_iter == 0) return 0;
   CU(cudaSetDevice(c->device));
   const int L = emax - emin + 1, VS = mgmb200_padded_labels(L);
   int pf = mgmb200_prefilter_index(p->prefilter), di = mgmb200_distance_index(p->distance);
   if (di == DIST_CENSUS) pf = PF_CENSUS;
   RET(upload(c, c->u, u, np * nch * 4));
   RET(upload(c, c->v, v, np * nch * 4));
   RET(upload(c, c->rg[0], dminI, np * 4));   // ranges of the cost vectors: fixed (mgm.cc:376 builds CC once)
   RET(upload(c, c->rg[1], dmaxI, np * 4));
   RET(upload(c, c->rg[2], dminI, np * 4));   // ranges of S: updated after every iteration
   RET(upload(c, c->rg[3], dmaxI, np * 4));
   RET(c->w.reserve(np * 8 * 4));
   RET(c->cc.reserve(np * VS * 4));
   RET(c->out.reserve(np * 4));
   RET(c->outcost.reserve(np * 4));
   RET(c->post[7].reserve(64));
   RET(clear_flags(c));
   CU(weights_launch(c->u.as<float>(), nx, ny, nch, p->aP, p->aThresh, c->w.as<float>(), c->flags.as<int>(), c->stream));
   RET(costvolume_dev_impl(c, c->u.as<float>(), c->v.as<float>(), nx, ny, nch, nx, ny, emin, emax, pf, di, p->truncDist,
                           p->census_ncc_win, c->rg[0].as<float>(), c->rg[1].as<float>(), c->cc.as<float>()));
   int fl = 0;
   RET(read_flags(c, &fl));
   const bool weighted = (fl & 1) != 0;
   if (weighted && !(p->aP >= 0.f && p->aP < INFINITY)) return fail(MGMB200_EUNSUPPORTED, "aP must be finite and >= 0");
   const int K = p->MGM, felz = p->use_felzenszwalb_potentials;
   const float P1 = p->P1 * nch, P2 = p->P2 * nch;   // mgm.cc:356-357
   const int refine = mgmb200_refinement_index(p->refinement);
   float *mm = c->post[7].as<float>();
   for (int it = 0; it < tsgm_iter; it++) {
      c->r_ccmin = c->rg[0].as<float>(); c->r_ccmax = c->rg[1].as<float>();
      c->r_smin = c->rg[2].as<float>(); c->r_smax = c->rg[3].as<float>();
      c->r_window = felz && rc.ragged && !(K == 2 && !weighted);
      c->r_emin = emin;
      const int r = mgmb200_aggregate_dev(c, c->cc.as<float>(), c->w.as<float>(), weighted ? 1 : 0, nx, ny, emin, emax, P1, P2,
                                          p->NDIR, K, felz, p->sgm_fix_overcount, refine, c->out.as<float>(),
                                          c->outcost.as<float>(), nullptr);
      c->r_ccmin = c->r_ccmax = c->r_smin = c->r_smax = nullptr;
      c->r_window = false;
      if (r) return r;
      // update_dmin_dmax + remove_nonfinite_values_Img on both range images, mgm.cc:390-392
      RET(mgmb200_update_dmin_dmax_dev(c, c->out.as<float>(), nx, ny, c->rg[2].as<float>(), c->rg[3].as<float>(), slack,
                                       radius, mm));
      CU(replace_nonfinite_launch(c->rg[2].as<float>(), (long long)np, mm + 0, c->stream));
      CU(replace_nonfinite_launch(c->rg[3].as<float>(), (long long)np, mm + 1, c->stream));
   }
   RET(download(c, out, c->out.p, np * 4));
   RET(download(c, outcost, c->outcost.p, np * 4));
   RET(download(c, dminI, c->rg[2].p, np * 4));
   RET(download(c, dmaxI, c->rg[3].p, np * 4));
   CU(cudaStreamSynchronize(c->stream));
   return 0;
}
