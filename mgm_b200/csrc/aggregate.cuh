// Parameters and launch plan of the aggregation kernel (aggregate.cu).
#pragma once
#include "common.cuh"
#include "wta.cuh"

#define MGM_AGG_GROUP 8           // lanes cooperating on one scan row
#define MGM_AGG_MAX_THREADS 512   // 56 rows x 8 lanes + two service warps (boundary consumer / publisher); 128 registers per thread
#define MGM_AGG_MAX_THREADS_RC 384   // register-chain kernels: 40 rows x 8 lanes + two service warps = 3 warps per SM sub-partition: 168 registers per thread
#define MGM_AGG_CREG 8            // 16-byte cost chunks per lane prefetched into registers (covers 256 labels)
#define MGM_MAX_SLABS 8           // row slabs of a message volume (multi-GPU: one per rank, peer mappings)
#define MGM_MAX_NDIR 16           // sweeps per stereo pair (8 of the reference + 8 defined here, common.cuh)

#ifndef MGM_TRUNC_LAY
#define MGM_TRUNC_LAY 1   // shared-memory vector layout of the lean truncated-linear kernels: 1 = 16-byte aligned rows and accesses,
                         // 0 = 8-byte halves with a row stride of 8 bytes modulo 128 (conflict-free LDS.64 gathers)
#endif

namespace mgm {

enum PotKind { POT_SGM = 0, POT_TRUNC = 1 };
enum SweepClass { CLS_AXIS = 0, CLS_DIAG = 1, CLS_KNIGHT = 2 };

// One sweep of one stereo pair ("virtual sweep").  A launch runs any number of them concurrently: the 8 or 16
// sweeps of a pair, of several pairs (batches, both directions of a left-right run), or the few sweeps a GPU owns
// in the sweep-sharded multi-GPU layout.  The table lives in device memory; a CTA copies the entry of the band it
// has claimed into shared memory.
struct SweepDesc {
   const float *cc;                  // matching costs [ny][nx][VS] of the sweep's pair, labels >= L hold +INF
   const float *w;                   // 8 weight planes [8][ny][nx] (weighted kernels only; nullptr = all ones)
   const float *win_lo, *win_hi;     // per-pixel cost ranges (float images, truncated) or nullptr: the weighted
                                     // truncated-linear kernels convolve inside the receiving pixel's range
   float *ldir[MGM_MAX_SLABS];       // message volume [ny][nx][VS]; with nslabs > 1 entry r is the volume that holds
                                     // image rows [r*slab_rows, (r+1)*slab_rows) -- on another GPU for r != own rank
   float *bnd;                       // boundary lines [nb][maxii][VS] (sheared: [nb][2][maxjj][VS])
   float *bndm;                      // boundary minima
   int *progress;                    // [nb] finished pixels of each band's last row
   int *band_done;                   // [nb] 1 once every message of the band is in memory (fused finish)
   int nb;                           // bands (0 = sweep not requested)
   int pass;                         // 0..15: scan geometry and predecessor order (common.cuh)
   int cls;                          // SweepClass: which band layout (T, TS) the sweep uses
   int filler;                       // 1: sheared diagonal sweep (short hand-off): claimed without the readiness test
   int pair;                         // index of the sweep's pair in the finish table
   int pad_[3];
};
static_assert(sizeof(SweepDesc) % 16 == 0, "staged into shared memory in 16-byte words");
static_assert(sizeof(WtaParams) % 16 == 0, "staged into shared memory in 16-byte words");

struct AggParams {
   const SweepDesc *sweeps;    // [nsweeps] device memory
   int nsweeps;
   int *next_band;             // [nsweeps] claim counters: bands of a sweep are claimed in order
   int nbands;                 // total
   int static_order;           // debugging knob: claim without the readiness test
   int win_emin;               // label origin of win_lo / win_hi
   int nx, ny, L, VS;
   int ndir;                   // sweeps per pair (the pairs' sweeps are consecutive table entries)
   int nslabs, slab_rows;      // row slabs of the message volumes (1 = whole volume in ldir[0])
   unsigned slab_magic;        // ceil(2^32 / slab_rows): y / slab_rows = umulhi(y, magic) for y < 2^16
   int T[3];                   // rows per band by SweepClass
   int TS[3];                  // per-row shared-memory stride in floats by SweepClass
   int ncb;                    // cost buffers per row (1: costs prefetched into registers, 2: cp.async ring)
   int shear;                  // 1: sweeps 4-7 run as sheared wavefronts (bands of anti-diagonals, run_band_shear)
   int ng[3];                  // row groups per band by SweepClass, each on its own named barrier
   int fused_sgm;              // 1: unweighted SGM kernels transform the message from registers (one barrier per step)
   int regchain;               // 1: unweighted truncated-linear kernels run the min-convolution from registers (chain_regs)
   int cc_pf;                  // > 0: matching costs of the pixel cc_pf steps ahead are prefetched into L2
   float P1, P2;
   // fused finish (optional): CTAs without a band to run take tiles of pixels whose sweeps are all complete and do
   // the ordered sum + over-count fix + WTA + sub-pixel there (wta_device.cuh), inside the same launch
   int fin_enabled;
   int fin_ntiles;             // tiles per pair; global tile id = pair * fin_ntiles + tile
   int fin_total;              // npairs * fin_ntiles
   int fin_tw, fin_th, fin_tiles_x;   // tiles of fin_tw x fin_th pixels, fin_tiles_x per image row of tiles
   const int *fin_order;       // global tile ids in the expected order of readiness
   int *fin_next;              // claim counter into fin_order
   const WtaParams *fins;      // [npairs] device memory
   int npairs;
   WtaParams fin0;             // npairs == 1: the pair's finish parameters as kernel parameters (constant bank operands
                               // in the per-pixel loop instead of shared-memory loads: 25.1 -> 19.9 ms on the headline step)
   unsigned long long *dbg;    // optional 8-word phase-timing accumulator (option "dbg"), or nullptr
   // dynamic shared memory carve-up (bytes)
   unsigned off_phase, off_cbar, off_vbar, off_ms, off_vms, off_virt, off_thr;
};

// Tuning / debugging knobs of the aggregation launch.  They live in the context: read ONCE from the environment
// (MGMB200_*) by mgmb200_create and changed afterwards only through mgmb200_set_option -- never per call.
struct AggTuning {
   int rows_axis = 0, rows_diag = 0;   // rows (workers) per band, 0 = as many as shared memory allows
   int groups = 1;                     // row groups per band (RowGroup)
   int no_creg = 0;                    // 1: cp.async cost ring instead of register-resident costs
   int no_fused_sgm = 0;               // 1: SGM transform through shared memory (two barriers per step)
   int reg_chains = 0;                 // 1: truncated-linear min-convolution from registers by the worker's own 8 lanes
                                       // (chain_regs; parity-tested, measured SLOWER than the lane-pair chains: the
                                       // hops between lanes repeat the additions in every lane)
   int lanes = 0;                      // 0 auto, 4 or 8 lanes per worker (unweighted SGM kernels)
   int no_shear = 0;                   // 1: diagonal sweeps row-per-worker
   int static_order = 0;               // 1: claim bands without the readiness test
   int no_fused_finish = 0;            // 1: finish stage as a separate launch
   int fused_finish = -1;              // 1: finish tiles inside the launch whenever possible, -1: where measured to gain
   int fin_tw = 128, fin_th = 16;      // finish tile
   int cc_pf = -1;                     // L2 prefetch distance (pixels) of the matching costs; 0 = off, -1 = auto (3 for
                                       // the SGM-potential kernels, whose steps are shorter than the HBM latency)
   int batch = 16;                     // stereo pairs in flight per launch (batch entry points; measured 32 KITTI-size
                                       // pairs: 88 / 74 / 69 / 68 ms with 4 / 8 / 16 / 32 pairs per launch)
   int no_lean_sgm = 0;                // 1: unweighted SGM launches take the generic kernel (aggregate.cu) instead of the lean
                                       // one (aggregate_sgm.cu); parity tests run both
   int no_lean_trunc = 0;              // 1: unweighted truncated-linear launches take the generic kernel instead of aggregate_trunc.cu
   int full_block = 0;                 // 1: 512 threads per CTA even when the bands hold fewer rows (spare warps serve the finish tiles)
   int lr_sequential = 0;              // 1: mgmb200_stereo_lr runs its two directions in two launches
   int verbose = 0;
   int dbg = 0;                        // 1: print the per-phase clock cycles of the register-chain band steps (axis sweeps)
};

struct AggPlan {
   int VS, VSP, T[3], TS[3], ncb, shear, ng[3], fused_sgm, regchain, lanes, block, num_sms, verbose;
   int lean_sgm;   // 1: the launch qualifies for the lean unweighted-SGM kernels (aggregate_sgm.cu)
   int lean_trunc; // 1: ... for the lean unweighted truncated-linear kernels (aggregate_trunc.cu)
   int lean_sgmw;  // 1: ... for the lean SGM kernels with per-edge weights (aggregate_sgmw.cu)
   size_t smem;
   size_t off_phase, off_cbar, off_vbar, off_ms, off_vms, off_virt, off_thr;
};

// bands of one sweep under a plan, and the floats of boundary lines / minima it needs
int agg_sweep_class(const AggPlan &plan, int pass);
void agg_sweep_bands(const AggPlan &plan, int pass, int nx, int ny, int *nb, size_t *bnd_floats, size_t *bndm_floats);

// knight: the launch holds sweeps 8-15 (their band layout then counts for the shared-memory size)
void agg_plan(AggPlan *plan, int nx, int ny, int L, int K, int pot, bool weighted, int max_smem, int num_sms,
              int t_override, bool knight, const AggTuning &tune);
cudaError_t agg_launch(const AggParams &P, const AggPlan &plan, int pot, int K, bool weighted, cudaStream_t st);
cudaError_t agg_launch_generic_sgm(const AggParams &P, const AggPlan &plan, int K, bool weighted, cudaStream_t st);     // aggregate.cu via aggregate_gsgm.cu
cudaError_t agg_launch_generic_trunc(const AggParams &P, const AggPlan &plan, int K, bool weighted, cudaStream_t st);   // aggregate.cu via aggregate_gtrunc.cu
// lean unweighted-SGM kernels (aggregate_sgm.cu): label layouts they are built for, and their launch
bool agg_sgm_lean_supported(int VS, int lanes);
cudaError_t agg_launch_sgm_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st);
bool agg_trunc_lean_supported(int VS);
cudaError_t agg_launch_trunc_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st);
cudaError_t agg_launch_sgmw_lean(const AggParams &P, const AggPlan &plan, int K, cudaStream_t st);

}  // namespace mgm
