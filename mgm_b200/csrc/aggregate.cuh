// Parameters and launch plan of the aggregation kernel (aggregate.cu).
#pragma once
#include "common.cuh"
#include "wta.cuh"

#define MGM_AGG_GROUP 8           // lanes cooperating on one scan row
#define MGM_AGG_MAX_THREADS 512   // 56 rows x 8 lanes + two service warps (boundary consumer / publisher); 128 registers per thread
#define MGM_AGG_CREG 8            // 16-byte cost chunks per lane prefetched into registers (covers 256 labels)

namespace mgm {

enum PotKind { POT_SGM = 0, POT_TRUNC = 1 };

struct AggParams {
   const float *cc;            // matching costs [ny][nx][VS], labels >= L hold +INF
   const float *w;             // 8 weight planes [8][ny][nx] (weighted kernels only; nullptr = all ones)
   const float *win_lo, *win_hi;   // per-pixel cost ranges (float images, truncated) or nullptr: the weighted
   int win_emin;                   // truncated-linear kernels convolve inside the receiving pixel's range
   float *ldir[8];             // per-sweep message volumes [ny][nx][VS], indexed by sweep id
   float *bnd[8];              // per-sweep boundary lines [nbands][maxii][VS]
   float *bndm[8];             // per-sweep boundary minima [nbands][maxii]
   int *progress[8];           // per-sweep [nbands]: finished pixels of each band's last row
   int *next_band;             // per-sweep claim counters [8]: bands of a sweep are claimed in order
   int nb[8];                  // bands per sweep (0 = sweep not requested)
   int nbands;                 // total
   int static_order;           // debugging knob: claim without the readiness test
   int nx, ny, L, VS;
   int T[2];                   // rows per band: [0] axis sweeps 0-3, [1] diagonal sweeps 4-7
   int TS[2];                  // per-row shared-memory stride in floats
   int ncb;                    // cost buffers per row (1: costs prefetched into registers, 2: cp.async ring)
   int shear;                  // 1: sweeps 4-7 run as sheared wavefronts (bands of anti-diagonals, run_band_shear)
   int ng[2];                  // row groups per band (axis / diagonal class), each on its own named barrier
   int fused_sgm;              // 1: unweighted SGM kernels transform the message from registers (one barrier per step)
   int cc_pf;                  // > 0: matching costs of the pixel cc_pf steps ahead are prefetched into L2
   float P1, P2;
   // fused finish (optional): CTAs without a band to run take tiles of pixels whose sweeps are all complete and do
   // the ordered sum + over-count fix + WTA + sub-pixel there (wta_device.cuh), inside the same launch
   int fin_enabled;
   int fin_ntiles, fin_tw, fin_th, fin_tiles_x;   // tiles of fin_tw x fin_th pixels, fin_tiles_x per image row of tiles
   const int *fin_order;       // tile ids in the expected order of readiness
   int *fin_next;              // claim counter into fin_order
   int *band_done[8];          // per-sweep [nbands]: 1 once every message of the band is in memory
   WtaParams fin;
   unsigned long long *dbg;    // optional 24-word phase-timing accumulator (profiling aid), or nullptr
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
   int lanes = 0;                      // 0 auto, 4 or 8 lanes per worker (unweighted SGM kernels)
   int no_shear = 0;                   // 1: diagonal sweeps row-per-worker
   int static_order = 0;               // 1: claim bands without the readiness test
   int no_fused_finish = 0;            // 1: finish stage as a separate launch
   int fin_tw = 128, fin_th = 16;      // finish tile
   int cc_pf = 0;                      // L2 prefetch distance (pixels) of the matching costs, 0 = off
   int verbose = 0;
};

struct AggPlan {
   int VS, T[2], TS[2], ncb, shear, ng[2], fused_sgm, lanes, block, num_sms, verbose;
   size_t smem;
   size_t off_phase, off_cbar, off_vbar, off_ms, off_vms, off_virt, off_thr;
};

void agg_plan(AggPlan *plan, int nx, int ny, int L, int K, int pot, bool weighted, int max_smem, int num_sms,
              int t_override, const AggTuning &tune);
cudaError_t agg_launch(const AggParams &P, const AggPlan &plan, int pot, int K, bool weighted, cudaStream_t st);

}  // namespace mgm
