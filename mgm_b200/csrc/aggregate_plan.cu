// Launch plan of the aggregation kernels (host side): rows per band, shared-memory carve-up, which kernel family a
// launch takes.  Kept apart from the kernels so that a change here does not recompile them.
#include <stdio.h>
#include <stdlib.h>

#include "aggregate.cuh"

namespace mgm {

static constexpr int RV = 8;   // virtual-row ring (aggregate_dev.cuh)

static int ring_slots(int cls, int K) { return ((cls != CLS_AXIS || K == 4) ? 2 : 1) + 2; }

static void agg_plan_try(AggPlan *plan, int nx, int ny, int L, int K, int pot, bool weighted, int max_smem, int num_sms,
                         int t_override, int shear, bool knight, const AggTuning &tune) {
   plan->verbose = tune.verbose;
   const int VS = (L + 31) & ~31;   // 8 lanes x 16 bytes per row and step
   plan->VS = VS;
   const int xtra = (weighted && pot == POT_TRUNC) ? K : 0;
   // lanes per worker: 8; 4 for the unweighted SGM kernels when the label vector is short (<= 128 labels), so that
   // a lane still owns 8 chunks and a band holds twice the workers for the same per-step overhead
   int want_groups = 1;   // row groups per band (RowGroup in the kernel); measured: no gain, the LSU is the shared limit
   if (!weighted && (tune.groups == 2 || tune.groups == 3)) want_groups = tune.groups;
   plan->lanes = MGM_AGG_GROUP;
   if (pot == POT_SGM && !weighted && shear && want_groups == 1 && VS <= 16 * MGM_AGG_CREG && !tune.no_creg &&
       !tune.no_fused_sgm && tune.lanes != 8) {
      // only when the image still yields enough bands of that size to fill the machine twice (measured: with
      // fewer, the longer hand-offs and the idle SMs cost more than the per-step overhead saved)
      const int t4 = (MGM_AGG_MAX_THREADS - 64) / 4;
      const long bands = 2L * ((ny + t4 - 1) / t4 + (nx + t4 - 1) / t4) + 4L * ((nx + ny + t4 - 2) / t4);
      if (bands >= 2L * num_sms || tune.lanes == 4) plan->lanes = 4;
   }
   // costs prefetched into registers when a lane's share fits (one cost buffer per row), else a cp.async ring of two
   plan->ncb = (VS / (4 * plan->lanes) <= MGM_AGG_CREG && !tune.no_creg) ? 1 : 2;
   // unweighted truncated-linear kernels with register-resident costs: min-convolution chains in registers by the
   // worker's own lanes (chain_regs), vectors in the padded layout (VSP floats per vector), smaller block
   plan->regchain = (!weighted && pot == POT_TRUNC && plan->ncb == 1 && plan->lanes == 8 && want_groups == 1 && tune.reg_chains) ? 1 : 0;
   const int tcap = ((plan->regchain ? MGM_AGG_MAX_THREADS_RC : MGM_AGG_MAX_THREADS) - 64) / plan->lanes;   // rows per CTA allowed by the thread budget
   plan->shear = shear;
   const int nvirt = shear ? 2 : 1;   // boundary workers kept per position
   // unweighted truncated-linear kernels with register-resident costs build the message in its ring slot and run
   // the min-convolution in place: no cost buffer
   // ... and the unweighted SGM kernels transform the message straight from registers (one group only)
   plan->fused_sgm = (!weighted && pot == POT_SGM && plan->ncb == 1 && want_groups == 1 && !tune.no_fused_sgm) ? 1 : 0;
   // ... through the lean kernels of aggregate_sgm.cu when the diagonal sweeps 4-7 run sheared
   // (TSGM <= 3) or row-per-worker (TSGM = 4), and the label layout is one they are built for
   plan->lean_trunc = (!weighted && pot == POT_TRUNC && plan->ncb == 1 && !plan->regchain && want_groups == 1 && plan->lanes == 8 && !knight &&
                       !tune.no_lean_trunc && (K == 4 || shear) && agg_trunc_lean_supported(VS)) ? 1 : 0;
   plan->lean_sgmw = (weighted && pot == POT_SGM && plan->ncb == 1 && want_groups == 1 && plan->lanes == 8 && !knight && !tune.no_creg &&
                      !tune.no_lean_sgm && agg_sgm_lean_supported(VS, 8)) ? 1 : 0;
   plan->lean_sgm = (plan->fused_sgm && !tune.no_lean_sgm && (K == 4 || shear) && agg_sgm_lean_supported(VS, plan->lanes)) ? 1 : 0;
   const int VSP = plan->regchain ? VS + (VS >> 3) : VS;
   plan->VSP = VSP;
   const int ncbuf = (!weighted && plan->ncb == 1 && (pot == POT_TRUNC || plan->fused_sgm)) ? 0 : plan->ncb;
   for (int cls = 0; cls < 3; ++cls) {
      int nbuf = ((cls == CLS_DIAG && shear) ? 2 : ring_slots(cls, K)) + ncbuf + xtra;
      int TS = nbuf * VSP;
      if ((plan->ncb == 1 && pot == POT_SGM) || plan->regchain || (plan->lean_trunc && MGM_TRUNC_LAY == 0)) TS += 2;   // 8 bytes modulo 128: 64-bit accesses of adjacent rows tile the banks
      else if (((TS >> 2) & 1) == 0) TS += 4;          // 16-byte aligned rows (truncated linear, cp.async mode): odd number of 16-byte units
      plan->TS[cls] = TS;
      size_t fixed = 1024 + (size_t)nvirt * RV * VSP * 4 + (size_t)tcap * (16 + 16 + 4) + RV * 16;
      long avail = (long)max_smem - (long)fixed;
      int Tc = avail > 0 ? (int)(avail / ((long)TS * 4)) : 0;
      if (Tc > tcap) Tc = tcap;
      {
         const int knob = (cls == CLS_AXIS) ? tune.rows_axis : tune.rows_diag;
         const int ov = knob > 0 ? knob : t_override;
         if (ov > 0 && Tc > ov) Tc = (cls == CLS_DIAG && shear && ov < 2) ? 2 : ov;
      }
      if (Tc < 1) Tc = 0;
      // groups need whole warps (4 rows) and at least two warps each for their chain pair
      int ng = want_groups;
      while (ng > 1 && Tc / (4 * ng) * 4 < 8) --ng;
      if (ng > 1) Tc = Tc / (4 * ng) * (4 * ng);
      plan->ng[cls] = ng;
      plan->T[cls] = Tc;
   }
   if (!knight) plan->T[CLS_KNIGHT] = min(plan->T[CLS_KNIGHT], max(plan->T[0], plan->T[1]));   // unused: keep max() below unchanged
   const int tm = max(max(plan->T[0], plan->T[1]), plan->T[2]);
   const int ncomp = (tm * plan->lanes + 31) & ~31;
   plan->block = ncomp + 64;   // + boundary-consumer warp + boundary-publisher warp
   // the lean truncated-linear kernels run the forward and backward halves of the min-convolution on two COMPUTE warps
   // (their service warps are off the step barriers): bands of up to 4 rows still get two compute warps
   if (plan->lean_trunc && ncomp < 64) plan->block = 64 + 64;
   // full_block: keep the largest block although the bands hold fewer rows -- the spare warps idle during the band steps
   // and work in the fused finish tiles (one warp per pixel)
   if (tune.full_block && !plan->regchain && plan->block < MGM_AGG_MAX_THREADS) plan->block = MGM_AGG_MAX_THREADS;
   size_t off = 0;
   plan->off_phase = off; off += (size_t)(tm + 1) * 4; off = (off + 15) & ~(size_t)15;
   plan->off_cbar = off; off += (size_t)tm * 16;
   plan->off_vbar = off; off += RV * 8;
   plan->off_ms = off; off += (size_t)tm * 16;
   plan->off_vms = off; off += (size_t)nvirt * RV * 4; off = (off + 127) & ~(size_t)127;
   plan->off_virt = off; off += (size_t)nvirt * RV * VSP * 4; off = (off + 127) & ~(size_t)127;
   plan->off_thr = off;
   // the rows region is sized for the classes this launch runs: the knight class (ring of 4 slots) only with more than
   // 8 sweeps -- shared memory the kernel does not need stays L1 (measured: 228 KB instead of 191 KB of shared memory
   // made the headline launch 13 % slower)
   size_t per_thr = 0;
   for (int cls = 0; cls < (knight ? 3 : 2); ++cls) per_thr = max(per_thr, (size_t)plan->TS[cls] * plan->T[cls] * 4);
   plan->smem = off + per_thr;
   plan->num_sms = num_sms;
}

void agg_plan(AggPlan *plan, int nx, int ny, int L, int K, int pot, bool weighted, int max_smem, int num_sms,
              int t_override, bool knight, const AggTuning &tune) {
   // diagonal sweeps as sheared wavefronts (run_band_shear): predecessors in the row above only, no image-dependent
   // weights, and room for at least two workers per band (the hand-off carries the last two)
   const bool want_shear = (K <= 3 && !weighted && !tune.no_shear);
   if (want_shear) {
      agg_plan_try(plan, nx, ny, L, K, pot, weighted, max_smem, num_sms, t_override, 1, knight, tune);
      if (plan->T[0] >= 1 && plan->T[1] >= 2 && plan->T[2] >= 1) return;
   }
   agg_plan_try(plan, nx, ny, L, K, pot, weighted, max_smem, num_sms, t_override, 0, knight, tune);
}

int agg_sweep_class(const AggPlan &plan, int pass) {
   (void)plan;
   return pass >= 8 ? CLS_KNIGHT : (pass >= 4 ? CLS_DIAG : CLS_AXIS);
}

void agg_sweep_bands(const AggPlan &plan, int pass, int nx, int ny, int *nb, size_t *bnd_floats, size_t *bndm_floats) {
   const PassGeom g = pass_geometry(pass, nx, ny);
   const int cls = agg_sweep_class(plan, pass);
   const int T = plan.T[cls];
   if (cls == CLS_DIAG && plan.shear) {
      // sheared wavefront: bands of T anti-diagonals, two boundary lines of maxjj positions per band
      *nb = (g.maxii + g.maxjj - 1 + T - 1) / T;
      *bnd_floats = (size_t)*nb * 2 * g.maxjj * plan.VSP;
      *bndm_floats = (size_t)*nb * 2 * g.maxjj;
   } else {
      *nb = (g.maxjj + T - 1) / T;
      *bnd_floats = (size_t)*nb * g.maxii * plan.VSP;
      *bndm_floats = (size_t)*nb * g.maxii;
   }
}

cudaError_t agg_launch(const AggParams &P, const AggPlan &plan, int pot, int K, bool weighted, cudaStream_t st) {
   if (pot == POT_SGM && !weighted && plan.lean_sgm && plan.ng[0] == 1 && plan.ng[1] == 1)
      return agg_launch_sgm_lean(P, plan, K, st);
   if (pot == POT_SGM && weighted && plan.lean_sgmw && plan.ng[0] == 1 && plan.ng[1] == 1) return agg_launch_sgmw_lean(P, plan, K, st);
   if (pot == POT_TRUNC && !weighted && plan.lean_trunc && plan.ng[0] == 1 && plan.ng[1] == 1)
      return agg_launch_trunc_lean(P, plan, K, st);
   return pot == POT_SGM ? agg_launch_generic_sgm(P, plan, K, weighted, st) : agg_launch_generic_trunc(P, plan, K, weighted, st);
}

}  // namespace mgm
