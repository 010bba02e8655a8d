// K4 -- MGM aggregation sweeps (replaces mgm_core.cc:489-580 with the message
// updates of :66-281) as ONE persistent sm_100a kernel running all requested
// sweeps concurrently.
//
// Formulation (checked against the reference by
// oracle/mgm_oracle.c): in the scan space of a sweep every pixel (xs,ys) reads the
// predecessors {(-1,0),(0,-1),(-1,-1),(+1,-1)}; sweeps 0-3 take them in that order,
// sweeps 4-7 in the order {(+1,-1),(-1,-1),(0,-1),(-1,0)}, TSGM=K keeps the first K.
// Pixels with xs==0, ys==0 or xs==maxii-1 keep their matching cost (mgm_core.cc:538-541).
//
// Mapping to the machine (DESIGN.md section 4.1 has the measurements behind each choice):
//   * work unit = a BAND of T consecutive workers of one sweep, run by one CTA in lock step, one label vector
//     per worker and step.  Axis sweeps 0-3 (and every sweep with TSGM=4 or image-dependent weights): worker =
//     scan row, row t trails row t-1 by SIGMA pixels (run_band).  Diagonal sweeps 4-7 with TSGM<=3: sheared
//     wavefront, worker = anti-diagonal, no lag inside a band (run_band_shear);
//   * 8 lanes per worker own interleaved 16-byte chunks of the label vector: phase 1 (gather) builds the message
//     from the matching costs (streamed into registers one step ahead) and the predecessors' transformed
//     vectors in a shared-memory ring, writes it to the sweep's volume and into the worker's ring slot;
//   * phase 2 transforms the message IN PLACE into what its consumers add (SGM min3, or the exact sequential
//     truncated-linear min-convolution by one lane pair per worker, minus its minimum) -- computed ONCE by the
//     producer instead of once per consumer;
//   * bands are chained through global boundary lines (publisher-warp copy -> release counter -> acquire ->
//     TMA bulk load into a virtual-row ring, mbarrier completion) and claimed dynamically, in order within a
//     sweep, ready axis bands first (claim_band), so any grid size is deadlock free.
// Precondition of this fast path (enforced by the host): every cost
// vector holds a finite value, no NaN / -INF, P1,P2 >= 0 and weights >= 0; then the
// hardware min is bit-identical to the reference's compare-select forms.
#include "aggregate_dev.cuh"

namespace mgm {


// Row groups.  The rows of a band are split into NG <= 3 groups of TG rows that run the two phases of a step on
// their OWN named barrier, out of phase with each other: while one group gathers (LSU-bound) the others run their
// min-convolution chains (latency-bound).  Adjacent groups are ordered by two producer/consumer named barriers
// (bar.arrive by the producer group, bar.sync by the consumer group -- no polling):
//   token: group g starts the gather of step s after group g-1 has finished ITS gather of step s (this also
//          covers the read-after-write on the rows of group g-1 and serialises the gathers on the LSU);
//   reuse: group g starts the gather of step s after group g+1 has finished the gather of step s-1 (the ring slot
//          it is about to overwrite was last read there).
// Each barrier sees exactly one arrival and one sync per step in alternation (the token of step s+1 can only be
// sent after the reuse barrier of step s+1 was passed, which needs the receiver's gather of step s, hence its
// token sync of step s), so phases never mix.  Barrier ids: 0 CTA, 1..3 groups, 4..9 chain pairs, 10..11 token,
// 12..13 reuse.
struct RowGroup {
   int gi, ng, bar_id, cnt, cnt_prev, cnt_next, first_tid;
   __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(cnt) : "memory"); }
   // top of step: wait for the token of the group above and for the reuse signal of the group below
   __device__ __forceinline__ void begin_step(bool first) const {
      if (ng > 1) {
         if (gi > 0) asm volatile("bar.sync %0, %1;" ::"r"(10 + gi - 1), "r"(cnt_prev + cnt) : "memory");
         if (gi < ng - 1 && !first) asm volatile("bar.sync %0, %1;" ::"r"(12 + gi), "r"(cnt + cnt_next) : "memory");
      }
   }
   // my part of the gather of this step is finished
   __device__ __forceinline__ void gather_done(bool last) const {
      if (ng > 1) {
         __threadfence_block();
         if (gi < ng - 1) asm volatile("bar.arrive %0, %1;" ::"r"(10 + gi), "r"(cnt + cnt_next) : "memory");
         if (gi > 0 && !last) asm volatile("bar.arrive %0, %1;" ::"r"(12 + gi - 1), "r"(cnt_prev + cnt) : "memory");
      }
   }
};
__device__ __forceinline__ RowGroup make_row_group(int tid, int ncomp, int r, int NG, int TG, int G) {
   RowGroup g;
   g.ng = NG;
   g.gi = (tid < ncomp) ? min(r / TG, NG - 1) : (tid < ncomp + 32 ? 0 : NG - 1);   // consumer warp -> 0, publisher -> last
   auto count = [&](int i) { return (i < NG - 1 ? TG * G : ncomp - (NG - 1) * TG * G) + (i == 0 ? 32 : 0) + (i == NG - 1 ? 32 : 0); };
   g.bar_id = 1 + g.gi;
   g.cnt = count(g.gi);
   g.cnt_prev = g.gi > 0 ? count(g.gi - 1) : 0;
   g.cnt_next = g.gi < NG - 1 ? count(g.gi + 1) : 0;
   g.first_tid = g.gi * TG * G;
   return g;
}

// Register-resident min-convolution by the 8 lanes of a worker (unweighted truncated-linear kernels, register-chain
// mode).  Lane gl owns the CONTIGUOUS labels of chunks gl*nj .. gl*nj+nj-1 (padded layout LAY 2: conflict-free).  The
// forward recurrence F[o] = min(F[o-1]+c, M[o]) and the backward recurrence B on the ORIGINAL values (identity 2:
// the reference's backward pass over F equals min(F, B)) run as segmented scans:
//   1. every lane scans its own labels from +INF (local F, local B);
//   2. the value at the end of a lane's segment hops to the next lane, which adds c once per label of its segment --
//      4*nj separately rounded additions, the reference's own sequence -- and takes the minimum with its local end
//      value; 7 hops per direction, both directions interleaved in one instruction stream;
//   3. every lane folds its carry into its labels: F[o] = min(local F[o], carry + c + ... + c).
// x -> RN(x+c) is monotone, so min distributes over it: each F[o] is the minimum over the same candidates, each built
// by the same chain of rounded additions, as in the sequential loop (mgm_core.cc:152-163).  Result written in place:
// min(F, B, cap) - sub.  No barrier: a worker's 8 lanes sit in one warp (the caller's __syncwarp orders the slot).
// __noinline__ on purpose: inlined into the band loops the compiler schedules the surrounding loads and address
// arithmetic across the chain and spills 2.5 KB per thread; as a call it is a scheduling barrier (87 registers of its own).
template <int NJR>
__device__ __noinline__ void chain_regs(float2 *slot, const int nj, const int gl, const unsigned gmask, const float c,
                                           const float cap, const float sub) {
   float4 f[NJR], b[NJR];
   const int q0 = gl * nj;
#pragma unroll
   for (int j = 0; j < NJR; ++j)
      if (j < nj) f[j] = b[j] = ld16<2>(slot, q0 + j);
   float runf = MGM_INF, runb = MGM_INF;
#pragma unroll
   for (int j = 0; j < NJR; ++j) {   // local scans: forward ascending, backward descending (independent: they interleave)
      if (j < nj) chain4(runf, f[j].x, f[j].y, f[j].z, f[j].w, c);
      const int jb = NJR - 1 - j;
      if (jb < nj) chain4(runb, b[jb].w, b[jb].z, b[jb].y, b[jb].x, c);
   }
   const float lastf = runf, lastb = runb;
   float ef = lastf, eb = lastb, cf = MGM_INF, cb = MGM_INF;
   const int nadd = 4 * nj;
#pragma unroll
   for (int r = 1; r < 8; ++r) {
      const float tf = __shfl_up_sync(gmask, ef, 1, 8);      // end value of the segment below
      const float tb = __shfl_down_sync(gmask, eb, 1, 8);    // start value of the segment above
      float uf = tf, ub = tb;
      for (int k = 0; k < nadd; k += 4) {
         uf = ((((uf + c) + c) + c) + c);
         ub = ((((ub + c) + c) + c) + c);
      }
      if (gl == r) { cf = tf; ef = fminf(lastf, uf); }
      if (gl == 7 - r) { cb = tb; eb = fminf(lastb, ub); }
   }
#pragma unroll
   for (int j = 0; j < NJR; ++j) {   // fold the carries in
      if (j < nj) {
         cf += c; f[j].x = fminf(f[j].x, cf);
         cf += c; f[j].y = fminf(f[j].y, cf);
         cf += c; f[j].z = fminf(f[j].z, cf);
         cf += c; f[j].w = fminf(f[j].w, cf);
      }
      const int jb = NJR - 1 - j;
      if (jb < nj) {
         cb += c; b[jb].w = fminf(b[jb].w, cb);
         cb += c; b[jb].z = fminf(b[jb].z, cb);
         cb += c; b[jb].y = fminf(b[jb].y, cb);
         cb += c; b[jb].x = fminf(b[jb].x, cb);
      }
   }
#pragma unroll
   for (int j = 0; j < NJR; ++j) {
      if (j < nj) {
         const float4 o = make_float4(fminf(fminf(f[j].x, b[j].x), cap), fminf(fminf(f[j].y, b[j].y), cap),
                                      fminf(fminf(f[j].z, b[j].z), cap), fminf(fminf(f[j].w, b[j].w), cap));
         st16<2>(slot, q0 + j, add4s(o, -sub));
      }
   }
}



// KN: sweeps 8-15 -- the scan of sweep pass-8, lag 2, predecessor ORDER by the parity of the scan coordinates
// (knight_pred_type, common.cuh); instantiated with DIAG = true for the lag and the ring depth.
// RC: register-chain mode of the unweighted truncated-linear kernels (chain_regs): padded slots, costs in registers,
// gather -> __syncwarp -> chain by the same 8 lanes -> ONE block barrier per step.
template <int POT, int K, bool WEIGHTED, bool DIAG, int GL, bool KN = false, bool RC = false>
__device__ void run_band(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem) {
   constexpr int G = GL;   // lanes per worker: 8, or 4 for short label vectors (unweighted SGM kernels, agg_plan)
   static_assert(!KN || DIAG, "knight sweeps use the lag-2 layout");
   static_assert(!RC || (!WEIGHTED && POT == POT_TRUNC && GL == 8), "register chains: unweighted truncated linear, 8 lanes");
   const int pass = D.pass;
   const bool kn_diag = (pass & 7) >= 4;
   constexpr int SIG = (DIAG || K == 4) ? 2 : 1;
   constexpr int R = SIG + 2;
   constexpr bool NEEDM = WEIGHTED || (POT == POT_TRUNC && K == 2);
   constexpr bool WTRUNC = WEIGHTED && POT == POT_TRUNC;
   constexpr bool CHAINS = !WEIGHTED && POT == POT_TRUNC;   // phase-2 transform done by lane pairs
   constexpr int CLS = KN ? CLS_KNIGHT : (DIAG ? CLS_DIAG : CLS_AXIS);
   constexpr int A16 = RC ? 2 : (POT == POT_TRUNC ? 1 : 0);   // shared-memory vector layout (ld16 / st16)
   constexpr int JB = (K <= 3) ? MGM_JB : 2;   // chunks per lane whose loads are issued together in the gather
   constexpr int NJR = MGM_AGG_CREG;      // cost chunks per lane that can be prefetched into registers

   const PassGeom g = pass_geometry(pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int T = P.T[CLS];
   const int TS = P.TS[CLS];
   const int VS = P.VS;
   const int VSP = RC ? VS + (VS >> 3) : VS;   // floats per vector in shared memory and in the boundary lines
   const int nq = VS >> 2;
#ifdef MGM_FORCE_NJ
   constexpr int nj = MGM_FORCE_NJ;   // experiment: compile-time chunks per lane
#else
   const int nj = nq / G;   // chunks per lane (VS is a multiple of 4*G)
#endif
   const int ncb = RC ? 1 : P.ncb;   // cost buffers per row: 1 = costs prefetched into registers (nj <= NJR), 2 = cp.async ring
   const bool creg_mode = (ncb == 1);
   const uint32_t vbytes = (uint32_t)VSP * 4u;
   const int tid = threadIdx.x;
   const int ncomp = blockDim.x - 64;   // two service warps follow the row threads
   const bool is_prod = (tid == ncomp);        // fetches the previous band's boundary row
   const bool pub_warp = (tid >= ncomp + 32);  // this warp stores and publishes the band's boundary row
   const int row0 = band * T;
   const int nrows = min(T, maxjj - row0);
   const bool has_prev = band > 0;
   const bool has_next = row0 + T < maxjj;
   const int nsteps = maxii + SIG * (nrows - 1);

   // group-per-row mapping (gather, label-parallel transforms)
   const int r = tid / G, gl = tid % G;
   const unsigned gmask = ((1u << G) - 1u) << ((tid & 31) & ~(G - 1));
   const bool rowok = (tid < ncomp) && (r < nrows);
   const int ys = row0 + r;
   // lane-pair mapping for the sequential min-convolution chains: pair p of a warp is lanes (p, p+16); the
   // forward halves of 8 consecutive rows then sit in one quarter-warp and hit 8 different bank groups
   // (the row stride is an odd number of 16-byte units), likewise the backward halves
   // (superseded mapping kept simple) chain work is distributed per WARP: warps [0,NCW) run the upward halves of
   // vectors 32w..32w+31, warps [NCW,2NCW) the downward halves of the same vectors; 8 consecutive rows in a
   // quarter-warp hit 8 different bank groups because the row stride is an odd number of 16-byte units
   const int warp_id = tid >> 5, lane_id = tid & 31;

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);   // persistent mbarrier parities

   // row groups (RowGroup above); the boundary consumer warp belongs to group 0, the publisher warp to the last
   const int NG = P.ng[CLS], TG = T / NG;
   const RowGroup grp = make_row_group(tid, ncomp, r, NG, TG, G);
   const int gw0 = grp.first_tid >> 5;          // first warp of my group
   const int ncw = (TG + 31) >> 5;              // chain warps per direction and group (2 only when NG == 1)
   const int pair_id0 = 4 + grp.gi * 2;         // named barriers of my group's chain pairs

   // per-row shared memory: ring slots [R][VS] | cost buffers [ncbuf][VS] | scratch [K][VS] (weighted trunc)
   auto row_base = [&](int rr) -> float * { return thr + (size_t)rr * TS; };
   auto slot_of = [&](int rr, int px) -> float * {
      return (rr < 0) ? virt + (px & (RV - 1)) * VSP : row_base(rr) + (px % R) * VSP;
   };
   auto m_of = [&](int rr, int px) -> float { return (rr < 0) ? vms[px & (RV - 1)] : msr[rr * 4 + (px % R)]; };
   // predecessor k of pixel (xs) of row rr: which row's slot, which pixel; returns the predecessor's type
   auto pred_of = [&](int rr, int xs, int k, int &prow, int &ppx) -> int {
      const int pt = KN ? knight_pred_type(kn_diag, k, xs, row0 + rr) : pred_type<DIAG>(k);
      ppx = (pt == PRED_UP) ? xs : (pt == PRED_UPR ? xs + 1 : xs - 1);
      prow = (pt == PRED_SAME) ? rr : rr - 1;   // -1 = virtual row fed from the previous band
      return pt;
   };
   auto plane_of = [&](int k, int pt) -> int {
      return KN ? pass_weight_plane_of_type(pass & 7, pt) : pass_weight_plane(pass, k);
   };

   const float *ccv = D.cc;
   float *bnd_out = D.bnd + (size_t)band * maxii * VSP;
   float *bndm_out = D.bndm + (size_t)band * maxii;
   const float *bnd_in = has_prev ? D.bnd + (size_t)(band - 1) * maxii * VSP : nullptr;
   const float *bndm_in = has_prev ? D.bndm + (size_t)(band - 1) * maxii : nullptr;
   int *prog_out = D.progress + band;
   const int *prog_in = has_prev ? D.progress + band - 1 : nullptr;
   const size_t wplane = (size_t)P.nx * P.ny;

   const long long pix0 = g.base0 + (long long)ys * g.dys;   // pixel of (0, ys) for my group's row

   uint32_t vph = 0;
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);
   if (tid < G) vph = phase[vph_idx];
   int next_px = 0;   // producer: next boundary pixel to fetch

   // Cost vectors: every lane prefetches, one pixel ahead, exactly the 16-byte chunks it will read itself (128-byte
   // coalesced per row).  Up to NJR chunks per lane (256 labels) they go straight into REGISTERS with streaming
   // loads issued at the end of the gather and in flight during phase 2: no shared-memory write + read-back, no
   // cp.async issue cost, one cost buffer per row instead of two (more rows per band).  Longer label vectors use
   // cp.async (LDGSTS) into a two-deep ring: pixel px lands in buffer px&1.
   float4 creg[NJR];
#pragma unroll
   for (int j = 0; j < NJR; ++j) creg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
   // CHAINS with register-resident costs: the message is written straight into its ring slot and the
   // min-convolution runs IN PLACE there (its second half no longer reads the source) -> no cost buffer at all
   // unweighted SGM kernels in register mode: gather and transform fused, the message stays in creg[] (see
   // sgm_transform_regs); one barrier per step, no cost buffer either
   const bool fused = !WEIGHTED && POT == POT_SGM && P.fused_sgm != 0;
   const int ncbuf = ((CHAINS && creg_mode) || fused) ? 0 : ncb;
   auto cbuf_of = [&](int rr, int px) -> float * { return row_base(rr) + (R + (ncb == 2 ? (px & 1) : 0)) * VSP; };
   auto prefetch_cost = [&](int s) {
      const int px = s - SIG * r + 1;   // the pixel my row handles in the NEXT step
      const bool go = rowok && px >= 0 && px < maxii;
      if (P.cc_pf > 0) {   // short steps (SGM potentials): one step does not cover the HBM latency, warm the L2 further ahead
         const int pp = px + P.cc_pf;
         if (rowok && pp >= 0 && pp < maxii) {
            const float *line = ccv + (size_t)(pix0 + (long long)pp * g.dxs) * VS;
            for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + l * 32));
         }
      }
      if (creg_mode) {
         if (go) {
            const float4 *src = reinterpret_cast<const float4 *>(ccv + (size_t)(pix0 + (long long)px * g.dxs) * VS);
#pragma unroll
            for (int j = 0; j < NJR; ++j)
               if (j < nj) creg[j] = (MGM_EXP == 2) ? make_float4((float)j, 1.f, 2.f, 3.f) : (RC ? __ldg(src + gl + G * j) : __ldcs(src + gl + G * j));
         }
         return;
      }
      if (go) {
         const float4 *src = reinterpret_cast<const float4 *>(ccv + (size_t)(pix0 + (long long)px * g.dxs) * VS);
         float *dst = cbuf_of(r, px);   // 16-byte aligned in this mode (agg_plan)
         for (int j = 0; j < nj; ++j) cp_async16(dst + 4 * (gl + G * j), src + gl + G * j);
      }
      cp_async_commit();   // one (possibly empty) group per step keeps the wait count uniform
   };
   auto prefetch_cost_l1 = [&](int s) {
      const int px = s - SIG * r + 1;
      if (rowok && px >= 0 && px < maxii) {
         const float *line = ccv + (size_t)(pix0 + (long long)px * g.dxs) * VS;
         for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L1 [%0];" ::"l"(line + l * 32));
      }
   };
   prefetch_cost(-1);   // pixel 0 of row 0 (the other rows start later)
   // cp.async mode: lanes of warps that run min-convolution chains in phase 2 issue their prefetch at the top of
   // the step; all other warps are idle in phase 2 and issue it there (the LSU is less busy then).
   // Register mode: every lane issues its loads right after its gather (they complete during phase 2).
   const bool late_prefetch = !creg_mode && CHAINS && (warp_id - gw0 >= 2 * ncw);

   // optional phase timing (option "dbg"): thread 0 of the block adds the clock cycles of each phase of its steps
   const bool timing = P.dbg != nullptr && tid == 0;
   long long tacc[6] = {0, 0, 0, 0, 0, 0}, tl = 0;
   auto tick = [&](int i) { if (timing) { const long long t = clock64(); tacc[i] += t - tl; tl = t; } };
   if (timing) tl = clock64();
   for (int s = 0; s < nsteps; ++s) {
      if (!late_prefetch && !creg_mode) prefetch_cost(s);
      if (is_prod && has_prev) {
         // boundary pixels up to s+2 are needed before the next step (blocking); up to s+1+PF are fetched
         // ahead of time when the previous band has already published them (non-blocking)
         const int need = min(maxii - 1, s + 2);
         const int lim = min(maxii - 1, s + 1 + PF);
         int avail = ld_acquire(prog_in);
         while (next_px <= lim) {
            if (avail < next_px + 1) {
               if (next_px > need) break;
               do { __nanosleep(20); avail = ld_acquire(prog_in); } while (avail < next_px + 1);
            }
            fence_proxy_async();
            const int slot = next_px & (RV - 1);
            if (NEEDM) vms[slot] = __ldcg(bndm_in + next_px);
            mbar_expect_tx(&vbar[slot], vbytes);
            tma_load_1d(virt + slot * VSP, bnd_in + (size_t)next_px * VSP, vbytes, &vbar[slot]);
            ++next_px;
         }
      }
      // virtual row: group 0 waits one step AHEAD for the boundary pixels (xs+2 is first read at step xs+1),
      // so the end-of-step barrier publishes them to every thread that reads row -1 in the next step
      if (tid < G && nrows > 0 && has_prev && s < maxii) {
         if (s == 0) {
            for (int px = 0; px <= min(2, maxii - 1); ++px) {
               mbar_wait(&vbar[px & (RV - 1)], (vph >> (px & (RV - 1))) & 1u);
               vph ^= 1u << (px & (RV - 1));
            }
         } else if (s + 2 < maxii) {
            const int sl = (s + 2) & (RV - 1);
            mbar_wait(&vbar[sl], (vph >> sl) & 1u);
            vph ^= 1u << sl;
         }
      }
      grp.begin_step(s == 0);   // group ordering (RowGroup): token from the group above, slot reuse below
      tick(0);
      const int xs = s - SIG * r;
      const bool act = rowok && xs >= 0 && xs < maxii;
      const long long pix = pix0 + (long long)xs * g.dxs;
      float2 *cur = reinterpret_cast<float2 *>(row_base(r) + (act ? (xs % R) : 0) * VSP);
      float *Cbf = (CHAINS && creg_mode) ? reinterpret_cast<float *>(cur) : cbuf_of(r, xs);   // where the message is built
      float2 *Cb = reinterpret_cast<float2 *>(Cbf);
#ifdef MGM_Y_FROM_PIX
      float4 *gout = reinterpret_cast<float4 *>(ldir_of_pix(P, D, act ? pix : 0) + (size_t)(act ? pix : 0) * VS);
#else
      float4 *gout = reinterpret_cast<float4 *>(ldir_of_row(P, D, act ? g.y0 + xs * g.ydxs + ys * g.ydys : 0) +
                                                (size_t)(act ? pix : 0) * VS);
#endif
      const bool border = (xs == 0) || (ys == 0) || (xs == maxii - 1);
      float m = MGM_INF;

      // ---------------- phase 0 (weighted truncated-linear only): one lane pair per (row, neighbour)
      if constexpr (WTRUNC) {
         const int ncw0 = (K * nrows + 31) >> 5;   // warps per direction (weighted kernels run as one group)
         if (warp_id < 2 * ncw0) {
            const int cw = warp_id % ncw0, cdir = warp_id / ncw0;
            const int pidx = cw * 32 + lane_id;
            const int crow = pidx / K, ck = pidx % K;
            const int cxs = s - SIG * crow;
            const bool on = pidx < K * nrows && cxs > 0 && cxs < maxii - 1 && (row0 + crow) != 0;
            const float *src = thr;
            float *dst = thr;
            float cw1 = 0.f, capv = 0.f, mk = 0.f;
            int mlo = 0, mhi = VS;
            if (on) {
               int prow, ppx;
               const int cpt = pred_of(crow, cxs, ck, prow, ppx);
               const long long cpix = g.base0 + (long long)(row0 + crow) * g.dys + (long long)cxs * g.dxs;
               const float wk = D.w ? __ldg(D.w + (size_t)plane_of(ck, cpt) * wplane + cpix) : 1.0f;
               if (D.win_lo) {   // per-pixel cost ranges: convolve inside the receiving pixel's range
                  mlo = (int)__ldg(D.win_lo + cpix) - P.win_emin;
                  mhi = (int)__ldg(D.win_hi + cpix) - P.win_emin;
               }
               mk = m_of(prow, ppx);
               cw1 = P.P1 * wk; capv = mk + P.P2 * wk;
               src = slot_of(prow, ppx);
               dst = row_base(crow) + (R + ncbuf + ck) * VS;
            }
            if (cdir == 0) minconv_half<0, true>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, cw1, capv, mk, 4 + cw, mlo, mhi);
            else minconv_half<1, true>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, cw1, capv, mk, 4 + cw, mlo, mhi);
         }
         grp.sync();
      }

      // ---------------- phase 1: gather the message of pixel (xs,ys), G lanes per row
      if (act) {
         if (!creg_mode) { if (late_prefetch) cp_async_wait<0>(); else cp_async_wait<1>(); }   // all but a prefetch issued this step

         if (border) {
            if (creg_mode) {
#pragma unroll
               for (int j = 0; j < NJR; ++j) {
                  if (j < nj) {
                     const int q = gl + G * j;
                     m = hmin4(m, creg[j]);
                     if (!fused) st16<A16>(Cb, q, creg[j]);
                     __stcs(gout + q, creg[j]);
                  }
               }
            } else {
               const float2 *Cin = reinterpret_cast<const float2 *>(cbuf_of(r, xs));
               for (int j = 0; j < nj; ++j) {
                  const int q = gl + G * j;
                  const float4 c = ld16<A16>(Cin, q);
                  m = hmin4(m, c);
                  if (CHAINS) st16<A16>(Cb, q, c);
                  __stcs(gout + q, c);
               }
            }
         } else {
            const float2 *Cin = reinterpret_cast<const float2 *>(cbuf_of(r, xs));   // costs (cp.async mode)
            const float2 *S[K];
            float mk[K], wk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
               int prow, ppx;
               const int pt = pred_of(r, xs, k, prow, ppx);
               S[k] = reinterpret_cast<const float2 *>(WTRUNC ? row_base(r) + (R + ncbuf + k) * VS : slot_of(prow, ppx));
               mk[k] = 0.f; wk[k] = 1.f;
               if (NEEDM) mk[k] = m_of(prow, ppx);
               if (WEIGHTED && !WTRUNC) wk[k] = __ldg(D.w + (size_t)plane_of(k, pt) * wplane + pix);
            }
            if constexpr (!WEIGHTED || WTRUNC) {
               // slots hold the neighbour-side transform already (producer row, or phase 0 scratch).
               // Straight-line batches: all loads of a batch are issued before the first use.
               auto batch = [&](auto jbc, int j0, auto regc) {
                  constexpr int B = decltype(jbc)::value;
                  constexpr bool REGC = decltype(regc)::value;   // costs come from creg[j0 .. j0+B) (j0 compile-time)
                  float4 c[B], a[K][B];
#pragma unroll
                  for (int jj = 0; jj < B; ++jj) {
                     const int q = gl + G * (j0 + jj);
                     if constexpr (!REGC) c[jj] = ld16<A16>(Cin, q);
#pragma unroll
                     for (int k = 0; k < K; ++k) a[k][jj] = ld16<A16>(S[k], q);
                  }
                  if constexpr (REGC) {
#pragma unroll
                     for (int jj = 0; jj < B; ++jj) c[jj] = creg[(j0 + jj) < NJR ? (j0 + jj) : 0];
                  }
#pragma unroll
                  for (int jj = 0; jj < B; ++jj) {
                     const int q = gl + G * (j0 + jj);
                     float4 o;
                     if constexpr (POT == POT_TRUNC && K == 2 && !WEIGHTED) {   // update_cost2_trunclinear :216
                        // c + (((a0 - m0) + a1) - m1) / 2
                        o = add4(c[jj], mul4s(add4s(add4(add4s(a[0][jj], -mk[0]), a[1 % K][jj]), -mk[1 % K]), 0.5f));
                     } else if constexpr (POT == POT_SGM && K == 2) {   // update_cost2: halves taken by the producer
                        o = add4(c[jj], add4(a[0][jj], a[1 % K][jj]));
                     } else if (MGM_EXP == 3) {
                        o = c[jj]; o.x += a[0][jj].x; o.y += a[1 % K][jj].y; o.z += a[2 % K][jj].z;
                     } else {
                        float4 e = a[0][jj];
#pragma unroll
                        for (int k = 1; k < K; ++k) e = add4(e, a[k][jj]);
                        o = add4(c[jj], div4_by_k<K>(e));
                     }
                     m = hmin4(m, o);
                     if (REGC && fused) creg[(j0 + jj) < NJR ? (j0 + jj) : 0] = o;   // transformed from registers below
                     else st16<A16>(Cb, q, o);
#if MGM_EXP != 1
                     __stcs(gout + q, o);
#endif
                  }
               };
               if (creg_mode) {
                  // register-resident costs: batch offsets are compile-time so creg[] stays in registers
                  constexpr int JR = (JB < NJR) ? JB : NJR;
#pragma unroll
                  for (int jb = 0; jb < NJR; jb += JR) {
                     if (jb + JR <= nj) batch(std::integral_constant<int, JR>{}, jb, std::true_type{});
                     else {
#pragma unroll
                        for (int jr = 0; jr < JR; ++jr)
                           if (jb + jr < nj) batch(std::integral_constant<int, 1>{}, jb + jr, std::true_type{});
                     }
                  }
               } else {
                  int j0 = 0;
                  for (; j0 + JB <= nj; j0 += JB) batch(std::integral_constant<int, JB>{}, j0, std::false_type{});
                  for (; j0 + 2 <= nj; j0 += 2) batch(std::integral_constant<int, 2>{}, j0, std::false_type{});
                  for (; j0 < nj; ++j0) batch(std::integral_constant<int, 1>{}, j0, std::false_type{});
               }
            } else {
               // update_costW with per-edge weights (mgm_core.cc:95-144); slots hold the raw messages
               float pw[K], cap[K];
#pragma unroll
               for (int k = 0; k < K; ++k) { pw[k] = P.P1 * wk[k]; cap[k] = mk[k] + P.P2 * wk[k]; }
               if (creg_mode) {   // this path reads the costs from the buffer: park the registers there first
#pragma unroll
                  for (int j = 0; j < NJR; ++j)
                     if (j < nj) st16<A16>(Cb, gl + G * j, creg[j]);
               }
               for (int j = 0; j < nj; ++j) {
                  const int q = gl + G * j;
                  const float4 c = ld16<A16>(Cb, q);
                  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int k = 0; k < K; ++k) {
                     const float *sf = reinterpret_cast<const float *>(S[k]);
                     const float4 v = ld16<A16>(S[k], q);
                     const float lft = (q > 0) ? sf[4 * q - 1] : MGM_INF;
                     const float rgt = (q + 1 < nq) ? sf[4 * q + 4] : MGM_INF;
                     e.x += sgm_x(lft, v.x, v.y, pw[k], cap[k], mk[k]);
                     e.y += sgm_x(v.x, v.y, v.z, pw[k], cap[k], mk[k]);
                     e.z += sgm_x(v.y, v.z, v.w, pw[k], cap[k], mk[k]);
                     e.w += sgm_x(v.z, v.w, rgt, pw[k], cap[k], mk[k]);
                  }
                  const float4 o = add4(c, div4_by_k<K>(e));
                  m = hmin4(m, o);
                  st16<A16>(Cb, q, o);
                  __stcs(gout + q, o);
               }
            }
         }
#pragma unroll
         for (int d = 1; d < G; d <<= 1) m = fminf(m, __shfl_xor_sync(gmask, m, d));
         if (gl == 0) msr[r * 4 + (xs % R)] = m;
         if constexpr (!WEIGHTED && POT == POT_SGM) {
            if (fused) sgm_transform_regs<K, NJR, A16, G>(creg, nj, nq, gl, gmask, m, P.P1, P.P2, cur);
         }
      }
      if (creg_mode && !RC) prefetch_cost(s);   // costs of the next pixel -> registers, in flight during phase 2
      if constexpr (RC) {
         // phase 2 by the worker's own 8 lanes, from registers (chain_regs): the slot only has to be ordered within the warp.
         // The chain needs the registers the prefetched costs would occupy: the next pixel's costs are pulled into L1
         // before the chain and loaded after it (an L1 hit by then).
         tick(1);
         prefetch_cost_l1(s);
         __syncwarp();
         tick(2);
         if (act) chain_regs<NJR>(cur, nj, gl, gmask, P.P1, m + P.P2, (K == 2) ? 0.0f : m);
         tick(3);
         prefetch_cost(s);
      } else if (!fused) {
         tick(1);
         grp.gather_done(s == nsteps - 1);
         grp.sync();
         tick(2);
      } else { tick(1); tick(2); }

      // ---------------- phase 2: build the neighbour-side transform of the message in the ring slot
      if (late_prefetch) prefetch_cost(s);
      if constexpr (RC) {
      } else if constexpr (CHAINS) {
         // minConvTruncatedLinear of the finished message, one lane pair per row (the first 2*ncw warps of the group)
         const int wg = warp_id - gw0;
         if (tid < ncomp + 64 && wg >= 0 && wg < 2 * ncw) {
            const int cw = wg % ncw, cdir = wg / ncw;
            const int crow = grp.gi * TG + cw * 32 + lane_id;
            const int cxs = s - SIG * crow;
            const bool on = cw * 32 + lane_id < TG && crow < nrows && cxs >= 0 && cxs < maxii;
            const float cm = on ? msr[crow * 4 + (cxs % R)] : 0.f;
            float *dst = on ? row_base(crow) + (cxs % R) * VS : thr;
            const float *src = on ? (creg_mode ? dst : cbuf_of(crow, cxs)) : thr;
            if (MGM_EXP == 4) {}
            else if (cdir == 0) minconv_half<0>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, P.P1, cm + P.P2, (K == 2) ? 0.0f : cm, pair_id0 + cw);
            else minconv_half<1>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, P.P1, cm + P.P2, (K == 2) ? 0.0f : cm, pair_id0 + cw);
         }
      } else if (act && !fused) {
         if constexpr (!WEIGHTED) {
            // SGM transform, label-parallel: A(o) = min3(L(o), min(L(o-1),L(o+1))+P1, m+P2) - m  [x 1/2 for K=2]
            const float p1 = P.P1, cap = m + P.P2;
            const float sc = (K == 2) ? 0.5f : 1.0f;
            auto tbatch = [&](auto jbc, int j0) {
               constexpr int B = decltype(jbc)::value;
               float4 v[B];
               float lft[B], rgt[B];
#pragma unroll
               for (int jj = 0; jj < B; ++jj) {
                  const int q = gl + G * (j0 + jj);
                  v[jj] = ld16<A16>(Cb, q);
                  lft[jj] = (q > 0) ? Cbf[4 * q - 1] : MGM_INF;
                  rgt[jj] = (q + 1 < nq) ? Cbf[4 * q + 4] : MGM_INF;
               }
#pragma unroll
               for (int jj = 0; jj < B; ++jj) {
                  const int q = gl + G * (j0 + jj);
                  float4 a;
                  a.x = sgm_x(lft[jj], v[jj].x, v[jj].y, p1, cap, m) * sc;
                  a.y = sgm_x(v[jj].x, v[jj].y, v[jj].z, p1, cap, m) * sc;
                  a.z = sgm_x(v[jj].y, v[jj].z, v[jj].w, p1, cap, m) * sc;
                  a.w = sgm_x(v[jj].z, v[jj].w, rgt[jj], p1, cap, m) * sc;
                  st16<A16>(cur, q, a);
               }
            };
            int j0 = 0;
            for (; j0 + 4 <= nj; j0 += 4) tbatch(std::integral_constant<int, 4>{}, j0);
            for (; j0 < nj; ++j0) tbatch(std::integral_constant<int, 1>{}, j0);
         } else {
            for (int j = 0; j < nj; ++j) st16<A16>(cur, gl + G * j, ld16<A16>(Cb, gl + G * j));   // weighted paths keep the raw message
         }
      }
      if constexpr (!RC) tick(3);
      tick(4);
      grp.sync();
      tick(5);

      // ---------------- hand the boundary row to the next band: the publisher warp copies the transformed vector
      // of the last row to the boundary line and releases the counter right away (pixels [0,xl] are in memory)
      if (pub_warp && has_next) {
         const int xl = s - SIG * (nrows - 1);
         if (xl >= 0 && xl < maxii) {
            if (NEEDM && lane_id == 0) bndm_out[xl] = msr[(nrows - 1) * 4 + (xl % R)];
            warp_copy_vector(bnd_out + (size_t)xl * VSP, slot_of(nrows - 1, xl), VSP, lane_id);
            warp_publish(prog_out, xl + 1, lane_id);
         }
      }
   }

   if (timing) {
      for (int i = 0; i < 6; ++i) atomicAdd(P.dbg + i, (unsigned long long)tacc[i]);
      atomicAdd(P.dbg + 6, (unsigned long long)nsteps);
      if (D.pass == 0 && band < 60) {   // per band of sweep 0: phases, steps, start and end time stamps
         unsigned long long *b = P.dbg + 8 + band * 10;
         for (int i = 0; i < 6; ++i) b[i] = (unsigned long long)tacc[i];
         b[6] = (unsigned long long)nsteps;
         unsigned long long gt;
         asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
         b[7] = gt;
      }
   }
   // epilogue: save mbarrier parities
   cp_async_wait<0>();
   if (tid == 0) phase[vph_idx] = vph;
   __syncthreads();
   band_finished(P, D, band);
}

// ---------------------------------------------------------------------------------------------------------
// Diagonal sweeps 4-7 with TSGM <= 3 and no image-dependent weights: SHEARED wavefront.
//
// Such a pixel (xs,ys) only reads the scan row above: (xs+1,ys-1), (xs-1,ys-1), (xs,ys-1) in that order.  With
// u = xs+ys (anti-diagonal) and v = ys these are (u,v-1), (u-2,v-1), (u-1,v-1): every predecessor sits at v-1 and
// at u, u-1 or u-2.  A band therefore holds T consecutive anti-diagonals ("workers"), one per 8-lane group, ALL
// at the same v in a step -- no lag between the workers of a band, a ring of two slots per worker -- and depends
// on the previous band only through its last two workers.  The dependency depth of the sweep drops from
// maxii + 2*maxjj steps (row-per-worker, lag 2) to maxjj + (band hand-off) * #bands.
// Everything else (register-resident cost prefetch, gather, exact min-convolution chains / SGM transform,
// boundary lines through TMA stores + release/acquire counters + TMA loads) is as in run_band.
template <int POT, int K, int GL, bool RC = false>
__device__ void run_band_shear(const AggParams &P, const SweepDesc &D, const int band, unsigned char *smem) {
   constexpr int G = GL;
   static_assert(!RC || (POT == POT_TRUNC && GL == 8), "register chains: unweighted truncated linear, 8 lanes");
   const int pass = D.pass;
   static_assert(K <= 3, "the sheared wavefront needs predecessors in the row above only");
   constexpr bool NEEDM = (POT == POT_TRUNC && K == 2);
   constexpr bool CHAINS = (POT == POT_TRUNC);
   constexpr int A16 = RC ? 2 : (POT == POT_TRUNC ? 1 : 0);   // shared-memory vector layout (ld16 / st16)
   constexpr int JB = MGM_JB;
   constexpr int NJR = MGM_AGG_CREG;

   const PassGeom g = pass_geometry(pass, P.nx, P.ny);
   const int maxii = g.maxii, maxjj = g.maxjj;
   const int nu = maxii + maxjj - 1;   // anti-diagonals
   const int T = P.T[1];
   const int TS = P.TS[1];
   const int VS = P.VS;
   const int VSP = RC ? VS + (VS >> 3) : VS;   // floats per vector in shared memory and in the boundary lines
   const int nq = VS >> 2;
#ifdef MGM_FORCE_NJ
   constexpr int nj = MGM_FORCE_NJ;
#else
   const int nj = nq / G;
#endif
   const int ncb = RC ? 1 : P.ncb;
   const bool creg_mode = (ncb == 1);
   const uint32_t vbytes = (uint32_t)VSP * 4u;
   const int tid = threadIdx.x;
   const int ncomp = blockDim.x - 64;
   const bool is_prod = (tid == ncomp);
   const bool pub_warp = (tid >= ncomp + 32);
   const int u0 = band * T;
   const int nrows = min(T, nu - u0);
   const bool has_prev = band > 0;
   const bool has_next = u0 + T < nu;
   auto vlo = [&](int u) { return max(0, u - (maxii - 1)); };
   auto vhi = [&](int u) { return min(maxjj - 1, u); };
   const int sb = vlo(u0), se = vhi(u0 + nrows - 1);   // step window of the band (v = step)
   // boundary positions this band ever reads from the previous band's last two workers
   const int pf_lo = max(sb - 1, 0), pf_hi = has_prev ? min(se - 1, vhi(u0 - 1)) : -1;

   const int r = tid / G, gl = tid % G;
   const unsigned gmask = ((1u << G) - 1u) << ((tid & 31) & ~(G - 1));
   const bool rowok = (tid < ncomp) && (r < nrows);
   const int u = u0 + r;
   const int my_lo = vlo(u), my_hi = vhi(u);
   const int warp_id = tid >> 5, lane_id = tid & 31;

   uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
   float *msr = reinterpret_cast<float *>(smem + P.off_ms);
   float *vms = reinterpret_cast<float *>(smem + P.off_vms);     // [2][RV]
   float *virt = reinterpret_cast<float *>(smem + P.off_virt);   // [2][RV][VS]: worker -1, worker -2
   float *thr = reinterpret_cast<float *>(smem + P.off_thr);
   uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);

   // row groups (RowGroup): the boundary consumer warp belongs to group 0, the publisher warp to the last
   const int NG = P.ng[1], TG = T / NG;
   const RowGroup grp = make_row_group(tid, ncomp, r, NG, TG, G);
   const int gw0 = grp.first_tid >> 5;
   const int ncw = (TG + 31) >> 5;
   const int pair_id0 = 4 + grp.gi * 2;
   const bool inplace = CHAINS && creg_mode;   // message built in its ring slot, min-convolution in place
   const bool fused = POT == POT_SGM && P.fused_sgm != 0;   // see run_band / sgm_transform_regs

   // per-worker shared memory: ring slots [2][VS] (position v&1) | cost buffers [ncb][VS] (none when in place)
   auto row_base = [&](int rr) -> float * { return thr + (size_t)rr * TS; };
   auto slot_of = [&](int rr, int v) -> float * {
      return (rr < 0) ? virt + ((-rr - 1) * RV + (v & (RV - 1))) * VSP : row_base(rr) + (v & 1) * VSP;
   };
   auto m_of = [&](int rr, int v) -> float { return (rr < 0) ? vms[(-rr - 1) * RV + (v & (RV - 1))] : msr[rr * 4 + (v & 1)]; };
   auto cbuf_of = [&](int rr, int v) -> float * { return row_base(rr) + (2 + (ncb == 2 ? (v & 1) : 0)) * VSP; };

   const float *ccv = D.cc;
   // boundary lines: [band][line][maxjj][VS], line 0 = last worker of the band, line 1 = the one before
   float *bnd_out = D.bnd + (size_t)band * 2 * maxjj * VSP;
   float *bndm_out = D.bndm + (size_t)band * 2 * maxjj;
   const float *bnd_in = has_prev ? D.bnd + (size_t)(band - 1) * 2 * maxjj * VSP : nullptr;
   const float *bndm_in = has_prev ? D.bndm + (size_t)(band - 1) * 2 * maxjj : nullptr;
   int *prog_out = D.progress + band;
   const int *prog_in = has_prev ? D.progress + band - 1 : nullptr;

   // pixel of (u, v): xs = u - v, ys = v
   const long long pix_u = g.base0 + (long long)u * g.dxs;
   const long long dv = g.dys - g.dxs;

   uint32_t vph = 0;
   const int vph_idx = max(max(P.T[0], P.T[1]), P.T[2]);
   if (tid < G) vph = phase[vph_idx];
   int next_p = pf_lo;   // producer: next boundary position to fetch

   float4 creg[NJR];
#pragma unroll
   for (int j = 0; j < NJR; ++j) creg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
   auto prefetch_cost = [&](int v) {   // costs of position v+1
      const int vn = v + 1;
      const bool go = rowok && vn >= my_lo && vn <= my_hi;
      if (P.cc_pf > 0) {
         const int vp = vn + P.cc_pf;
         if (rowok && vp >= my_lo && vp <= my_hi) {
            const float *line = ccv + (size_t)(pix_u + (long long)vp * dv) * VS;
            for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L2 [%0];" ::"l"(line + l * 32));
         }
      }
      if (creg_mode) {
         if (go) {
            const float4 *src = reinterpret_cast<const float4 *>(ccv + (size_t)(pix_u + (long long)vn * dv) * VS);
#pragma unroll
            for (int j = 0; j < NJR; ++j)
               if (j < nj) creg[j] = RC ? __ldg(src + gl + G * j) : __ldcs(src + gl + G * j);
         }
         return;
      }
      if (go) {
         const float4 *src = reinterpret_cast<const float4 *>(ccv + (size_t)(pix_u + (long long)vn * dv) * VS);
         float *dst = cbuf_of(r, vn);   // 16-byte aligned in this mode (agg_plan)
         for (int j = 0; j < nj; ++j) cp_async16(dst + 4 * (gl + G * j), src + gl + G * j);
      }
      cp_async_commit();
   };
   auto prefetch_cost_l1 = [&](int v) {
      const int vn = v + 1;
      if (rowok && vn >= my_lo && vn <= my_hi) {
         const float *line = ccv + (size_t)(pix_u + (long long)vn * dv) * VS;
         for (int l = gl; l < (VS >> 5); l += G) asm volatile("prefetch.global.L1 [%0];" ::"l"(line + l * 32));
      }
   };
   prefetch_cost(sb - 1);
   const bool late_prefetch = !creg_mode && CHAINS && (warp_id - gw0 >= 2 * ncw);

   for (int v = sb; v <= se; ++v) {
      if (!late_prefetch && !creg_mode) prefetch_cost(v);
      if (is_prod && pf_hi >= pf_lo) {
         // position v is read in the NEXT step (blocking); up to v+PF are fetched ahead when already published
         const int need = min(pf_hi, v);
         const int lim = min(pf_hi, v + PF);
         int avail = ld_acquire(prog_in);
         while (next_p <= lim) {
            if (avail < next_p + 1) {
               if (next_p > need) break;
               do { __nanosleep(20); avail = ld_acquire(prog_in); } while (avail < next_p + 1);
            }
            fence_proxy_async();
            const int slot = next_p & (RV - 1);
            if (NEEDM) {
               vms[slot] = __ldcg(bndm_in + next_p);
               vms[RV + slot] = __ldcg(bndm_in + maxjj + next_p);
            }
            mbar_expect_tx(&vbar[slot], 2 * vbytes);
            tma_load_1d(virt + slot * VSP, bnd_in + (size_t)next_p * VSP, vbytes, &vbar[slot]);
            tma_load_1d(virt + (RV + slot) * VSP, bnd_in + ((size_t)maxjj + next_p) * VSP, vbytes, &vbar[slot]);
            ++next_p;
         }
      }
      // group 0 waits one step AHEAD for the boundary positions; the end-of-step barriers publish them to all
      if (tid < G && pf_hi >= pf_lo) {
         auto wait_pos = [&](int p) {
            if (p >= pf_lo && p <= pf_hi) {
               const int sl = p & (RV - 1);
               mbar_wait(&vbar[sl], (vph >> sl) & 1u);
               vph ^= 1u << sl;
            }
         };
         if (v == sb) { wait_pos(sb - 1); wait_pos(sb); } else wait_pos(v);
      }

      grp.begin_step(v == sb);   // group ordering (RowGroup): token from the group above, slot reuse below
      const bool act = rowok && v >= my_lo && v <= my_hi;
      const int xs = u - v;
      const long long pix = pix_u + (long long)v * dv;
      float2 *cur = reinterpret_cast<float2 *>(row_base(r) + (v & 1) * VSP);
      float *Cbf = inplace ? reinterpret_cast<float *>(cur) : cbuf_of(r, v);   // where the message is built
      float2 *Cb = reinterpret_cast<float2 *>(Cbf);
#ifdef MGM_Y_FROM_PIX
      float4 *gout = reinterpret_cast<float4 *>(ldir_of_pix(P, D, act ? pix : 0) + (size_t)(act ? pix : 0) * VS);
#else
      float4 *gout = reinterpret_cast<float4 *>(ldir_of_row(P, D, act ? g.y0 + xs * g.ydxs + v * g.ydys : 0) +
                                                (size_t)(act ? pix : 0) * VS);
#endif
      const bool border = (xs == 0) || (v == 0) || (xs == maxii - 1);
      float m = MGM_INF;

      // ---------------- phase 1: gather
      if (act) {
         if (!creg_mode) { if (late_prefetch) cp_async_wait<0>(); else cp_async_wait<1>(); }
         if (border) {
            if (creg_mode) {
#pragma unroll
               for (int j = 0; j < NJR; ++j) {
                  if (j < nj) {
                     const int q = gl + G * j;
                     m = hmin4(m, creg[j]);
                     if (!fused) st16<A16>(Cb, q, creg[j]);
                     __stcs(gout + q, creg[j]);
                  }
               }
            } else {
               const float2 *Cin = reinterpret_cast<const float2 *>(cbuf_of(r, v));
               for (int j = 0; j < nj; ++j) {
                  const int q = gl + G * j;
                  const float4 c = ld16<A16>(Cin, q);
                  m = hmin4(m, c);
                  __stcs(gout + q, c);
               }
            }
         } else {
            const float2 *Cin = reinterpret_cast<const float2 *>(cbuf_of(r, v));   // costs (cp.async mode)
            const float2 *S[K];
            float mk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
               const int prow = (k == 0) ? r : (k == 1 ? r - 2 : r - 1);   // (+1,-1), (-1,-1), (0,-1)
               S[k] = reinterpret_cast<const float2 *>(slot_of(prow, v - 1));
               mk[k] = NEEDM ? m_of(prow, v - 1) : 0.f;
            }
            auto batch = [&](auto jbc, int j0, auto regc) {
               constexpr int B = decltype(jbc)::value;
               constexpr bool REGC = decltype(regc)::value;
               float4 c[B], a[K][B];
#pragma unroll
               for (int jj = 0; jj < B; ++jj) {
                  const int q = gl + G * (j0 + jj);
                  if constexpr (!REGC) c[jj] = ld16<A16>(Cin, q);
#pragma unroll
                  for (int k = 0; k < K; ++k) a[k][jj] = ld16<A16>(S[k], q);
               }
               if constexpr (REGC) {
#pragma unroll
                  for (int jj = 0; jj < B; ++jj) c[jj] = creg[(j0 + jj) < NJR ? (j0 + jj) : 0];
               }
#pragma unroll
               for (int jj = 0; jj < B; ++jj) {
                  const int q = gl + G * (j0 + jj);
                  float4 o;
                  if constexpr (POT == POT_TRUNC && K == 2) {   // update_cost2_trunclinear :216
                     o = add4(c[jj], mul4s(add4s(add4(add4s(a[0][jj], -mk[0]), a[1 % K][jj]), -mk[1 % K]), 0.5f));
                  } else if constexpr (POT == POT_SGM && K == 2) {   // update_cost2: halves taken by the producer
                     o = add4(c[jj], add4(a[0][jj], a[1 % K][jj]));
                  } else {
                     float4 e = a[0][jj];
#pragma unroll
                     for (int k = 1; k < K; ++k) e = add4(e, a[k][jj]);
                     o = add4(c[jj], div4_by_k<K>(e));
                  }
                  m = hmin4(m, o);
                  if (REGC && fused) creg[(j0 + jj) < NJR ? (j0 + jj) : 0] = o;
                  else st16<A16>(Cb, q, o);
                  __stcs(gout + q, o);
               }
            };
            if (creg_mode) {
               constexpr int JR = (JB < NJR) ? JB : NJR;
#pragma unroll
               for (int jb = 0; jb < NJR; jb += JR) {
                  if (jb + JR <= nj) batch(std::integral_constant<int, JR>{}, jb, std::true_type{});
                  else {
#pragma unroll
                     for (int jr = 0; jr < JR; ++jr)
                        if (jb + jr < nj) batch(std::integral_constant<int, 1>{}, jb + jr, std::true_type{});
                  }
               }
            } else {
               int j0 = 0;
               for (; j0 + JB <= nj; j0 += JB) batch(std::integral_constant<int, JB>{}, j0, std::false_type{});
               for (; j0 + 2 <= nj; j0 += 2) batch(std::integral_constant<int, 2>{}, j0, std::false_type{});
               for (; j0 < nj; ++j0) batch(std::integral_constant<int, 1>{}, j0, std::false_type{});
            }
         }
#pragma unroll
         for (int d = 1; d < G; d <<= 1) m = fminf(m, __shfl_xor_sync(gmask, m, d));
         if (gl == 0) msr[r * 4 + (v & 1)] = m;
         if constexpr (POT == POT_SGM) {
            if (fused) sgm_transform_regs<K, NJR, A16, G>(creg, nj, nq, gl, gmask, m, P.P1, P.P2, cur);
         }
      }
      if (creg_mode && !RC) prefetch_cost(v);
      if constexpr (RC) {
         prefetch_cost_l1(v);   // see run_band
         __syncwarp();
         if (act) chain_regs<NJR>(cur, nj, gl, gmask, P.P1, m + P.P2, (K == 2) ? 0.0f : m);
         prefetch_cost(v);
      } else if (!fused) {
         grp.gather_done(v == se);
         grp.sync();
      }

      // ---------------- phase 2: neighbour-side transform of the message into ring slot v&1
      if (late_prefetch) prefetch_cost(v);
      if constexpr (RC) {
      } else if constexpr (CHAINS) {
         const int wg = warp_id - gw0;
         if (wg >= 0 && wg < 2 * ncw) {
            const int cw = wg % ncw, cdir = wg / ncw;
            const int crow = grp.gi * TG + cw * 32 + lane_id;
            const bool on = cw * 32 + lane_id < TG && crow < nrows && v >= vlo(u0 + crow) && v <= vhi(u0 + crow);
            const float cm = on ? msr[crow * 4 + (v & 1)] : 0.f;
            float *dst = on ? row_base(crow) + (v & 1) * VS : thr;
            const float *src = on ? (inplace ? dst : cbuf_of(crow, v)) : thr;
            if (cdir == 0) minconv_half<0>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, P.P1, cm + P.P2, (K == 2) ? 0.0f : cm, pair_id0 + cw);
            else minconv_half<1>(on, reinterpret_cast<const float2 *>(src), reinterpret_cast<float2 *>(dst), nq, P.P1, cm + P.P2, (K == 2) ? 0.0f : cm, pair_id0 + cw);
         }
      } else if (act && !fused) {
         // SGM transform, label-parallel: A(o) = min3(L(o), min(L(o-1),L(o+1))+P1, m+P2) - m  [x 1/2 for K=2]
         const float p1 = P.P1, cap = m + P.P2;
         const float sc = (K == 2) ? 0.5f : 1.0f;
         for (int j = 0; j < nj; ++j) {
            const int q = gl + G * j;
            const float4 x = ld16<A16>(Cb, q);
            const float lft = (q > 0) ? Cbf[4 * q - 1] : MGM_INF;
            const float rgt = (q + 1 < nq) ? Cbf[4 * q + 4] : MGM_INF;
            float4 a;
            a.x = sgm_x(lft, x.x, x.y, p1, cap, m) * sc;
            a.y = sgm_x(x.x, x.y, x.z, p1, cap, m) * sc;
            a.z = sgm_x(x.y, x.z, x.w, p1, cap, m) * sc;
            a.w = sgm_x(x.z, x.w, rgt, p1, cap, m) * sc;
            st16<A16>(cur, q, a);
         }
      }
      grp.sync();

      // ---------------- hand the two boundary workers to the next band (publisher warp, see run_band)
      if (pub_warp && has_next) {
#pragma unroll
         for (int line = 0; line < 2; ++line) {
            const int br = nrows - 1 - line;
            if (v >= vlo(u0 + br) && v <= vhi(u0 + br)) {
               if (NEEDM && lane_id == 0) bndm_out[(size_t)line * maxjj + v] = msr[br * 4 + (v & 1)];
               warp_copy_vector(bnd_out + ((size_t)line * maxjj + v) * VSP, slot_of(br, v), VSP, lane_id);
            }
         }
         warp_publish(prog_out, v + 1, lane_id);   // positions <= v of both boundary workers are in global memory
      }
   }

   if (pub_warp && has_next) warp_publish(prog_out, 0x7fffffff, lane_id);
   cp_async_wait<0>();
   if (tid == 0) phase[vph_idx] = vph;
   __syncthreads();
   band_finished(P, D, band);
}


// RC: the register-chain kernels (unweighted truncated linear, 8 lanes per worker) are separate entry points with a
// smaller block (48 workers) so that a thread can hold two label segments and the prefetched costs (144 registers)
template <int POT, int K, bool WEIGHTED, int GL, bool RC = false>
__global__ void __launch_bounds__(RC ? MGM_AGG_MAX_THREADS_RC : MGM_AGG_MAX_THREADS, 1) mgm_aggregate_kernel(const AggParams P) {
   extern __shared__ __align__(128) unsigned char smem[];
   __shared__ int2 s_ticket;
   __shared__ AggStage s_stage;
   __shared__ __align__(16) unsigned char s_tab_raw[MGM_MAX_NDIR * sizeof(SweepDesc)];
   const int t = threadIdx.x;
   const int ncomp = blockDim.x - 64;   // two service warps follow the row threads
   // the sweep table of a single pair (up to 16 entries) is kept in shared memory: band claims and tile readiness
   // tests read it at every poll
   SweepDesc *s_tab = reinterpret_cast<SweepDesc *>(s_tab_raw);
   const bool small_tab = P.nsweeps <= MGM_MAX_NDIR;
   if (small_tab) {
      const uint4 *src = reinterpret_cast<const uint4 *>(P.sweeps);
      uint4 *dst = reinterpret_cast<uint4 *>(s_tab_raw);
      for (int i = t; i < P.nsweeps * (int)(sizeof(SweepDesc) / 16); i += blockDim.x) dst[i] = src[i];
   }
   const SweepDesc *tab = small_tab ? s_tab : P.sweeps;

   // one-time barrier setup: RV boundary barriers
   {
      uint64_t *vbar = reinterpret_cast<uint64_t *>(smem + P.off_vbar);
      uint32_t *phase = reinterpret_cast<uint32_t *>(smem + P.off_phase);
      const int tmax = max(max(P.T[0], P.T[1]), P.T[2]);
      if (t == ncomp) { for (int i = 0; i < RV; ++i) mbar_init(&vbar[i], 1); }
      if (t == 0) phase[tmax] = 0;
      mbar_fence_init();
      __syncthreads();
   }

   int pending = -1;   // warp 0: claimed finish tile
   for (;;) {
      if (t < 32) {
         const int2 tk = claim_band(P, tab, pending, t);
         if (t == 0) s_ticket = tk;
         // stage the descriptor (16-byte words, warp 0)
         if (tk.x >= 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tab + tk.x);
            uint4 *dst = reinterpret_cast<uint4 *>(&s_stage.d);
            for (int i = t; i < (int)(sizeof(SweepDesc) / 16); i += 32) dst[i] = src[i];
         } else if (tk.x == -2 && P.npairs > 1) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.fins + tk.y / P.fin_ntiles);
            uint4 *dst = reinterpret_cast<uint4 *>(&s_stage.f);
            for (int i = t; i < (int)(sizeof(WtaParams) / 16); i += 32) dst[i] = src[i];
         }
      }
      __syncthreads();
      const int2 pb = s_ticket;
      if (pb.x == -1) break;
      if (pb.x == -2) {
         if (P.npairs == 1) run_finish_tile(P, P.fin0, pb.y, smem);
         else run_finish_tile(P, s_stage.f, pb.y % P.fin_ntiles, smem);
      }
      else {
         const SweepDesc &D = s_stage.d;
         const int pass = D.pass;
         {
            if (pass >= 8) run_band<POT, K, WEIGHTED, true, GL, true, RC>(P, D, pb.y, smem);
            else if (pass < 4) run_band<POT, K, WEIGHTED, false, GL, false, RC>(P, D, pb.y, smem);
            else if constexpr (!WEIGHTED && K <= 3) {
               if (P.shear) run_band_shear<POT, K, GL, RC>(P, D, pb.y, smem);
               else run_band<POT, K, WEIGHTED, true, GL, false, RC>(P, D, pb.y, smem);
            } else run_band<POT, K, WEIGHTED, true, GL, false, RC>(P, D, pb.y, smem);
         }
      }
      __syncthreads();   // the staged descriptor and the ticket are rewritten by the next claim
   }
}

// ---------------------------------------------------------------- host side

template <int POT, int K, bool WEIGHTED, int GL = MGM_AGG_GROUP, bool RC = false>
static cudaError_t launch_t(const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   auto kern = mgm_aggregate_kernel<POT, K, WEIGHTED, GL, RC>;
   cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
   if (e != cudaSuccess) return e;
   int per_sm = 0;
   e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan.block, plan.smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   int grid = min(P.nbands, plan.num_sms * per_sm);
   if (grid < 1) grid = 1;
   if (plan.verbose)
      fprintf(stderr, "[mgmb200] aggregate: grid=%d block=%d smem=%zu CTAs/SM=%d sweeps=%d bands=%d rows=%d/%d/%d groups=%d/%d shear=%d\n",
              grid, plan.block, plan.smem, per_sm, P.nsweeps, P.nbands, plan.T[0], plan.T[1], plan.T[2], plan.ng[0], plan.ng[1],
              plan.shear);
   kern<<<grid, plan.block, plan.smem, st>>>(P);
   return cudaGetLastError();
}

template <int POT, bool WEIGHTED>
static cudaError_t launch_k(int K, const AggParams &P, const AggPlan &plan, cudaStream_t st) {
   if constexpr (POT == POT_SGM && !WEIGHTED) {
      if (plan.lanes == 4) {   // short label vectors: 4 lanes per worker, twice the workers per band
         switch (K) {
         case 1: return launch_t<POT, 1, WEIGHTED, 4>(P, plan, st);
         case 2: return launch_t<POT, 2, WEIGHTED, 4>(P, plan, st);
         case 3: return launch_t<POT, 3, WEIGHTED, 4>(P, plan, st);
         default: return launch_t<POT, 4, WEIGHTED, 4>(P, plan, st);
         }
      }
   }
   if constexpr (POT == POT_TRUNC && !WEIGHTED) {
      if (plan.regchain) {
         switch (K) {
         case 1: return launch_t<POT, 1, WEIGHTED, 8, true>(P, plan, st);
         case 2: return launch_t<POT, 2, WEIGHTED, 8, true>(P, plan, st);
         case 3: return launch_t<POT, 3, WEIGHTED, 8, true>(P, plan, st);
         default: return launch_t<POT, 4, WEIGHTED, 8, true>(P, plan, st);
         }
      }
   }
   switch (K) {
   case 1: return launch_t<POT, 1, WEIGHTED>(P, plan, st);
   case 2: return launch_t<POT, 2, WEIGHTED>(P, plan, st);
   case 3: return launch_t<POT, 3, WEIGHTED>(P, plan, st);
   default: return launch_t<POT, 4, WEIGHTED>(P, plan, st);
   }
}

// every variant behind runtime switches (the lean kernels of aggregate_sgm.cu / aggregate_trunc.cu are dispatched by
// agg_launch in aggregate_plan.cu)
// This file is compiled twice, once per potential (aggregate_gsgm.cu / aggregate_gtrunc.cu define MGM_GENERIC_POT and
// include it): the two halves of the instantiations build in parallel.
#ifndef MGM_GENERIC_POT
#error "compile through aggregate_gsgm.cu / aggregate_gtrunc.cu"
#endif
#if MGM_GENERIC_POT == 0
cudaError_t agg_launch_generic_sgm(const AggParams &P, const AggPlan &plan, int K, bool weighted, cudaStream_t st) {
   return weighted ? launch_k<POT_SGM, true>(K, P, plan, st) : launch_k<POT_SGM, false>(K, P, plan, st);
}
#else
cudaError_t agg_launch_generic_trunc(const AggParams &P, const AggPlan &plan, int K, bool weighted, cudaStream_t st) {
   return weighted ? launch_k<POT_TRUNC, true>(K, P, plan, st) : launch_k<POT_TRUNC, false>(K, P, plan, st);
}
#endif

}  // namespace mgm
