// Device helpers shared by the aggregation translation units (aggregate.cu: every variant; aggregate_sgm.cu: the
// lean unweighted SGM kernels): packed fp32 arithmetic, shared-memory vector layouts, boundary publishing, dynamic
// band claiming and the fused finish tiles.  See aggregate.cu for the design notes.
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "aggregate.cuh"
#include "wta_device.cuh"

#ifndef MGM_JB
#define MGM_JB 2   // chunks per lane whose loads are issued together in the gather (measured: 2 < 4 < 8)
#endif
#ifndef MGM_CHAIN_PF
#define MGM_CHAIN_PF 1   // chunks loaded ahead of the min-convolution chain (see minconv_half)
#endif
#ifndef MGM_EXP
#define MGM_EXP 0   // timing experiments only (tools/micro): 1 no message store, 2 no cost load, 3 no gather arithmetic, 4 no min-convolution
#endif

namespace mgm {

static constexpr int RV = 8;   // virtual-row ring (pixels of the previous band's last row)
static constexpr int PF = 3;   // boundary prefetch distance in pixels
static constexpr int G = MGM_AGG_GROUP;   // lanes cooperating on one scan row (the band functions shadow it with their GL)

template <bool DIAG>
__device__ __forceinline__ constexpr int pred_type(int k) {
   return DIAG ? (k == 0 ? PRED_UPR : k == 1 ? PRED_UPL : k == 2 ? PRED_UP : PRED_SAME)
               : (k == 0 ? PRED_SAME : k == 1 ? PRED_UP : k == 2 ? PRED_UPL : PRED_UPR);
}

__device__ __forceinline__ float hmin4(float m, const float4 &v) {
   return fminf(fminf(fminf(m, v.x), fminf(v.y, v.z)), v.w);
}

// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each lane is an independently rounded IEEE operation, so
// the results are bit-identical to the scalar forms; one issue slot per two labels in the ALU-bound gather.
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) {
   const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
   const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
   return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 add4s(const float4 &a, const float s) {   // a + s per label (s = -m: a - m exactly)
   const float2 ss = make_float2(s, s);
   const float2 lo = __fadd2_rn(make_float2(a.x, a.y), ss);
   const float2 hi = __fadd2_rn(make_float2(a.z, a.w), ss);
   return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 mul4s(const float4 &a, const float s) {
   const float2 ss = make_float2(s, s);
   const float2 lo = __fmul2_rn(make_float2(a.x, a.y), ss);
   const float2 hi = __fmul2_rn(make_float2(a.z, a.w), ss);
   return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// e/3 for FINITE e (div3_exact without its non-finite guard; common.cuh).  The sums of neighbour terms are finite
// under the fast-path preconditions (finite caps m+P2*w or a finite entry in every vector, P1 finite).
__device__ __forceinline__ float4 div3_4(const float4 &e) {
   const float2 c3 = make_float2(0x1.555556p-2f, 0x1.555556p-2f), m3 = make_float2(-3.0f, -3.0f);
   const float2 elo = make_float2(e.x, e.y), ehi = make_float2(e.z, e.w);
   const float2 qlo = __fmul2_rn(elo, c3), qhi = __fmul2_rn(ehi, c3);
   const float2 rlo = __ffma2_rn(m3, qlo, elo), rhi = __ffma2_rn(m3, qhi, ehi);
   const float2 lo = __ffma2_rn(rlo, c3, qlo), hi = __ffma2_rn(rhi, c3, qhi);
   return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Shared-memory vectors are moved as 8-byte halves.  Measured on B200 (tools/micro/smem_bw.cu, smem_pat.cu):
// LDS.128 delivers 64 B/clk per SM, LDS.64 / STS.64 the full 128 B/clk when the 16 lanes of a half-warp hit 16
// different 8-byte banks.  With a row stride of 8 bytes modulo 128 (register-cost mode, agg_plan) two adjacent
// rows read by 8 lanes each -- and 16 consecutive rows read by one lane each (the chains) -- do exactly that;
// the kernel is bound by the LSU, so this layout is worth 2x on the gather loads and 1.3x on the chain loads.
// Rows are then only 8-byte aligned: chunk q of a vector is the pair of float2 elements 2q, 2q+1 and no float4
// pointer is ever formed on row memory.
// A16: the rows are 16-byte aligned and the chunk moves as one 16-byte access (truncated-linear kernels: their
// chain lanes read one chunk per ROW, a pattern that gains little from 8-byte halves and pays for the extra
// instructions in its dependent stream -- measured).
// LAY: 0 = 8-byte halves (SGM kernels), 1 = 16-byte aligned rows and accesses (truncated-linear kernels with lane-pair
// chains), 2 = 8-byte halves in the PADDED layout of the register-chain kernels: 16 bytes of padding after every 8
// chunks (128 bytes), i.e. chunk q sits at float offset 4q + 4(q >> 3).  With a row stride of 8 bytes modulo 128 this
// layout is conflict-free at the LSU's full 128 B/clk for BOTH ways the 8 lanes of a worker walk a vector: interleaved
// (lane g takes chunks g, g+8, ...: the gather, coalesced with the global accesses) and contiguous (lane g takes chunks
// g*nj .. g*nj+nj-1: the min-convolution chain, nj = 1, 2, 4, 8) -- the transposition between the two costs nothing.
template <int LAY>
__device__ __forceinline__ int chunk_pos(int q) { return LAY == 2 ? 2 * q + 2 * (q >> 3) : 2 * q; }   // in float2 units
template <int LAY>
__device__ __forceinline__ float4 ld16(const float2 *p, int q) {
   if (LAY == 1) return *reinterpret_cast<const float4 *>(p + 2 * q);
   const int u = chunk_pos<LAY>(q);
   const float2 a = p[u], b = p[u + 1];
   return make_float4(a.x, a.y, b.x, b.y);
}
template <int LAY>
__device__ __forceinline__ void st16(float2 *p, int q, const float4 &v) {
   if (LAY == 1) { *reinterpret_cast<float4 *>(p + 2 * q) = v; return; }
   const int u = chunk_pos<LAY>(q);
   p[u] = make_float2(v.x, v.y);
   p[u + 1] = make_float2(v.z, v.w);
}

// The publisher warp copies a boundary vector from its ring slot to the global boundary line (rows may be only
// 8-byte aligned, so no bulk copy here; the consuming band TMA-loads the line into its 128-byte aligned ring).
__device__ __forceinline__ void warp_copy_vector(float *gdst, const float *ssrc, int VS, int lane) {
   const float2 *s2 = reinterpret_cast<const float2 *>(ssrc);
   float2 *d2 = reinterpret_cast<float2 *>(gdst);
   for (int i = lane; i < (VS >> 1); i += 32) d2[i] = s2[i];
}
// all lanes' stores -> gpu-scope fence -> warp barrier -> release by lane 0
__device__ __forceinline__ void warp_publish(int *prog, int value, int lane) {
   __threadfence();
   __syncwarp();
   if (lane == 0) st_release(prog, value);
}

// SGM neighbour transform of one label: min3(L(o), min(L(o-1),L(o+1))+P1, m+P2) - m   (mgm_core.cc:113-116)
__device__ __forceinline__ float sgm_x(float l, float c, float r, float p1, float cap, float m) {
   return fminf(fminf(c, fminf(l, r) + p1), cap) - m;
}

// SGM transform of a whole message held in registers (register-cost mode, unweighted SGM kernels): lane gl of the
// 8-lane group holds the chunks gl + 8j in v[j]; the labels next to a chunk live in the neighbouring lanes (same
// j) or, at the ends of the group, in lane 7 / lane 0 of the previous / next j.  The message never goes through
// shared memory and the step needs one barrier instead of two.
template <int K, int NJR, int A16, int GL>
__device__ __forceinline__ void sgm_transform_regs(const float4 (&v)[NJR], int nj, int nq, int gl, unsigned gmask, float m,
                                                   float p1, float p2, float2 *cur) {
   constexpr int G = GL;
   const float cap = m + p2;
   const float sc = (K == 2) ? 0.5f : 1.0f;
#pragma unroll
   for (int j = 0; j < NJR; ++j) {
      if (j < nj) {
         const int q = gl + G * j;
         // last label of chunk q-1, first label of chunk q+1: one rotation of the group each way, the SENDER picks
         // the value (the last lane hands its chunk j-1 to lane 0, lane 0 its chunk j+1 to the last lane) -- the
         // shuffles share the LSU pipe with the ring traffic that bounds this kernel
         const float snd_l = (gl == G - 1) ? v[j > 0 ? j - 1 : 0].w : v[j].w;
         const float snd_r = (gl == 0) ? v[j + 1 < NJR ? j + 1 : j].x : v[j].x;
         const float got_l = __shfl_sync(gmask, snd_l, (gl + G - 1) & (G - 1), G);
         const float got_r = __shfl_sync(gmask, snd_r, (gl + 1) & (G - 1), G);
         const float lft = (q == 0) ? MGM_INF : got_l;
         const float rgt = (q + 1 >= nq) ? MGM_INF : got_r;
         float4 a;
         a.x = sgm_x(lft, v[j].x, v[j].y, p1, cap, m) * sc;
         a.y = sgm_x(v[j].x, v[j].y, v[j].z, p1, cap, m) * sc;
         a.z = sgm_x(v[j].y, v[j].z, v[j].w, p1, cap, m) * sc;
         a.w = sgm_x(v[j].z, v[j].w, rgt, p1, cap, m) * sc;
         st16<A16>(cur, q, a);
      }
   }
}

// Truncated-linear min-convolution (minConvTruncatedLinear, mgm_core.cc:152-163) of src into dst by a PAIR
// of adjacent lanes: lane dir=0 runs the forward recurrence F[o] = min(F[o-1]+c, M[o]) upwards, lane dir=1
// the same recurrence downwards on the ORIGINAL values (Bp).  Because x -> RN(x+c) is monotone and
// inflationary (c >= 0) the reference's backward pass over F equals min(F, Bp) bit for bit, so the two
// sequential chains can run concurrently and meet in the middle: each lane writes its partial values for
// its first half of the labels, then finishes the other half with the partner's partials.  Every addition
// is performed in the same order as the reference's sequential loops.  The result written to dst is
// min(minconv, cap) - sub.  Loads are issued one chunk ahead of the dependent add/min chain.
// One chunk of four labels of the recurrence F[o] = min(F[o-1]+c, a[o]), two labels per dependent step.  By
// monotonicity of x -> RN(x+c):  F1 = min(r+c, a0),  F2 = min((r+c)+c, a0+c, a1)  with every "+c" a separately
// rounded addition -- bit-identical to the step-by-step form.  The loop-carried dependency per pair is two
// additions and one 3-input minimum (~13 cycles) instead of two add->min pairs (~20) for five instructions
// instead of four; the kernel is issue-bound in this phase, so the cheaper two-label form beats deeper unrolling.
__device__ __forceinline__ void chain4(float &run, float &a0, float &a1, float &a2, float &a3, const float c) {
   {
      const float u1 = run + c, u2 = u1 + c, a0p = a0 + c;
      a0 = fminf(u1, a0);
      a1 = fminf(fminf(u2, a0p), a1);
   }
   {
      const float u1 = a1 + c, u2 = u1 + c, a2p = a2 + c;
      a2 = fminf(u1, a2);
      a3 = fminf(fminf(u2, a2p), a3);
   }
   run = a3;
}

// named barrier shared by one forward warp and its backward partner warp (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// Half of the meet-in-the-middle min-convolution of `src` into `dst`, executed by a whole WARP whose lanes
// hold the same direction (DIR=0 upwards, DIR=1 downwards) of 32 different vectors; the partner warp runs
// the other direction of the same vectors.  Warp-uniform direction: no divergence, no lane swizzles.
// `on` = this lane has a vector to process (the barrier is executed by every lane regardless).
//
// Second half without re-reading the source: with F the upward recurrence and B the downward one, both on the
// original values M, the result is Q[o] = min(F[o], B[o]).  Since F[o] <= M[o] and x -> RN(x+c) is monotone,
//    min(F[o], B[o]) = min(F[o], min(B[o+1]+c, M[o])) = min(F[o], B[o+1]+c)   and   F[o+1]+c >= min(F[o], B[o+1]+c),
// hence Q[o] = min(Q[o+1]+c, F[o]) (and symmetrically Q[o] = min(Q[o-1]+c, B[o])): each lane continues ITS
// recurrence over the partner's partial values only -- same additions in the same order as the reference,
// one shared-memory read per label less (checked bit for bit in tools/micro/chain_bench.cu, form 6).
// MASK: labels outside [mlo,mhi] of the source read as +INF -- the min-convolution of a truncated-linear update
// with per-pixel ranges runs inside the RECEIVING pixel's range (mgm_core.cc:229-281), which on dense vectors is
// the convolution of the vector masked to that range.
__device__ __forceinline__ float4 mask_labels(float4 v, int q, int mlo, int mhi) {
   const int o = 4 * q;
   v.x = (o >= mlo && o <= mhi) ? v.x : MGM_INF;
   v.y = (o + 1 >= mlo && o + 1 <= mhi) ? v.y : MGM_INF;
   v.z = (o + 2 >= mlo && o + 2 <= mhi) ? v.z : MGM_INF;
   v.w = (o + 3 >= mlo && o + 3 <= mhi) ? v.w : MGM_INF;
   return v;
}
template <int DIR, bool MASK = false, int LAY = 1>
__device__ __forceinline__ void minconv_half(bool on, const float2 *src, float2 *dst, int nq, float c, float cap,
                                             float sub, int bar_id, int mlo = 0, int mhi = 0) {
   // D = MGM_CHAIN_PF chunks are loaded ahead of the dependent add/min chain (register ring).  One chunk covers the
   // shared-memory latency of an otherwise idle LSU; when other warps gather at the same time (row groups) the
   // queueing delay is longer than one chunk of chain work (~29 cycles) and a deeper ring keeps the chain fed.
   constexpr int D = MGM_CHAIN_PF;
   const int h = nq >> 1;   // nq is even (a multiple of 8)
   const int dq = DIR ? -1 : 1;
   const int qend = DIR ? 0 : (nq - 1);
   int q = DIR ? (nq - 1) : 0;
   float run = MGM_INF;
   float4 sb[D];
   auto clampq = [&](int qq) { return DIR ? max(qq, qend) : min(qq, qend); };
   auto ldm = [&](const float2 *p, int qq, bool msk) {
      float4 v = ld16<LAY>(p, qq);
      if (MASK && msk) v = mask_labels(v, qq, mlo, mhi);
      return v;
   };
   if (on) {
#pragma unroll
      for (int d = 0; d < D; ++d) sb[d] = ldm(src, clampq(q + d * dq), true);   // own half only needs chunks < h: h >= D or unused
      for (int i = 0; i < h; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            if (i + d < h) {
               float4 v = sb[d];
               // never beyond this lane's half: in place, the partner is overwriting the other half
               if (i + d + D < h) sb[d] = ldm(src, q + D * dq, true);
               if (DIR) { chain4(run, v.w, v.z, v.y, v.x, c); } else { chain4(run, v.x, v.y, v.z, v.w, c); }
               st16<LAY>(dst, q, v);
               q += dq;
            }
         }
      }
   }
   pair_barrier(bar_id);   // partner's partial values are now in dst
   if (on) {
#pragma unroll
      for (int d = 0; d < D; ++d) sb[d] = ld16<LAY>(dst, clampq(q + d * dq));
      for (int i = h; i < nq; i += D) {
#pragma unroll
         for (int d = 0; d < D; ++d) {
            if (i + d < nq) {
               float4 v = sb[d];
               sb[d] = ld16<LAY>(dst, clampq(q + D * dq));   // beyond the end: a redundant re-load, never used
               if (DIR) { chain4(run, v.w, v.z, v.y, v.x, c); } else { chain4(run, v.x, v.y, v.z, v.w, c); }
               v = add4s(make_float4(fminf(v.x, cap), fminf(v.y, cap), fminf(v.z, cap), fminf(v.w, cap)), -sub);
               st16<LAY>(dst, q, v);
               q += dq;
            }
         }
      }
   }
}

template <int K>
__device__ __forceinline__ float4 div4_by_k(const float4 &e) {   // e finite (see div3_4)
   if (K == 1) return e;
   if (K == 2) return mul4s(e, 0.5f);    // exact: same real quotient, same rounding
   if (K == 4) return mul4s(e, 0.25f);
   return div3_4(e);
}

// Every message of the band is in memory (all threads have passed the barrier that follows their last store):
// one thread publishes the completion flag the fused finish tiles wait for.
__device__ __forceinline__ void band_finished(const AggParams &P, const SweepDesc &D, int band) {
   if (P.fin_enabled && threadIdx.x == 0) {
      __threadfence();
      st_release(D.band_done + band, 1);
   }
}

// Where the message of a pixel in image row y is stored: the sweep's volume, or -- sweep-sharded multi-GPU layout --
// the volume of the rank that finishes the row slab y belongs to (a peer mapping: the store crosses NVLink while the
// sweep runs, so the ordered finish later reads local memory only).
__device__ __forceinline__ float *ldir_of_row(const AggParams &P, const SweepDesc &D, int y) {
   return (P.nslabs > 1) ? D.ldir[__umulhi((unsigned)y, P.slab_magic)] : D.ldir[0];
}
// the same from the pixel index: the row is only computed in the slab layout (nothing extra stays live otherwise)
__device__ __forceinline__ float *ldir_of_pix(const AggParams &P, const SweepDesc &D, long long pix) {
   if (P.nslabs > 1) return D.ldir[__umulhi((unsigned)pix / (unsigned)P.nx, P.slab_magic)];
   return D.ldir[0];
}

// Dynamic band scheduling (warp 0 of every CTA).  A launch holds any number of sweeps (SweepDesc table): the 8 or 16
// sweeps of a pair, of several pairs, or the few sweeps this GPU owns.  Bands of a sweep are claimed strictly in
// order through a per-sweep counter, so a claimed band's predecessor is always running or finished: any grid size
// is deadlock free.  A band of a row-per-worker sweep (axis, knight, unsheared diagonal) trails its predecessor by a
// whole band of steps; claiming it before the predecessor has published anything would park an SM for
// milliseconds.  Hence:
//   1. a row-per-worker band whose predecessor has started publishing its boundary row (or a first band), longest
//      remaining chain first -- these sweeps are the critical path;
//   2. else the next band of the sheared diagonal sweep that is least advanced (short hand-off: the filler work);
//   3. else (fused finish) the finish tile this CTA holds, if the bands that cover it are complete;
//   4. else any remaining row-per-worker band (it waits inside run_band);
//   5. else, with a tile still pending, wait for it; without, the CTA is done.
// Finish tiles never block: they are claimed one per CTA from a host-built order and only run when ready, so the
// band argument above is unchanged.  The table is scanned by the 32 lanes in parallel (a batch of 32 pairs holds
// 256 sweeps).

// A tile can be finished once every band (of every sweep of its pair) that holds one of its pixels is complete.
// Band indices are monotone in the scan coordinates, which are affine in (x,y): the extremes are at the corners.
// Lane p checks sweep p of the pair.
static __device__ bool tile_ready(const AggParams &P, const SweepDesc *tab, int gtile, int lane) {
   const int pair = gtile / P.fin_ntiles, tile = gtile % P.fin_ntiles;
   const int x0 = (tile % P.fin_tiles_x) * P.fin_tw, y0 = (tile / P.fin_tiles_x) * P.fin_th;
   const int x1 = min(x0 + P.fin_tw, P.nx) - 1, y1 = min(y0 + P.fin_th, P.ny) - 1;
   int done = 1;
   if (lane < P.ndir) {
      const SweepDesc &d = tab[pair * P.ndir + lane];
      if (d.nb) {
         const int p = d.pass & 7;
         const int rm = (0x53 >> p) & 1, incx = (0xC5 >> p) & 1, incy = (0x99 >> p) & 1;   // pass_geometry
         const int ax0 = incx ? x0 : P.nx - 1 - x1, ax1 = incx ? x1 : P.nx - 1 - x0;       // ascending along the scan
         const int ay0 = incy ? y0 : P.ny - 1 - y1, ay1 = incy ? y1 : P.ny - 1 - y0;
         const int xs0 = rm ? ax0 : ay0, xs1 = rm ? ax1 : ay1, ys0 = rm ? ay0 : ax0, ys1 = rm ? ay1 : ax1;
         int b0, b1;
         if (d.filler) { b0 = (xs0 + ys0) / P.T[CLS_DIAG]; b1 = (xs1 + ys1) / P.T[CLS_DIAG]; }
         else { const int T = P.T[d.cls]; b0 = ys0 / T; b1 = ys1 / T; }
         // relaxed loads (they pipeline), ordered before the tile's reads by the fence below
         for (int b = b1; b >= b0; --b) done &= *reinterpret_cast<volatile const int *>(d.band_done + b);
      }
   }
   if (!__all_sync(0xffffffffu, done)) return false;
   __threadfence();
   return true;
}

// Executed by the 32 lanes of warp 0 in lock step; returns (sweep index, band), (-2, global tile) or (-1, 0) = done.
// `pending`: a finish tile this CTA has claimed but not run yet (-1 none, -2 no tiles left); lane 0's copy counts.
static __device__ int2 claim_band(const AggParams &P, const SweepDesc *tab, int &pending, int lane) {
   for (;;) {
      // candidates of this lane: best ready row-per-worker band, best filler band, best not-ready row-per-worker band
      int rdy_v = -1, rdy_b = 0, rdy_rem = 0, fil_v = -1, fil_b = 0x7fffffff, any_v = -1, any_b = 0, any_rem = 0;
      for (int v = lane; v < P.nsweeps; v += 32) {
         const SweepDesc &d = tab[v];
         const int nbp = d.nb;
         if (!nbp) continue;
         const int b = *reinterpret_cast<volatile int *>(P.next_band + v);
         if (b >= nbp) continue;
         if (d.filler) {
            if (b < fil_b) { fil_v = v; fil_b = b; }
         } else {
            const int rem = nbp - b;
            const bool ready = P.static_order || b == 0 || ld_acquire(d.progress + b - 1) >= 1;
            if (ready && rem > rdy_rem) { rdy_v = v; rdy_b = b; rdy_rem = rem; }
            if (rem > any_rem) { any_v = v; any_b = b; any_rem = rem; }
         }
      }
      // warp reductions (ties: smaller sweep index)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
         const int orem = __shfl_xor_sync(0xffffffffu, rdy_rem, o), ov = __shfl_xor_sync(0xffffffffu, rdy_v, o),
                   ob = __shfl_xor_sync(0xffffffffu, rdy_b, o);
         if (ov >= 0 && (rdy_v < 0 || orem > rdy_rem || (orem == rdy_rem && ov < rdy_v))) { rdy_rem = orem; rdy_v = ov; rdy_b = ob; }
         const int fb = __shfl_xor_sync(0xffffffffu, fil_b, o), fv = __shfl_xor_sync(0xffffffffu, fil_v, o);
         if (fv >= 0 && (fil_v < 0 || fb < fil_b || (fb == fil_b && fv < fil_v))) { fil_b = fb; fil_v = fv; }
         const int arem = __shfl_xor_sync(0xffffffffu, any_rem, o), av = __shfl_xor_sync(0xffffffffu, any_v, o),
                   ab = __shfl_xor_sync(0xffffffffu, any_b, o);
         if (av >= 0 && (any_v < 0 || arem > any_rem || (arem == any_rem && av < any_v))) { any_rem = arem; any_v = av; any_b = ab; }
      }
      int best = rdy_v, bb = rdy_b;
      if (best < 0 && fil_v >= 0) { best = fil_v; bb = fil_b; }
      if (best < 0 && P.fin_enabled) {
         // 3. a finish tile (keeps the SM busy instead of parking on a not-ready band): tiles are claimed one by one in
         // the expected order of readiness and run once the bands that hold their pixels are complete
         if (pending == -1) {
            int t = 0;
            if (lane == 0) t = atomicAdd(P.fin_next, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            pending = (t < P.fin_total) ? P.fin_order[t] : -2;
         }
         if (pending >= 0 && tile_ready(P, tab, pending, lane)) {
            const int tile = pending;
            pending = -1;
            return make_int2(-2, tile);
         }
      }
      if (best < 0 && any_v >= 0) { best = any_v; bb = any_b; }
      if (best < 0) {
         if (!P.fin_enabled || pending < 0) return make_int2(-1, 0);
         __nanosleep(500);   // every band is claimed: wait for the bands that still hold this CTA's tile back
         continue;
      }
      int got = 0;
      if (lane == 0) got = (atomicCAS(P.next_band + best, bb, bb + 1) == bb);
      if (__shfl_sync(0xffffffffu, got, 0)) return make_int2(best, bb);
   }
}

// One finish tile: fin_tw x fin_th pixels, one warp per pixel (wta_device.cuh); the rows region of the shared
// memory is free between bands and holds one label vector per warp.  F: the pair's finish parameters (shared memory).
template <int LP>
__device__ __forceinline__ void run_finish_tile_lp(const AggParams &P, const WtaParams &F, int tile, unsigned char *smem) {
   constexpr int NP = 32 / LP;   // pixels per warp (wta_pixel)
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5, sub = lane / LP;
   float *sS = reinterpret_cast<float *>(smem + P.off_thr) + ((size_t)warp * NP + sub) * P.VS;
   const int x0 = (tile % P.fin_tiles_x) * P.fin_tw, y0 = (tile / P.fin_tiles_x) * P.fin_th;
   const int w = min(P.fin_tw, P.nx - x0), h = min(P.fin_th, P.ny - y0);
   // 128-byte lines of one pixel: (ndir + 1) vectors of VS floats; the warp's next pixels are prefetched into L2
   // while the current ones are reduced (16 warps per SM cannot keep enough loads in flight otherwise)
   const int lpv = max(1, P.VS >> 5), nlines = (F.ndir + 1) * lpv;
   auto prefetch_pixels = [&](int i0) {   // pixels i0 .. i0+NP-1 of the tile
      for (int k = 0; k < NP; ++k) {
         const int i = i0 + k;
         if (i >= w * h) break;
         const long long pix = (long long)(y0 + i / w) * P.nx + x0 + i % w;
         for (int l = lane; l < nlines; l += 32) {
            const int v = l / lpv;
            const float *base = (v < F.ndir) ? F.ldir[v] : F.cc;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)pix * P.VS + (size_t)(l % lpv) * 32));
         }
      }
   };
#ifndef MGM_FIN_PF
#define MGM_FIN_PF 1   // prefetch distance in pixel groups of the warp
#endif
   for (int d = 0; d < MGM_FIN_PF; ++d)
      if ((warp + d * nwarps) * NP < w * h) prefetch_pixels((warp + d * nwarps) * NP);
   for (int i0 = warp * NP; i0 < w * h; i0 += nwarps * NP) {
      if (i0 + MGM_FIN_PF * nwarps * NP < w * h) prefetch_pixels(i0 + MGM_FIN_PF * nwarps * NP);
      const int i = i0 + sub;
      const bool valid = i < w * h;
      const long long pix = valid ? (long long)(y0 + i / w) * P.nx + x0 + i % w : 0;
      wta_pixel<true, LP>(F, pix, sS, lane, valid);
   }
   __syncthreads();
}
__device__ __forceinline__ void run_finish_tile(const AggParams &P, const WtaParams &F, int tile, unsigned char *smem) {
   const int lp = wta_lanes_per_pixel(P.VS);
   if (lp == 32) run_finish_tile_lp<32>(P, F, tile, smem);
   else if (lp == 16) run_finish_tile_lp<16>(P, F, tile, smem);
   else run_finish_tile_lp<8>(P, F, tile, smem);
}

// the claimed band's sweep (or the claimed tile's pair) staged in shared memory
union __align__(16) AggStage {
   SweepDesc d;
   WtaParams f;
   __device__ AggStage() {}
};

}  // namespace mgm
