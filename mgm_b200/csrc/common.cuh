// mgm_b200 -- shared device/host helpers for the sm_100a kernels.
//
// Numerics contract (DESIGN.md "parity"): the reference is strict IEEE fp32
// compiled without FMA (Makefile:1), so every translation unit here is built
// with -fmad=false and default -prec-div/-prec-sqrt; nothing uses fast-math
// intrinsics.  Minima follow the reference's compare-and-select macros
// (mgm_core.cc:47-60); where a kernel states the "no NaN" precondition the
// hardware FMNMX form is used because it is then bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define MGM_INF CUDART_INF_F

namespace mgm {

// ---------------------------------------------------------------- float helpers
// mgm_core.cc:48   __min(a,b) = (a<b)?a:b   (NaN in b survives, NaN in a is dropped)
__device__ __forceinline__ float sel_min(float a, float b) { return (a < b) ? a : b; }
__device__ __forceinline__ float sel_max(float a, float b) { return (a > b) ? a : b; }
// mgm_core.cc:54-60  fmin3
__device__ __forceinline__ float min3_gt(float a, float b, float c) {
   float m = a;
   if (m > b) m = b;
   if (m > c) m = c;
   return m;
}

// Correctly rounded x/3 without the generic division sequence: q0 = x*RN(1/3), one exact
// residual (fma) and one correction.  Checked exhaustively on the host against x/3.0f for
// every finite float, denormals included (tests/test_host_math.py): the only difference in
// 2^32 inputs is the sign of the zero returned for x = -0.  Non-finite x is returned as is
// (INF/3 = INF, NaN/3 = NaN).  Branch-free on purpose: it sits in the gather inner loop.
__device__ __forceinline__ float div3_exact(float x) {
   const float c3 = 0x1.555556p-2f;
   const float q = __fmul_rn(x, c3);
   const float r = __fmaf_rn(-3.0f, q, x);
   const float q2 = __fmaf_rn(r, c3, q);
   return (fabsf(x) < MGM_INF) ? q2 : x;
}

// edge_potentials / howmany  (mgm_core.cc:141,278): float divided by an int
template <int K>
__device__ __forceinline__ float div_by_k(float e) {
   if (K == 1) return e;
   if (K == 2) return e * 0.5f;    // exact: same real quotient, same rounding
   if (K == 4) return e * 0.25f;
   return div3_exact(e);
}

// ---------------------------------------------------------------- shared-memory / TMA plumbing
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
   return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
   uint32_t ok;
   asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
   return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
   while (!mbar_try_wait(bar, parity)) {}
}

// generic-proxy accesses before this point are ordered before async-proxy (TMA) accesses after it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() {
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
                "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}
// 1-D bulk copy shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                "r"(smem_u32(smem_src)), "r"(bytes)
                : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
   asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all() {
   asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Ampere-style asynchronous 16-byte copy global -> shared (SASS: LDGSTS), L2 only
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
   asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ int ld_acquire(const int *p) {
   int v;
   asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
   asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- sweep geometry (mgm_core.cc:463-471,500-523)
// Scan space: xs in [0,maxii) fast, ys in [0,maxjj) slow.  Image pixel of (xs,ys):
//   p0 (xs,ys)  p1 (W-1-xs,H-1-ys)  p2 (ys,H-1-xs)  p3 (W-1-ys,xs)
//   p4 (W-1-xs,ys)  p5 (W-1-ys,H-1-xs)  p6 (xs,H-1-ys)  p7 (ys,xs)
// Sweeps 8-15 (-O 16) scan like sweeps 0-7 and differ in the ORDER of the predecessors only (knight_pred_type).
struct PassGeom {
   int maxii, maxjj;
   long long base0;   // pixel index of (xs=0, ys=0)
   long long dxs;     // pixel index increment per xs
   long long dys;     // pixel index increment per ys
   int y0, ydxs, ydys;   // image row of (xs,ys) = y0 + xs*ydxs + ys*ydys
};
__host__ __device__ inline PassGeom pass_geometry(int pass, int nx, int ny) {
   pass &= 7;
   // row_major / inc_x / inc_y per sweep, restated from the reference table
   const int rm = (0x53 >> pass) & 1;     // passes 0,1,4,6 scan image rows
   const int incx = (0xC5 >> pass) & 1;   // passes 0,2,6,7 ascend in x
   const int incy = (0x99 >> pass) & 1;   // passes 0,3,4,7 ascend in y
   PassGeom g;
   g.maxii = rm ? nx : ny;
   g.maxjj = rm ? ny : nx;
   long long x0 = incx ? 0 : nx - 1, y0 = incy ? 0 : ny - 1;
   long long sx = incx ? 1 : -1, sy = incy ? 1 : -1;
   g.base0 = x0 + y0 * nx;
   g.y0 = (int)y0;
   if (rm) { g.dxs = sx; g.dys = sy * nx; g.ydxs = 0; g.ydys = (int)sy; }
   else    { g.dxs = sy * nx; g.dys = sx; g.ydxs = (int)sy; g.ydys = 0; }
   return g;
}

// Scan-space predecessors {(-1,0),(0,-1),(-1,-1),(+1,-1)} of every sweep
enum PredType { PRED_SAME = 0, PRED_UP = 1, PRED_UPL = 2, PRED_UPR = 3 };

// Sweeps 8-15 (NDIR = 16).  The reference advertises -O 16 (mgm.cc:223) but its table ends after the eight sweeps
// above with the stub comment "// 22.5 deg" (mgm_core.cc:472-473), and Pass_setup carries the note "use dir1 if y odd
// dir2 otherwise" (mgm_core.cc:386): running it reads past the table (undefined).  DEFINED HERE:
// sweep 8+b scans exactly like sweep b and visits the same four predecessors, but takes them in an order that
// depends on the parity of the scan coordinates, so that the k-th neighbour chains follow the eight knight-move
// (22.5 degree) directions -- the first neighbour alternates between the axis direction and the adjacent diagonal:
//   base sweeps 0-3:  k0 = xs odd ? (-1,-1) : (-1,0)    net (-2,-1) per two pixels
//                     k1 = ys odd ? (+1,-1) : (0,-1)    net (+1,-2), perpendicular to k0
//                     k2, k3 = the other member of {(-1,0),(-1,-1)} resp. {(0,-1),(+1,-1)}
//   base sweeps 4-7:  k0 = ys odd ? (0,-1) : (+1,-1)    net (+1,-2)
//                     k1 = xs odd ? (-1,0) : (-1,-1)    net (-2,-1)
//                     k2, k3 = the other member of {(+1,-1),(0,-1)} resp. {(-1,-1),(-1,0)}
// In image space the first-neighbour chains of sweeps 8..15 run along (-2,-1) (2,1) (-1,2) (1,-2) (-1,-2) (2,-1)
// (1,2) (-2,1).  Border rule, weights (the plane of the neighbour's offset, read at the pixel) and everything else
// are as for sweeps 0-7.  oracle/mgm_oracle.c restates the same definition; parity against the reference is
// necessarily unpinned for these sweeps.
__host__ __device__ inline int knight_pred_type(bool diag_base, int k, int xs, int ys) {
   const int xo = xs & 1, yo = ys & 1;
   if (!diag_base) {
      switch (k) {
      case 0: return xo ? PRED_UPL : PRED_SAME;
      case 1: return yo ? PRED_UPR : PRED_UP;
      case 2: return xo ? PRED_SAME : PRED_UPL;
      default: return yo ? PRED_UP : PRED_UPR;
      }
   }
   switch (k) {
   case 0: return yo ? PRED_UP : PRED_UPR;
   case 1: return xo ? PRED_SAME : PRED_UPL;
   case 2: return yo ? PRED_UPR : PRED_UP;
   default: return xo ? PRED_UPL : PRED_SAME;
   }
}

// weight plane of neighbour k (0..3) for each sweep 0-7, read AT the pixel (mgm_core.cc:481-484,550-554)
__host__ __device__ inline int pass_weight_plane(int pass, int k) {
   const unsigned tab[4] = {0x76543210u, 0x47651023u, 0x02135764u, 0x30216475u};
   return (tab[k] >> (4 * pass)) & 0xF;
}
// weight plane of the predecessor of scan-space type `pt` for base sweep 0-7: sweeps 0-3 list their neighbours in the
// order SAME, UP, UPL, UPR (k = pt), sweeps 4-7 in the order UPR, UPL, UP, SAME (k = 3 - pt)
__host__ __device__ inline int pass_weight_plane_of_type(int base_pass, int pt) {
   return pass_weight_plane(base_pass, base_pass < 4 ? pt : 3 - pt);
}

}  // namespace mgm
