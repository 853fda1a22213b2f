// tcgen05 / TMEM edge kernel: the second Linear of every edge / coord MLP (85-91 % of the path's FLOPs,
// SURVEY.md section 8(d)) on the 5th-generation tensor cores, fused with its producer (first Linear + SiLU) and its
// consumer (bias + SiLU + edge weight + segment sum), so no per-edge tensor ever reaches HBM or L2.
//
// Numerics: fp32 operands are split into fp16 hi + lo (x = hi + lo, |lo| <= 2^-11 |x|) and the product is
// accumulated as  W_hi s_hi + W_hi s_lo + W_lo s_hi  in fp32 TMEM accumulators ("3xFP16", ~2^-22 relative
// per product) -- fp32-grade, which is what the 1e-4 end-to-end tolerance needs; plain bf16/tf32 operands
// do not meet it.  Range: W is pre-scaled per matrix, and every edge's activation row per tile, by exact powers
// of two chosen from a-priori bounds so that |operand| < 2^14 (fp16 tops out at 65504 and diverging samples do
// exceed it); the combined descale is folded into the epilogue's first FMA, so scaling costs no accuracy.
//
// Layouts (no swizzle, K-major, canonical "interleave" form, cute/atom/mma_traits_sm100.hpp):
//   operand tile in smem = [kc 0..15][row][8 halves]   (kc = 16-byte chunk along K=128)
//   core matrix = 8 rows x 16 B contiguous (SBO = 128 B between 8-row groups, LBO = slab pitch between kc)
// GCL  (swapped operands): D[c, e] = sum_k W2[c,k] s[e,k]  -> TMEM lane = output channel, column = edge;
//        each epilogue thread owns one channel and walks the edges of a row sequentially, so the segment sum
//        over j is an in-register, order-deterministic reduction (no shuffles, no atomics).
// COORD (natural operands): D[e, c]: TMEM lane = edge, columns = channels; the thread reduces over channels
//        with w5 in registers (coord_mlp.4), then a 3-wide segment sum through shared memory.
#pragma once
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "../../include/difflinker_b200.h"
#include "kernels_simt.cuh"

namespace dl {
namespace tc {

constexpr bool AVAILABLE = true;
constexpr int TN = 128;                    // edges per tile (UMMA N for the GCL, UMMA M for the coord variant)
constexpr int MAXR = 32;                   // rows per tile
constexpr int KC = 16;                     // 16-byte chunks per K=128 row of fp16
constexpr int W_LBO = H * 16;              // 2048 B: W slab pitch
constexpr int B_LBO = TN * 16 + 16;        // 2064 B: activation slab pitch (+16 B: conflict-free producer stores)
constexpr int SBO = 128;
constexpr int W_BYTES = KC * W_LBO;        // 32 KB per fp16 copy
constexpr int B_BYTES = KC * B_LBO;        // 33,024 B per fp16 copy
constexpr float F16_TARGET = 16384.0f;     // operands are scaled by exact powers of two to stay below 2^14

constexpr int N_STAGE = 2;                 // activation operand stages in shared memory
constexpr int N_ACC = 4;                   // TMEM accumulator stages (4 x 128 columns) = table ring depth

// warp roles of k_edge_tc: 28 warps = 7 warpgroups; registers are re-balanced per warpgroup with setmaxnreg
// (launch: 72/thread; producers grow to 88, epilogue shrinks to 56, the MMA/table group to 40 -- 64,512 in total).
constexpr int W_EPI = 0;                   // warps 0-7   (WG 0,1): epilogue (warp % 4 = TMEM lane quarter, warp / 4 = row parity)
constexpr int N_EPI_WARPS = 8;
constexpr int W_MMA = 8;                   // warp  8     (WG 2)  : MMA issuer (one thread), TMEM alloc/dealloc
constexpr int W_TBL = 9;                   // warps 9-11  (WG 2)  : per-edge table builders (run ahead of everyone)
constexpr int N_TBL_WARPS = 3;
constexpr int W_PROD = 12;                 // warps 12-27 (WG 3-6): producers (first Linear + SiLU -> fp16 operand tile)
constexpr int N_PROD_WARPS = 16;
constexpr int EDGE_TC_THREADS = 32 * (W_PROD + N_PROD_WARPS);   // 896
constexpr int REGS_EPI = 48, REGS_CTRL = 56, REGS_PROD = 88;

// shared memory map (bytes from a 1024-aligned base)
constexpr int OFF_WHI = 0;
constexpr int OFF_WLO = OFF_WHI + W_BYTES;
constexpr int OFF_ST = OFF_WLO + W_BYTES;                // N_STAGE x [hi | lo]
constexpr int STAGE_BYTES = 2 * B_BYTES;
constexpr int OFF_TBL = OFF_ST + N_STAGE * STAGE_BYTES;  // N_ACC x per-tile tables
constexpr int TBL_HDR = 0;                               // int [16]: Et, nrt, ncc, flags(first|last<<1), gb_lo, gb_hi, end
constexpr int TBL_ROWOFF = 64;                           // int  [TN]  AB offset (floats) of the edge's row node
constexpr int TBL_COLOFF = TBL_ROWOFF + TN * 4;          // int  [TN]  AB offset of the column node's B half
constexpr int TBL_D = TBL_COLOFF + TN * 4;               // f32  [TN]  |x_i-x_j|^2 of this block
constexpr int TBL_D0 = TBL_D + TN * 4;                   // f32  [TN]  |x0_i-x0_j|^2 of the call's input
constexpr int TBL_SC = TBL_D0 + TN * 4;                  // f32  [TN]  power-of-two scale of the edge's fp16 operand row
constexpr int TBL_EM = TBL_SC + TN * 4;                  // f32x2[TN]  (edge weight, accumulator descale)
constexpr int TBL_CD = TBL_EM + TN * 8;                  // f32  [TN][3] normalised difference (COORD)
constexpr int TBL_ROWNODE = TBL_CD + TN * 12;            // int  [MAXR] node index of each tile row
constexpr int TBL_ROWSTART = TBL_ROWNODE + MAXR * 4;     // int  [MAXR+1] first tile column of each row (neighbour-list tiles)
constexpr int TBL_BYTES = TBL_ROWSTART + (MAXR + 4) * 4;
constexpr int OFF_B2W5 = OFF_TBL + N_ACC * TBL_BYTES;    // float2 [128] (b2, w5)
constexpr int OFF_TX = OFF_B2W5 + H * 8;                 // f32 [TN][3] per-edge translation (COORD epilogue)
constexpr int OFF_BAR = OFF_TX + TN * 12;                // mbarriers + tmem ptr
constexpr int BAR_W = 0, BAR_FULL = 8, BAR_EMPTY = BAR_FULL + 8 * N_STAGE, BAR_TBL = BAR_EMPTY + 8 * N_STAGE,
              BAR_TFULL = BAR_TBL + 8 * N_ACC, BAR_TEMPTY = BAR_TFULL + 8 * N_ACC, BAR_TMEMSLOT = BAR_TEMPTY + 8 * N_ACC;
constexpr int SMEM_BYTES = OFF_BAR + BAR_TMEMSLOT + 16 + 1024;   // + alignment slack
static_assert(SMEM_BYTES <= 232448, "k_edge_tc exceeds the 227 KB of shared memory a CTA can opt into");
static_assert(TBL_BYTES % 16 == 0, "table slots must keep 16-byte alignment");

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must surface as a launch error, never as a hung GPU.
// How a waiting thread waits (set once per process from DL_WAIT_MODE, experiment switch):
//   0  bare try_wait loop: a failed try_wait returns after a few cycles on this part, so a waiting warp keeps issuing
//      (try_wait, select, compare, branch) and competes with the warps it is waiting FOR: a third of all instructions the
//      edge kernel issued were such polls
//   1  try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the time limit passes
//   2  try_wait, then nanosleep(32) between polls
__device__ __constant__ int c_wait_mode = 1;
constexpr uint32_t WAIT_HINT_NS = 20000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const int mode = c_wait_mode;
  if (mode == 1) {
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity), "r"(WAIT_HINT_NS)
          : "memory");
      if (done) return;
    }
    __trap();
  }
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (mode == 2) __nanosleep(32);
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Same, for roles that run far ahead of (or lag behind) the critical path: back off between polls so the spin does
// not steal issue slots from the producer / epilogue warps sharing the scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    __nanosleep(64);
  }
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 1-D bulk copy global -> shared through the TMA engine, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=SWIZZLE_NONE [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32 [4,6)=1, A=B=f16 (0), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from TENSOR MEMORY (128 lanes = M rows; one 32-bit column holds two consecutive K
// elements of a row, so a K=16 step spans 8 columns): a stationary weight matrix then costs no shared-memory space and
// no shared-memory read bandwidth per MMA.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: lane l of warp w writes 16 consecutive columns of TMEM lane 32 * (w % 4) + l
#define TMEM_ST_X16(taddr, r)                                                                                     \
  asm volatile(                                                                                                   \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"    \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),       \
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])             \
      : "memory")
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2-D tiled TMA load (cp.async.bulk.tensor, SASS: UTMALDG): box at (c0 = innermost coordinate, c1 = row) of the tensor
// described by `tmap` -> dense [rows][inner] box in shared memory, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// add to the barrier's pending transaction bytes without arriving (the arrive follows once the tables are written)
__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Warp-convergent issue: ALL 32 lanes of the issuing warp execute this with identical (warp-uniform) operands and exactly one
// elected lane issues the MMA. Issued from a divergent `if (lane == 0)` region instead, every tcgen05.mma is wrapped by the
// compiler in an ELECT / BRA.U.ANY loop and its descriptor arithmetic runs on the per-thread datapath.
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_LD_X8(taddr, r)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                        \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))
#define TMEM_LD_X16(taddr, r)                                                                                    \
  asm volatile(                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"   \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),          \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])     \
      : "r"(taddr))
#define TMEM_LD_X4(taddr, r) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr))
#define TMEM_LD_X2(taddr, r) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr))
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 pair -> fp16x2 hi and lo words (packed subtract)
__device__ __forceinline__ void split2v(float2 v, uint32_t& hi, uint32_t& lo) {
  __half2 h = __float22half2_rn(v);
  float2 back = __half22float2(h);
  __half2 l = __float22half2_rn(__fadd2_rn(v, make_float2(-back.x, -back.y)));
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
// fp32 -> fp16 hi/lo pair of two values
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(a, b);
  float2 back = __half22float2(h);
  __half2 l = __floats2half2_rn(a - back.x, b - back.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
// ---------------------------------------------------------------------------------------------------------
// Tile iteration (table warps only): a CTA walks its work items (static round-robin) -> row groups ->
// 128-column chunks. A tile is whole rows x all live columns (or one row x a 128-column chunk when nc > 128), so
// the segment sum over j never crosses CTAs.
// ---------------------------------------------------------------------------------------------------------
struct Tile {
  int b, nc, slot0, nrt, c0, ncc;
  bool first_chunk, last_chunk;
  const int* rows;
  // neighbour-list (SPARSE) tiles only; per-lane values: lane l describes tile row l
  int Et;        // edges in the tile (uniform)
  int start;     // first tile column of row l (exclusive prefix of the row degrees)
  int node;      // node index of row l
};

// Warp-synchronous (all 32 lanes of a table warp call it together): the CTA's contiguous slice of the work-item
// list is read 32 items at a time with one coalesced load and served from registers by shuffles, so walking the
// list costs no dependent global round trip per tile (items carry their molecule's live column count).
template <bool COORD>
struct TileIter {
  const int4* list;
  const int* rowlist;
  int N, wi_end, wi, rt, c0, cache_base, lane;
  int4 cache;
  __device__ TileIter(const Plan& p, int N_) : N(N_), rt(0), c0(0), cache_base(-(1 << 30)), lane(threadIdx.x & 31) {
    // blocked distribution: CTA c owns the contiguous work items [lo, hi) -- consecutive tiles then mostly belong to
    // the same molecule, so the table warps' and producers' L2 lines are reused while they are hot.
    list = COORD ? p.xitems : p.items;
    rowlist = COORD ? p.xrowidx : p.rowidx;
    const int total = COORD ? *p.n_xitems : *p.n_items;
    const int per = total / (int)gridDim.x, extra = total % (int)gridDim.x, c = (int)blockIdx.x;
    wi = c * per + min(c, extra);
    wi_end = wi + per + (c < extra ? 1 : 0);
    cache = make_int4(0, 0, 0, 0);
  }
  __device__ bool next(Tile& t) {
    while (wi < wi_end) {
      if (wi - cache_base >= 32) {                        // refill (warp-uniform)
        cache_base = wi;
        cache = list[min(wi + lane, wi_end - 1)];
      }
      const int src = wi - cache_base;
      const int b = __shfl_sync(0xffffffffu, cache.x, src), r_begin = __shfl_sync(0xffffffffu, cache.y, src);
      const int r_count = __shfl_sync(0xffffffffu, cache.z, src), nc = __shfl_sync(0xffffffffu, cache.w, src);
      int per = nc >= TN ? 1 : TN / max(nc, 1);
      if (per > MAXR) per = MAXR;
      if (rt >= r_count || nc <= 0) { wi += 1; rt = 0; c0 = 0; continue; }
      t.b = b; t.nc = nc; t.slot0 = r_begin + rt; t.nrt = min(per, r_count - rt);
      t.c0 = c0; t.ncc = min(TN, nc - c0);
      t.first_chunk = c0 == 0; t.last_chunk = c0 + TN >= nc;
      t.rows = rowlist + (size_t)b * N;
      c0 += TN;
      if (c0 >= nc) { c0 = 0; rt += per; }
      return true;
    }
    return false;
  }
};

// Cut-off graphs: the CTA walks the tile records k_nbr packed for this call (round-robin over CTAs: records cost one
// tile, or ceil(degree / 128) chunk tiles for a row with more than 128 neighbours -- its sum is carried in the epilogue's
// registers, so the chunks stay on one CTA). Warp-synchronous; lane l holds int l of the 128-byte record and the
// next record is prefetched while the current one is served.
struct RecIter {
  const int* recs;
  int n, k, G, lane, cur, nxt, chunk;
  __device__ RecIter(const int* recs_, const int* n_recs) : recs(recs_), lane(threadIdx.x & 31), cur(0), chunk(0) {
    n = *n_recs; G = (int)gridDim.x; k = (int)blockIdx.x;
    nxt = k < n ? recs[(size_t)k * CUT_REC + lane] : 0;
  }
  __device__ bool next(Tile& t) {
    if (chunk == 0) {
      if (k >= n) return false;
      cur = nxt;
      const int kn = k + G;
      nxt = kn < n ? recs[(size_t)kn * CUT_REC + lane] : 0;
    }
    const int w1 = __shfl_sync(0xffffffffu, cur, 1), v2 = __shfl_sync(0xffffffffu, cur, 2);
    const int info = __shfl_sync(0xffffffffu, cur, (4 + lane) & 31);
    t.b = __shfl_sync(0xffffffffu, cur, 0);
    t.nc = 0; t.slot0 = 0; t.rows = nullptr;
    if (((w1 >> 8) & 1) == 0) {
      t.nrt = w1 & 0xff; t.Et = v2; t.ncc = v2; t.c0 = 0; t.first_chunk = true; t.last_chunk = true;
      t.node = info & 0xffff; t.start = info >> 16;
      k += G;
    } else {
      const int deg = v2, c0 = chunk * TN;
      t.nrt = 1; t.Et = min(TN, deg - c0); t.ncc = t.Et; t.c0 = c0;
      t.first_chunk = chunk == 0; t.last_chunk = c0 + TN >= deg;
      t.node = __shfl_sync(0xffffffffu, cur, 4) & 0xffff;
      t.start = lane == 0 ? 0 : t.Et;
      ++chunk;
      if (t.last_chunk) { chunk = 0; k += G; }
    }
    return true;
  }
};

template <bool COORD, bool SPARSE>
__device__ __forceinline__ typename std::conditional<SPARSE, RecIter, TileIter<COORD>>::type make_iter(const EdgeArgs& a, int N) {
  if constexpr (SPARSE) return RecIter(a.recs, a.n_recs);
  else return TileIter<COORD>(a.plan, N);
}

// ---------------------------------------------------------------------------------------------------------
// The kernel: persistent, 1 CTA / SM, 17 warps in four roles connected by mbarrier rings.
//
//   table warps (4)  : walk the CTA's tile list; per edge (i,j): node offsets, d_ij, d0_ij, edge weight, operand
//                      scale, normalised difference -> table ring slot a = t % 4          [tbl_full[a]]
//   producers (8)    : A_i + B_j + d w_d + d0 w_0 -> SiLU -> fp16 hi/lo operand tile, stage s = t % 2
//                      (register prefetch of the next item's A/B rows)                      [full[s]]
//   MMA issuer (1 thr): 24 x tcgen05.mma (3xFP16, K=128) into TMEM accumulator a; commits     [empty[s], tfull[a]]
//   epilogue (4)     : tcgen05.ld accumulator a, bias + SiLU + edge weight + segment sum        [tempty[a]]
//
// Every role works on a different tile at any moment, so L2 latency (producers), TMEM latency (epilogue) and
// the tensor pipe overlap; the two SiLUs per edge-channel (MUFU) are the shared bottleneck by design.
// ---------------------------------------------------------------------------------------------------------
// PROF = true adds clock64() accounting per role (wait vs work cycles) into `prof` (16 x u64 per CTA); debug only.
// SPARSE = true: cut-off (pocket) graphs -- tiles are packed from the per-row neighbour lists k_nbr built for this
// forward call, so only edges the reference creates are processed (egnn.py:554-596); FC graphs use SPARSE = false.
// SPLIT = true: single-row tiles (chunks of rows longer than a tile: N > 64 on FC graphs, long rows of cut-off graphs)
// are split column-wise between the two epilogue warp halves; the SPLIT = false instantiation keeps the leaner epilogue for
// workloads whose tiles are (almost) all multi-row (N <= 64: the headline config).
template <bool COORD, bool PROF = false, bool SPARSE = false, bool SPLIT = SPARSE>
__global__ void __launch_bounds__(EDGE_TC_THREADS, 1) k_edge_tc(Geom gm, EdgeArgs a, const __half* __restrict__ w2tc,
                                                                unsigned long long* __restrict__ prof = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (no integer round trip) so accesses compile to LDS/STS
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = gm.N;

  const uint32_t bars = sbase + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + BAR_TMEMSLOT);
  unsigned long long pc[4] = {0, 0, 0, 0};               // PROF: [0..2] wait cycles by barrier kind, [3] tiles
  const long long t_begin = PROF ? clock64() : 0;
  auto wait_on = [&](uint32_t bar, uint32_t parity, int slot) {
    if (PROF) { const long long c0 = clock64(); mbar_wait(bar, parity); pc[slot] += (unsigned long long)(clock64() - c0); }
    else mbar_wait(bar, parity);
  };
  auto wait_relaxed = [&](uint32_t bar, uint32_t parity, int slot) {
    if (PROF) { const long long c0 = clock64(); mbar_wait_relaxed(bar, parity); pc[slot] += (unsigned long long)(clock64() - c0); }
    else mbar_wait_relaxed(bar, parity);
  };
  auto prof_flush = [&](int base) {
    if (PROF && lane == 0) {
      unsigned long long* o = prof + (size_t)blockIdx.x * 16 + base;
      o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = (unsigned long long)(clock64() - t_begin);
    }
  };
  float2* b2w5 = reinterpret_cast<float2*>(sm + OFF_B2W5);

  if (tid == 0) {
    mbar_init(bars + BAR_W, 1);
    for (int i = 0; i < N_STAGE; ++i) { mbar_init(bars + BAR_FULL + 8 * i, N_PROD_WARPS); mbar_init(bars + BAR_EMPTY + 8 * i, 1); }
    for (int i = 0; i < N_ACC; ++i) {
      mbar_init(bars + BAR_TBL + 8 * i, 1);
      mbar_init(bars + BAR_TFULL + 8 * i, 1);
      mbar_init(bars + BAR_TEMPTY + 8 * i, COORD ? 4 : N_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (tid < H) b2w5[tid] = make_float2(a.b2[tid], COORD ? a.w5[tid] : 0.f);
  if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Register re-balancing happens at the top of each role branch (warpgroup-aligned: every warp of a group runs the
  // same setmaxnreg; ptxas allocates each branch against the budget set by the instruction that dominates it).
  if (warp >= W_MMA && warp < W_PROD) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
  if (warp >= W_TBL && warp < W_TBL + N_TBL_WARPS) {
    // =================================== table warps ===================================================================
    // Tile-parallel: table warp k builds the tables of tiles t = k, k+3, ... on its own (4 edges per lane), so the
    // three warps overlap their dependent L2 round trips (index -> coordinates / maxima / mask) across tiles.
    const int tw = warp - W_TBL;
    typename std::conditional<SPARSE, RecIter, TileIter<COORD>>::type iter = make_iter<COORD, SPARSE>(a, N);
    Tile cur;
    for (int t = 0;; ++t) {
      const int acc = t & (N_ACC - 1);
      const bool more = iter.next(cur);
      if (t % N_TBL_WARPS != tw) { if (!more) break; continue; }
      if (t >= N_ACC) wait_relaxed(bars + BAR_TEMPTY + 8 * acc, ((t - N_ACC) / N_ACC) & 1, 0);   // slot's previous tile fully consumed
      uint8_t* tb = sm + OFF_TBL + acc * TBL_BYTES;
      int* hdr = reinterpret_cast<int*>(tb + TBL_HDR);
      if (more) {
        const int Et = SPARSE ? cur.Et : cur.nrt * cur.ncc;
        const size_t gb = (size_t)cur.b * N;
        if (lane == 0) {
          hdr[0] = Et; hdr[1] = cur.nrt; hdr[2] = cur.ncc;
          hdr[3] = (cur.first_chunk ? 1 : 0) | (cur.last_chunk ? 2 : 0);
          hdr[4] = cur.b;
        }
        if (SPARSE) {
          if (lane < cur.nrt) {
            reinterpret_cast<int*>(tb + TBL_ROWNODE)[lane] = cur.node;
            reinterpret_cast<int*>(tb + TBL_ROWSTART)[lane] = cur.start;
          }
          if (lane == 0) reinterpret_cast<int*>(tb + TBL_ROWSTART)[cur.nrt] = Et;
        } else if (lane < cur.nrt) {
          reinterpret_cast<int*>(tb + TBL_ROWNODE)[lane] = cur.rows[cur.slot0 + lane];
        }
        bool any_rescale = false;
#pragma unroll
        for (int e = lane; e < TN; e += 32) {
          const int ev = min(e, Et - 1);                   // slots past Et mirror the last edge: producers may prefetch them
          int i, j;
          float ew_list = 1.f;
          if (SPARSE) {
            int rr = 0;                                    // tile row of this edge: rows start at ascending columns
            for (int r = 1; r < cur.nrt; ++r) rr += ev >= __shfl_sync(0xffffffffu, cur.start, r) ? 1 : 0;
            i = __shfl_sync(0xffffffffu, cur.node, rr);
            const int first_col = __shfl_sync(0xffffffffu, cur.start, rr);
            const int raw = a.nbr[(gb + i) * N + cur.c0 + (ev - first_col)];
            j = raw & 0x7fffffff;
            ew_list = raw < 0 ? 0.f : 1.f;                 // padding edge of an isolated row
          } else {
            const int rr = ev / cur.ncc, jj = ev - rr * cur.ncc;
            i = cur.rows[cur.slot0 + rr];
            j = a.plan.colidx[gb + cur.c0 + jj];
          }
          const float4 xi = a.x4[gb + i], xj = a.x4[gb + j], yi = a.x04[gb + i], yj = a.x04[gb + j];   // 16-byte gathers
          const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
          const float d = dx * dx + dy * dy + dz * dz;                       // egnn.py:297-298
          const float ex = yi.x - yj.x, ey = yi.y - yj.y, ez = yi.z - yj.z;
          const float d0 = ex * ex + ey * ey + ez * ez;                      // egnn.py:220
          int ci = 0, cj = 0;
          if (!SPARSE && gm.graph_type != 0) { ci = a.cls[gb + i]; cj = a.cls[gb + j]; }
          reinterpret_cast<int*>(tb + TBL_ROWOFF)[e] = (int)((gb + i) * 2 * H);
          reinterpret_cast<int*>(tb + TBL_COLOFF)[e] = (int)((gb + j) * 2 * H + H);
          reinterpret_cast<float*>(tb + TBL_D)[e] = d;
          reinterpret_cast<float*>(tb + TBL_D0)[e] = d0;
          // |silu(pre)| <= |pre| <= max|A_i| + max|B_j| + d max|wd| + d0 max|w0|: exact power-of-two scale that keeps the
          // fp16 hi/lo operands of this edge below 2^14 (activations of diverging samples exceed fp16's 65504).
          const float bound = a.ABmax[(gb + i) * 2] + a.ABmax[(gb + j) * 2 + 1] + d * a.wdmax + d0 * a.w0max;
          float sc = 1.0f;
          if (!(bound <= F16_TARGET)) {
            const int ex2 = ((__float_as_int(bound) >> 23) & 0xff) - 127;
            sc = __int_as_float(max(127 + 13 - ex2, 1) << 23);
          }
          reinterpret_cast<float*>(tb + TBL_SC)[e] = sc;
          any_rescale |= (sc != 1.0f);
          // GCL epilogue works in the log2 domain: u = -log2(e) * (D*descale + b2) is one FMA, sigmoid = 1/(1+2^u), and
          // the -ln(2) that turns u*sigmoid back into silu rides on the edge weight (one rounding of a constant).
          const float ew = SPARSE ? ew_list
                                  : edge_weight(gm.graph_type, a.edge_mask ? a.edge_mask + gb * N : nullptr, N, i, j, ci, cj, d0);
          reinterpret_cast<float2*>(tb + TBL_EM)[e] =
              COORD ? make_float2(ew, a.w2_descale / sc)
                    : make_float2(ew * -0.6931471805599453f, (a.w2_descale / sc) * -1.4426950408889634f);
          if (COORD) {
            const float inv = 1.0f / (sqrtf(d + 1e-8f) + gm.norm_constant);  // egnn.py:299-300
            float* cds = reinterpret_cast<float*>(tb + TBL_CD);
            cds[e * 3 + 0] = dx * inv; cds[e * 3 + 1] = dy * inv; cds[e * 3 + 2] = dz * inv;
          }
        }
        any_rescale = __any_sync(0xffffffffu, any_rescale);
        if (lane == 0) hdr[5] = any_rescale ? 1 : 0;       // tile-level flag: producers skip the scale multiply
      } else if (lane == 0) {
        hdr[0] = 0;                                      // end marker travels through the whole pipeline
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + BAR_TBL + 8 * acc);
      if (!more) break;
    }
    if (warp == W_TBL) prof_flush(0);
  } else if (warp == W_MMA) {
    // =================================== MMA issuer =====================================================================
    if (lane == 0) {
      mbar_expect_tx(bars + BAR_W, 2 * W_BYTES);           // W2 hi|lo tiles: one 64 KB TMA bulk copy, resident for the launch
      bulk_g2s(sbase + OFF_WHI, w2tc, 2 * W_BYTES, bars + BAR_W);
      mbar_wait(bars + BAR_W, 0);
      const uint32_t whi = sbase + OFF_WHI, wlo = sbase + OFF_WLO;
      for (int t = 0;; ++t) {
        const int acc = t & (N_ACC - 1), s = t & (N_STAGE - 1);
        wait_relaxed(bars + BAR_FULL + 8 * s, (t / N_STAGE) & 1, 0);
        const int Et = reinterpret_cast<const int*>(sm + OFF_TBL + acc * TBL_BYTES + TBL_HDR)[0];
        if (Et <= 0) { mbar_arrive(bars + BAR_TFULL + 8 * acc); break; }
        if (t >= N_ACC) wait_on(bars + BAR_TEMPTY + 8 * acc, ((t - N_ACC) / N_ACC) & 1, 1);   // epilogue(t-4) drained this accumulator
        tc_fence_after();
        const uint32_t bhi = sbase + OFF_ST + s * STAGE_BYTES, blo = bhi + B_BYTES;
        const uint32_t dcol = tmem + acc * TN;
        if (!COORD) {
          const uint32_t idesc = umma_idesc(128, max(16, (Et + 15) & ~15));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a_hi = umma_desc(whi + ks * 2 * W_LBO, W_LBO, SBO), a_lo = umma_desc(wlo + ks * 2 * W_LBO, W_LBO, SBO);
            const uint64_t b_hi = umma_desc(bhi + ks * 2 * B_LBO, B_LBO, SBO), b_lo = umma_desc(blo + ks * 2 * B_LBO, B_LBO, SBO);
            umma_f16(dcol, a_lo, b_hi, idesc, ks > 0);
            umma_f16(dcol, a_hi, b_lo, idesc, 1);
            umma_f16(dcol, a_hi, b_hi, idesc, 1);
          }
        } else {
          const uint32_t idesc = umma_idesc(128, 128);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t a_hi = umma_desc(bhi + ks * 2 * B_LBO, B_LBO, SBO), a_lo = umma_desc(blo + ks * 2 * B_LBO, B_LBO, SBO);
            const uint64_t b_hi = umma_desc(whi + ks * 2 * W_LBO, W_LBO, SBO), b_lo = umma_desc(wlo + ks * 2 * W_LBO, W_LBO, SBO);
            umma_f16(dcol, a_lo, b_hi, idesc, ks > 0);
            umma_f16(dcol, a_hi, b_lo, idesc, 1);
            umma_f16(dcol, a_hi, b_hi, idesc, 1);
          }
        }
        umma_commit(bars + BAR_EMPTY + 8 * s);             // operand stage reusable once these MMAs have read it
        umma_commit(bars + BAR_TFULL + 8 * acc);           // accumulator ready for the epilogue
        pc[2] += 1;
      }
      prof_flush(8);
    }
  }
  } else if (warp >= W_PROD) {
    // =================================== producers =====================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    const int pw = warp - W_PROD;
    const int kc = lane & 15, esub = lane >> 4;            // this thread always owns k = kc*8 .. kc*8+7
    float2 wdr[4], w0r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      wdr[q] = make_float2(__ldg(a.wd + kc * 8 + 2 * q), __ldg(a.wd + kc * 8 + 2 * q + 1));
      w0r[q] = make_float2(__ldg(a.w0 + kc * 8 + 2 * q), __ldg(a.w0 + kc * 8 + 2 * q + 1));
    }
    for (int t = 0;; ++t) {
      const int acc = t & (N_ACC - 1), s = t & (N_STAGE - 1);
      wait_on(bars + BAR_TBL + 8 * acc, (t / N_ACC) & 1, 0);
      const uint8_t* tb = sm + OFF_TBL + acc * TBL_BYTES;
      const int Et = reinterpret_cast<const int*>(tb + TBL_HDR)[0];
      if (t >= N_STAGE) wait_on(bars + BAR_EMPTY + 8 * s, ((t - N_STAGE) / N_STAGE) & 1, 1);   // MMA(t-2) has read this stage
      if (Et > 0) {
        const int* rowoff = reinterpret_cast<const int*>(tb + TBL_ROWOFF);
        const int* coloff = reinterpret_cast<const int*>(tb + TBL_COLOFF);
        const float* dv = reinterpret_cast<const float*>(tb + TBL_D);
        const float* d0v = reinterpret_cast<const float*>(tb + TBL_D0);
        const float* scv = reinterpret_cast<const float*>(tb + TBL_SC);
        uint8_t* bhi = sm + OFF_ST + s * STAGE_BYTES + kc * B_LBO;
        uint8_t* blo = bhi + B_BYTES;
        const bool rescale = reinterpret_cast<const int*>(tb + TBL_HDR)[5] != 0;
        constexpr int ITEMS = TN / (2 * N_PROD_WARPS);       // 8 edges per thread per tile
        // Two B_j chunks are kept in flight in registers (L1 is ~5 KB next to 222 KB of shared memory, so these
        // loads are L2 round trips). Table slots past Et mirror the last edge: no clamping here.
        // A thread takes ITEMS consecutive edges of the tile: they share the row (almost always), so the A_i chunk is
        // loaded once and only the B_j chunks stream -- L2->SM traffic of this role drops from 64 to ~40 bytes/item.
        float4 pa0, pa1, pb[2][2];
        const int e_base = ITEMS * (2 * pw + esub);
        int ro_cur = rowoff[e_base];
        {
          const float* ap = a.AB + ro_cur + kc * 8;
          pa0 = __ldg(reinterpret_cast<const float4*>(ap)); pa1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
        }
#pragma unroll
        for (int pf = 0; pf < 2; ++pf) {
          const float* bp = a.AB + coloff[e_base + pf] + kc * 8;
          pb[pf][0] = __ldg(reinterpret_cast<const float4*>(bp)); pb[pf][1] = __ldg(reinterpret_cast<const float4*>(bp + 4));
        }
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int e = e_base + it;
          const int slot = it & 1;
          const float4 b0 = pb[slot][0], b1 = pb[slot][1];
          if (it + 2 < ITEMS) {
            const float* bp = a.AB + coloff[e + 2] + kc * 8;
            pb[slot][0] = __ldg(reinterpret_cast<const float4*>(bp)); pb[slot][1] = __ldg(reinterpret_cast<const float4*>(bp + 4));
          }
          const int ro = rowoff[e];
          if (ro != ro_cur) {                              // row boundary inside this thread's run (rare): reload A_i
            ro_cur = ro;
            const float* ap = a.AB + ro + kc * 8;
            pa0 = __ldg(reinterpret_cast<const float4*>(ap)); pa1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
          }
          const float4 a0 = pa0, a1 = pa1;
          if (e < Et) {
            const float d = dv[e], d0 = d0v[e];
            const float2 dd = make_float2(d, d), dd0 = make_float2(d0, d0);
            const float2 av[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
            const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
            float2 sv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)                     // egnn.py:49-50, two channels per packed instruction
              sv[q] = usig2(__ffma2_rn(dd0, w0r[q], __ffma2_rn(dd, wdr[q], __fadd2_rn(av[q], bv[q]))));   // log2 domain (see pack_w2)
            if (rescale) {                                 // rare: diverging samples only (tile-uniform)
              const float sc = scv[e];
#pragma unroll
              for (int q = 0; q < 4; ++q) sv[q] = __fmul2_rn(sv[q], make_float2(sc, sc));
            }
            uint4 hi, lo;
            split2v(sv[0], hi.x, lo.x); split2v(sv[1], hi.y, lo.y);
            split2v(sv[2], hi.z, lo.z); split2v(sv[3], hi.w, lo.w);
            *reinterpret_cast<uint4*>(bhi + e * 16) = hi;
            *reinterpret_cast<uint4*>(blo + e * 16) = lo;
          }
        }
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + BAR_FULL + 8 * s);
      if (Et <= 0) break;
    }
    if (warp == W_PROD) prof_flush(4);
  } else {
    // =================================== epilogue warps ==================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    const int q = warp & 3;                                // TMEM lane quarter of this warp
    const int hw = warp >> 2;                              // GCL: this warp takes the tile rows with (rr + t) % 2 == hw
    if (COORD && hw != 0) goto edge_tc_done;               // coord variant: lanes are edges, 4 warps cover the tile
    float* txs = reinterpret_cast<float*>(sm + OFF_TX);
    float run = 0.f;   // GCL: row sum carried across column chunks (thread = channel); COORD: (row,dim) running sum
    int xbuf = 0;      // GCL: parity of the half-to-half exchange buffer (reuses the COORD-only txs region)
    for (int t = 0;; ++t) {
      const int acc = t & (N_ACC - 1);
      wait_on(bars + BAR_TBL + 8 * acc, (t / N_ACC) & 1, 0);
      wait_on(bars + BAR_TFULL + 8 * acc, (t / N_ACC) & 1, 1);
      tc_fence_after();
      const uint8_t* tb = sm + OFF_TBL + acc * TBL_BYTES;
      const int* hdr = reinterpret_cast<const int*>(tb + TBL_HDR);
      const int Et = hdr[0];
      if (Et <= 0) break;
      const int nrt = hdr[1], ncc_tile = hdr[2];
      const bool first_chunk = hdr[3] & 1, last_chunk = hdr[3] & 2;
      const size_t gb = (size_t)hdr[4] * N;
      const float2* emds = reinterpret_cast<const float2*>(tb + TBL_EM);
      const int* rownode = reinterpret_cast<const int*>(tb + TBL_ROWNODE);
      const int* rowstart = reinterpret_cast<const int*>(tb + TBL_ROWSTART);
      if (!COORD) {
        const int c = q * 32 + lane;
        const float bias = b2w5[c].x * -1.4426950408889634f;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + acc * TN;
        // Multi-row tiles: the two warp halves take alternate rows. Single-row tiles (a row with more live columns than
        // half a tile, or one 128-column chunk of a longer row): the halves split the row's columns, each carries its own
        // partial sum across the chunks, and they meet once per row through shared memory (fixed order: deterministic).
        const bool single = SPLIT && nrt == 1;
        for (int rr = single ? 0 : (nrt == 1 ? hw : ((hw + t) & 1)); rr < nrt; rr += 2) {   // !SPLIT: single-row tiles on half 0
          float2 acc2 = make_float2((nrt == 1 && !first_chunk) ? run : 0.f, 0.f);   // (even, odd) column partial sums
          int col0 = SPARSE ? rowstart[rr] : rr * ncc_tile;
          int ncc = SPARSE ? rowstart[rr + 1] - col0 : ncc_tile;
          if (single) {
            const int split = min(ncc, (((ncc + 1) >> 1) + 15) & ~15);
            if (hw) { col0 += split; ncc -= split; } else ncc = split;
          }
          int jj = 0;
          for (; jj + 16 <= ncc; jj += 16) {
            uint32_t r[16];
            TMEM_LD_X16(tlane + col0 + jj, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const float2 e0 = emds[col0 + jj + u], e1 = emds[col0 + jj + u + 1];
              const float2 uu = __ffma2_rn(make_float2(__uint_as_float(r[u]), __uint_as_float(r[u + 1])),
                                           make_float2(e0.y, e1.y), make_float2(bias, bias));
              acc2 = __ffma2_rn(usig2(uu), make_float2(e0.x, e1.x), acc2);
            }
          }
          if (ncc - jj >= 8) {
            uint32_t r[8];
            TMEM_LD_X8(tlane + col0 + jj, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const float2 e0 = emds[col0 + jj + u], e1 = emds[col0 + jj + u + 1];
              const float2 uu = __ffma2_rn(make_float2(__uint_as_float(r[u]), __uint_as_float(r[u + 1])),
                                           make_float2(e0.y, e1.y), make_float2(bias, bias));
              acc2 = __ffma2_rn(usig2(uu), make_float2(e0.x, e1.x), acc2);
            }
            jj += 8;
          }
          if (ncc - jj >= 4) {
            uint32_t r[4];
            TMEM_LD_X4(tlane + col0 + jj, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const float2 e0 = emds[col0 + jj + u], e1 = emds[col0 + jj + u + 1];
              const float2 uu = __ffma2_rn(make_float2(__uint_as_float(r[u]), __uint_as_float(r[u + 1])),
                                           make_float2(e0.y, e1.y), make_float2(bias, bias));
              acc2 = __ffma2_rn(usig2(uu), make_float2(e0.x, e1.x), acc2);
            }
            jj += 4;
          }
          for (; jj < ncc; ++jj) {
            const uint32_t r = tmem_ld_x1(tlane + col0 + jj);
            tmem_ld_wait();
            const float2 ed = emds[col0 + jj];
            acc2.x = fmaf(usig_f(fmaf(__uint_as_float(r), ed.y, bias)), ed.x, acc2.x);
          }
          const float accv = acc2.x + acc2.y;
          if (single) {
            run = accv;
            if (last_chunk) {
              float* xch = txs + (xbuf & 1) * H;             // double-buffered: a half may run one row ahead of the other
              ++xbuf;
              if (hw) xch[c] = accv;
              named_sync(4 + q, 64);                         // warps q and q + 4: same TMEM lane quarter, same channels
              if (!hw) a.agg[(gb + rownode[rr]) * H + c] = (accv + xch[c]) / gm.normalization_factor;
            }
          } else {
            if (nrt == 1) run = accv;
            if (last_chunk) a.agg[(gb + rownode[rr]) * H + c] = accv / gm.normalization_factor;   // egnn.py:312-313
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BAR_TEMPTY + 8 * acc);
      } else {
        const int e = q * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + acc * TN;
        if (q * 32 < Et) {                                 // warp-uniform
          const float2 ed = emds[min(e, Et - 1)];
          float2 phi2 = make_float2(0.f, 0.f);
#pragma unroll 1
          for (int c0 = 0; c0 < H; c0 += 16) {
            uint32_t r[16];
            TMEM_LD_X16(tlane + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const float2 w0 = b2w5[c0 + u], w1 = b2w5[c0 + u + 1];
              const float2 v = __ffma2_rn(make_float2(__uint_as_float(r[u]), __uint_as_float(r[u + 1])),
                                          make_float2(ed.y, ed.y), make_float2(w0.x, w1.x));
              phi2 = __ffma2_rn(silu2(v), make_float2(w0.y, w1.y), phi2);                  // coord_mlp.2 + .4
            }
          }
          const float phi = phi2.x + phi2.y;
          if (e < Et) {
            const float* cd = reinterpret_cast<const float*>(tb + TBL_CD) + e * 3;
            const float w = phi * ed.x;                    // egnn.py:107-109
            txs[e * 3 + 0] = cd[0] * w; txs[e * 3 + 1] = cd[1] * w; txs[e * 3 + 2] = cd[2] * w;
          }
        }
        tc_fence_before();
        named_sync(2, 128);                                // all four epilogue warps: txs complete
        const int et = warp * 32 + lane;                   // epilogue-group thread id 0..127
        if (et < nrt * 3) {
          const int rr = et / 3, dim = et - rr * 3;
          float sacc = (nrt == 1 && !first_chunk) ? run : 0.f;
          const int col0 = SPARSE ? rowstart[rr] : rr * ncc_tile;
          const int ncc = SPARSE ? rowstart[rr + 1] - col0 : ncc_tile;
          for (int jj = 0; jj < ncc; ++jj) sacc += txs[(col0 + jj) * 3 + dim];
          if (nrt == 1) run = sacc;
          if (last_chunk) {
            const int i = rownode[rr];
            const float lm = a.linker_mask ? a.linker_mask[gb + i] : 1.f;
            const float xv = a.x[(gb + i) * 3 + dim];
            const float xn = (xv + (sacc / gm.normalization_factor) * lm) * a.nm[gb + i];                 // egnn.py:110-124
            a.x_out[(gb + i) * 3 + dim] = xn;
            reinterpret_cast<float*>(a.x4_out + gb + i)[dim] = xn;
          }
        }
        named_sync(2, 128);                                // txs and the table slot may be reused
        if (lane == 0) mbar_arrive(bars + BAR_TEMPTY + 8 * acc);
      }
    }
    if (warp == W_EPI) prof_flush(12);
  }
edge_tc_done:
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
template <typename K>
inline bool opt_in_smem(K kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
}
inline dl_status configure() {
  const bool ok = opt_in_smem(k_edge_tc<true, false, false>) && opt_in_smem(k_edge_tc<true, false, true>) &&
                  opt_in_smem(k_edge_tc<false, false, false, false>) && opt_in_smem(k_edge_tc<false, false, false, true>) &&
                  opt_in_smem(k_edge_tc<false, false, true>) && opt_in_smem(k_edge_tc<false, true, false, false>) &&
                  opt_in_smem(k_edge_tc<false, true, false, true>) && opt_in_smem(k_edge_tc<false, true, true>);
  return ok ? DL_OK : DL_ERR_CUDA;
}

// Log2-domain first layer (tcgen05 path): the node kernel's projection weights, b1, wd and w0 are pre-multiplied by
// -log2(e) (dl_finalize_weights), so the producers form u = -log2(e) * pre directly and s' = u / (1 + 2^u) =
// -log2(e) * silu(pre) costs no scaling multiply; the -ln(2) that undoes it is folded into this operand
// (W2' = -ln2 * W2, computed in double before the fp16 hi/lo split).
constexpr double NEG_LN2 = -0.6931471805599453094;
constexpr double NEG_LOG2E = -1.4426950408889634074;

// edge_mlp.2 / coord_mlp.2 weight (out=128, in=128, row-major) -> [hi|lo][kc][out][8] fp16, scaled by the power of
// two that puts max|W| in [2^13, 2^14). Returns the offset (in halves) inside `blob`; *descale = 1/scale.
inline size_t pack_w2(const std::vector<float>& W_in, std::vector<__half>& blob, float* descale) {
  std::vector<float> W(W_in.size());
  for (size_t i = 0; i < W.size(); ++i) W[i] = (float)((double)W_in[i] * NEG_LN2);
  while (blob.size() % 64) blob.push_back(__float2half(0.f));           // keep 128-byte alignment for the bulk copy
  const size_t off = blob.size();
  blob.resize(off + 2 * (size_t)KC * H * 8);
  float mx = 0.f;
  for (float v : W) mx = std::max(mx, std::fabs(v));
  int ex = 0;
  if (mx > 0.f && std::isfinite(mx)) { std::frexp(mx, &ex); ex -= 1; }   // mx in [2^ex, 2^(ex+1))
  const int sh = std::min(std::max(13 - ex, -40), 40);
  const float scale = std::ldexp(1.0f, sh);
  *descale = std::ldexp(1.0f, -sh);
  for (int kc = 0; kc < KC; ++kc)
    for (int c = 0; c < H; ++c)
      for (int u = 0; u < 8; ++u) {
        const float v = W[(size_t)c * H + kc * 8 + u] * scale;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        blob[off + ((size_t)kc * H + c) * 8 + u] = hi;
        blob[off + (size_t)KC * H * 8 + ((size_t)kc * H + c) * 8 + u] = lo;
      }
  return off;
}

// FC graphs: rows longer than half a tile (N > 64) make single-row tiles the norm -> column-split epilogue.
inline bool split_epilogue(const Geom& gm) { return gm.N > 64; }

inline dl_status launch_edge_tc(const Geom& gm, const EdgeArgs& ea, bool coord, const void* w2_tc, int num_sms,
                                cudaStream_t st) {
  const __half* w = reinterpret_cast<const __half*>(w2_tc);
  if (ea.recs != nullptr) {
    if (coord) k_edge_tc<true, false, true><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, w, nullptr);
    else k_edge_tc<false, false, true><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, w, nullptr);
  } else if (coord) k_edge_tc<true, false, false><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, w, nullptr);
  else if (split_epilogue(gm)) k_edge_tc<false, false, false, true><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, w, nullptr);
  else k_edge_tc<false, false, false, false><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, w, nullptr);
  return DL_OK;
}

// Debug: one profiled GCL launch; prints per-role wait/total cycles averaged over CTAs to stderr.
inline dl_status profile_edge_tc(const Geom& gm, const EdgeArgs& ea, const void* w2_tc, int num_sms, cudaStream_t st) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, (size_t)num_sms * 16 * 8) != cudaSuccess) return DL_ERR_CUDA;
  cudaMemsetAsync(d, 0, (size_t)num_sms * 16 * 8, st);
  if (ea.recs != nullptr)
    k_edge_tc<false, true, true><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, reinterpret_cast<const __half*>(w2_tc), d);
  else if (split_epilogue(gm))
    k_edge_tc<false, true, false, true><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, reinterpret_cast<const __half*>(w2_tc), d);
  else
    k_edge_tc<false, true, false, false><<<num_sms, EDGE_TC_THREADS, SMEM_BYTES, st>>>(gm, ea, reinterpret_cast<const __half*>(w2_tc), d);
  if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(d); return DL_ERR_CUDA; }
  std::vector<unsigned long long> h((size_t)num_sms * 16);
  cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  double avg[16] = {0};
  for (int b = 0; b < num_sms; ++b) for (int i = 0; i < 16; ++i) avg[i] += (double)h[(size_t)b * 16 + i] / num_sms;
  unsigned long long tmin = ~0ull, tmax = 0, cmin = ~0ull, cmax = 0;
  for (int b = 0; b < num_sms; ++b) {
    tmin = std::min(tmin, h[(size_t)b * 16 + 10]); tmax = std::max(tmax, h[(size_t)b * 16 + 10]);
    cmin = std::min(cmin, h[(size_t)b * 16 + 15]); cmax = std::max(cmax, h[(size_t)b * 16 + 15]);
  }
  fprintf(stderr, "[dl prof] cycles per CTA (avg over %d): tiles %.1f (min %llu max %llu), epilogue total min %llu max %llu\n",
          num_sms, avg[10], tmin, tmax, cmin, cmax);
  fprintf(stderr, "[dl prof]  table   : wait tempty %.0f | total %.0f\n", avg[0], avg[3]);
  fprintf(stderr, "[dl prof]  producer: wait tbl %.0f, wait empty %.0f | total %.0f\n", avg[4], avg[5], avg[7]);
  fprintf(stderr, "[dl prof]  mma     : wait full %.0f, wait tempty %.0f | total %.0f\n", avg[8], avg[9], avg[11]);
  fprintf(stderr, "[dl prof]  epilogue: wait tbl %.0f, wait tfull %.0f | total %.0f\n", avg[12], avg[13], avg[15]);
  return DL_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Self test: one 128 x 256 x 128 3xFP16 UMMA chain against a CPU fp64 result. Exercises descriptors, the
// canonical layout, TMEM addressing and the commit/wait protocol in isolation.
// ---------------------------------------------------------------------------------------------------------
constexpr int PN = 256;                          // probe: B operand rows (UMMA N = 256)
constexpr int P_LBO = PN * 16 + 16;
constexpr int P_OFF_BHI = 2 * W_BYTES, P_OFF_BLO = P_OFF_BHI + KC * P_LBO, P_OFF_BAR = P_OFF_BLO + KC * P_LBO;
constexpr int P_SMEM_BYTES = P_OFF_BAR + 64 + 1024;
// B_MN = true: the B operand is laid out MN-major (canonical no-swizzle form ((8,n),(8,k)):((1,SBO),(8,LBO)) in halves: 8
// consecutive N-rows of one k are contiguous 16 bytes; b_major bit of the instruction descriptor set) -- the layout a
// lane = channel producer could fill with 16-byte stores (DESIGN.md, next-round analysis).
template <bool B_MN>
__global__ void __launch_bounds__(128, 1) k_umma_probe(const __half* __restrict__ A /*[2][kc][128][8]*/,
                                                       const __half* __restrict__ Bm /*[2][kc][256][8]*/,
                                                       float* __restrict__ D /*[128][256]*/) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (no integer round trip) so accesses compile to LDS/STS
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = sbase + P_OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P_OFF_BAR + 32);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  // A -> W slots (pitch W_LBO), B -> activation slots (pitch P_LBO) with ordinary stores
  for (int idx = tid; idx < 2 * KC * H; idx += 128) {
    const int copy = idx / (KC * H), rem = idx % (KC * H), kcx = rem / H, row = rem % H;
    *reinterpret_cast<uint4*>(sm + (copy ? OFF_WLO : OFF_WHI) + kcx * W_LBO + row * 16) =
        *reinterpret_cast<const uint4*>(A + (size_t)idx * 8);
  }
  constexpr int MN_LBO = (PN / 8) * 128;             // MN-major: pitch between 8-wide K groups; SBO = 128 B between 8-row N groups
  for (int idx = tid; idx < 2 * KC * PN; idx += 128) {
    const int copy = idx / (KC * PN), rem = idx % (KC * PN), kcx = rem / PN, row = rem % PN;
    if (!B_MN) {
      *reinterpret_cast<uint4*>(sm + (copy ? P_OFF_BLO : P_OFF_BHI) + kcx * P_LBO + row * 16) =
          *reinterpret_cast<const uint4*>(Bm + (size_t)idx * 8);
    } else {
      for (int u = 0; u < 8; ++u)                      // element (n = row, k = kcx*8 + u)
        *reinterpret_cast<__half*>(sm + (copy ? P_OFF_BLO : P_OFF_BHI) + kcx * MN_LBO + (row >> 3) * 128 + u * 16 + (row & 7) * 2) =
            Bm[(size_t)idx * 8 + u];
    }
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(128, 256) | (B_MN ? (1u << 16) : 0u);      // bit 16: B operand MN-major
    const uint32_t blbo = B_MN ? MN_LBO : P_LBO;
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t a_hi = umma_desc(sbase + OFF_WHI + ks * 2 * W_LBO, W_LBO, SBO), a_lo = umma_desc(sbase + OFF_WLO + ks * 2 * W_LBO, W_LBO, SBO);
      const uint64_t b_hi = umma_desc(sbase + P_OFF_BHI + ks * 2 * blbo, blbo, SBO), b_lo = umma_desc(sbase + P_OFF_BLO + ks * 2 * blbo, blbo, SBO);
      umma_f16(tmem, a_lo, b_hi, idesc, ks > 0);
      umma_f16(tmem, a_hi, b_lo, idesc, 1);
      umma_f16(tmem, a_hi, b_hi, idesc, 1);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < PN; c0 += 8) {
    uint32_t r[8];
    TMEM_LD_X8(tlane + c0, r);
    tmem_ld_wait();
    for (int u = 0; u < 8; ++u) D[(size_t)(warp * 32 + lane) * PN + c0 + u] = __uint_as_float(r[u]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

inline dl_status selftest(int /*num_sms*/, float* max_abs_err, float* max_rel_err, bool b_mn_major = false) {
  std::vector<float> A((size_t)H * H), Bv((size_t)PN * H);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : A) v = rnd() * 0.2f;
  for (auto& v : Bv) v = rnd() * 3.0f;
  auto pack = [&](const std::vector<float>& M, int rows) {
    std::vector<__half> out(2 * (size_t)KC * rows * 8);
    for (int kc = 0; kc < KC; ++kc)
      for (int r = 0; r < rows; ++r)
        for (int u = 0; u < 8; ++u) {
          const float v = M[(size_t)r * H + kc * 8 + u];
          const __half hi = __float2half_rn(v);
          out[((size_t)kc * rows + r) * 8 + u] = hi;
          out[(size_t)KC * rows * 8 + ((size_t)kc * rows + r) * 8 + u] = __float2half_rn(v - __half2float(hi));
        }
    return out;
  };
  std::vector<__half> Ap = pack(A, H), Bp = pack(Bv, PN);
  __half *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  if (cudaMalloc(&dA, Ap.size() * 2) != cudaSuccess || cudaMalloc(&dB, Bp.size() * 2) != cudaSuccess ||
      cudaMalloc(&dD, (size_t)H * PN * 4) != cudaSuccess)
    return DL_ERR_CUDA;
  cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)H * PN * 4);
  if (b_mn_major) {
    cudaFuncSetAttribute(k_umma_probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    k_umma_probe<true><<<1, 128, P_SMEM_BYTES>>>(dA, dB, dD);
  } else {
    cudaFuncSetAttribute(k_umma_probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    k_umma_probe<false><<<1, 128, P_SMEM_BYTES>>>(dA, dB, dD);
  }
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> D((size_t)H * PN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  if (err != cudaSuccess) {
    fprintf(stderr, "[dl selftest] k_umma_probe failed: %s\n", cudaGetErrorString(err));
    return DL_ERR_CUDA;
  }
  double ma = 0, mr = 0, mref = 0;
  for (int m = 0; m < H; ++m)
    for (int n = 0; n < PN; ++n) {
      double ref = 0;
      for (int k = 0; k < H; ++k) ref += (double)A[(size_t)m * H + k] * (double)Bv[(size_t)n * H + k];
      ma = std::max(ma, std::fabs(ref - (double)D[(size_t)m * PN + n]));
      mref = std::max(mref, std::fabs(ref));
    }
  mr = ma / std::max(mref, 1e-30);
  if (max_abs_err) *max_abs_err = (float)ma;
  if (max_rel_err) *max_rel_err = (float)mr;
  fprintf(stderr, "[dl selftest] 3xFP16 UMMA 128x256x128 (B %s-major): max abs err %.3e (rel to max |ref| %.3e)\n",
          b_mn_major ? "MN" : "K", ma, mr);
  return DL_OK;
}

}  // namespace tc
}  // namespace dl
