// tcgen05 node kernel: GCL.node_model (egnn.py:62-72) + node mask + the first-layer projections of the next edge
// MLP(s), as a chain of UMMA GEMMs on one 128-node tile per CTA.
//
//   G1: hid = silu([h, agg] W3^T + b3)        K = 256  (two 128x128 weight blocks)
//   G2: h'  = (h + hid W4^T + b4) * node_mask  K = 128
//   P : A = h' W1a^T + b1 ; B = h' W1b^T       K = 128 each, for one or two consumers
//
// Swapped orientation D[channel, node] (weights are the A operand, the node tile is the B operand): TMEM lane =
// output channel, column = node. Thread c therefore owns one output channel: its bias is a register, and every
// global access of the epilogues (h, h', A|B rows) has the 32 lanes of a warp on 32 consecutive floats of one node
// row -- fully coalesced. (The row-per-thread orientation measured 36-43 us per launch, two thirds of it in
// uncoalesced row I/O.) The price is a 2-byte scatter when an epilogue writes the next GEMM's operand tile
// ([kc][node][8 halves]); the slab pitch is padded by 16 B so those stores are bank-conflict free.
// Same 3xFP16 numerics as the edge kernel (kernels_tc.cuh). Range scaling is per TILE here (one exact power of
// two chosen from the tile's maximum, only ever != 1 for diverging samples): operands are first written unscaled
// while the maximum is reduced, and rewritten only in the rare case the bound exceeds 2^14.
// Roles (k_node_tc2): 24 worker warps (row staging, epilogues; thread = output channel of one lane quarter), one loader
// warp streaming K=64 weight half-blocks (2 x 16 KB cp.async.bulk each) through a ring of up to 4 stages, and one
// warp-convergent MMA issuer (elect.sync) that walks a table of half-blocks and is fed by mbarriers only: `rdy` (the
// workers have written the GEMM's operand tile) and the ring's `full`; it commits to `acc` per GEMM and to `empty` per
// ring stage. The tile width is chosen at launch so that one wave of CTAs covers all nodes (72 nodes on cfg 2).
// Measured: the GEMM phases take (4096 + 32 N)/64 cycles per MMA (shared-memory operand fetch at 64 B/clk), so splitting
// the tile into pipelined column chunks doubles the tensor time (nchunk stays 1).
#pragma once
#include "kernels_tc.cuh"

namespace dl {
namespace tcn {

using namespace dl::tc;

constexpr int TM = 128;                        // nodes per tile = UMMA N
constexpr int X_LBO_MAX = TM * 16 + 16;        // 2064 B: padded slab pitch of a 128-node operand tile (run time: n_pad * 16 + 16)
constexpr int X_BYTES_MAX = KC * X_LBO_MAX;    // 33,024 B per fp16 copy
constexpr int HALF_BYTES = 8 * W_LBO;          // 16 KB: kc 0..7 of one fp16 copy of a 128x128 block
constexpr int STAGE_BYTES = 2 * HALF_BYTES;    // hi | lo
constexpr int BLOCK_BYTES = 2 * W_BYTES;       // one packed 128x128 block: [hi 32 KB | lo 32 KB]
constexpr int N_RING = 4;                      // ring slots that have barriers; how many are USED depends on the tile width:
                                               // the operand tiles are sized for the run-time tile (72 nodes on cfg 2: 83 KB
                                               // instead of 132 KB), and the space goes to the weight ring (4 x 32 KB in flight
                                               // instead of 2: the next GEMM's weights land while the current epilogue runs)
constexpr int N_OFF_XA = 0;                    // hi | lo, then XB hi | lo, then the weight ring (run-time offsets)
constexpr int N_OFF_MISC = 226304;             // nm[128] f32, nodemax[2][128] i32, tilemax i32
constexpr int N_OFF_BAR = N_OFF_MISC + 128 * 4 * 3 + 16;      // full[4], empty[4], acc[4]; tmem slot
constexpr int N_SMEM_BYTES = N_OFF_BAR + 128 + 1024;
static_assert(N_SMEM_BYTES <= 232448 && 4 * X_BYTES_MAX + 2 * STAGE_BYTES <= N_OFF_MISC, "k_node_tc shared memory");
#ifndef DL_NODE_WORKER_WARPS
#define DL_NODE_WORKER_WARPS 24
#endif
constexpr int NW = DL_NODE_WORKER_WARPS;       // worker warps: warp w serves TMEM lane quarter w % 4 and node columns
constexpr int NPART = NW / 4;                  //   [part*cw, (part+1)*cw) with part = w / 4 -- 6 warps per scheduler hide each
constexpr int CW = ((TM + NPART - 1) / NPART + 7) & ~7;   // other's TMEM / shared / global latencies (the epilogues of this kernel
                                               //   are dependency-latency bound: 24 instead of 16 warps shorten every phase)
constexpr int NODE_TC_THREADS = 32 * (NW + 1); // + 1 weight-loader warp

struct NodeTcArgs {
  float* h;              // (n,128) in/out
  const float* agg;      // (n,128)
  const float* nm;       // (n)
  const __half* w3;      // 2 packed blocks (h part, agg part), common scale
  const __half* w4;      // 1 packed block
  const float *b3, *b4;
  float w3_descale, w4_descale;
  int tile_nodes;        // nodes per CTA (multiple of 8, <= 128; 0 = 128): chosen by the host so that one wave fills the SMs
  int proj_only;         // 1: skip the node MLP -- h is only projected (first layer of a forward: replaces the SIMT projection of k_prep)
  int n_proj;            // 1 or 2
  const __half* pw[2];   // 2 packed blocks each (W1a, W1b), common scale
  const float* pb1[2];
  float p_descale[2];
  float* AB[2];
  float* ABmax[2];
};

__device__ __forceinline__ float pow2_scale_for(float bound) {
  float sc = 1.0f;
  if (!(bound <= F16_TARGET)) {
    const int ex = ((__float_as_int(bound) >> 23) & 0xff) - 127;
    sc = __int_as_float(max(127 + 13 - ex, 1) << 23);
  }
  return sc;
}

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory"); }

// one value of channel c, node n -> fp16 hi/lo operand tile
__device__ __forceinline__ void store_elem(uint8_t* xhi, uint8_t* xlo, int x_lbo, int c, int n, float v) {
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int off = (c >> 3) * x_lbo + n * 16 + (c & 7) * 2;
  *reinterpret_cast<__half*>(xhi + off) = hi;
  *reinterpret_cast<__half*>(xlo + off) = lo;
}

// ---------------------------------------------------------------------------------------------------------
// k_node_tc2: three roles. 24 worker warps (rows -> operand tiles, the epilogues), one weight-loader warp, and one
// warp-convergent MMA issuer that is fed by mbarriers: the workers signal "operand ready" (rows / hid / h') and wait on the
// accumulator barriers the issuer commits, so MMA issue never waits for a block barrier or for a worker thread's own work
// (the first version issued from worker thread 0 after a block barrier: 22.9 / 29.1 us per launch vs 21.5 / 24.9 us).
// The column range is written as "chunks" because a two-chunk pipelined variant (tensor core on one chunk, workers on the
// other) was built on this structure; it measured slower (see the comment at `nchunk`).
// ---------------------------------------------------------------------------------------------------------
constexpr int W2_LOADER = NW, W2_MMA = NW + 1;
constexpr int NODE_TC2_THREADS = 32 * (NW + 2);
constexpr int NB2_RDY = 8 * (2 * N_RING);                  // byte offsets inside the barrier area: full[4], empty[4], then
constexpr int NB2_ACC = NB2_RDY + 8 * 6;                   //   rdy[3 phases][2 chunks], acc[6 units][2 chunks], tmem slot
constexpr int NB2_TMEM = NB2_ACC + 8 * 12;
constexpr int N2_SMEM_BYTES = N_OFF_BAR + NB2_TMEM + 16 + 1024;
static_assert(N2_SMEM_BYTES <= 232448, "k_node_tc2 shared memory");

__global__ void __launch_bounds__(NODE_TC2_THREADS, 1) k_node_tc2(int n_total, NodeTcArgs a, long long* __restrict__ prof = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t0 = prof ? clock64() : 0;
  auto mark = [&](int i) { if (prof && tid == 0) prof[(size_t)blockIdx.x * 8 + i] = clock64() - t0; };
  const int tile = a.tile_nodes > 0 ? a.tile_nodes : TM;
  const int g0 = blockIdx.x * tile;
  const int n_live = min(tile, n_total - g0);
  const int n_pad = (tile + 15) & ~15;
  // chunks: [0, csplit) and [csplit, n_pad), both multiples of 16 columns (UMMA N)
  // ONE chunk: splitting the node columns into two pipelined chunks was built and measured SLOWER (27 vs 23 us) -- an MMA of
  // this shape costs about the same for N = 48 as for N = 80 (the 4 KB weight-operand read and the issue overhead dominate), so
  // halving N doubles the tensor time. The chunk plumbing below is kept generic (nchunk is a compile-time 1).
  const int csplit = n_pad;
  constexpr int nchunk = 1;
  auto cbeg = [&](int c) { return c == 0 ? 0 : csplit; };
  auto cend = [&](int c) { return (c == 0 && nchunk == 2) ? csplit : n_pad; };

  const int X_LBO = n_pad * 16 + 16;
  const int X_BYTES = KC * X_LBO;
  const int off_ws = 4 * X_BYTES;
  const int n_ring = min(N_RING, (N_OFF_MISC - off_ws) / STAGE_BYTES);
  const uint32_t bars = sbase + N_OFF_BAR;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * N_RING, bar_rdy = bars + NB2_RDY, bar_acc = bars + NB2_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + N_OFF_BAR + NB2_TMEM);
  uint8_t* xa_hi = sm + N_OFF_XA; uint8_t* xa_lo = xa_hi + X_BYTES;
  uint8_t* xb_hi = sm + 2 * X_BYTES; uint8_t* xb_lo = xb_hi + X_BYTES;
  float* nms = reinterpret_cast<float*>(sm + N_OFF_MISC);
  int* tilemax = reinterpret_cast<int*>(sm + N_OFF_MISC + 512);

  if (tid == 0) {
    for (int i = 0; i < N_RING; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 6; ++i) mbar_init(bar_rdy + 8 * i, NW);
    for (int i = 0; i < 12; ++i) mbar_init(bar_acc + 8 * i, 1);
    fence_barrier_init();
  }
  if (tid < 16) tilemax[tid] = 0;
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);      // worker warp 0 allocates and, at the end, frees
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // Up to here, and the loader warp's first weight stages below, nothing depends on the previous kernel of the forward: it
  // overlaps that kernel's tail. The workers are the only role that touches chain-produced memory (h, agg, node mask in;
  // h, projections out), so they wait; the node mask is first used after a workers_sync (tile_max) of the row pass.
  if (warp < NW) {
    chain_wait();
    if (tid == 0) chain_release();
    if (tid < TM) nms[tid] = tid < n_live ? a.nm[g0 + tid] : 0.f;
  }

  const int blk0 = a.proj_only ? 3 : 0;
  const int n_blocks = 3 + 2 * a.n_proj - blk0;
  const int n_half = 2 * n_blocks;
  // accumulator units: 0 = G1 (TMEM col 0), 1 = G2 (128), 2 = A of consumer 0 (256), 3 = B of consumer 0 (384),
  //                    4 = A of consumer 1 (0, reuse), 5 = B of consumer 1 (128, reuse)
  auto unit_col = [](int u) { return u == 0 ? 0 : u == 1 ? 128 : u == 2 ? 256 : u == 3 ? 384 : u == 4 ? 0 : 128; };

  if (warp == W2_LOADER) {
    if (lane == 0) {
      for (int i = 0; i < n_half; ++i) {
        const int s = i % n_ring, blk = (i >> 1) + blk0, hf = i & 1;
        if (i >= n_ring) mbar_wait(bar_empty + 8 * s, ((i - n_ring) / n_ring) & 1);
        const __half* base = blk < 2 ? a.w3 + (size_t)blk * (BLOCK_BYTES / 2)
                             : blk == 2 ? a.w4
                                        : a.pw[(blk - 3) >> 1] + (size_t)((blk - 3) & 1) * (BLOCK_BYTES / 2);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(base);
        const uint32_t dst = sbase + off_ws + s * STAGE_BYTES;
        mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
        bulk_g2s(dst, src + hf * HALF_BYTES, HALF_BYTES, bar_full + 8 * s);
        bulk_g2s(dst + HALF_BYTES, src + W_BYTES + hf * HALF_BYTES, HALF_BYTES, bar_full + 8 * s);
      }
    }
    return;
  }

  if (warp == W2_MMA) {
    // All 32 lanes run this loop with warp-uniform values; one elected lane issues each tcgen05 instruction. One generic pass
    // over the streamed half-blocks (weights block blk, K half hf): which operand tile, accumulator and barrier it touches
    // follows from blk alone, so the 12 MMAs of a half-block exist once in the instruction stream.
    const uint32_t idesc = umma_idesc(128, n_pad);
#pragma unroll 1
    for (int i = 0; i < n_half; ++i) {
      const int s = i % n_ring, blk = (i >> 1) + blk0, hf = i & 1;
      // operand readiness: rows before G1, hid before G2, h' before the first projection
      if (hf == 0 && (blk == 0 || blk == 2 || blk == 3)) { mbar_wait(bar_rdy + 8 * ((blk == 0 ? 0 : blk - 1) * 2), 0); tc_fence_after(); }
      const int unit = blk < 2 ? 0 : blk - 1;              // accumulator unit (0 G1, 1 G2, 2.. projections)
      const uint8_t* xop = (blk == 1 || blk == 2) ? xb_hi : xa_hi;
      const bool first = hf == 0 && blk != 1;              // first half-block of its accumulator (G1 spans blocks 0 and 1)
      mbar_wait(bar_full + 8 * s, (i / n_ring) & 1);
      tc_fence_after();
      const uint32_t xh = smem_u32(xop) + (8 * hf) * X_LBO, xl = xh + X_BYTES;
      const uint32_t wh = sbase + off_ws + s * STAGE_BYTES, wl = wh + HALF_BYTES;
      const uint32_t dcol = tmem + unit_col(unit);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a_hi = umma_desc(wh + ks * 2 * W_LBO, W_LBO, SBO), a_lo = umma_desc(wl + ks * 2 * W_LBO, W_LBO, SBO);
        const uint64_t b_hi = umma_desc(xh + ks * 2 * X_LBO, X_LBO, SBO), b_lo = umma_desc(xl + ks * 2 * X_LBO, X_LBO, SBO);
        umma_f16_elect(dcol, a_lo, b_hi, idesc, !(first && ks == 0));
        umma_f16_elect(dcol, a_hi, b_lo, idesc, 1);
        umma_f16_elect(dcol, a_hi, b_hi, idesc, 1);
      }
      umma_commit_elect(bar_empty + 8 * s);                // ring slot reusable once these MMAs have read it
      if (hf == 1 && blk != 0) umma_commit_elect(bar_acc + 8 * (unit * 2));   // accumulator complete
    }
    return;
  }

  // =================================== workers ===========================================================================
  const int c = tid & (TM - 1);                          // output channel = TMEM lane
  const int part = tid >> 7;
  const uint32_t tq = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  int tm_calls = 0;
  auto tile_max = [&](float mx) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int* slot = tilemax + (tm_calls++ & 15);
    if (lane == 0) atomicMax(slot, __float_as_int(mx));
    workers_sync();
    return __int_as_float(*slot);
  };
  auto operand_ready = [&](int phase, int ch) {            // generic-proxy operand writes -> visible to the tensor core
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_rdy + 8 * (phase * 2 + ch));
  };
  auto wait_acc = [&](int unit, int ch) { mbar_wait(bar_acc + 8 * (unit * 2 + ch), 0); tc_fence_after(); };
  // this thread's columns of chunk ch: [col0, col0 + ncols), ng groups of 8
  struct Slice { int col0, ncols, ng; };
  auto slice_of = [&](int ch) {
    const int w = cend(ch) - cbeg(ch);
    const int cw = (((w + NPART - 1) / NPART) + 7) & ~7;
    Slice s;
    s.col0 = cbeg(ch) + part * cw;
    s.ncols = max(0, min(cw, cend(ch) - s.col0));
    s.ng = (s.ncols + 7) >> 3;
    return s;
  };

  // ---- operand rows of chunk ch: h -> XA, agg -> XB ---------------------------------------------------------------------
  auto load_rows = [&](int ch, float scale) -> float {
    float mx = 0.f;
    const int r_end = cend(ch);
#pragma unroll 1
    for (int rb = cbeg(ch) + warp; rb < r_end; rb += 4 * NW) {
      float4 hv[4], av[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rb + NW * j;
        hv[j] = make_float4(0, 0, 0, 0); av[j] = hv[j];
        if (r < n_live && r < r_end) {
          hv[j] = *reinterpret_cast<const float4*>(a.h + (size_t)(g0 + r) * H + lane * 4);
          if (!a.proj_only) av[j] = __ldg(reinterpret_cast<const float4*>(a.agg + (size_t)(g0 + r) * H + lane * 4));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rb + NW * j;
        if (r >= r_end) continue;
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(hv[j].x), fabsf(hv[j].y)), fmaxf(fabsf(hv[j].z), fabsf(hv[j].w))));
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(av[j].x), fabsf(av[j].y)), fmaxf(fabsf(av[j].z), fabsf(av[j].w))));
        const int off = (lane >> 1) * X_LBO + r * 16 + (lane & 1) * 8;
        uint2 hi, lo;
        split2(hv[j].x * scale, hv[j].y * scale, hi.x, lo.x); split2(hv[j].z * scale, hv[j].w * scale, hi.y, lo.y);
        *reinterpret_cast<uint2*>(xa_hi + off) = hi; *reinterpret_cast<uint2*>(xa_lo + off) = lo;
        split2(av[j].x * scale, av[j].y * scale, hi.x, lo.x); split2(av[j].z * scale, av[j].w * scale, hi.y, lo.y);
        *reinterpret_cast<uint2*>(xb_hi + off) = hi; *reinterpret_cast<uint2*>(xb_lo + off) = lo;
      }
    }
    return mx;
  };

  float s1[2] = {1.f, 1.f}, s2[2] = {1.f, 1.f}, s3[2] = {1.f, 1.f};
  for (int ch = 0; ch < nchunk; ++ch) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {               // second pass: rare (diverging samples), rewrites the operands scaled
      const float mx = load_rows(ch, s1[ch]);
      if (pass == 1) break;
      s1[ch] = pow2_scale_for(tile_max(mx));
      if (s1[ch] == 1.0f) break;
    }
    operand_ready(0, ch);
    if (a.proj_only) { s3[ch] = s1[ch]; operand_ready(2, ch); }
  }
  mark(0);

  if (!a.proj_only) {
    // ---- epilogue 1: hid = silu(D ds + b3) -> XB (the agg operand of this chunk is dead: its G1 is complete) -------------
    const float bias3 = __ldg(a.b3 + c);
    for (int ch = 0; ch < nchunk; ++ch) {
      wait_acc(0, ch);
      if (ch == 0) mark(1);
      const Slice sl = slice_of(ch);
      const float ds = a.w3_descale / s1[ch];
      auto epi1 = [&](float scale) -> float {
        float mx = 0.f;
        uint32_t r[2][8];
        if (sl.ng > 0) TMEM_LD_X8(tq + sl.col0, r[0]);
#pragma unroll
        for (int k = 0; k < CW / 8; ++k) {
          if (k < sl.ng) {
            tmem_ld_wait();
            if (k + 1 < sl.ng) TMEM_LD_X8(tq + sl.col0 + (k + 1) * 8, r[(k + 1) & 1]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float v = fmaf(__uint_as_float(r[k & 1][u]), ds, bias3);
              mx = fmaxf(mx, fabsf(v));
              store_elem(xb_hi, xb_lo, X_LBO, c, sl.col0 + k * 8 + u, silu_f(v) * scale);
            }
          }
        }
        return mx;
      };
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const float mx = epi1(s2[ch]);
        if (pass == 1) break;
        s2[ch] = pow2_scale_for(tile_max(mx));
        if (s2[ch] == 1.0f) break;
      }
      operand_ready(1, ch);
    }
    mark(2);
    // ---- epilogue 2: h' = (h + D ds + b4) * node_mask -> global h and XA ---------------------------------------------------
    const float bias4 = __ldg(a.b4 + c);
    float* hcol = a.h + (size_t)g0 * H + c;
    for (int ch = 0; ch < nchunk; ++ch) {
      wait_acc(1, ch);
      if (ch == 0) mark(3);
      const Slice sl = slice_of(ch);
      const float ds = a.w4_descale / s2[ch];
      float mx = 0.f;
      uint32_t r[2][8];
      float hv[2][8];
      if (sl.ng > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) hv[0][u] = (sl.col0 + u < n_live) ? hcol[(size_t)(sl.col0 + u) * H] : 0.f;
        TMEM_LD_X8(tq + 128 + sl.col0, r[0]);
      }
#pragma unroll
      for (int k = 0; k < CW / 8; ++k) {
        if (k < sl.ng) {
          tmem_ld_wait();
          if (k + 1 < sl.ng) {
            TMEM_LD_X8(tq + 128 + sl.col0 + (k + 1) * 8, r[(k + 1) & 1]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int nn = sl.col0 + (k + 1) * 8 + u;
              hv[(k + 1) & 1][u] = (nn < n_live) ? hcol[(size_t)nn * H] : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int n = sl.col0 + k * 8 + u;
            const float o = (hv[k & 1][u] + fmaf(__uint_as_float(r[k & 1][u]), ds, bias4)) * nms[n];     // egnn.py:71,78-79
            if (n < n_live) hcol[(size_t)n * H] = o;
            mx = fmaxf(mx, fabsf(o));
            store_elem(xa_hi, xa_lo, X_LBO, c, n, o);
          }
        }
      }
      s3[ch] = pow2_scale_for(tile_max(mx));
      if (s3[ch] != 1.0f) {                                // rare: rewrite h' scaled (own global writes, program order)
        for (int n = sl.col0; n < sl.col0 + 8 * sl.ng; ++n)
          store_elem(xa_hi, xa_lo, X_LBO, c, n, (n < n_live ? hcol[(size_t)n * H] : 0.f) * s3[ch]);
      }
      operand_ready(2, ch);
    }
    mark(4);
  }

  // ---- projection epilogues: A | B rows of the next edge MLP(s) ------------------------------------------------------------
  for (int p = 0; p < a.n_proj; ++p) {
#pragma unroll 1
    for (int pt = 0; pt < 2; ++pt) {
      const int unit = 2 + 2 * p + pt;
      const float bias = pt == 0 ? __ldg(a.pb1[p] + c) : 0.f;
      float* abcol = a.AB[p] + (size_t)g0 * 2 * H + pt * H + c;
      float mx = 0.f;
      for (int ch = 0; ch < nchunk; ++ch) {
        wait_acc(unit, ch);
        const Slice sl = slice_of(ch);
        const float ds = a.p_descale[p] / s3[ch];
        const uint32_t tcol = tq + unit_col(unit) + sl.col0;
        uint32_t r[2][8];
        if (sl.ng > 0) TMEM_LD_X8(tcol, r[0]);
#pragma unroll
        for (int k = 0; k < CW / 8; ++k) {
          if (k < sl.ng) {
            tmem_ld_wait();
            if (k + 1 < sl.ng) TMEM_LD_X8(tcol + (k + 1) * 8, r[(k + 1) & 1]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int n = sl.col0 + k * 8 + u;
              const float o = fmaf(__uint_as_float(r[k & 1][u]), ds, bias);
              if (n < n_live) { abcol[(size_t)n * 2 * H] = o; mx = fmaxf(mx, fabsf(o)); }
            }
          }
        }
      }
      // range bound for the consumer's fp16 operands: the TILE maximum of |A| (|B|), written for every node of the tile
      const float tmx = tile_max(mx);
      if (tid < n_live) a.ABmax[p][(size_t)(g0 + tid) * 2 + pt] = tmx;
    }
  }
  mark(6);
  tc_fence_before();
  workers_sync();
  if (warp == 0) tmem_dealloc(tmem, 512);                    // every accumulator has been waited for above
}

// pack `nblk` horizontally adjacent 128x128 blocks of a (128 x in_stride) row-major matrix with one common
// power-of-two scale. Returns the offset in halves.
inline size_t pack_blocks(const std::vector<float>& W, int in_stride, int nblk, std::vector<__half>& blob,
                          float* descale) {
  while (blob.size() % 64) blob.push_back(__float2half(0.f));
  const size_t off = blob.size();
  const size_t per = 2 * (size_t)KC * H * 8;
  blob.resize(off + nblk * per);
  float mx = 0.f;
  for (int c = 0; c < H; ++c)
    for (int k = 0; k < nblk * H; ++k) mx = std::max(mx, std::fabs(W[(size_t)c * in_stride + k]));
  int ex = 0;
  if (mx > 0.f && std::isfinite(mx)) { std::frexp(mx, &ex); ex -= 1; }
  const int sh = std::min(std::max(13 - ex, -40), 40);
  const float scale = std::ldexp(1.0f, sh);
  *descale = std::ldexp(1.0f, -sh);
  for (int b = 0; b < nblk; ++b)
    for (int kc = 0; kc < KC; ++kc)
      for (int c = 0; c < H; ++c)
        for (int u = 0; u < 8; ++u) {
          const float v = W[(size_t)c * in_stride + b * H + kc * 8 + u] * scale;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          blob[off + b * per + ((size_t)kc * H + c) * 8 + u] = hi;
          blob[off + b * per + (size_t)KC * H * 8 + ((size_t)kc * H + c) * 8 + u] = lo;
        }
  return off;
}

// Nodes per CTA: the smallest multiple of 8 that lets one wave of `num_sms` CTAs cover all n nodes (<= 128). Fewer, fuller
// tiles would leave SMs idle (cfg 2: 80 tiles of 128 on 148 SMs); every phase of the per-tile chain shortens with the tile.
inline int pick_tile_nodes(int n, int num_sms) {
  const int per = (n + num_sms - 1) / std::max(num_sms, 1);
  return std::min(TM, std::max(8, (per + 7) & ~7));
}

// Debug: one timed launch; prints the phase boundaries (cycles from kernel entry, thread 0, averaged over CTAs).
inline void launch_node(int n, const NodeTcArgs& ta, cudaStream_t st, long long* prof);
inline void profile_node(int n, const NodeTcArgs& ta_in, cudaStream_t st, int num_sms = 148) {
  NodeTcArgs ta = ta_in;
  ta.tile_nodes = pick_tile_nodes(n, num_sms);             // the configuration the forward actually launches
  const int grid = (n + ta.tile_nodes - 1) / ta.tile_nodes;
  long long* d = nullptr;
  if (cudaMalloc(&d, (size_t)grid * 8 * 8) != cudaSuccess) return;
  cudaMemsetAsync(d, 0, (size_t)grid * 64, st);
  launch_node(n, ta, st, d);
  cudaStreamSynchronize(st);
  std::vector<long long> h((size_t)grid * 8);
  cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  double avg[8] = {0};
  for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) avg[i] += (double)h[(size_t)b * 8 + i] / grid;
  fprintf(stderr, "[dl prof node] tile %d, grid %d, n_proj %d: rows %.0f | G1 ready %.0f | epi1 %.0f | G2 ready %.0f | epi2 %.0f | proj done %.0f\n",
          ta.tile_nodes, grid, ta.n_proj, avg[0], avg[1], avg[2], avg[3], avg[4], avg[6]);
}

inline dl_status configure_node() {
  if (cudaFuncSetAttribute(k_node_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, N2_SMEM_BYTES) != cudaSuccess) return DL_ERR_CUDA;
  return DL_OK;
}

inline void launch_node(int n, const NodeTcArgs& ta, cudaStream_t st, long long* prof) {
  const int grid = (n + ta.tile_nodes - 1) / ta.tile_nodes;
  launch_chain(k_node_tc2, dim3(grid), dim3(NODE_TC2_THREADS), N2_SMEM_BYTES, st, n, ta, prof);
}

}  // namespace tcn
}  // namespace dl
