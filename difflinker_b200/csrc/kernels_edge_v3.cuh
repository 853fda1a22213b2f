// GCL edge kernel, third generation (fully connected graphs with N <= 64: BASELINE configs 1-3 and the small end of
// the padded-N sweep). Same algorithm and numerics as tc::k_edge_tc (kernels_tc.cuh; reference egnn.py:45-60,
// 62-72, 295-320), restructured so that no role of the pipeline waits on an L2 round trip:
//
//   * the molecule's column projections B_j (N x 128 fp32) are staged in shared memory by ONE 2-D tiled TMA load
//     (cp.async.bulk.tensor, SASS UTMALDG) per molecule, double-buffered so the next molecule's panel lands while the
//     current one is consumed; the producers gather B_j with conflict-free LDS.128 instead of per-edge __ldg from L2;
//   * everything about a tile that does not depend on the layer lives in global "tile tables" (static part per plan,
//     d0 per forward, d per block: k_tiles_static / k_tiles_d) and the table warps only STREAM it into their ring
//     slots with bulk copies (UBLKCP), together with the 512-byte A_i rows of the tile; they compute nothing per edge;
//   * W2 (hi | lo fp16) lives in TENSOR MEMORY as the A operand of tcgen05.mma (TS form): 64 KB of shared memory and half
//     of the tensor core's shared-memory read traffic are gone, and with them the 64 KB bulk load at every launch;
//   * both SiLUs use the four-sigmoids-per-reciprocal form (usig4: 1.25 MUFU per SiLU instead of 1.5), the first layer
//     runs in the log2 domain (weights pre-scaled by -log2 e: no scaling multiply), tile rows are padded to a multiple
//     of four columns (weight-0 mirror edges) so the epilogue only has the 8- and 4-column paths;
//   * 12 epilogue warps (three per TMEM lane quarter, rotating over the tile's rows) next to 16 producer warps: both
//     roles are dependency-latency bound per warp, so the split is chosen to balance them (measured: 8 epilogue warps
//     with deeper software pipelining per warp were 35 % slower);
//   * the producer-side and epilogue-side table slots are separate rings (3 and 8 deep) so the table warps run as far
//     ahead of the epilogue as the epilogue-side ring allows instead of being throttled by the slowest consumer;
//   * the producer work of a tile is ceil(Et / 8) warp-tasks, dealt round-robin to the 16 producer warps ACROSS tiles (a
//     15-task tile of 3 x 40 edges leaves one warp free to start on the next tile, the short last tile of a molecule
//     occupies 5 warps instead of 16);
//   * launched with programmatic stream serialization (common.cuh launch_chain): barrier init, TMEM allocation and the
//     W2 -> tensor memory copy run while the previous kernel of the forward drains; only the table warps -- the one role
//     that reads what other kernels produce -- execute griddepcontrol.wait.
//
// K permutation: producer thread kc owns channels {4kc..4kc+3} u {64+4kc..64+4kc+3} (two conflict-free 16-byte reads of
// a 512-byte panel row per half-warp) and writes them as operand positions 8kc..8kc+7; W2 is packed with the same
// permutation of its K index (pack_w2_v3), so the contraction is unchanged.
#pragma once
#include <cuda.h>

#include "kernels_tc.cuh"

namespace dl {
namespace tc3 {

using namespace dl::tc;

constexpr int PROF_SLOTS = 32;                 // u64 per CTA of the profiling variant: 16 role counters + 16 timeline marks
constexpr int MAXR3 = 8;                       // rows per tile (A rows staged per producer-side slot)
constexpr int NACC3 = 3;                       // TMEM accumulator stages: columns 128 + 128 a (columns 0..127 hold W2 hi | lo)
constexpr int NPS3 = 3;                        // ring of producer-side table slots (freed by the producers)
constexpr int NES3 = 8;                        // ring of epilogue-side table slots (freed by the epilogue): how far the table
                                               // warps may run ahead of the pipeline's last stage
constexpr int PANEL_N = 64;                    // largest molecule whose B panel is double-buffered in shared memory
constexpr int PANEL_BYTES = PANEL_N * H * 4;   // 32 KB per buffer
constexpr int TM_W = 0, TM_ACC = 128;          // TMEM columns

constexpr int O3_ST = 0;                                   // 2 x [hi | lo] activation operand stages
constexpr int O3_PANEL = O3_ST + N_STAGE * STAGE_BYTES;    // 2 x B panel
constexpr int O3_PT = O3_PANEL + 2 * PANEL_BYTES;          // NPS3 x producer-side slot
constexpr int P3_D = 0;                                    // f32[TN]  |x_i - x_j|^2 of this block          (bulk copy of td[tile])
constexpr int P3_D0 = P3_D + TN * 4;                       // f32[TN]  |x0_i - x0_j|^2 of the call's input  (bulk copy of td0[tile])
constexpr int P3_BOFF = P3_D0 + TN * 4;                    // int[TN]  byte offset of B_j inside the panel   } one bulk copy of the
constexpr int P3_GRP = P3_BOFF + TN * 4;                   // int[TN/4] A-row byte offset of each 4-edge group } tile's static part
constexpr int P3_A = P3_GRP + TN;                          // f32[MAXR3][128]: A_i rows of the tile (bulk copies from AB)
constexpr int P3_BYTES = P3_A + MAXR3 * H * 4;
constexpr int TS_P_BYTES = P3_A - P3_BOFF;                 // 640: static producer-side part of a tile in global memory
constexpr int O3_ET = O3_PT + NPS3 * P3_BYTES;             // NES3 x epilogue-side slot
constexpr int E3_HDR = 0;                                  // int[16]: Et, nrt, ncc4, b                      } one bulk copy of the
constexpr int E3_EW = 64;                                  // f32[TN]  edge weight * -ln2                     } tile's static part
constexpr int E3_ROWNODE = E3_EW + TN * 4;                 // int[16]  node of each tile row                  }
constexpr int TS_E_BYTES = E3_ROWNODE + 64;                // 640: static epilogue-side part of a tile in global memory
constexpr int E3_DYN = TS_E_BYTES;                         // int[8]: first-of-molecule, q, rescale, f32 operand scale, f32 descale * -log2 e
constexpr int E3_BYTES = E3_DYN + 32;
constexpr int E3_CD = E3_BYTES;                            // COORD only: f32[3][TN] edge weight * normalised difference * -ln2 (bulk copy of tcd[tile])
constexpr int E3C_BYTES = E3_CD + 3 * TN * 4;
constexpr int NES3_COORD = 4;                              // COORD launches have few tiles per CTA and bigger slots
constexpr int E3_REGION = (NES3 * E3_BYTES > NES3_COORD * E3C_BYTES) ? NES3 * E3_BYTES : NES3_COORD * E3C_BYTES;
constexpr int TS_BYTES = TS_P_BYTES + TS_E_BYTES;          // per-tile static record in global memory: [P part | E part]
constexpr int O3_B2 = O3_ET + E3_REGION;                   // f32[128]: b2 * -log2 e (COORD: b2, natural domain)
constexpr int O3_W5 = O3_B2 + H * 4;                       // f32[128]: coord_mlp.4 weight (COORD)
constexpr int O3_PART = O3_W5 + H * 4;                     // f32[3 groups][4 quarters][4]: per-row channel sums of the COORD epilogue
constexpr int O3_BAR = O3_PART + 256;
constexpr int B3_FULL = 0, B3_EMPTY = B3_FULL + 8 * N_STAGE, B3_TBL = B3_EMPTY + 8 * N_STAGE, B3_TFREE = B3_TBL + 8 * NES3,
              B3_PFREE = B3_TFREE + 8 * NES3, B3_TFULL = B3_PFREE + 8 * NPS3, B3_TEMPTY = B3_TFULL + 8 * NACC3,
              B3_PFULL = B3_TEMPTY + 8 * NACC3, B3_PEMPTY = B3_PFULL + 16, B3_W = B3_PEMPTY + 16, B3_TMEMSLOT = B3_W + 8;
constexpr int SMEM3_BYTES = O3_BAR + B3_TMEMSLOT + 16 + 1024;
static_assert(SMEM3_BYTES <= 232448, "k_edge_v3 exceeds the 227 KB of shared memory a CTA can opt into");
static_assert(O3_PANEL % 128 == 0 && P3_A % 16 == 0 && P3_BYTES % 16 == 0 && O3_PT % 16 == 0 && O3_ET % 16 == 0 && E3_BYTES % 16 == 0 &&
              E3C_BYTES % 16 == 0 && TS_P_BYTES % 16 == 0 && TS_E_BYTES % 16 == 0, "TMA destinations need their alignment");

// Per-tile tables in global memory (tile = GCL work item of the plan): see the header comment.
struct TileTables {
  const uint8_t* ts;      // [tiles][TS_BYTES] static records
  const float* td;        // [tiles][TN] squared distances of the current block
  const float* td0;       // [tiles][TN] squared distances of the call's input coordinates
  const float* tdmax;     // [tiles] max of td over the tile
  const float* td0max;    // [tiles]
  const float* tcd;       // COORD tiles: [tiles][3][TN] edge weight * normalised coordinate difference * -ln2 of the current block
  const int4* items;      // the tile list these tables describe (GCL: plan.items over rowidx; COORD: plan.xitems over xrowidx)
  const int* n_items;
  const int* rowidx;
  const int* cta_begin;   // [gridDim.x + 1] cost-balanced contiguous slices of the tile list (k_plan_items)
  int stream_tasks;       // producers: deal the warp-tasks of consecutive tiles round-robin (0: warp w always takes task w)
};

// Warp layout: producers 0..15, MMA issuer 16, table warps 17..19, epilogue 20..31 (20 % 4 == 0: an epilogue warp's TMEM lane
// quarter is warp % 4): 1024 threads, 64 registers at launch. Registers after setmaxnreg:
// 16 x 32 x 80 (producers) + 4 x 32 x 48 (control) + 12 x 32 x 48 (epilogue) = 65,536.
// (8 epilogue warps at 64 registers with double-buffered accumulator loads were measured 35 % slower: 102 vs 76 us.)
constexpr int W3_PROD = 0, W3_MMA = 16, W3_TBL = 17, W3_EPI = 20, NEPI = 12;
struct Regs3 { static constexpr int EPI = 48, CTRL = 48, PROD = 80; };

// operand position p (0..127) -> channel (see "K permutation" above)
__host__ __device__ constexpr int chan_of_pos(int p) { return (p & 4) ? 64 + 4 * (p >> 3) + (p & 3) : 4 * (p >> 3) + (p & 3); }

__device__ __forceinline__ float4 lds128(const void* p) {   // volatile: stays where it is written relative to the tcgen05 asm
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

struct Tile3 {
  int b, nc, ncc4, slot0, nrt, gidx;
  const int* rows;
};

// Warp-synchronous walk over the CTA's contiguous slice of the GCL work items (one item = one tile: k_plan_items
// builds them with the same rows-per-tile rule, col_pad = 4, max_rows = MAXR3).
struct TileIter3 {
  const int4* list;
  const int* rowlist;
  int N, wi_end, wi, rt, cache_base, lane;
  int4 cache;
  __device__ TileIter3(const TileTables& tt, int N_) : N(N_), rt(0), cache_base(-(1 << 30)), lane(threadIdx.x & 31) {
    list = tt.items;
    rowlist = tt.rowidx;
    wi = tt.cta_begin[blockIdx.x];
    wi_end = tt.cta_begin[blockIdx.x + 1];
    cache = make_int4(0, 0, 0, 0);
  }
  __device__ bool next(Tile3& t) {
    while (wi < wi_end) {
      if (wi - cache_base >= 32) {
        cache_base = wi;
        cache = list[min(wi + lane, wi_end - 1)];
      }
      const int src = wi - cache_base;
      const int b = __shfl_sync(0xffffffffu, cache.x, src), r_begin = __shfl_sync(0xffffffffu, cache.y, src);
      const int r_count = __shfl_sync(0xffffffffu, cache.z, src), nc = __shfl_sync(0xffffffffu, cache.w, src);
      const int ncc4 = (nc + 3) & ~3;
      const int per = min(TN / max(ncc4, 4), MAXR3);
      if (rt >= r_count || nc <= 0) { wi += 1; rt = 0; continue; }
      t.b = b; t.nc = nc; t.ncc4 = ncc4; t.slot0 = r_begin + rt; t.nrt = min(per, r_count - rt);
      t.rows = rowlist + (size_t)b * N;
      t.gidx = wi;                                         // one item = one tile (k_plan_items uses the same rows-per-tile rule)
      rt += per;
      return true;
    }
    return false;
  }
};

// COORD = true: EquivariantUpdate (egnn.py:101-125) on the same pipeline. With phi_ij = w5 . silu(...) the update is
//   x_i += sum_j cd_ij ew_ij phi_ij / norm = sum_c w5[c] * ( sum_j (cd_ij ew_ij) m_ij[c] ) / norm,
// i.e. the GCL epilogue's per-channel segment sum with THREE edge weights (cd_x ew, cd_y ew, cd_z ew: tile table tcd) and one
// reduction over the 128 channels per ROW (warp shuffles, then the four lane quarters meet in shared memory in a fixed order)
// instead of one per edge.
template <bool PROF, bool COORD>
__global__ void __launch_bounds__(32 * (W3_EPI + NEPI), 1) k_edge_v3(Geom gm, EdgeArgs a, const uint32_t* __restrict__ w2p,
                                                                    const __grid_constant__ CUtensorMap tm_b, TileTables tt,
                                                                    unsigned long long* __restrict__ prof = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = gm.N;
  constexpr int NG = NEPI / 4;                             // epilogue warps per TMEM lane quarter: they take the tile rows round-robin
  constexpr int NES = COORD ? NES3_COORD : NES3;           // epilogue-side ring depth
  constexpr int EB = COORD ? E3C_BYTES : E3_BYTES;         // epilogue-side slot size
  using RG = Regs3;
  const uint32_t bars = sbase + O3_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + O3_BAR + B3_TMEMSLOT);
  float* b2s = reinterpret_cast<float*>(sm + O3_B2);

  unsigned long long pc[4] = {0, 0, 0, 0};
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
      unsigned long long* o = prof + (size_t)blockIdx.x * PROF_SLOTS + base;
      o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = (unsigned long long)(clock64() - t_begin);
    }
  };
  // timeline marks (cycles since kernel entry; each costs the marking warp a global round trip -- read them as an ordering with
  // ~1 K cycles of overhead per mark on the same warp, not as exact times)
  auto mark = [&](int i) {
    if (PROF && lane == 0) { unsigned long long* o = prof + (size_t)blockIdx.x * PROF_SLOTS + 16 + i; if (*o == 0) *o = (unsigned long long)(clock64() - t_begin); }
  };

  if (tid == 0) {
    for (int i = 0; i < N_STAGE; ++i) { mbar_init(bars + B3_FULL + 8 * i, N_PROD_WARPS); mbar_init(bars + B3_EMPTY + 8 * i, 1); }
    for (int i = 0; i < NES; ++i) { mbar_init(bars + B3_TBL + 8 * i, 1); mbar_init(bars + B3_TFREE + 8 * i, NEPI); }
    for (int i = 0; i < NPS3; ++i) mbar_init(bars + B3_PFREE + 8 * i, N_PROD_WARPS);
    for (int i = 0; i < NACC3; ++i) { mbar_init(bars + B3_TFULL + 8 * i, 1); mbar_init(bars + B3_TEMPTY + 8 * i, NEPI); }
    for (int i = 0; i < 2; ++i) { mbar_init(bars + B3_PFULL + 8 * i, 1); mbar_init(bars + B3_PEMPTY + 8 * i, N_PROD_WARPS); }
    mbar_init(bars + B3_W, 8);
    fence_barrier_init();
  }
  if (tid < H) {
    b2s[tid] = a.b2[tid] * -1.4426950408889634f;
    if (COORD) reinterpret_cast<float*>(sm + O3_W5)[tid] = a.w5[tid];
  }
  if (warp == W3_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= W3_MMA && warp < W3_EPI) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::CTRL));
    if (warp >= W3_TBL) {
      // =================================== table warps =================================================================
      // Tile-parallel: table warp k serves tiles t = k, k+3, ...; all three walk the same tile sequence.
      const int tw = warp - W3_TBL;
      // Everything this CTA reads that another kernel of the forward produces (projections, tile tables, coordinates) enters
      // through the table warps, and everything it writes leaves after tiles have travelled the pipeline: the table warps
      // are the only ones that have to wait for the previous kernel. The prologue above and the W2 -> tensor memory copy of
      // the epilogue warps overlap the previous kernel's tail.
      if (warp == W3_TBL) mark(0);                          // prologue done
      chain_wait();
      if (warp == W3_TBL) mark(1);                          // previous kernel complete
      if (warp == W3_TBL && lane == 0) chain_release();
      // (Walking the per-plan tile list for the first tile BEFORE the wait -- it does not depend on the chain -- was measured
      // slower, 77.9 vs 75.3 us per GCL launch on the same box.)
      TileIter3 iter(tt, N);
      Tile3 cur;
      int q = -1, prev_b = -1, ps = 0, pu = 0;              // ps = t % NPS3, pu = t / NPS3
      float mol_abmax = 0.f;                                // max |A_i| + max |B_j| over the current molecule (range bound)
      for (int t = 0;; ++t, ps = (ps + 1 == NPS3 ? 0 : ps + 1), pu += (ps == 0)) {
        const int slot = t & (NES - 1);
        const bool more = iter.next(cur);
        bool first = false;
        if (more && cur.b != prev_b) { first = true; ++q; prev_b = cur.b; }
        if (more && first) {                                 // every table warp tracks the molecule bound (cheap, keeps them in step)
          float ma = 0.f, mb = 0.f;
          for (int j = lane; j < N; j += 32) {
            const float2 v = *reinterpret_cast<const float2*>(a.ABmax + ((size_t)cur.b * N + j) * 2);
            ma = fmaxf(ma, v.x); mb = fmaxf(mb, v.y);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o)); }
          mol_abmax = ma + mb;
        }
        if (t % N_TBL_WARPS != tw) { if (!more) break; continue; }
        if (t >= NES) wait_relaxed(bars + B3_TFREE + 8 * slot, ((t - NES) / NES) & 1, 0);        // epilogue done with tile t - NES
        if (pu > 0) wait_relaxed(bars + B3_PFREE + 8 * ps, (pu - 1) & 1, 2);                         // producers done with tile t - 3
        uint8_t* tb = sm + O3_ET + slot * EB;                  // epilogue-side slot (+ headers)
        uint8_t* pb = sm + O3_PT + ps * P3_BYTES;              // producer-side slot
        const uint32_t bar = bars + B3_TBL + 8 * slot;
        if (more) {
          const size_t gb = (size_t)cur.b * N;
          if (first) {                                       // this molecule's B panel: one tiled TMA load
            const int buf = q & 1;
            if (q >= 2) wait_relaxed(bars + B3_PEMPTY + 8 * buf, ((q - 2) >> 1) & 1, 1);   // producers left molecule q-2
            if (lane == 0) {
              mbar_expect_tx(bars + B3_PFULL + 8 * buf, (uint32_t)N * H * 4);
              tma_load_2d(sbase + O3_PANEL + buf * PANEL_BYTES, &tm_b, H, cur.b * N, bars + B3_PFULL + 8 * buf);
            }
          }
          if (lane == 0) mbar_expect_tx_only(bar, (uint32_t)(cur.nrt * H * 4 + TS_BYTES + 2 * TN * 4 + (COORD ? 3 * TN * 4 : 0)));
          __syncwarp();
          if (lane < cur.nrt) {                              // A_i rows of the tile: 512-byte bulk copies into the slot
            const int node_r = cur.rows[cur.slot0 + lane];
            bulk_g2s(smem_u32(pb + P3_A) + lane * (H * 4), a.AB + (gb + node_r) * 2 * H, H * 4, bar);
          } else if (lane == 8) {
            bulk_g2s(smem_u32(pb + P3_BOFF), tt.ts + (size_t)cur.gidx * TS_BYTES, TS_P_BYTES, bar);
          } else if (lane == 9) {
            bulk_g2s(smem_u32(tb + E3_HDR), tt.ts + (size_t)cur.gidx * TS_BYTES + TS_P_BYTES, TS_E_BYTES, bar);
          } else if (lane == 10) {
            bulk_g2s(smem_u32(pb + P3_D), tt.td + (size_t)cur.gidx * TN, TN * 4, bar);
          } else if (lane == 11) {
            bulk_g2s(smem_u32(pb + P3_D0), tt.td0 + (size_t)cur.gidx * TN, TN * 4, bar);
          } else if (COORD && lane == 12) {
            bulk_g2s(smem_u32(tb + E3_CD), tt.tcd + (size_t)cur.gidx * 3 * TN, 3 * TN * 4, bar);
          }
          if (lane == 0) {
            // |u| <= max|A_i| + max|B_j| (over the molecule) + d max|wd| + d0 max|w0| (over the tile): ONE exact power-of-two
            // scale per tile keeps the fp16 hi/lo operands below 2^14 (only ever != 1 for diverging samples)
            const float bound = mol_abmax + tt.tdmax[cur.gidx] * a.wdmax + tt.td0max[cur.gidx] * a.w0max;
            float sc = 1.0f;
            if (!(bound <= F16_TARGET)) {
              const int ex2 = ((__float_as_int(bound) >> 23) & 0xff) - 127;
              sc = __int_as_float(max(127 + 13 - ex2, 1) << 23);
            }
            int* dyn = reinterpret_cast<int*>(tb + E3_DYN);
            dyn[0] = first ? 1 : 0; dyn[1] = q; dyn[2] = sc != 1.0f ? 1 : 0;
            dyn[3] = __float_as_int(sc);
            dyn[4] = __float_as_int((a.w2_descale / sc) * -1.4426950408889634f);
          }
        } else if (lane == 0) {
          reinterpret_cast<int*>(tb + E3_HDR)[0] = 0;         // end marker travels through the whole pipeline
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
        if (t == 0) mark(2);                                 // first tile's copies issued
        if (!more) break;
      }
      if (warp == W3_TBL) prof_flush(0);
    } else {
      // =================================== MMA issuer ===================================================================
      // warp-convergent: all lanes walk the loop with uniform values, one elected lane issues each tcgen05 instruction
      {
        mbar_wait(bars + B3_W, 0);                           // W2 hi | lo are in tensor memory
        tc_fence_after();
        int acc = 0, use = 0;
        for (int t = 0;; ++t) {
          const int s = t & (N_STAGE - 1), slot = t & (NES - 1);
          wait_relaxed(bars + B3_FULL + 8 * s, (t / N_STAGE) & 1, 0);
          // the epilogue drained this accumulator (also before the end marker: its plain arrive below must not land in the
          // phase a still-running commit of tile t-3 is about to complete)
          if (use > 0) wait_on(bars + B3_TEMPTY + 8 * acc, (use - 1) & 1, 1);
          const int Et = reinterpret_cast<const int*>(sm + O3_ET + slot * EB + E3_HDR)[0];
          if (Et <= 0) { if (lane == 0) mbar_arrive(bars + B3_TFULL + 8 * acc); break; }
          tc_fence_after();
          const uint32_t bhi = sbase + O3_ST + s * STAGE_BYTES, blo = bhi + B_BYTES;
          const uint32_t dcol = tmem + TM_ACC + acc * TN;
          const uint32_t idesc = umma_idesc(128, max(16, (Et + 15) & ~15));
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t b_hi = umma_desc(bhi + ks * 2 * B_LBO, B_LBO, SBO), b_lo = umma_desc(blo + ks * 2 * B_LBO, B_LBO, SBO);
            const uint32_t a_hi = tmem + TM_W + ks * 8, a_lo = tmem + TM_W + 64 + ks * 8;
            umma_f16_ts_elect(dcol, a_lo, b_hi, idesc, ks > 0);
            umma_f16_ts_elect(dcol, a_hi, b_lo, idesc, 1);
            umma_f16_ts_elect(dcol, a_hi, b_hi, idesc, 1);
          }
          umma_commit_elect(bars + B3_EMPTY + 8 * s);
          umma_commit_elect(bars + B3_TFULL + 8 * acc);
          if (t == 0) mark(5);                               // first tile's MMAs issued
          pc[2] += 1;
          if (++acc == NACC3) { acc = 0; ++use; }
        }
        prof_flush(8);
      }
    }
  } else if (warp < W3_MMA) {
    // =================================== producers =======================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RG::PROD));
    const int pw = warp - W3_PROD;
    const int kc = lane & 15, esub = lane >> 4;
    float2 wdr[4], w0r[4];
    {
      const float4 d0v = __ldg(reinterpret_cast<const float4*>(a.wd + 4 * kc)), d1v = __ldg(reinterpret_cast<const float4*>(a.wd + 64 + 4 * kc));
      const float4 z0v = __ldg(reinterpret_cast<const float4*>(a.w0 + 4 * kc)), z1v = __ldg(reinterpret_cast<const float4*>(a.w0 + 64 + 4 * kc));
      wdr[0] = make_float2(d0v.x, d0v.y); wdr[1] = make_float2(d0v.z, d0v.w); wdr[2] = make_float2(d1v.x, d1v.y); wdr[3] = make_float2(d1v.z, d1v.w);
      w0r[0] = make_float2(z0v.x, z0v.y); w0r[1] = make_float2(z0v.z, z0v.w); w0r[2] = make_float2(z1v.x, z1v.y); w0r[3] = make_float2(z1v.z, z1v.w);
    }
    const uint8_t* panel = sm + O3_PANEL + kc * 16;
    // A tile's producer work is ceil(Et / 8) warp-tasks (two groups of four consecutive edges x 16 channel chunks). The tasks of
    // consecutive tiles form one stream dealt round-robin to the 16 producer warps, so a tile with 15 tasks (3 rows of 40)
    // leaves one warp free to start on the next tile and the short last tile of a molecule (5 tasks) occupies 5 warps, not 16.
    int task_base = 0;                                       // tasks of all earlier tiles, mod 16
    int ps = 0;
    for (int t = 0;; ++t, ps = (ps + 1 == NPS3 ? 0 : ps + 1)) {
      const int s = t & (N_STAGE - 1), slot = t & (NES - 1);
      wait_on(bars + B3_TBL + 8 * slot, (t / NES) & 1, 0);
      if (t == 0 && warp == W3_PROD) mark(3);               // first tile's tables have landed
      const uint8_t* tb = sm + O3_PT + ps * P3_BYTES;        // producer-side slot
      const int* hdr = reinterpret_cast<const int*>(sm + O3_ET + slot * EB + E3_HDR);
      const int* dyn = reinterpret_cast<const int*>(sm + O3_ET + slot * EB + E3_DYN);
      const int Et = hdr[0];
      if (Et > 0 && dyn[0]) {                                // first tile of a molecule: switch panel buffers
        const int q = dyn[1], buf = q & 1;
        if (q >= 1) { __syncwarp(); if (lane == 0) mbar_arrive(bars + B3_PEMPTY + 8 * (buf ^ 1)); }
        wait_on(bars + B3_PFULL + 8 * buf, (q >> 1) & 1, 2);
        panel = sm + O3_PANEL + buf * PANEL_BYTES + kc * 16;
      }
      if (t >= N_STAGE) wait_on(bars + B3_EMPTY + 8 * s, ((t - N_STAGE) / N_STAGE) & 1, 1);
      if (Et > 0) {
        const int grp = 2 * ((pw - task_base) & (N_PROD_WARPS - 1)) + esub;   // this thread's group of four consecutive edges
        const int e_base = 4 * grp;
        if (tt.stream_tasks) task_base = (task_base + ((Et + 7) >> 3)) & (N_PROD_WARPS - 1);
        if (e_base < Et) {
          uint8_t* bhi = sm + O3_ST + s * STAGE_BYTES + kc * B_LBO + e_base * 16;
          uint8_t* blo = bhi + B_BYTES;
          // the whole group's scalars in three 16-byte loads; B_j of the NEXT edge is fetched before the current one is
          // processed (the compiler cannot hoist those loads over the operand stores itself: same address space)
          const float4 dq = *reinterpret_cast<const float4*>(tb + P3_D + e_base * 4);
          const float4 d0q = *reinterpret_cast<const float4*>(tb + P3_D0 + e_base * 4);
          const int4 bo = *reinterpret_cast<const int4*>(tb + P3_BOFF + e_base * 4);
          const uint8_t* arow = tb + P3_A + kc * 16 + reinterpret_cast<const int*>(tb + P3_GRP)[grp];
          const float4 a0 = *reinterpret_cast<const float4*>(arow), a1 = *reinterpret_cast<const float4*>(arow + 256);
          const float2 av[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
          const float dq_[4] = {dq.x, dq.y, dq.z, dq.w}, d0q_[4] = {d0q.x, d0q.y, d0q.z, d0q.w};
          const int bo_[4] = {bo.x, bo.y, bo.z, bo.w};
          auto run_items = [&](auto rs_tag) {
            constexpr bool RS = decltype(rs_tag)::value;
            float4 b0 = *reinterpret_cast<const float4*>(panel + bo_[0]), b1 = *reinterpret_cast<const float4*>(panel + bo_[0] + 256);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              float4 n0 = b0, n1 = b1;
              if (it < 3) { n0 = *reinterpret_cast<const float4*>(panel + bo_[it + 1]); n1 = *reinterpret_cast<const float4*>(panel + bo_[it + 1] + 256); }
              const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
              const float2 dd = make_float2(dq_[it], dq_[it]), dd0 = make_float2(d0q_[it], d0q_[it]);
              float2 u[4], sv[4];
#pragma unroll
              for (int k = 0; k < 4; ++k)                     // egnn.py:49-50 in the log2 domain, two channels per instruction
                u[k] = __ffma2_rn(dd0, w0r[k], __ffma2_rn(dd, wdr[k], __fadd2_rn(av[k], bv[k])));
              usig4(u[0], u[1], sv[0], sv[1]);
              usig4(u[2], u[3], sv[2], sv[3]);
              if (RS) {                                      // rare: diverging samples only (tile-uniform power of two)
                const float sc = __int_as_float(dyn[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) sv[k] = __fmul2_rn(sv[k], make_float2(sc, sc));
              }
              uint4 hi, lo;
              split2v(sv[0], hi.x, lo.x); split2v(sv[1], hi.y, lo.y);
              split2v(sv[2], hi.z, lo.z); split2v(sv[3], hi.w, lo.w);
              *reinterpret_cast<uint4*>(bhi + it * 16) = hi;
              *reinterpret_cast<uint4*>(blo + it * 16) = lo;
              b0 = n0; b1 = n1;
            }
          };
          if (dyn[2] != 0) run_items(std::true_type{}); else run_items(std::false_type{});
        }
        fence_proxy_async();
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(bars + B3_FULL + 8 * s); mbar_arrive(bars + B3_PFREE + 8 * ps); }
      if (t == 0 && warp == W3_PROD) mark(4);               // first operand tile written
      if (Et <= 0) { if (warp == W3_PROD) mark(8); break; }
    }
    if (warp == W3_PROD) prof_flush(4);
  } else {
    // =================================== epilogue warps ====================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RG::EPI));   // first: the producers' increase waits for it
    const int q = warp & 3;                                  // TMEM lane quarter
    const int g = (warp - W3_EPI) >> 2;                      // prologue: 0 -> W2 hi, 1 -> W2 lo; main loop: row phase
    const int c = q * 32 + lane;                             // output channel = TMEM lane
    if (g < 2) {
      // W2 (this thread's output row, hi or lo half: 64 packed words) -> tensor memory columns [64 g, 64 g + 64)
      const uint4* src = reinterpret_cast<const uint4*>(w2p + ((size_t)g * H + c) * 64);
      const uint32_t tw = tmem + ((uint32_t)(q * 32) << 16) + TM_W + g * 64;
#pragma unroll 1
      for (int i4 = 0; i4 < 4; ++i4) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint4 v = __ldg(src + i4 * 4 + i);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        TMEM_ST_X16(tw + i4 * 16, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B3_W);
      if (warp == W3_EPI) mark(9);                           // W2 in tensor memory
    }
    const float bias = b2s[c];
    const float2 bias2 = make_float2(bias, bias);
    int acc = 0, use = 0, rot = g;                           // rot = (g + t) % NG: the warps of a lane quarter rotate over the rows
    for (int t = 0;; ++t) {
      const int slot = t & (NES - 1);
      wait_on(bars + B3_TBL + 8 * slot, (t / NES) & 1, 0);
      wait_on(bars + B3_TFULL + 8 * acc, use & 1, 1);
      if (t == 0 && warp == W3_EPI) mark(6);                // first accumulator complete
      tc_fence_after();
      const uint8_t* tb = sm + O3_ET + slot * EB;
      const int* hdr = reinterpret_cast<const int*>(tb + E3_HDR);
      const int Et = hdr[0];
      if (Et <= 0) break;
      const int nrt = hdr[1], ncc4 = hdr[2];
      const size_t gb = (size_t)hdr[3] * N;
      const float* ews = reinterpret_cast<const float*>(tb + E3_EW);
      const int* rownode = reinterpret_cast<const int*>(tb + E3_ROWNODE);
      const float ds0 = __int_as_float(reinterpret_cast<const int*>(tb + E3_DYN)[4]);   // tile-uniform descale * -log2 e
      const float2 ds2 = make_float2(ds0, ds0);
      const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + TM_ACC + acc * TN;
      // One row: this thread's channel over the row's columns. u = -log2e (D descale + b2) (packed FMA), u * sigmoid (usig4: four
      // columns per reciprocal), times the edge weight (which carries the -ln2) into packed partial sums -- fixed order,
      // deterministic.
      auto do_row = [&](int rr) {
        const int col0 = rr * ncc4;
        const int n8 = ncc4 >> 3;
        float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
        auto quad = [&](const uint32_t* r, const float4 ew0, float2& sa, float2& sb) {
          const float2 u0 = __ffma2_rn(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), ds2, bias2);
          const float2 u1 = __ffma2_rn(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), ds2, bias2);
          float2 g0, g1;
          usig4(u0, u1, g0, g1);
          sa = __ffma2_rn(g0, make_float2(ew0.x, ew0.y), sa);
          sb = __ffma2_rn(g1, make_float2(ew0.z, ew0.w), sb);
        };
        for (int k = 0; k < n8; ++k) {                       // 48 registers per thread: parallelism comes from the 12 warps
          uint32_t ra[8];
          TMEM_LD_X8(tacc + col0 + k * 8, ra);
          const float4 ew0 = lds128(ews + col0 + k * 8), ew1 = lds128(ews + col0 + k * 8 + 4);
          tmem_ld_wait();
          quad(ra, ew0, s0, s1);
          quad(ra + 4, ew1, s2, s3);
        }
        if (ncc4 & 4) {
          uint32_t r4[4];
          const int col = col0 + n8 * 8;
          TMEM_LD_X4(tacc + col, r4);
          const float4 ew0 = lds128(ews + col);
          tmem_ld_wait();
          quad(r4, ew0, s0, s1);
        }
        const float accv = ((s0.x + s0.y) + (s1.x + s1.y)) + ((s2.x + s2.y) + (s3.x + s3.y));
        a.agg[(gb + rownode[rr]) * H + c] = accv / gm.normalization_factor;   // egnn.py:312-313
      };
      // COORD: three weighted segment sums per channel, then one reduction over the channels per row (see the kernel comment)
      auto do_row_coord = [&](int rr) {
        const int col0 = rr * ncc4;
        const float* wx = reinterpret_cast<const float*>(tb + E3_CD);
        float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
        for (int k = 0; k < ncc4; k += 4) {
          uint32_t r[4];
          TMEM_LD_X4(tacc + col0 + k, r);
          const float4 w0 = lds128(wx + col0 + k), w1 = lds128(wx + TN + col0 + k), w2 = lds128(wx + 2 * TN + col0 + k);
          tmem_ld_wait();
          const float2 u0 = __ffma2_rn(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), ds2, bias2);
          const float2 u1 = __ffma2_rn(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), ds2, bias2);
          float2 g0, g1;
          usig4(u0, u1, g0, g1);
          ax = __ffma2_rn(g0, make_float2(w0.x, w0.y), ax); ax = __ffma2_rn(g1, make_float2(w0.z, w0.w), ax);
          ay = __ffma2_rn(g0, make_float2(w1.x, w1.y), ay); ay = __ffma2_rn(g1, make_float2(w1.z, w1.w), ay);
          az = __ffma2_rn(g0, make_float2(w2.x, w2.y), az); az = __ffma2_rn(g1, make_float2(w2.z, w2.w), az);
        }
        const float w5c = reinterpret_cast<const float*>(sm + O3_W5)[c];
        float vx = (ax.x + ax.y) * w5c, vy = (ay.x + ay.y) * w5c, vz = (az.x + az.y) * w5c;   // coord_mlp.4 (egnn.py:96-97, no bias)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          vx += __shfl_xor_sync(0xffffffffu, vx, o); vy += __shfl_xor_sync(0xffffffffu, vy, o); vz += __shfl_xor_sync(0xffffffffu, vz, o);
        }
        float* part = reinterpret_cast<float*>(sm + O3_PART) + g * 16;
        named_sync(1 + g, 128);                              // the group's previous row has been read
        if (lane == 0) { part[q * 4 + 0] = vx; part[q * 4 + 1] = vy; part[q * 4 + 2] = vz; }
        named_sync(1 + g, 128);                              // the four lane quarters of this group meet (fixed order below: deterministic)
        if (q == 0 && lane < 3) {
          const float sacc = ((part[lane] + part[4 + lane]) + part[8 + lane]) + part[12 + lane];
          const int i = rownode[rr];
          const float lm = a.linker_mask ? a.linker_mask[gb + i] : 1.f;
          const float xv = a.x[(gb + i) * 3 + lane];
          const float xn = (xv + (sacc / gm.normalization_factor) * lm) * a.nm[gb + i];   // egnn.py:110-124
          a.x_out[(gb + i) * 3 + lane] = xn;
          reinterpret_cast<float*>(a.x4_out + gb + i)[lane] = xn;
        }
      };
      if constexpr (COORD) { for (int rr = rot; rr < nrt; rr += NG) do_row_coord(rr); }
      else { for (int rr = rot; rr < nrt; rr += NG) do_row(rr); }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bars + B3_TEMPTY + 8 * acc); mbar_arrive(bars + B3_TFREE + 8 * slot); }
      if (t == 0 && warp == W3_EPI) mark(7);                // first tile's epilogue done
      if (++acc == NACC3) { acc = 0; ++use; }
      if (++rot == NG) rot = 0;
    }
    if (warp == W3_EPI) prof_flush(12);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W3_MMA) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------
// Tile tables (see TileTables): one CTA of TN threads per tile, thread = tile column (edge slot)
// ---------------------------------------------------------------------------------------------------------
// Static part, once per plan (masks are constant over a sampling chain): panel offsets, A-row offsets, the header, the
// edge weights (caller's int8 edge_mask value, egnn.py:55-58, times -ln2; padding slots mirror a real edge with weight 0)
// and the (row node, column node) pair of every slot for k_tiles_d.
__global__ void __launch_bounds__(TN) k_tiles_static(int N, const int4* __restrict__ items, const int* __restrict__ n_items,
                                                    const int* __restrict__ rowidx, const int* __restrict__ colidx,
                                                    const int8_t* __restrict__ edge_mask, uint8_t* __restrict__ ts, int2* __restrict__ tij) {
  const int g = blockIdx.x, e = threadIdx.x;
  if (g >= *n_items) return;
  const int4 it = items[g];
  const int b = it.x, r0 = it.y, nrt = it.z, nc = it.w, ncc4 = (nc + 3) & ~3;
  const size_t gb = (size_t)b * N;
  int rr = e / ncc4;
  int jj = e - rr * ncc4;
  const bool valid = rr < nrt && jj < nc;
  rr = min(rr, nrt - 1); jj = min(jj, nc - 1);
  const int i = rowidx[gb + r0 + rr], j = colidx[gb + jj];
  uint8_t* rec = ts + (size_t)g * TS_BYTES;
  reinterpret_cast<int*>(rec)[e] = j * (H * 4);
  if (e < TN / 4) reinterpret_cast<int*>(rec + TN * 4)[e] = min((4 * e) / ncc4, nrt - 1) * (H * 4);
  uint8_t* er = rec + TS_P_BYTES;
  if (e < 16) {
    reinterpret_cast<int*>(er + E3_HDR)[e] = e == 0 ? nrt * ncc4 : e == 1 ? nrt : e == 2 ? ncc4 : e == 3 ? b : 0;
    reinterpret_cast<int*>(er + E3_ROWNODE)[e] = e < nrt ? rowidx[gb + r0 + e] : 0;
  }
  float ew = 0.f;
  if (valid) ew = edge_mask ? (float)edge_mask[gb * N + (size_t)i * N + j] : 1.0f;
  reinterpret_cast<float*>(er + E3_EW)[e] = ew * -0.6931471805599453f;
  tij[(size_t)g * TN + e] = make_int2((int)(gb + i), (int)(gb + j));
}

// Squared distances of every tile slot from the coordinates `x4` (egnn.py:297-298; with the call's input coordinates also
// egnn.py:220 -- at block 0 both coincide, so td0 is written by the same launch) and their per-tile maxima, for the GCL tile
// set (blockIdx.y = 0) and the COORD tile set (blockIdx.y = 1); the latter also gets its three weighted edge-weight arrays
// ew * (x_i - x_j) / (sqrt(d + 1e-8) + norm_constant) (egnn.py:299-300, 107-109; ew already carries the -ln2). The y = 0 CTAs
// also carry the x -> x_next copy that precedes a block's coordinate update (rows the update does not touch keep x).
struct TileDSet {
  const int* n_items; const int2* tij; float* td; float* tdmax; float* td0; float* td0max;
  const uint8_t* ts; float* tcd;   // COORD set only (else null)
};
__global__ void __launch_bounds__(TN) k_tiles_d(TileDSet s0, TileDSet s1, const float4* __restrict__ x4, int write_d0, float norm_constant,
                                               int n3, const float* __restrict__ xsrc, float* __restrict__ xdst,
                                               const float4* __restrict__ x4src, float4* __restrict__ x4dst) {
  __shared__ float red[2][TN / 32];
  const int e = threadIdx.x;
  const TileDSet& s = blockIdx.y == 0 ? s0 : s1;
  chain_wait();
  chain_release();
  if (blockIdx.y == 0 && xdst != nullptr) {                // grid-stride copy (the grid is a few CTAs per SM, not one per tile)
    for (int i = blockIdx.x * TN + e; i < n3; i += gridDim.x * TN) {
      xdst[i] = xsrc[i];
      if (i * 3 < n3) x4dst[i] = x4src[i];
    }
  }
  if (s.n_items == nullptr) return;
  const int n_tiles = *s.n_items;
  int par = 0;
  for (int g = blockIdx.x; g < n_tiles; g += gridDim.x, par ^= 1) {
  const int2 ij = s.tij[(size_t)g * TN + e];
  const float4 xi = x4[ij.x], xj = x4[ij.y];
  const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
  const float d = dx * dx + dy * dy + dz * dz;
  s.td[(size_t)g * TN + e] = d;
  if (write_d0) s.td0[(size_t)g * TN + e] = d;
  if (s.tcd != nullptr) {
    const float ew = reinterpret_cast<const float*>(s.ts + (size_t)g * TS_BYTES + TS_P_BYTES + E3_EW)[e];
    const float inv = ew / (sqrtf(d + 1e-8f) + norm_constant);
    float* o = s.tcd + (size_t)g * 3 * TN;
    o[e] = dx * inv; o[TN + e] = dy * inv; o[2 * TN + e] = dz * inv;
  }
  float m = d;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((e & 31) == 0) red[par][e >> 5] = m;                 // double-buffered: one barrier per tile
  __syncthreads();
  if (e == 0) {
    m = fmaxf(fmaxf(red[par][0], red[par][1]), fmaxf(red[par][2], red[par][3]));
    s.tdmax[g] = m;
    if (write_d0) s.td0max[g] = m;
  }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
// edge_mlp.2 weight (out = 128 rows, in = 128, row-major) -> [hi | lo][out][64 words], word w of a row = fp16 pair of
// operand positions (2w, 2w+1), positions permuted by chan_of_pos, values scaled by -ln2 (log2-domain first layer,
// kernels_tc.cuh) and by the power of two that puts max|W| in [2^13, 2^14). Returns the offset in halves; *descale = 1/scale.
inline size_t pack_w2_v3(const std::vector<float>& W_in, std::vector<__half>& blob, float* descale) {
  while (blob.size() % 64) blob.push_back(__float2half(0.f));
  const size_t off = blob.size();
  blob.resize(off + 2 * (size_t)H * H);
  std::vector<float> W(W_in.size());
  float mx = 0.f;
  for (size_t i = 0; i < W.size(); ++i) { W[i] = (float)((double)W_in[i] * NEG_LN2); mx = std::max(mx, std::fabs(W[i])); }
  int ex = 0;
  if (mx > 0.f && std::isfinite(mx)) { std::frexp(mx, &ex); ex -= 1; }
  const int sh = std::min(std::max(13 - ex, -40), 40);
  const float scale = std::ldexp(1.0f, sh);
  *descale = std::ldexp(1.0f, -sh);
  for (int c = 0; c < H; ++c)
    for (int p = 0; p < H; ++p) {
      const float v = W[(size_t)c * H + chan_of_pos(p)] * scale;
      const __half hi = __float2half_rn(v);
      blob[off + (size_t)c * H + p] = hi;
      blob[off + (size_t)H * H + (size_t)c * H + p] = __float2half_rn(v - __half2float(hi));
    }
  return off;
}

inline bool supports(const Geom& gm) { return gm.graph_type == 0 && gm.N <= PANEL_N; }

// Tensor map of one (B*N, 256) fp32 projection buffer: box = the B half (128 floats) of one molecule's N rows.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<encode_tiled_fn>(p);
  return fn;
}
inline dl_status make_panel_map(CUtensorMap* out, const float* AB, int B, int N) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return DL_ERR_CUDA;
  const cuuint64_t gdim[2] = {(cuuint64_t)(2 * H), (cuuint64_t)B * N};
  const cuuint64_t gstride[1] = {(cuuint64_t)(2 * H) * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)H, (cuuint32_t)N};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(AB), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DL_OK : DL_ERR_CUDA;
}

inline dl_status configure3() {
  const bool ok = cudaFuncSetAttribute(k_edge_v3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES) == cudaSuccess &&
                  cudaFuncSetAttribute(k_edge_v3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES) == cudaSuccess &&
                  cudaFuncSetAttribute(k_edge_v3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES) == cudaSuccess &&
                  cudaFuncSetAttribute(k_edge_v3<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM3_BYTES) == cudaSuccess;
  return ok ? DL_OK : DL_ERR_CUDA;
}

inline dl_status profile_edge_v3(const Geom& gm, const EdgeArgs& ea, const void* w2_v3, const CUtensorMap& tm, const TileTables& tt, int num_sms,
                                 cudaStream_t st, bool coord);

inline void launch_edge_v3(const Geom& gm, const EdgeArgs& ea, bool coord, const void* w2_v3, const CUtensorMap& tm, const TileTables& tt,
                           int num_sms, cudaStream_t st) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(w2_v3);
  // DL_PROFILE_EDGE_LIVE=k: the launches k .. k+17 of the process (one forward's worth at L=6) run the profiling variant,
  // each synchronised and printed -- eager Dynamics.forward only (not inside a stream capture)
  static const int live0 = [] { const char* v = getenv("DL_PROFILE_EDGE_LIVE"); return v ? atoi(v) : -1; }();
  static int n_launch = 0;
  if (live0 >= 0) {
    const int i = n_launch++;
    if (i >= live0 && i < live0 + 18) { profile_edge_v3(gm, ea, w2_v3, tm, tt, num_sms, st, coord); return; }
  }
  if (coord) launch_chain(k_edge_v3<false, true>, dim3(num_sms), dim3(32 * (W3_EPI + NEPI)), SMEM3_BYTES, st, gm, ea, w, tm, tt, nullptr);
  else launch_chain(k_edge_v3<false, false>, dim3(num_sms), dim3(32 * (W3_EPI + NEPI)), SMEM3_BYTES, st, gm, ea, w, tm, tt, nullptr);
}

// Debug: one profiled launch (clock64 accounting per role: wait vs total cycles, and a timeline of the first tile), averaged
// over the CTAs, to stderr. Synchronises the stream: eager calls only.
inline dl_status profile_edge_v3(const Geom& gm, const EdgeArgs& ea, const void* w2_v3, const CUtensorMap& tm, const TileTables& tt, int num_sms,
                                 cudaStream_t st, bool coord) {
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, (size_t)num_sms * PROF_SLOTS * 8) != cudaSuccess) return DL_ERR_CUDA;
  cudaMemsetAsync(d, 0, (size_t)num_sms * PROF_SLOTS * 8, st);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(w2_v3);
  if (coord) launch_chain(k_edge_v3<true, true>, dim3(num_sms), dim3(32 * (W3_EPI + NEPI)), SMEM3_BYTES, st, gm, ea, w, tm, tt, d);
  else launch_chain(k_edge_v3<true, false>, dim3(num_sms), dim3(32 * (W3_EPI + NEPI)), SMEM3_BYTES, st, gm, ea, w, tm, tt, d);
  if (cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(d); return DL_ERR_CUDA; }
  std::vector<unsigned long long> h((size_t)num_sms * PROF_SLOTS);
  cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  double avg[PROF_SLOTS] = {0}, mx[PROF_SLOTS] = {0};
  for (int b = 0; b < num_sms; ++b)
    for (int i = 0; i < PROF_SLOTS; ++i) { const double v = (double)h[(size_t)b * PROF_SLOTS + i]; avg[i] += v / num_sms; mx[i] = std::max(mx[i], v); }
  const char* k = coord ? "COORD" : "GCL";
  fprintf(stderr, "[dl prof v3 %s] cycles per CTA (avg over %d), tiles %.1f (max %.0f)\n", k, num_sms, avg[10], mx[10]);
  fprintf(stderr, "[dl prof v3 %s]  table   : wait tfree %.0f, wait pempty %.0f, wait pfree %.0f | total %.0f\n", k, avg[0], avg[1], avg[2], avg[3]);
  fprintf(stderr, "[dl prof v3 %s]  producer: wait tbl %.0f, wait empty %.0f, wait panel %.0f | total %.0f\n", k, avg[4], avg[5], avg[6], avg[7]);
  fprintf(stderr, "[dl prof v3 %s]  mma     : wait full %.0f, wait tempty %.0f | total %.0f\n", k, avg[8], avg[9], avg[11]);
  fprintf(stderr, "[dl prof v3 %s]  epilogue: wait tbl %.0f, wait tfull %.0f | total %.0f (max %.0f)\n", k, avg[12], avg[13], avg[15], mx[15]);
  fprintf(stderr, "[dl prof v3 %s]  timeline: prologue %.0f | chain wait over %.0f | copies of tile 0 issued %.0f | tables landed %.0f | operand tile 0 %.0f | "
                  "MMAs of tile 0 issued %.0f | W2 in TMEM %.0f | accumulator 0 complete %.0f | epilogue of tile 0 done %.0f | producers done %.0f\n",
          k, avg[16], avg[17], avg[18], avg[19], avg[20], avg[21], avg[25], avg[22], avg[23], avg[24]);
  return DL_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Self test of the TS form: A (128 x 128, hi | lo) written to tensor memory with tcgen05.st in the layout k_edge_v3
// uses, B (256 rows, K-major canonical layout) in shared memory, 3xFP16 chain, against fp64 on the host.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_umma_probe_ts(const uint32_t* __restrict__ A /*[2][128][64] words*/,
                                                          const __half* __restrict__ Bm /*[2][kc][256][8]*/,
                                                          float* __restrict__ D /*[128][256]*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar = sbase + P_OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P_OFF_BAR + 32);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  for (int idx = tid; idx < 2 * KC * PN; idx += 128) {
    const int copy = idx / (KC * PN), rem = idx % (KC * PN), kcx = rem / PN, row = rem % PN;
    *reinterpret_cast<uint4*>(sm + (copy ? P_OFF_BLO : P_OFF_BHI) + kcx * P_LBO + row * 16) =
        *reinterpret_cast<const uint4*>(Bm + (size_t)idx * 8);
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  {
    const int c = warp * 32 + lane;
    for (int half = 0; half < 2; ++half) {
      const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16) + 256 + half * 64;   // A at columns 256..383, D at 0..255
      for (int g = 0; g < 4; ++g) {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = A[((size_t)half * H + c) * 64 + g * 16 + i];
        TMEM_ST_X16(tw + g * 16, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(128, 256);
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t b_hi = umma_desc(sbase + P_OFF_BHI + ks * 2 * P_LBO, P_LBO, SBO), b_lo = umma_desc(sbase + P_OFF_BLO + ks * 2 * P_LBO, P_LBO, SBO);
      const uint32_t a_hi = tmem + 256 + ks * 8, a_lo = tmem + 256 + 64 + ks * 8;
      umma_f16_ts(tmem, a_lo, b_hi, idesc, ks > 0);
      umma_f16_ts(tmem, a_hi, b_lo, idesc, 1);
      umma_f16_ts(tmem, a_hi, b_hi, idesc, 1);
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

inline dl_status selftest_ts(float* max_abs_err, float* max_rel_err) {
  std::vector<float> A((size_t)H * H), Bv((size_t)PN * H);
  uint32_t s = 4321u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : A) v = rnd() * 0.2f;
  for (auto& v : Bv) v = rnd() * 3.0f;
  std::vector<__half> Ap(2 * (size_t)H * H), Bp(2 * (size_t)KC * PN * 8);
  for (int c = 0; c < H; ++c)
    for (int k = 0; k < H; ++k) {
      const float v = A[(size_t)c * H + k];
      const __half hi = __float2half_rn(v);
      Ap[(size_t)c * H + k] = hi;
      Ap[(size_t)H * H + (size_t)c * H + k] = __float2half_rn(v - __half2float(hi));
    }
  for (int kc = 0; kc < KC; ++kc)
    for (int r = 0; r < PN; ++r)
      for (int u = 0; u < 8; ++u) {
        const float v = Bv[(size_t)r * H + kc * 8 + u];
        const __half hi = __float2half_rn(v);
        Bp[((size_t)kc * PN + r) * 8 + u] = hi;
        Bp[(size_t)KC * PN * 8 + ((size_t)kc * PN + r) * 8 + u] = __float2half_rn(v - __half2float(hi));
      }
  __half *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  if (cudaMalloc(&dA, Ap.size() * 2) != cudaSuccess || cudaMalloc(&dB, Bp.size() * 2) != cudaSuccess ||
      cudaMalloc(&dD, (size_t)H * PN * 4) != cudaSuccess)
    return DL_ERR_CUDA;
  cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, (size_t)H * PN * 4);
  cudaFuncSetAttribute(k_umma_probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
  k_umma_probe_ts<<<1, 128, P_SMEM_BYTES>>>(reinterpret_cast<const uint32_t*>(dA), dB, dD);
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> D((size_t)H * PN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  if (err != cudaSuccess) {
    fprintf(stderr, "[dl selftest] k_umma_probe_ts failed: %s\n", cudaGetErrorString(err));
    return DL_ERR_CUDA;
  }
  double ma = 0, mref = 0;
  for (int m = 0; m < H; ++m)
    for (int n = 0; n < PN; ++n) {
      double ref = 0;
      for (int k = 0; k < H; ++k) ref += (double)A[(size_t)m * H + k] * (double)Bv[(size_t)n * H + k];
      ma = std::max(ma, std::fabs(ref - (double)D[(size_t)m * PN + n]));
      mref = std::max(mref, std::fabs(ref));
    }
  if (max_abs_err) *max_abs_err = (float)ma;
  if (max_rel_err) *max_rel_err = (float)(ma / std::max(mref, 1e-30));
  fprintf(stderr, "[dl selftest] 3xFP16 UMMA 128x256x128, A in tensor memory: max abs err %.3e (rel to max |ref| %.3e)\n", ma,
          ma / std::max(mref, 1e-30));
  return DL_OK;
}

}  // namespace tc3
}  // namespace dl
