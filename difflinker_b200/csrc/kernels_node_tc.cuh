// tcgen05 node kernel: GCL.node_model (egnn.py:62-72) + node mask + the first-layer projections of the next edge
// MLP(s), as a chain of 128-row UMMA GEMMs on one 128-node tile per CTA.
//
//   G1: hid = silu([h, agg] W3^T + b3)        K = 256  (two 128x128 weight blocks)
//   G2: h'  = (h + hid W4^T + b4) * node_mask  K = 128
//   P : A = h' W1a^T + b1 ; B = h' W1b^T       K = 128 each, for one or two consumers
//
// Natural orientation D[node, channel]: TMEM lane = node row, so thread r owns row r end to end: it loads the
// row, scales/splits it into the fp16 hi/lo operand tile ([kc][row][8 halves], 16-byte vector stores), and in every
// epilogue reads its own accumulator row, applies bias/SiLU/residual/mask and writes the next operand row.
// Same 3xFP16 numerics and exact power-of-two range scaling as the edge kernel (kernels_tc.cuh): every operand
// row is scaled so |x| < 2^14 using the row's own maximum; the descale is a per-thread scalar.
// Weights stream through a 3-stage ring of K=64 half-blocks (2 x 16 KB cp.async.bulk each) fed by a dedicated
// loader thread; the MMA issuer thread only waits on the ring's full barriers.
#pragma once
#include "kernels_tc.cuh"

namespace dl {
namespace tcn {

using namespace dl::tc;

constexpr int TM = 128;                        // nodes per tile = UMMA M
constexpr int X_LBO = TM * 16;                 // 2048 B
constexpr int X_BYTES = KC * X_LBO;            // 32 KB per fp16 copy
constexpr int HALF_BYTES = 8 * W_LBO;          // 16 KB: kc 0..7 of one fp16 copy of a 128x128 block
constexpr int STAGE_BYTES = 2 * HALF_BYTES;    // hi | lo
constexpr int BLOCK_BYTES = 2 * W_BYTES;       // one packed 128x128 block: [hi 32 KB | lo 32 KB]

constexpr int N_OFF_XA = 0;                    // hi | lo
constexpr int N_OFF_XB = N_OFF_XA + 2 * X_BYTES;
constexpr int N_OFF_WS = N_OFF_XB + 2 * X_BYTES;        // N_RING stages
constexpr int N_RING = 3;                      // ring depth: 96 KB of weights in flight per CTA
constexpr int N_OFF_BAR = N_OFF_WS + N_RING * STAGE_BYTES;   // full[3], empty[3], acc[4]; tmem slot
constexpr int N_SMEM_BYTES = N_OFF_BAR + 128 + 1024;

struct NodeTcArgs {
  float* h;              // (n,128) in/out
  const float* agg;      // (n,128)
  const float* nm;       // (n)
  const __half* w3;      // 2 packed blocks (h part, agg part), common scale
  const __half* w4;      // 1 packed block
  const float *b3, *b4;
  float w3_descale, w4_descale;
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

// write 16 consecutive channels (two kc chunks) of this thread's row into an operand tile
__device__ __forceinline__ void store_row16(uint8_t* xhi, uint8_t* xlo, int row, int c0, const float (&v)[16]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint4 hi, lo;
    split2(v[half * 8 + 0], v[half * 8 + 1], hi.x, lo.x); split2(v[half * 8 + 2], v[half * 8 + 3], hi.y, lo.y);
    split2(v[half * 8 + 4], v[half * 8 + 5], hi.z, lo.z); split2(v[half * 8 + 6], v[half * 8 + 7], hi.w, lo.w);
    const int kc = (c0 >> 3) + half;
    *reinterpret_cast<uint4*>(xhi + kc * X_LBO + row * 16) = hi;
    *reinterpret_cast<uint4*>(xlo + kc * X_LBO + row * 16) = lo;
  }
}

constexpr int NODE_TC_THREADS = 160;          // warps 0-3: row owners (thread r = node row r = TMEM lane r); warp 4: weight loader

__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(NODE_TC_THREADS, 1) k_node_tc(int n_total, NodeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer derived from the __shared__ array (no integer round trip) so accesses compile to LDS/STS
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x * TM + tid;
  const bool live = g < n_total;

  const uint32_t bar_full = sbase + N_OFF_BAR, bar_empty = bar_full + 8 * N_RING, bar_acc = bar_empty + 8 * N_RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + N_OFF_BAR + 8 * (2 * N_RING + 4));
  uint8_t* xa_hi = sm + N_OFF_XA; uint8_t* xa_lo = xa_hi + X_BYTES;
  uint8_t* xb_hi = sm + N_OFF_XB; uint8_t* xb_lo = xb_hi + X_BYTES;

  if (tid == 0) {
    for (int i = 0; i < N_RING; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 4; ++i) mbar_init(bar_acc + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);

  const int n_blocks = 3 + 2 * a.n_proj;          // 128x128 weight blocks consumed by this tile, in order
  const int n_half = 2 * n_blocks;

  // ---- weight loader: a dedicated warp streams the half-blocks through the 2-stage ring; it never joins the
  //      workers' barriers (the ring's empty barriers are released by MMA completion, which needs the workers) -----
  if (warp == 4) {
    if (tid == 128) {
    for (int i = 0; i < n_half; ++i) {
      const int s = i % N_RING, blk = i >> 1, hf = i & 1;
      if (i >= N_RING) mbar_wait(bar_empty + 8 * s, ((i - N_RING) / N_RING) & 1);
      const __half* base = blk < 2 ? a.w3 + (size_t)blk * (BLOCK_BYTES / 2)
                           : blk == 2 ? a.w4
                                      : a.pw[(blk - 3) >> 1] + (size_t)((blk - 3) & 1) * (BLOCK_BYTES / 2);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(base);
      const uint32_t dst = sbase + N_OFF_WS + s * STAGE_BYTES;
      mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
      bulk_g2s(dst, src + hf * HALF_BYTES, HALF_BYTES, bar_full + 8 * s);                      // hi, kc 8hf..8hf+7
      bulk_g2s(dst + HALF_BYTES, src + W_BYTES + hf * HALF_BYTES, HALF_BYTES, bar_full + 8 * s);  // lo
    }
    }
    return;
  }

  // ---- operand rows: h -> XA, agg -> XB, common row scale ----------------------------------------------------------
  float s1 = 1.f;
  {
    const float* hr = a.h + (size_t)g * H;
    const float* ar = a.agg + (size_t)g * H;
    float mx = 0.f;
    if (live) {
#pragma unroll 4
      for (int c = 0; c < H; c += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hr + c);
        const float4 av = __ldg(reinterpret_cast<const float4*>(ar + c));
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(hv.x), fabsf(hv.y)), fmaxf(fabsf(hv.z), fabsf(hv.w))));
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(av.x), fabsf(av.y)), fmaxf(fabsf(av.z), fabsf(av.w))));
      }
    }
    s1 = pow2_scale_for(mx);
#pragma unroll 1
    for (int c0 = 0; c0 < H; c0 += 16) {
      float hv[16], av[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 x = make_float4(0, 0, 0, 0), y = x;
        if (live) {
          x = *reinterpret_cast<const float4*>(hr + c0 + 4 * q);
          y = __ldg(reinterpret_cast<const float4*>(ar + c0 + 4 * q));
        }
        hv[4 * q] = x.x * s1; hv[4 * q + 1] = x.y * s1; hv[4 * q + 2] = x.z * s1; hv[4 * q + 3] = x.w * s1;
        av[4 * q] = y.x * s1; av[4 * q + 1] = y.y * s1; av[4 * q + 2] = y.z * s1; av[4 * q + 3] = y.w * s1;
      }
      store_row16(xa_hi, xa_lo, tid, c0, hv);
      store_row16(xb_hi, xb_lo, tid, c0, av);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  workers_sync();

  // ---- MMA issue helper (thread 0): consume `nh` half-blocks of the ring starting at ring index i0 ------------------
  const uint32_t idesc = umma_idesc(128, 128);
  auto issue = [&](int i0, int nh, const uint8_t* const* xhi_of, uint32_t acc_col, int acc_bar) {
    // xhi_of[j]: operand hi base for half-block j of this GEMM (lo = hi + X_BYTES); kc offset = 8*(j&1)
    tc_fence_after();
    for (int j = 0; j < nh; ++j) {
      const int i = i0 + j, s = i % N_RING;
      mbar_wait(bar_full + 8 * s, (i / N_RING) & 1);
      tc_fence_after();
      const uint32_t xh = smem_u32(xhi_of[j]) + (8 * (j & 1)) * X_LBO, xl = xh + X_BYTES;
      const uint32_t wh = sbase + N_OFF_WS + s * STAGE_BYTES, wl = wh + HALF_BYTES;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a_hi = umma_desc(xh + ks * 2 * X_LBO, X_LBO, SBO), a_lo = umma_desc(xl + ks * 2 * X_LBO, X_LBO, SBO);
        const uint64_t b_hi = umma_desc(wh + ks * 2 * W_LBO, W_LBO, SBO), b_lo = umma_desc(wl + ks * 2 * W_LBO, W_LBO, SBO);
        umma_f16(tmem + acc_col, a_lo, b_hi, idesc, (j | ks) != 0);
        umma_f16(tmem + acc_col, a_hi, b_lo, idesc, 1);
        umma_f16(tmem + acc_col, a_hi, b_hi, idesc, 1);
      }
      umma_commit(bar_empty + 8 * s);            // ring slot reusable once these MMAs have read it
    }
    umma_commit(bar_acc + 8 * acc_bar);
  };

  // ---- G1: [h, agg] W3^T --------------------------------------------------------------------------------------------
  if (tid == 0) {
    const uint8_t* xs[4] = {xa_hi, xa_hi, xb_hi, xb_hi};
    issue(0, 4, xs, 0, 0);
  }
  mbar_wait(bar_acc, 0);
  tc_fence_after();
  float s2;
  {
    const float ds = a.w3_descale / s1;
    float mx = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < H; c0 += 16) {           // pass 1: bound |silu(v)| <= |v|
      uint32_t r[16];
      TMEM_LD_X16(trow + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 16; ++u) mx = fmaxf(mx, fabsf(fmaf(__uint_as_float(r[u]), ds, __ldg(a.b3 + c0 + u))));
    }
    s2 = pow2_scale_for(mx);
#pragma unroll 1
    for (int c0 = 0; c0 < H; c0 += 16) {           // pass 2: hid = silu(v) -> XB (agg no longer needed: G1 is complete)
      uint32_t r[16];
      TMEM_LD_X16(trow + c0, r);
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = silu_f(fmaf(__uint_as_float(r[u]), ds, __ldg(a.b3 + c0 + u))) * s2;
      store_row16(xb_hi, xb_lo, tid, c0, v);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  workers_sync();

  // ---- G2: hid W4^T, residual, mask -----------------------------------------------------------------------------------
  if (tid == 0) {
    const uint8_t* xs[2] = {xb_hi, xb_hi};
    issue(4, 2, xs, 128, 1);
  }
  mbar_wait(bar_acc + 8, 0);
  tc_fence_after();
  float s3;
  {
    const float ds = a.w4_descale / s2;
    const float m = live ? a.nm[g] : 0.f;
    float* hr = a.h + (size_t)g * H;
    float mx = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < H; c0 += 16) {
      uint32_t r[16];
      TMEM_LD_X16(trow + 128 + c0, r);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 hv = *reinterpret_cast<const float4*>(hr + c0 + 4 * q);
          float4 o;
          o.x = (hv.x + fmaf(__uint_as_float(r[4 * q + 0]), ds, __ldg(a.b4 + c0 + 4 * q + 0))) * m;   // egnn.py:71,78-79
          o.y = (hv.y + fmaf(__uint_as_float(r[4 * q + 1]), ds, __ldg(a.b4 + c0 + 4 * q + 1))) * m;
          o.z = (hv.z + fmaf(__uint_as_float(r[4 * q + 2]), ds, __ldg(a.b4 + c0 + 4 * q + 2))) * m;
          o.w = (hv.w + fmaf(__uint_as_float(r[4 * q + 3]), ds, __ldg(a.b4 + c0 + 4 * q + 3))) * m;
          *reinterpret_cast<float4*>(hr + c0 + 4 * q) = o;
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
        }
      }
    }
    s3 = pow2_scale_for(mx);
#pragma unroll 1
    for (int c0 = 0; c0 < H; c0 += 16) {           // h' (own writes, program order) -> XA
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 x = make_float4(0, 0, 0, 0);
        if (live) x = *reinterpret_cast<const float4*>(hr + c0 + 4 * q);
        v[4 * q] = x.x * s3; v[4 * q + 1] = x.y * s3; v[4 * q + 2] = x.z * s3; v[4 * q + 3] = x.w * s3;
      }
      store_row16(xa_hi, xa_lo, tid, c0, v);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  workers_sync();

  // ---- projections: A -> accumulator 2, B -> accumulator 3 (second consumer reuses 0 / 1) --------------------------------
  if (tid == 0) {
    const uint8_t* xs[2] = {xa_hi, xa_hi};
    issue(6, 2, xs, 256, 2);
    issue(8, 2, xs, 384, 3);
    if (a.n_proj > 1) {
      issue(10, 2, xs, 0, 0);
      issue(12, 2, xs, 128, 1);
    }
  }
  for (int p = 0; p < a.n_proj; ++p) {
    const float ds = a.p_descale[p] / s3;
    float* ab = a.AB[p] + (size_t)g * 2 * H;
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
      const int accn = p == 0 ? 2 + part : part;
      mbar_wait(bar_acc + 8 * accn, p == 0 ? 0 : 1);
      tc_fence_after();
      float mx = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < H; c0 += 16) {
        uint32_t r[16];
        TMEM_LD_X16(trow + accn * 128 + c0, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o;
            if (part == 0) {
              o.x = fmaf(__uint_as_float(r[4 * q + 0]), ds, __ldg(a.pb1[p] + c0 + 4 * q + 0));
              o.y = fmaf(__uint_as_float(r[4 * q + 1]), ds, __ldg(a.pb1[p] + c0 + 4 * q + 1));
              o.z = fmaf(__uint_as_float(r[4 * q + 2]), ds, __ldg(a.pb1[p] + c0 + 4 * q + 2));
              o.w = fmaf(__uint_as_float(r[4 * q + 3]), ds, __ldg(a.pb1[p] + c0 + 4 * q + 3));
            } else {
              o.x = __uint_as_float(r[4 * q + 0]) * ds; o.y = __uint_as_float(r[4 * q + 1]) * ds;
              o.z = __uint_as_float(r[4 * q + 2]) * ds; o.w = __uint_as_float(r[4 * q + 3]) * ds;
            }
            *reinterpret_cast<float4*>(ab + part * H + c0 + 4 * q) = o;
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
          }
        }
      }
      if (live) a.ABmax[p][(size_t)g * 2 + part] = mx;
    }
  }
  tc_fence_before();
  workers_sync();
  if (warp == 0) tmem_dealloc(tmem, 512);
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

inline dl_status configure_node() {
  if (cudaFuncSetAttribute(k_node_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, N_SMEM_BYTES) != cudaSuccess)
    return DL_ERR_CUDA;
  return DL_OK;
}

}  // namespace tcn
}  // namespace dl
