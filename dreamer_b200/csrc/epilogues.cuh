// Epilogue functors for fused_gemm_kernel.
//
// Thread mapping: EPI_PARTS (= 4) threads share one accumulator row (one start state / frame);
// thread `part` owns columns [64 * part, 64 * part + 64) of the tile.  Row-wise reductions
// (LayerNorm statistics, the bucket softmax) are done per part and combined through a small smem
// exchange (Chan's parallel variance / log-sum-exp merge).  Per-tile constants (biases, LN affine,
// buckets) are staged in shared memory by `stage` while the main loop runs and read back as
// warp-wide 16-byte broadcasts.  Outputs go through a shared-memory tile (gemm.cuh: tile_put /
// tile_copy_out) so that global stores are coalesced.
#pragma once

#include "gemm.cuh"

namespace drm {

constexpr int XCHG = 1024;  // float offset of the exchange scratch inside the epilogue smem

// v[j] += c[j] for 32 consecutive constants (16-byte aligned) read as float4 broadcasts
__device__ __forceinline__ void add_const32(float (&v)[32], const float* c) {
  const float4* c4 = reinterpret_cast<const float4*>(c);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = c4[j];
    v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
  }
}
__device__ __forceinline__ void load_const32(float (&v)[32], const float* c) {
  const float4* c4 = reinterpret_cast<const float4*>(c);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = c4[j];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

// ------------------------------------------------------------------------------------------
// plain: out = act(acc + bias) with an optional output-row remap; used by the test hook, the dense
// encoder / decoder layers and the (transposed) convolutions (VariationalAutoEncoder.py:33-42,
// 119-137), whose GEMM rows are pixels.
// ------------------------------------------------------------------------------------------
struct EpiPlain {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0, MIN_CTAS = 1, CLUSTER_N = 1;
  struct Params {
    const float* bias;       // [slots_or_tiles * bn] when bias_per_slot, else [N]; or NULL
    float* out_f32;          // [rows, ld_f32] or NULL
    __nv_bfloat16* out_bf16; // [rows, ld_bf16] or NULL
    long ld_f32, ld_bf16;
    int N;                   // valid output columns (n-tiling: tile y covers columns [y * bn, ..))
    int act;                 // 0 none, 1 SiLU, 2 tanh
    int phases;              // 0: blockIdx.y tiles N;  1: blockIdx.y = transposed-conv phase (all tiles write columns [0, N))
    RowMap rm;               // output row mapping (rm.p2 is overwritten with the phase when phases = 1)
  };
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    const int n0 = p.phases ? 0 : slot * g.bn;
    for (int i = tid; i < 256; i += EPI_THREADS) sm[i] = (p.bias && i < g.bn && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    const int n0 = p.phases ? 0 : slot * g.bn;
    const int ncols = (g.bn + 31) & ~31;   // whole 32-column chunks are staged
    const int pitch = ncols + 4;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int c = part * 64 + 32 * h;
      if (c >= g.bn) break;
      float v[32];
      tmem_ld32(taddr + c, v);
      add_const32(v, sm + c);
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = siluf_(v[j]);
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
      }
      tile_put<32>(tile, pitch, row, c, v);
    }
    epi_bar_sync();
    RowMap rm = p.rm;
    if (p.phases) rm.p2 = slot;
    tile_copy_out(tile, pitch, ncols, min(g.bn, p.N - n0), (m - row), g.M, p.out_f32 ? p.out_f32 + n0 : nullptr, p.ld_f32,
                  p.out_bf16 ? p.out_bf16 + n0 : nullptr, p.ld_bf16, tid, rm);
  }
};

// EpiPlain for narrow outputs (bn <= 64: the conv layers with <= 64 output channels).  Those GEMMs are thousands of tiles with
// 1 .. 8 k-blocks each, so a tile's life is mostly fixed cost (barrier init, TMEM allocation, pipeline fill, epilogue); with a
// third of the shared memory and a quarter of the TMEM columns two CTAs share an SM and hide each other's fixed cost.
struct EpiPlainS : EpiPlain {
  static constexpr int B_ROWS_MAX = 64, STAGES = 3, TMEM_COLS = 64, MIN_CTAS = 2;
};

// ------------------------------------------------------------------------------------------
// Linear + LayerNorm(eps) + SiLU -> bf16 (the hidden layers of every MLP head:
// DynamicsPredictors.py:15-23, 52-60, 85-93; Agent.py:178-185, 219-227;
// VariationalAutoEncoder.py:50-53, 119-122).  One tile holds the whole feature row (<= 256).
//
// bn (the MMA N) is the feature count rounded up to 32 with zero weight rows, and bias / gamma /
// beta are staged as zeros beyond n_valid, so pad columns are exact zeros end to end and the inner
// loops carry no masks: the statistics pass corrects for the pads analytically, and the
// normalise pass yields silu(0 * t + 0) = 0 there.
// ------------------------------------------------------------------------------------------
template <bool HAS_ADD>
struct EpiLnSiluT {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0, MIN_CTAS = 1, CLUSTER_N = 1;
  struct Params {
    const float* bias;   // [slots * bn]
    const float* gamma;  // [slots * bn]
    const float* beta;   // [slots * bn]
    const float* addend; // HAS_ADD: fp32 [M, ld_addend] added before LN (pre-computed partial product)
    long ld_addend;
    __nv_bfloat16* out;  // rows (out_row0 + slot * out_y_stride + m), ld_out columns
    int ld_out, out_row0, out_y_stride;
    int n_valid;         // features (<= bn)
    float eps;
    int cstride;         // bias / gamma / beta entries per slot
  };
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    for (int i = tid; i < 256; i += EPI_THREADS) {
      const bool ok = i < p.n_valid;
      sm[i] = ok ? __ldg(p.bias + slot * p.cstride + i) : 0.f;
      sm[256 + i] = ok ? __ldg(p.gamma + slot * p.cstride + i) : 0.f;
      sm[512 + i] = ok ? __ldg(p.beta + slot * p.cstride + i) : 0.f;
    }
  }
  // x = acc + bias (+ addend) for the 32 columns starting at tile column c (pads are exact zeros)
  static __device__ __forceinline__ void load_x32(const float* sm, uint32_t taddr, const float* add, int c, int nv, float (&v)[32]) {
    tmem_ld32(taddr + c, v);
    add_const32(v, sm + c);
    if constexpr (HAS_ADD) {
      if (add) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < nv) v[j] += __ldg(add + c + j);
      }
    }
  }
  // LayerNorm + SiLU of this thread's 64 columns; every 32-column chunk of the ceil64(n_valid) output columns goes to emit(c, v).
  template <class Emit>
  static __device__ __forceinline__ void normalise(const Params& p, const TileG& g, float* sm, uint32_t taddr, int m, int row,
                                                   int part, int ncols, Emit&& emit) {
    const int nv = p.n_valid;
    const int c0 = part * 64;
    const int cnt = max(0, min(64, nv - c0));          // this part's valid columns
    const int nld = min(64, max(0, g.bn - c0));        // columns this part loads (multiple of 32)
    const float* add = (HAS_ADD && p.addend && m < g.M) ? p.addend + (long)m * p.ld_addend : nullptr;
    // single statistics pass: sums of (x - shift), (x - shift)^2 with shift = the part's first column
    float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (32 * h >= nld) break;
      float v[32];
      load_x32(sm, taddr, add, c0 + 32 * h, nv, v);
      if (h == 0) shift = v[0];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
      }
    }
    const float npad = (float)(nld - cnt);             // loaded pad columns contributed (0 - shift)
    s1 = fmaf(npad, shift, s1);
    s2 = fmaf(-npad * shift, shift, s2);
    const float inv_cnt = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    float* xs = sm + XCHG;
    xs[part * 128 + row] = shift + s1 * inv_cnt;                       // local mean
    xs[512 + part * 128 + row] = fmaxf(s2 - s1 * s1 * inv_cnt, 0.f);   // local M2
    epi_bar_sync();
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < EPI_PARTS; ++q) tot += xs[q * 128 + row] * (float)max(0, min(64, nv - q * 64));
    const float mean = tot / (float)nv;
    float M2 = 0.f;
#pragma unroll
    for (int q = 0; q < EPI_PARTS; ++q) {
      const float d = xs[q * 128 + row] - mean;
      M2 += xs[512 + q * 128 + row] + d * d * (float)max(0, min(64, nv - q * 64));
    }
    const float rstd = rsqrtf(M2 / (float)nv + p.eps);
    const float nmr = -mean * rstd;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int c = c0 + 32 * h;
      if (c >= ncols) break;
      float v[32];
      if (c < g.bn) load_x32(sm, taddr, add, c, nv, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const float4* ga = reinterpret_cast<const float4*>(sm + 256 + c);
      const float4* be = reinterpret_cast<const float4*>(sm + 512 + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 G = ga[j], Bt = be[j];
        v[4 * j] = siluf_(fmaf(fmaf(v[4 * j], rstd, nmr), G.x, Bt.x));
        v[4 * j + 1] = siluf_(fmaf(fmaf(v[4 * j + 1], rstd, nmr), G.y, Bt.y));
        v[4 * j + 2] = siluf_(fmaf(fmaf(v[4 * j + 2], rstd, nmr), G.z, Bt.z));
        v[4 * j + 3] = siluf_(fmaf(fmaf(v[4 * j + 3], rstd, nmr), G.w, Bt.w));
      }
      emit(c, v);
    }
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    const int ncols = min(p.ld_out, (p.n_valid + 63) & ~63);  // the consumer reads ceil64(n_valid) columns: pad with zeros
    const int pitch = ncols + 4;
    normalise(p, g, sm, taddr, m, row, part, ncols, [&](int c, float (&v)[32]) { tile_put<32>(tile, pitch, row, c, v); });
    epi_bar_sync();
    tile_copy_out(tile, pitch, ncols, ncols, (m - row), g.M, nullptr, 0,
                  opnd_at(p.out, (long)(p.out_row0 + slot * p.out_y_stride) * p.ld_out, g.wide), p.ld_out, tid, RowMap{0, 0, 0, 0}, g.wide);
  }
};
using EpiLnSilu = EpiLnSiluT<false>;
using EpiLnSiluAdd = EpiLnSiluT<true>;

// ------------------------------------------------------------------------------------------
// The same Linear + LayerNorm + SiLU stage for SMALL grids (few m-tiles): the tile's 256 columns are
// split over a cluster of four CTAs (gridDim.z), 64 columns = one N = 64 MMA each, so the main loop
// streams a quarter of the weight tile and the epilogue does a quarter of the row work per SM.
// Row statistics: 4 threads x 16 columns inside a CTA (smem merge), then the four CTAs publish
// (mean, M2) of their 64 columns into every peer's shared memory (st.shared::cluster) and meet at
// one barrier.cluster; the 16 values per thread stay in registers across it.
// ------------------------------------------------------------------------------------------
template <bool HAS_ADD>
struct EpiLnSiluN4T {
  static constexpr int B_ROWS_MAX = 64, STAGES = 2, KPS = 4, TMEM_COLS = 64, GRU_U = 0, MIN_CTAS = 1, CLUSTER_N = 4;   // 2 stages x 4 k-blocks x 24 KB in flight; one full / empty handshake per 4 k-blocks
  using Params = typename EpiLnSiluT<HAS_ADD>::Params;
  static constexpr int XST = 2560;   // float offset of the cross-CTA statistics [4 ranks][128 rows][2]
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    const int cr = (int)cluster_ctarank();
    for (int i = tid; i < 64; i += EPI_THREADS) {
      const int col = 64 * cr + i;
      const bool ok = col < p.n_valid;
      sm[i] = ok ? __ldg(p.bias + slot * p.cstride + col) : 0.f;
      sm[256 + i] = ok ? __ldg(p.gamma + slot * p.cstride + col) : 0.f;
      sm[512 + i] = ok ? __ldg(p.beta + slot * p.cstride + col) : 0.f;
    }
  }
  static __device__ __forceinline__ int cnt_of(int nv, int first, int width) { return max(0, min(width, nv - first)); }
  // accumulator -> LayerNorm + SiLU of this thread's 16 columns (v); contains the cluster's ONE statistics barrier
  static __device__ __forceinline__ void compute(const Params& p, const TileG& g, float* sm, uint32_t taddr, int m, int row, int part,
                                                 float (&v)[16]) {
    const int nv = p.n_valid;
    const int cr = (int)cluster_ctarank();
    const int c0 = part * 16;                  // CTA-local column of this thread's 16 values
    const int gc0 = 64 * cr + c0;              // tile column
    const int cnt = cnt_of(nv, gc0, 16);
    tmem_ld16(taddr + c0, v);
    {
      const float4* b4 = reinterpret_cast<const float4*>(sm + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = b4[j];
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
      }
    }
    if constexpr (HAS_ADD) {
      if (p.addend && m < g.M) {
        const float* add = p.addend + (long)m * p.ld_addend + gc0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < cnt) v[j] += __ldg(add + j);
      }
    }
    // pads (columns >= n_valid) are exact zeros: shifted sums over all 16, corrected analytically
    const float shift = v[0];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float d = v[j] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    const float npad = (float)(16 - cnt);
    s1 = fmaf(npad, shift, s1);
    s2 = fmaf(-npad * shift, shift, s2);
    const float inv_cnt = cnt > 0 ? 1.0f / (float)cnt : 0.f;
    float* xs = sm + XCHG;
    xs[part * 128 + row] = shift + s1 * inv_cnt;
    xs[512 + part * 128 + row] = fmaxf(s2 - s1 * s1 * inv_cnt, 0.f);
    epi_bar_sync();
    float* xst = sm + XST;
    if (part == 0) {   // merge the CTA's four parts, publish to every CTA of the cluster
      const int cnt_c = cnt_of(nv, 64 * cr, 64);
      float tot = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_PARTS; ++q) tot += xs[q * 128 + row] * (float)cnt_of(nv, 64 * cr + 16 * q, 16);
      const float mean_c = cnt_c > 0 ? tot / (float)cnt_c : 0.f;
      float M2c = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_PARTS; ++q) {
        const float d = xs[q * 128 + row] - mean_c;
        M2c += xs[512 + q * 128 + row] + d * d * (float)cnt_of(nv, 64 * cr + 16 * q, 16);
      }
#pragma unroll
      for (uint32_t dst = 0; dst < 4; ++dst) st_cluster_v2f32(xst + (cr * 128 + row) * 2, dst, mean_c, M2c);
    }
    __syncwarp();
    cluster_arrive_release();
    cluster_wait_acquire();
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) tot += xst[(q * 128 + row) * 2] * (float)cnt_of(nv, 64 * q, 64);
    const float mean = tot / (float)nv;
    float M2 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float d = xst[(q * 128 + row) * 2] - mean;
      M2 += xst[(q * 128 + row) * 2 + 1] + d * d * (float)cnt_of(nv, 64 * q, 64);
    }
    const float rstd = rsqrtf(M2 / (float)nv + p.eps);
    const float nmr = -mean * rstd;
    {
      const float4* ga = reinterpret_cast<const float4*>(sm + 256 + c0);
      const float4* be = reinterpret_cast<const float4*>(sm + 512 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 G = ga[j], Bt = be[j];
        const float x0 = fmaf(fmaf(v[4 * j], rstd, nmr), G.x, Bt.x), x1 = fmaf(fmaf(v[4 * j + 1], rstd, nmr), G.y, Bt.y);
        const float x2 = fmaf(fmaf(v[4 * j + 2], rstd, nmr), G.z, Bt.z), x3 = fmaf(fmaf(v[4 * j + 3], rstd, nmr), G.w, Bt.w);
        v[4 * j] = siluf_(x0);
        v[4 * j + 1] = siluf_(x1);
        v[4 * j + 2] = siluf_(x2);
        v[4 * j + 3] = siluf_(x3);
      }
    }
  }
  // v -> bf16 columns [64 * rank + 16 * part, + 16) of the output row
  static __device__ __forceinline__ void store(const Params& p, const TileG& g, float* tile, int m, int row, int part, int slot, int tid,
                                               const float (&v)[16]) {
    const int cr = (int)cluster_ctarank();
    const int c0 = part * 16;
    const int ncols_tot = min(p.ld_out, (p.n_valid + 63) & ~63);   // the consumer reads ceil64(n_valid) columns
    const int mine = max(0, min(64, ncols_tot - 64 * cr));  // columns this CTA writes (0 or 64)
    constexpr int pitch = 64 + 4;
    __nv_bfloat16* orow = p.out + (long)(p.out_row0 + slot * p.out_y_stride + m) * p.ld_out + 64 * cr + c0;
    if ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0 && (p.ld_out & 7) == 0) {      // uniform over the CTA: every row is 16-byte aligned
      // this thread's 16 activations are 32 contiguous bytes of its row: two 16-byte stores straight from registers (whole
      // sectors), instead of a shared-memory transpose, a barrier and a coalesced copy-out -- the tile is only 64 columns wide
      if (mine > 0 && m < g.M) {
        *reinterpret_cast<uint4*>(orow) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        *reinterpret_cast<uint4*>(orow + 8) = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
      }
      return;
    }
    tile_put<16>(tile, pitch, row, c0, v);
    epi_bar_sync();
    if (mine > 0)
      tile_copy_out(tile, pitch, 64, mine, (m - row), g.M, nullptr, 0,
                    p.out + (long)(p.out_row0 + slot * p.out_y_stride) * p.ld_out + 64 * cr, p.ld_out, tid);
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    float v[16];
    compute(p, g, sm, taddr, m, row, part, v);
    store(p, g, tile, m, row, part, slot, tid, v);
  }
};
using EpiLnSiluN4 = EpiLnSiluN4T<false>;
using EpiLnSiluAddN4 = EpiLnSiluN4T<true>;

// ------------------------------------------------------------------------------------------
// GRU gates + state update (nn.GRUCell, SequenceModel.py:13-24), U hidden units per tile,
// U / 4 units per thread.  TMEM columns: [r | z | n_x | n_h], each U wide.
// ------------------------------------------------------------------------------------------
template <int U>
struct EpiGru {
  static constexpr int CLUSTER_N = 1;
  // Two CTAs per SM (3 x 28 KB or 2 x 40 KB of stages, 128 / 256 TMEM columns each): one CTA's epilogue and prologue overlap
  // the other's main loop, which keeps the per-SM operand ingress -- the limiter of this stage -- busy.
  static constexpr int B_ROWS_MAX = 3 * U, STAGES = (U == 32 ? 3 : 2), TMEM_COLS = 4 * U, GRU_U = U, MIN_CTAS = 2;
  static constexpr int UP = 8;                          // hidden units per thread per pass
  static constexpr int PASSES = U / (EPI_PARTS * UP);   // 1 (U = 32) or 2 (U = 64)
  struct Params {
    const float* b_ih;   // [3D] reference layout [r; z; n]
    const float* b_hh;   // [3D]
    const float* h_prev; // fp32 [M, ld_hprev]
    float* h_out;        // fp32 [M, ld_hout]
    __nv_bfloat16* s_h;  // bf16 h columns of the next state buffer, [M, ld_s]
    long ld_hprev, ld_hout;
    int ld_s, D;
  };
  // sm: [b_r (U) | b_z (U) | b_in (U) | b_hn (U)] with b_r = b_ir + b_hr, b_z = b_iz + b_hz
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    const int D = p.D;
    for (int i = tid; i < U; i += EPI_THREADS) {
      const int u = slot * U + i;
      const bool ok = u < D;
      sm[i] = ok ? __ldg(p.b_ih + u) + __ldg(p.b_hh + u) : 0.f;
      sm[U + i] = ok ? __ldg(p.b_ih + D + u) + __ldg(p.b_hh + D + u) : 0.f;
      sm[2 * U + i] = ok ? __ldg(p.b_ih + 2 * D + u) : 0.f;
      sm[3 * U + i] = ok ? __ldg(p.b_hh + 2 * D + u) : 0.f;
    }
  }
  // h_prev tile [128 x U] -> smem (pitch U + 4), coalesced: consecutive threads on consecutive 16 bytes of a row
  static constexpr int HP_OFF = 4 * U;   // float offset inside the epilogue scratch, behind the tile constants (U = 32: 18 KB)
  static __device__ __forceinline__ void load_hprev(const Params& p, const TileG& g, int u0, int m0, float* hp_tile, int tid) {
    constexpr int pitch = U + 4;
    const int nvalid = min(U, p.D - u0);
    for (int i = tid; i < BM * (U / 4); i += EPI_THREADS) {
      const int r = i / (U / 4), cc = (i % (U / 4)) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < g.M && cc < nvalid) {
        const float* src = p.h_prev + (long)(m0 + r) * p.ld_hprev + u0 + cc;
        if (cc + 4 <= nvalid && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) x = __ldg(reinterpret_cast<const float4*>(src));
        else { x.x = __ldg(src); if (cc + 1 < nvalid) x.y = __ldg(src + 1); if (cc + 2 < nvalid) x.z = __ldg(src + 2); if (cc + 3 < nvalid) x.w = __ldg(src + 3); }
      }
      *reinterpret_cast<float4*>(hp_tile + r * pitch + cc) = x;
    }
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    const int u0 = slot * U;
    const int D = p.D;
    const int m0 = m - row;   // first row of this tile (the CTA-pair kernel remaps tiles, so not blockIdx.x * BM)
    constexpr int pitch = U + 4;
    const int nvalid = min(U, D - u0);
    // h_prev tile [128 x U]: already in the epilogue scratch (fetched under the main loop, load_hprev below), else staged now
    // in the idle pipeline buffers
    float* hp_tile = g.hp_pre ? sm + HP_OFF : tile + BM * pitch;
    if (!g.hp_pre) {
      load_hprev(p, g, u0, m0, hp_tile, tid);
      epi_bar_sync();           // h_prev tile visible
    }
#pragma unroll 1
    for (int ps = 0; ps < PASSES; ++ps) {
      const int c = (ps * EPI_PARTS + part) * UP;
      float r_[UP], z_[UP], nx[UP], nh[UP];
      tmem_ld8_nowait(taddr + c, r_);
      tmem_ld8_nowait(taddr + U + c, z_);
      tmem_ld8_nowait(taddr + 2 * U + c, nx);
      tmem_ld8_nowait(taddr + 3 * U + c, nh);
      tmem_ld_wait();
      float hn[UP];
#pragma unroll
      for (int j = 0; j < UP; ++j) {
        const float rr = sigmoidf_(r_[j] + sm[c + j]);
        const float zz = sigmoidf_(z_[j] + sm[U + c + j]);
        const float nn = tanhf_(nx[j] + sm[2 * U + c + j] + rr * (nh[j] + sm[3 * U + c + j]));
        hn[j] = (1.0f - zz) * nn + zz * hp_tile[row * pitch + c + j];
      }
      tile_put<UP>(tile, pitch, row, c, hn);
    }
    epi_bar_sync();
    tile_copy_out(tile, pitch, U, nvalid, m0, g.M, p.h_out + u0, p.ld_hout, opnd_at(p.s_h, u0, g.wide), p.ld_s, tid, RowMap{0, 0, 0, 0}, g.wide);
  }
};

// ------------------------------------------------------------------------------------------
// prior / posterior logits -> 32-class categorical (softmax, 1% unimix, inverse-CDF sample from a
// host-supplied uniform, one-hot, straight-through).  DynamicsPredictors.py:31-40;
// VariationalAutoEncoder.py:85-99.  A 256-column tile = 8 latent rows x 32 classes; each thread
// owns 2 of them.
// ------------------------------------------------------------------------------------------
struct EpiCat {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0, MIN_CTAS = 1, CLUSTER_N = 1;
  struct Params {
    const float* bias;     // [R * 32]
    const float* uniforms; // [M, R] for this step, or NULL (logits only)
    float* latent;         // straight-through value, fp32 [M, ld_latent] (R * 32 columns) or NULL
    float* logits;         // fp32 [M, ld_logits] or NULL
    uint8_t* idx;          // [M, ld_idx] or NULL
    __nv_bfloat16* s_z;    // bf16 one-hot into the z columns of a state buffer [M, ld_s] or NULL
    const float* addend;   // optional fp32 [M, ld_addend] added to the logits (unused by the prior)
    long ld_latent, ld_logits, ld_idx, ld_addend;
    int ld_s, R;
    RowMap rm;             // row mapping of the fp32 outputs (latent, logits, idx); s_z always uses the GEMM row
  };
  // A tile is g.bn = 128 or 256 columns = G = 4 or 8 latent rows of 32 classes; thread (row, part) owns groups part, part + 4.
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    const int G = g.bn >> 5;
    for (int i = tid; i < g.bn; i += EPI_THREADS) sm[i] = (slot * g.bn + i < p.R * 32) ? __ldg(p.bias + slot * g.bn + i) : 0.f;
    // this tile's uniforms [128 rows x G latent rows] -> sm[256 ..), coalesced row segments
    for (int i = tid; i < BM * G; i += EPI_THREADS) {
      const int r = i / G, gi = i - r * G;
      const int lrow = slot * G + gi;
      sm[256 + r * 8 + gi] = (p.uniforms && m0 + r < g.M && lrow < p.R) ? __ldg(p.uniforms + (long)(m0 + r) * p.R + lrow) : 0.f;
    }
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    const int m0 = (m - row);
    const int G = g.bn >> 5;
    const int col0 = slot * g.bn;                                  // first logit column of this tile
    const int ncols = max(0, min(g.bn, p.R * 32 - col0));
    const int pitch = g.bn + 4;
    uint8_t* idx_sm = reinterpret_cast<uint8_t*>(sm + 2048);       // [128 rows][8] (sm[256, 1280) holds the uniforms)
    if (p.logits) {   // pass A: logits tile out (training paths only)
#pragma unroll 1
      for (int gi = part; gi < G; gi += EPI_PARTS) {
        float v[32];
        tmem_ld32(taddr + gi * 32, v);
        add_const32(v, sm + gi * 32);
        tile_put<32>(tile, pitch, row, gi * 32, v);
      }
      epi_bar_sync();
      tile_copy_out(tile, pitch, g.bn, ncols, m0, g.M, p.logits + col0, p.ld_logits, nullptr, 0, tid, p.rm);
      epi_bar_sync();
    }
    if (p.uniforms == nullptr) return;
#pragma unroll 1
    for (int gi = part; gi < G; gi += EPI_PARTS) {
      float v[32];
      tmem_ld32(taddr + gi * 32, v);
      add_const32(v, sm + gi * 32);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fexpf_(v[j] - mx);
        s += v[j];
      }
      const float u = sm[256 + row * 8 + gi];
      const float k = 0.99f / s;
      float cdf = 0.f;
      int idx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = fmaf(v[j], k, 0.01f * (1.0f / 32.0f));
        cdf += v[j];
        idx += (cdf <= u) ? 1 : 0;
      }
      idx = idx > 31 ? 31 : idx;
      idx_sm[row * 8 + gi] = (uint8_t)idx;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ((j == idx ? 1.0f : 0.0f) + v[j]) - v[j];   // (onehot + p) - p
      tile_put<32>(tile, pitch, row, gi * 32, v);
    }
    epi_bar_sync();
    if (p.latent) tile_copy_out(tile, pitch, g.bn, ncols, m0, g.M, p.latent + col0, p.ld_latent, nullptr, 0, tid, p.rm);
    // idx [128 x G] bytes and the bf16 one-hot [128 x bn] straight from the staged indices
    const int ngrp = ncols >> 5;
    if (p.idx) {
      for (int i = tid; i < BM * 8; i += EPI_THREADS) {
        const int r = i >> 3, gi = i & 7;
        if (m0 + r < g.M && gi < ngrp) p.idx[map_row(p.rm, m0 + r) * p.ld_idx + slot * G + gi] = idx_sm[i];
      }
    }
    if (p.s_z && g.wide) {   // TF32 mode: the state buffer holds fp32 -- eight floats (two 16-byte stores) per item
      float* sz = reinterpret_cast<float*>(p.s_z);
      for (int i = tid; i < BM * 32; i += EPI_THREADS) {
        const int r = i >> 5, w = i & 31, gi = w >> 2, j0 = (w & 3) * 8;
        if (m0 + r >= g.M || gi >= ngrp) continue;
        const int idx = idx_sm[r * 8 + gi];
        float4* o = reinterpret_cast<float4*>(sz + (long)(m0 + r) * p.ld_s + col0 + w * 8);
        o[0] = make_float4(idx == j0 ? 1.f : 0.f, idx == j0 + 1 ? 1.f : 0.f, idx == j0 + 2 ? 1.f : 0.f, idx == j0 + 3 ? 1.f : 0.f);
        o[1] = make_float4(idx == j0 + 4 ? 1.f : 0.f, idx == j0 + 5 ? 1.f : 0.f, idx == j0 + 6 ? 1.f : 0.f, idx == j0 + 7 ? 1.f : 0.f);
      }
    } else if (p.s_z) {
      for (int i = tid; i < BM * 32; i += EPI_THREADS) {    // one uint4 (8 bf16) per item
        const int r = i >> 5, w = i & 31, gi = w >> 2, j0 = (w & 3) * 8;
        if (m0 + r >= g.M || gi >= ngrp) continue;
        const int idx = idx_sm[r * 8 + gi];
        uint32_t q[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = (idx == j0 + 2 * e ? 0x3F80u : 0u) | (idx == j0 + 2 * e + 1 ? 0x3F800000u : 0u);
        *reinterpret_cast<uint4*>(p.s_z + (long)(m0 + r) * p.ld_s + col0 + w * 8) = make_uint4(q[0], q[1], q[2], q[3]);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// output layers of the [h, z] heads: bucket readout (reward / critic), Bernoulli logit
// (continue), tanh-Normal actor.  DynamicsPredictors.py:64-74, 95-105; Agent.py:191-210, 237-241.
// slot = head id.
// ------------------------------------------------------------------------------------------
enum HeadKind { HEAD_BUCKET = 0, HEAD_SIGMOID = 1, HEAD_ACTOR = 2 };
constexpr int MAX_HEADS = 5;

struct EpiHeads {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0, MIN_CTAS = 1, CLUSTER_N = 1;
  struct Params {
    const float* bias;  // [MAX_HEADS * 256]
    int kind[MAX_HEADS];
    const float* buckets[MAX_HEADS];
    float* value[MAX_HEADS];   // bucket: symexp(E[bucket]); sigmoid: probability;  element (m * ld_value)
    float* logits[MAX_HEADS];  // bucket: [M, ld_logits] fp32; sigmoid: the logit (m * ld_value); or NULL
    long ld_value[MAX_HEADS], ld_logits[MAX_HEADS];
    int NB, A;
    // actor
    const float* normals;  // [M, ld_normals] or NULL (then no action is produced)
    float *mu, *sigma, *action;  // [M, ld_act]
    long ld_act, ld_normals;
    __nv_bfloat16* s_a;    // bf16 action into the a columns of a state buffer [M, ld_s] or NULL
    int ld_s;
    RowMap rm;             // row mapping of the bucket / sigmoid outputs
  };
  // sm: [bias (256) | buckets (256)]
  static __device__ __forceinline__ void stage(const Params& p, const TileG& g, int slot, float* sm, int tid, int m0) {
    const bool bucket = p.kind[slot] == HEAD_BUCKET;
    for (int i = tid; i < 256; i += EPI_THREADS) {
      sm[i] = __ldg(p.bias + slot * 256 + i);
      sm[256 + i] = (bucket && i < p.NB) ? __ldg(p.buckets[slot] + i) : 0.f;
    }
  }
  static __device__ __forceinline__ void run(const Params& p, const TileG& g, float* sm, float* tile, uint32_t taddr, int m,
                                             int row, int part, int slot, int tid) {
    const bool live = m < g.M;
    const int kind = p.kind[slot];
    const int m0 = (m - row);
    if (kind == HEAD_BUCKET) {
      const int NB = p.NB;
      const int c0 = part * 64;
      const int cnt = max(0, min(64, NB - c0));
      constexpr int pitch = 256 + 4;
      const bool want_logits = p.logits[slot] != nullptr;
      float mx = -INFINITY, s = 0.f, ws = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {          // online (max, sum, weighted sum) over two 32-column halves
        const int c = c0 + 32 * h;
        const int n = cnt - 32 * h;
        if (n <= 0) break;
        float v[32], bk[32];
        tmem_ld32(taddr + c, v);
        add_const32(v, sm + c);
        load_const32(bk, sm + 256 + c);
        if (want_logits) tile_put<32>(tile, pitch, row, c, v);
        float hm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) hm = fmaxf(hm, j < n ? v[j] : -INFINITY);
        const float nm = fmaxf(mx, hm);
        const float f = mx == -INFINITY ? 0.f : fexpf_(mx - nm);
        s *= f; ws *= f; mx = nm;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = j < n ? fexpf_(v[j] - mx) : 0.f;
          s += e;
          ws = fmaf(e, bk[j], ws);
        }
      }
      float* xs = sm + XCHG;
      xs[part * 128 + row] = mx;
      xs[512 + part * 128 + row] = s;
      xs[1024 + part * 128 + row] = ws;
      epi_bar_sync();
      if (want_logits) tile_copy_out(tile, pitch, 256, NB, m0, g.M, p.logits[slot], p.ld_logits[slot], nullptr, 0, tid, p.rm);
      if (part == 0 && live && p.value[slot]) {
        float M = -INFINITY;
#pragma unroll
        for (int q = 0; q < EPI_PARTS; ++q) M = fmaxf(M, xs[q * 128 + row]);
        float S = 0.f, WS = 0.f;
#pragma unroll
        for (int q = 0; q < EPI_PARTS; ++q) {
          const float mq = xs[q * 128 + row];
          const float f = mq == -INFINITY ? 0.f : fexpf_(mq - M);
          S = fmaf(xs[512 + q * 128 + row], f, S);
          WS = fmaf(xs[1024 + q * 128 + row], f, WS);
        }
        p.value[slot][map_row(p.rm, m) * p.ld_value[slot]] = symexpf_(WS / S);
      }
    } else if (kind == HEAD_SIGMOID) {
      if (part != 0) return;
      float v[16];
      tmem_ld16(taddr, v);
      const float x = v[0] + sm[0];
      if (live) {
        if (p.value[slot]) p.value[slot][map_row(p.rm, m) * p.ld_value[slot]] = 1.0f / (1.0f + expf(-x));
        if (p.logits[slot]) p.logits[slot][map_row(p.rm, m) * p.ld_value[slot]] = x;
      }
    } else {
      if (part != 0) return;
      float v[32];
      tmem_ld32(taddr, v);
      if (live) {
        const int A = p.A;
        float act[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          act[j] = 0.f;
          if (j < A) {
            const float mu = v[j] + sm[j];
            float ls = v[16 + j] + sm[16 + j];  // log-sigma rows are packed at 16..16+A
            ls = fminf(fmaxf(ls, -5.0f), 2.0f);
            const float sg = softplusf_(ls) + 1e-3f;
            if (p.mu) p.mu[(long)m * p.ld_act + j] = mu;
            if (p.sigma) p.sigma[(long)m * p.ld_act + j] = sg;
            if (p.normals) {
              act[j] = tanhf(mu + sg * __ldg(p.normals + (long)m * p.ld_normals + j));
              if (p.action) p.action[(long)m * p.ld_act + j] = act[j];
            }
          }
        }
        if (p.s_a && p.normals) {
          if (g.wide) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.s_a) + (long)m * p.ld_s);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(tf32_rn(act[4 * j]), tf32_rn(act[4 * j + 1]), tf32_rn(act[4 * j + 2]), tf32_rn(act[4 * j + 3]));
          } else {
            store_bf16_row<16>(p.s_a + (long)m * p.ld_s, act, 16);
          }
        }
      }
    }
  }
};

}  // namespace drm
