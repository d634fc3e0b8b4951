// Implicit-GEMM convolutions: the A operand of every k-block is fetched by an im2col-mode TMA load straight from the NHWC bf16
// activation of the previous layer -- no patch matrix is written to or read back from HBM.  Included from rssm.cu.
//
//   Conv2d(k4, s2, p1)            (VariationalAutoEncoder.py:33-42):   16 taps, traversal stride 2, filter positions start at -1
//   ConvTranspose2d(k4, s2, p1)   (VariationalAutoEncoder.py:128-137): four sub-pixel phases, each a 2 x 2-tap stride-1 convolution whose
//                                 output pixel (q, r) lands at (2q + py, 2r + px); one im2col tensor map per phase (the phases differ
//                                 in where the 2 x 2 window sits: rows {q - 1, q} for py = 0, {q, q + 1} for py = 1)
//
// A k-block is ONE filter tap x `chunk` input channels (chunk = min(C_in padded, 64): 32-, 64- or 128-byte swizzled rows), i.e. the
// GEMM K order is (tap, channel) exactly as the packed weights already have it.  The TMA unit walks the 128 output pixels of a tile
// across image rows and frames and zero-fills the padding ring, so tile boundaries need no special cases.
//
// Structure = conv_persist.cuh: one persistent CTA per SM walks the tiles, the TMA ring runs across tile boundaries, two TMEM
// accumulators alternate so tile i + 1 accumulates while the 16 epilogue warps drain tile i (bias + SiLU, 64 columns at a time
// through a transposed shared-memory tile, coalesced row-remapped stores).  Output widths up to 256 channels.
#pragma once

namespace drm {

constexpr int CI_RING_BYTES = 160 * 1024;
constexpr int CI_TILE_OFF = CI_RING_BYTES;
constexpr int CI_TILE_BYTES = BM * (64 + 4) * 4;            // epilogue transpose tile, pitch 68 floats
constexpr int CI_BAR_OFF = CI_TILE_OFF + CI_TILE_BYTES;
constexpr int CI_CONST_OFF = CI_BAR_OFF + 256;              // 256 bias values
constexpr int CI_SMEM = CI_CONST_OFF + 1024 + 1024;
constexpr int CI_MAX_STAGES = 8;

struct ConvImplicit {
  CUtensorMap tmA[4];          // im2col maps, one per phase
  CUtensorMap tmB;             // weights [phases * bn, K], box {chunk, bn}
  int M, n_mtiles, phases;     // output pixels per phase, 128-pixel tiles per phase, 1 or 4 sub-pixel phases
  int Wo, Ho;                  // filter positions per image row / column (per phase)
  int stride;                  // traversal stride
  int lower_w[4], lower_h[4];  // base pixel of filter position 0 per phase
  int n_base;                  // first frame of this chunk inside the mapped activation
  int ntap, cpt, chunk;        // taps, k-blocks per tap, channels per k-block
  unsigned char off_w[4][16], off_h[4][16];   // im2col offsets of tap t in phase p
  int bn;                      // MMA N (padded output channels, multiple of 16, <= 256)
  int stage_bytes, n_stages, tmem_cols;
  int kps, sub_bytes;          // k-blocks per pipeline stage (a full / empty handshake costs ~0.3 us whatever it carries: DESIGN.md 4a) and bytes of one
  const float* bias;
  __nv_bfloat16* out;          // [rows, ld] bf16
  long ld;
  int n_valid, act;            // output columns to store; 0 none, 1 SiLU
  RowMap rm;                   // output row mapping (p2 = phase is filled in per tile when phases > 1)
};

__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const CUtensorMap* tm, int c, int w, int h, int n, uint16_t ow, uint16_t oh,
                                                   uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
      : "memory");
}
// K-major shared-memory descriptor for rows of 32 / 64 / 128 bytes (swizzle span = row length): 8-row groups are 8 x row bytes apart
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_implicit_kernel(const __grid_constant__ ConvImplicit c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CI_BAR_OFF);
  uint64_t* empty = full + CI_MAX_STAGES;
  uint64_t* tfull = empty + CI_MAX_STAGES;   // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* tile = reinterpret_cast<float*>(smem + CI_TILE_OFF);
  float* bias_s = reinterpret_cast<float*>(smem + CI_CONST_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = c.n_mtiles * c.phases;
  const int nk = c.ntap * c.cpt;
  const int NS = c.n_stages;
  const int a_bytes = BM * c.chunk * 2;

  if (threadIdx.x == 0) {
    for (int p = 0; p < c.phases; ++p) tma_prefetch_desc(&c.tmA[p]);
    tma_prefetch_desc(&c.tmB);
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EPI_THREADS / 32); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)c.tmem_cols);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)a_bytes + (uint32_t)c.bn * c.chunk * 2;
      const int ppf = c.Wo * c.Ho;   // filter positions per frame
      int it = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int ph = t / c.n_mtiles, mt = t - ph * c.n_mtiles;
        const int m0 = mt * BM;
        const int n = m0 / ppf, rem = m0 - n * ppf, oy = rem / c.Wo, ox = rem - oy * c.Wo;
        const int cw = ox * c.stride + c.lower_w[ph], chh = oy * c.stride + c.lower_h[ph], cn = n + c.n_base;
        const int b_row = ph * c.bn;
        int tap = 0, cc = 0;
        for (int kb = 0; kb < nk; ++it) {
          const int s = it % NS;
          mbar_wait(&empty[s], ((it / NS) & 1) ^ 1u);
          const int n_sub = min(c.kps, nk - kb);
          mbar_expect_tx(&full[s], (uint32_t)n_sub * tx);
          for (int u = 0; u < n_sub; ++u, ++kb) {
            uint8_t* sa = smem + s * c.stage_bytes + u * c.sub_bytes;
            tma_load_im2col_4d(sa, &c.tmA[ph], cc * c.chunk, cw, chh, cn, c.off_w[ph][tap], c.off_h[ph][tap], &full[s]);
            tma_load_2d(sa + a_bytes, &c.tmB, kb * c.chunk, b_row, &full[s]);
            if (++cc == c.cpt) { cc = 0; ++tap; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(c.bn);
      const int row_bytes = c.chunk * 2, ksteps = c.chunk / 16;
      int it = 0, i = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(&tempty[buf], ((i >> 1) & 1) ^ 1u);      // the epilogue has drained this accumulator's previous tile
        tc_fence_after();
        for (int kb = 0; kb < nk; ++it) {
          const int s = it % NS;
          mbar_wait(&full[s], (it / NS) & 1);
          tc_fence_after();
          const int n_sub = min(c.kps, nk - kb);
          for (int u = 0; u < n_sub; ++u, ++kb) {
            const uint32_t a_addr = smem_u32(smem + s * c.stage_bytes + u * c.sub_bytes);
            const uint64_t adesc = umma_desc_kmajor(a_addr, row_bytes), bdesc = umma_desc_kmajor(a_addr + a_bytes, row_bytes);
            for (int k = 0; k < ksteps; ++k) umma_bf16(tmem + (uint32_t)(buf * c.bn), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    for (int i = tid; i < 256; i += EPI_THREADS) bias_s[i] = (c.bias && i < c.n_valid) ? __ldg(c.bias + i) : 0.f;
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int pw = c.bn < 64 ? c.bn : 64;            // columns per pass
    const int cpp = pw >> 2;                          // columns per thread and pass: 4, 8 or 16
    const int n_pass = (c.bn + 63) / 64;
    const int pitch = pw + 4;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    int i = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++i) {
      const int buf = i & 1;
      const int ph = t / c.n_mtiles, mt = t - ph * c.n_mtiles;
      mbar_wait(&tfull[buf], (i >> 1) & 1);
      tc_fence_after();
      RowMap rm = c.rm;
      if (c.phases > 1) rm.p2 = ph;
      for (int ps = 0; ps < n_pass; ++ps) {
        const int c0 = ps * 64 + part * cpp;          // this thread's first accumulator column of the pass
        float v[16];
        if (cpp == 16) {
          tmem_ld16(tlane + (uint32_t)(buf * c.bn + c0), v);
        } else {
          float w[8];
          tmem_ld8_nowait(tlane + (uint32_t)(buf * c.bn + (c0 & ~7)), w);   // (cpp = 4: two threads share an 8-column load)
          tmem_ld_wait();
          const int sh = cpp == 4 ? (c0 & 4) : 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = w[(j + sh) & 7];
        }
        if (ps == n_pass - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);    // the accumulator may be overwritten: its values are in registers
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < cpp) {
            const float x = v[j] + bias_s[c0 + j];
            v[j] = c.act == 1 ? siluf_(x) : x;
          }
        }
        float4* dst = reinterpret_cast<float4*>(tile + row * pitch + part * cpp);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        if (cpp >= 8) dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        if (cpp == 16) {
          dst[2] = make_float4(v[8], v[9], v[10], v[11]);
          dst[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
        epi_bar_sync();
        const int nv = min(pw, c.n_valid - ps * 64);
        if (nv > 0) tile_copy_out(tile, pitch, pw, nv, mt * BM, c.M, nullptr, 0, c.out + ps * 64, c.ld, tid, rm);
        epi_bar_sync();                                // the tile is free for the next pass's writes
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, (uint32_t)c.tmem_cols);
}

}  // namespace drm
