// Persistent GEMM for the NARROW conv layers (<= 64 output channels, a few k-blocks, thousands of 128-row tiles).
// Included from rssm.cu (before vae.cuh).
//
// One tile per CTA makes those layers pay a CTA's whole fixed cost (launch slot, barrier init, TMEM allocation, pipeline fill and
// drain, epilogue) per 128 rows: ncu showed 3 - 8 % tensor-pipe utilisation and ~1.9 TB/s of DRAM reads on the widest decoder
// layers (profiles/prof_r1_conv_c3_summary.txt).  Here a CTA is resident for the whole layer and walks tiles blockIdx.x,
// blockIdx.x + gridDim.x, ...:
//   * the TMA producer runs across tile boundaries, so the ring never drains;
//   * the MMA thread alternates between two TMEM accumulators (2 x 64 columns): tile i + 1 accumulates while
//   * the 16 epilogue warps drain tile i (4 threads per row, 8 or 16 columns each): bias + SiLU, transposed through a dedicated
//     shared-memory tile, coalesced (row-remapped) stores.
// Barriers: full / empty per ring stage, tfull / tempty per accumulator.  Same operands, same k order, same epilogue arithmetic
// as fused_gemm_kernel<EpiPlainS>: results are bit-identical.
#pragma once

namespace drm {

constexpr int CP_STAGES = 4;
constexpr int CP_STAGE = A_STAGE_BYTES + 64 * BK * 2;      // A 16 KB + B up to 64 rows = 24 KB
constexpr int CP_TILE_OFF = CP_STAGES * CP_STAGE;           // 96 KB ring
constexpr int CP_TILE_BYTES = BM * (64 + 4) * 4;            // epilogue transpose tile, pitch 68 floats
constexpr int CP_BAR_OFF = CP_TILE_OFF + CP_TILE_BYTES;
constexpr int CP_CONST_OFF = CP_BAR_OFF + 256;              // 64 bias values
constexpr int CP_SMEM = CP_CONST_OFF + 256 + 1024;

struct ConvPersist {
  CUtensorMap tmA, tmB;        // patches [rows, K] box {64, 128}; weights [phases * bn, K] box {64, bn}
  int M, n_mtiles, phases;     // rows per phase, 128-row tiles per phase, 1 or 4 sub-pixel phases
  int a_phase_rows;            // A rows between consecutive phases
  int bn, nk;                  // MMA N (32 or 64), k-blocks
  const float* bias;           // [bn] (zero beyond n_valid)
  __nv_bfloat16* out;          // [rows, ld] bf16
  long ld;
  int n_valid, act;            // valid output channels; 0 none, 1 SiLU
  RowMap rm;                   // output row mapping (p2 = phase is filled in per tile when phases > 1)
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) conv_persist_kernel(const __grid_constant__ ConvPersist c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CP_BAR_OFF);
  uint64_t* empty = full + CP_STAGES;
  uint64_t* tfull = empty + CP_STAGES;     // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* tile = reinterpret_cast<float*>(smem + CP_TILE_OFF);
  float* bias_s = reinterpret_cast<float*>(smem + CP_CONST_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = c.n_mtiles * c.phases;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&c.tmA);
    tma_prefetch_desc(&c.tmB);
    for (int s = 0; s < CP_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EPI_THREADS / 32); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)A_STAGE_BYTES + (uint32_t)c.bn * BK * 2;
      int it = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x) {
        const int ph = t / c.n_mtiles, mt = t - ph * c.n_mtiles;
        const int a_row = ph * c.a_phase_rows + mt * BM, b_row = ph * c.bn;
        for (int kb = 0; kb < c.nk; ++kb, ++it) {
          const int s = it % CP_STAGES;
          mbar_wait(&empty[s], ((it / CP_STAGES) & 1) ^ 1u);
          uint8_t* sa = smem + s * CP_STAGE;
          mbar_expect_tx(&full[s], tx);
          tma_load_2d(sa, &c.tmA, kb * BK, a_row, &full[s]);
          tma_load_2d(sa + A_STAGE_BYTES, &c.tmB, kb * BK, b_row, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(c.bn);
      int it = 0, i = 0;
      for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(&tempty[buf], ((i >> 1) & 1) ^ 1u);      // the epilogue has drained this accumulator's previous tile
        tc_fence_after();
        for (int kb = 0; kb < c.nk; ++kb, ++it) {
          const int s = it % CP_STAGES;
          mbar_wait(&full[s], (it / CP_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * CP_STAGE);
          const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem + (uint32_t)(buf * 64), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[buf]);
      }
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    for (int i = tid; i < 64; i += EPI_THREADS) bias_s[i] = (c.bias && i < c.n_valid) ? __ldg(c.bias + i) : 0.f;
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int cpp = c.bn >> 2;                       // columns per part: 8 or 16
    const int c0 = part * cpp;
    const int pitch = c.bn + 4;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    int i = 0;
    for (int t = (int)blockIdx.x; t < n_tiles; t += (int)gridDim.x, ++i) {
      const int buf = i & 1;
      const int ph = t / c.n_mtiles, mt = t - ph * c.n_mtiles;
      mbar_wait(&tfull[buf], (i >> 1) & 1);
      tc_fence_after();
      float v[16];
      if (cpp == 16) {
        tmem_ld16(tlane + (uint32_t)(buf * 64 + c0), v);
      } else {
        float w[8];
        tmem_ld8_nowait(tlane + (uint32_t)(buf * 64 + c0), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = w[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);      // the accumulator may be overwritten: its values are in registers
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < cpp) {
          const float x = v[j] + bias_s[c0 + j];
          v[j] = c.act == 1 ? siluf_(x) : x;
        }
      }
      float4* dst = reinterpret_cast<float4*>(tile + row * pitch + c0);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      if (cpp == 16) {
        dst[2] = make_float4(v[8], v[9], v[10], v[11]);
        dst[3] = make_float4(v[12], v[13], v[14], v[15]);
      }
      epi_bar_sync();
      RowMap rm = c.rm;
      if (c.phases > 1) rm.p2 = ph;
      tile_copy_out(tile, pitch, c.bn, min(c.bn, c.n_valid), mt * BM, c.M, nullptr, 0, c.out, c.ld, tid, rm);
      epi_bar_sync();                                // the tile is free for the next iteration's writes
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

}  // namespace drm
