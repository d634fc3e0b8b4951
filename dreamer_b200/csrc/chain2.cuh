// Two dependent layers in ONE CTA (no cluster, no cross-CTA exchange):  Linear-LN-SiLU  ->  an output stage (EpiB).
// Included from rssm.cu.
//
//   layer A   A = activations via TMA, B = the LN layer's weight tile (<= 256 rows)          -> TMEM columns [0, 256)
//             epilogue: LayerNorm + SiLU (EpiLnSiluT::normalise) written as bf16, 128-byte swizzled, into shared memory
//   layer B   A = that shared-memory tile (4 k-blocks of [128 x 64]), B = EpiB's weight tile -> TMEM columns [256, 512)
//             epilogue: EpiB::run (categorical sample, head outputs, ...)
//
// Used for (prior L2 -> logits + sample), (head L2 -> head outputs) and (posterior L1 -> logits + sample) on SMALL grids, where
// each stage is a few microseconds of work behind ~1.3 us of launch boundary plus a global-memory round trip of the hidden
// activations.  When layer B is tiled over columns (blockIdx.y = column tile of the 1024 logits) every column tile recomputes
// layer A for its rows: K = 256 (or 640), a few k-blocks -- cheaper than the boundary it removes.
// The chained-CLUSTER experiment (chain.cuh) showed that exchanging activations between CTAs costs more than a launch
// boundary; here nothing leaves the SM.
#pragma once

namespace drm {

constexpr int C2_STAGES = 2;
constexpr int C2_STAGE = A_STAGE_BYTES + 256 * BK * 2;     // A 16 KB + B up to 256 rows
constexpr int C2_Y_OFF = C2_STAGES * C2_STAGE;              // 96 KB
constexpr int C2_Y_BYTES = 4 * A_STAGE_BYTES;               // 4 k-blocks of [128 x 64] bf16
constexpr int C2_BAR_OFF = C2_Y_OFF + C2_Y_BYTES;
constexpr int C2_SB_OFF = C2_BAR_OFF + 256;                 // EpiB scratch (16 KB)
constexpr int C2_SA_OFF = C2_SB_OFF + 16384;                // LN constants + partial statistics (8 KB)
constexpr int C2_SMEM = C2_SA_OFF + 8192 + 1024;

struct Chain2Common {
  CUtensorMap tmA, tmWA, tmWB;
  int M, a_row0, a_y_stride;      // layer-A rows: a_row0 + slotA * a_y_stride + blockIdx.x * 128
  int ka0, nka0, ka1, nka1;       // layer-A k-block ranges of tmA
  int a_bytes;                    // bytes per A k-block load (0 = full tile; short-box maps for few rows)
  int bnA, wa_slot_rows;          // layer-A MMA N (hidden width padded to 32); weight rows per slot
  int nkB, bnB, wb_tile_rows;     // layer-B k-blocks, MMA N, weight rows per blockIdx.y step
  int per_slot;                   // 1: blockIdx.y selects a head slot (both layers' weights and the A rows); 0: a column tile of layer B
  int n_slots;
  int y_slot[8];
};

template <class EpiB, bool HAS_ADD>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
chain2_kernel(const __grid_constant__ Chain2Common c, const __grid_constant__ typename EpiLnSiluT<HAS_ADD>::Params pa,
              const __grid_constant__ typename EpiB::Params pb) {
  using EpiA = EpiLnSiluT<HAS_ADD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C2_BAR_OFF);
  uint64_t* empty = full + C2_STAGES;
  uint64_t* tfull = empty + C2_STAGES;    // [2]: layer A, layer B accumulators complete
  uint64_t* ybar = tfull + 2;             // the shared-memory activations are complete (16 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ybar + 1);
  uint8_t* ybuf = smem + C2_Y_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int by = (int)blockIdx.y;
  const int slot = c.per_slot ? (c.n_slots > 0 ? c.y_slot[by] : by) : 0;     // head slot (weights, constants, A rows)
  const int slotB = c.per_slot ? slot : by;                                  // what EpiB calls "slot"
  const int a_row = c.a_row0 + slot * c.a_y_stride + (int)blockIdx.x * BM;
  const int wa_row = slot * c.wa_slot_rows;
  const int wb_row = (c.per_slot ? slot : by) * c.wb_tile_rows;
  const int nkA = c.nka0 + c.nka1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&c.tmA); tma_prefetch_desc(&c.tmWA); tma_prefetch_desc(&c.tmWB);
    for (int s = 0; s < C2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(ybar, EPI_THREADS / 32);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_a = c.a_bytes ? (uint32_t)c.a_bytes : (uint32_t)A_STAGE_BYTES;
      for (int it = 0; it < nkA + c.nkB; ++it) {
        const int s = it % C2_STAGES;
        const uint32_t ph = (it / C2_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * C2_STAGE;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (it < nkA) {
          const int ka = it < c.nka0 ? c.ka0 + it : c.ka1 + (it - c.nka0);
          mbar_expect_tx(&full[s], tx_a + (uint32_t)c.bnA * BK * 2);
          tma_load_2d(sa, &c.tmA, ka * BK, a_row, &full[s]);
          tma_load_2d(sb, &c.tmWA, it * BK, wa_row, &full[s]);
        } else {
          mbar_expect_tx(&full[s], (uint32_t)c.bnB * BK * 2);
          tma_load_2d(sb, &c.tmWB, (it - nkA) * BK, wb_row, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idescA = umma_idesc_bf16(c.bnA), idescB = umma_idesc_bf16(c.bnB);
      for (int it = 0; it < nkA + c.nkB; ++it) {
        const int s = it % C2_STAGES;
        const uint32_t ph = (it / C2_STAGES) & 1;
        const bool layerB = it >= nkA;
        const int kb = layerB ? it - nkA : it;
        if (layerB && kb == 0) {        // the LayerNorm epilogue has written the whole activation tile
          mbar_wait(ybar, 0);
          tc_fence_after();
        }
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t st_addr = smem_u32(smem + s * C2_STAGE);
        const uint64_t adesc = umma_desc_sw128(layerB ? smem_u32(ybuf) + kb * A_STAGE_BYTES : st_addr);
        const uint64_t bdesc = umma_desc_sw128(st_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem + (layerB ? 256u : 0u), adesc + 2 * k, bdesc + 2 * k, layerB ? idescB : idescA, (kb | k) != 0);
        umma_commit(&empty[s]);
        if (it == nkA - 1) umma_commit(&tfull[0]);
        if (it == nkA + c.nkB - 1) umma_commit(&tfull[1]);
      }
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    float* smA = reinterpret_cast<float*>(smem + C2_SA_OFF);
    float* smB = reinterpret_cast<float*>(smem + C2_SB_OFF);
    GemmCommon gA, gB;                // the epilogues read only bn and M
    gA.bn = c.bnA; gA.M = c.M;
    gB.bn = c.bnB; gB.M = c.M;
    EpiA::stage(pa, gA, slot, smA, tid);
    EpiB::stage(pb, gB, slotB, smB, tid);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int m = (int)blockIdx.x * BM + row;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    // ---- layer A epilogue: LayerNorm + SiLU -> bf16 activations in shared memory (the A operand of layer B) ----
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const uint32_t ybase = smem_u32(ybuf) + row * 128;
    EpiA::normalise(pa, gA, smA, tlane, m, row, part, 64 * c.nkB, [&](int col, float (&v)[32]) {
      // 32 columns = 4 chunks of 16 bytes of k-block col / 64; SWIZZLE_128B: chunk j of row r sits at chunk j ^ (r & 7)
      const uint32_t kb_base = ybase + (uint32_t)(col >> 6) * A_STAGE_BYTES;
      const int j0 = (col & 63) >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 w;
        w.x = pack_bf16x2(v[8 * j], v[8 * j + 1]); w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(kb_base + (uint32_t)(((j0 + j) ^ (row & 7)) << 4)), "r"(w.x), "r"(w.y),
                     "r"(w.z), "r"(w.w)
                     : "memory");
      }
    });
    fence_proxy_async();          // generic-proxy writes -> visible to the tensor cores' operand reads
    __syncwarp();
    if (lane == 0) mbar_arrive(ybar);
    // ---- layer B epilogue ----
    mbar_wait(&tfull[1], 0);
    tc_fence_after();
    EpiB::run(pb, gB, smB, reinterpret_cast<float*>(smem), tlane + 256, m, row, part, slotB, tid);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace drm
