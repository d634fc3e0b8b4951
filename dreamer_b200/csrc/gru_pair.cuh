// GRU stage on CTA PAIRS (tcgen05 cta_group::2).  Included from rssm.cu.
//
// The GRU stage is bound by operand delivery into shared memory: with one CTA per 128 x 3U tile every
// SM receives the whole weight tile.  Here two CTAs of a cluster (two neighbouring m-tiles, same n-tile) issue ONE
// M = 256 MMA: each CTA stages its own 128 state rows and only HALF of the weight tile's rows -- the tensor cores read the
// other half from the peer's shared memory -- so the weight bytes per SM halve while the accumulator layout per CTA
// ([r | z | n_x | n_h] in its own TMEM lanes) and therefore the epilogue (EpiGru<U>::run) stay exactly as they are.
//
//   B rows per stage   rank 0: [ r (U) | n first half (U/2) ]      rank 1: [ z (U) | n second half (U/2) ]
//   MMAs per k-step    N = 2U over the gate rows -> TMEM columns [0, 2U);  N = U over the n halves -> [2U, 3U) for x
//                      k-blocks, [3U, 4U) for h k-blocks (r multiplies only W_hn h, SequenceModel.py:13)
//   TMA                both CTAs load (cp.async.bulk.tensor ... cta_group::2), all bytes are counted on the LEADER's full barrier
//   MMA / commits      leader only; tcgen05.commit ... multicast::cluster releases the stage / publishes the accumulator in both
#pragma once

namespace drm {

__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
// TMA tile load whose completion bytes are counted on the LEADER CTA's mbarrier (peer bit 24 of the barrier address cleared)
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {   // arrives on this barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

template <int U>
struct GruPairSmem {
  static constexpr int STAGES = U == 64 ? 3 : 4;
  static constexpr int B_BYTES = (3 * U / 2) * BK * 2;            // this CTA's half of the weight tile per k-block
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;     // 28 KB (U = 64) / 22 KB (U = 32); multiples of 1024
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = EPI_OFF + 16384 + 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stages must keep 1024-byte alignment");
};

// grid (2 * ceil(m_tiles / 2), n_tiles), cluster (2, 1, 1).  g.tmB must have box rows U / 2.
template <int U>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gru_pair_kernel(const __grid_constant__ GemmCommon g, const __grid_constant__ typename EpiGru<U>::Params ep) {
  using Epi = EpiGru<U>;
  using SL = GruPairSmem<U>;
  constexpr int STAGES = SL::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int nk = g.nka0 + g.nka1;
  // Tile order: CTAs become resident in linear blockIdx order (x fastest).  Walking all m-tiles of one n-tile first makes
  // a resident wave (2 x 148 CTAs) cover every state row but only ~2 weight tiles, so the state buffer streams from DRAM
  // once per ~2 n-tiles (ncu: 5.6 GB read for a 1 GB problem).  Bands of GM m-tiles x all n-tiles keep a wave's state rows
  // and weight tiles both L2-resident.  GM is even, so the two CTAs of a cluster stay neighbouring m-tiles of one n-tile.
  int m_tile = (int)blockIdx.x, slot = (int)blockIdx.y;
  {
    const int GM = g.band > 0 ? g.band : 16;
    const int GX = (int)gridDim.x, NT = (int)gridDim.y;
    const int full_bands = GX / GM;
    const int id = (int)blockIdx.x + GX * (int)blockIdx.y;
    if (id < full_bands * GM * NT) {
      const int band = id / (GM * NT), r = id - band * GM * NT;
      slot = r / GM;
      m_tile = band * GM + (r - slot * GM);
    } else {                      // the remaining (GX % GM) m-tiles: plain order over the tail columns of the grid
      const int rem = GX - full_bands * GM, r = id - full_bands * GM * NT;
      slot = r / rem;
      m_tile = full_bands * GM + (r - slot * rem);
    }
  }
  const int a_row = g.a_row0 + m_tile * BM;
  const int b_row = slot * 3 * U;

  if (threadIdx.x == 0) {
    cta_probe(g.cta_times, 0);
    tma_prefetch_desc(&g.tmA);
    tma_prefetch_desc(&g.tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 4 * U);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  cluster_sync_all();   // both CTAs' barriers exist and both allocations are done before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      cta_probe(g.cta_times, 1);
      // rows of the packed tile [r (U) | z (U) | n (U)] this CTA stages: its gate block, then its half of n
      const int gate_row = b_row + rank * U, n_row = b_row + 2 * U + rank * (U / 2);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * SL::STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        const int ka = kb < g.nka0 ? g.ka0 + kb : g.ka1 + (kb - g.nka0);
        if (rank == 0) mbar_expect_tx(&full[s], 2u * SL::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
        tma_load_2d_2cta(sa, &g.tmA, ka * BK, a_row, &full[s]);
        tma_load_2d_2cta(sb, &g.tmB, kb * BK, gate_row, &full[s]);
        tma_load_2d_2cta(sb + (U / 2) * BK * 2, &g.tmB, kb * BK, gate_row + U / 2, &full[s]);
        tma_load_2d_2cta(sb + U * BK * 2, &g.tmB, kb * BK, n_row, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc_g = umma_idesc_bf16_m256(2 * U), idesc_n = umma_idesc_bf16_m256(U);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc_g = umma_desc_sw128(a_addr + A_STAGE_BYTES);
        const uint64_t bdesc_n = umma_desc_sw128(a_addr + A_STAGE_BYTES + U * BK * 2);
        const bool xpart = kb < g.nka0;
        const uint32_t ncol = xpart ? 2 * U : 3 * U;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_bf16_2cta(tmem, adesc + 2 * k, bdesc_g + 2 * k, idesc_g, (kb | k) != 0);
          const uint32_t acc_n = xpart ? (uint32_t)((kb | k) != 0) : (uint32_t)(kb > g.nka0 || k > 0);
          umma_bf16_2cta(tmem + ncol, adesc + 2 * k, bdesc_n + 2 * k, idesc_n, acc_n);
        }
        umma_commit_2cta(&empty[s]);
      }
      umma_commit_2cta(tmem_full);
    }
  } else {
    float* epi_sm = reinterpret_cast<float*>(smem + SL::EPI_OFF);
    Epi::stage(ep, g, slot, epi_sm, (int)threadIdx.x - 64, m_tile * BM);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int m = m_tile * BM + row;
    Epi::run(ep, g, epi_sm, reinterpret_cast<float*>(smem), tmem + ((uint32_t)(q * 32) << 16), m, row, part, slot, (int)threadIdx.x - 64);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM: nobody leaves early
  if (warp == 1) tmem_dealloc_2cta(tmem, 4 * U);
  if (threadIdx.x == 32) cta_probe(g.cta_times, 2);
}

}  // namespace drm
