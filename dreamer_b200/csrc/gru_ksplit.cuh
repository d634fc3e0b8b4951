// GRU stage for TINY grids (one m-tile: the 16-sequence posterior scan, the B = 1 acting path), split over K across a 2-CTA
// cluster.  Included from rssm.cu.
//
// A GRU tile's main loop (26 k-blocks, 148 MMA issues by ONE thread) takes ~10 us no matter how the operands arrive
// (knock-out experiments).  The lever is fewer MMA issues per CTA:
//   rank 0 accumulates the x part (z | a k-blocks:  gi = x W_ih^T),  rank 1 the h part (gh = h W_hh^T) of the SAME 128 x 3U tile,
// each with one N = 3U MMA per k-step into its own TMEM accumulator [r | z | n] -- the h part no longer needs the split
// r,z / n MMAs because n_h lives in the other CTA.  Then the CTAs swap halves: rank 0 finishes rows 0..63, rank 1 rows 64..127;
// every epilogue thread whose row belongs to the peer writes its 3 x 8 partial sums into the peer's shared memory
// (st.shared::cluster), one cluster barrier, and the owner combines  r = sig(r_x + r_h + b), z likewise,
// n = tanh(n_x + b_in + r (n_h + b_hn)),  h' = (1 - z) n + z h  (nn.GRUCell, SequenceModel.py:13-24).
// Measured (D = 600, per imagined step): 128 rows 62.1 -> 59.4 us; from two m-tiles on it is slower (256 rows 62.9 -> 65.1 us,
// 1024 rows with 48-unit tiles +8 us), so the launch takes this path only for single-m-tile grids of the posterior scan and of
// drm_gru_step (imagination rollouts keep batch-size-independent per-row arithmetic: shards concatenate bit-exactly).
#pragma once

namespace drm {

template <int U>
struct GruKsSmem {
  static constexpr int STAGES = U == 32 ? 3 : 2;
  static constexpr int SUB = A_STAGE_BYTES + 3 * U * BK * 2;       // 28 KB (U = 32), 34 KB (U = 48)
  static constexpr int XBUF_OFF = STAGES * SUB;                    // incoming peer partials: [64 rows][3U] fp32
  static constexpr int XBUF_BYTES = 64 * 3 * U * 4;
  static constexpr int BAR_OFF = XBUF_OFF + XBUF_BYTES;
  static constexpr int CONST_OFF = BAR_OFF + 128;
  static constexpr int TOTAL = CONST_OFF + 4 * U * 4;
  static_assert(SUB % 1024 == 0, "stages must keep 1024-byte alignment");
};

template <int U>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gru_ksplit_kernel(const __grid_constant__ GemmCommon g, const __grid_constant__ typename EpiGru<U>::Params ep) {
  using SL = GruKsSmem<U>;
  constexpr int STAGES = SL::STAGES;
  constexpr int UP = U / 4;                            // units per epilogue thread: 8 (U = 32) or 12 (U = 48)
  constexpr int TCOLS = 3 * U <= 128 ? 128 : 256;      // TMEM columns (power of two >= 3U)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  float* xbuf = reinterpret_cast<float*>(smem + SL::XBUF_OFF);
  float* cst = reinterpret_cast<float*>(smem + SL::CONST_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();            // 0: x part, 1: h part
  const int slot = (int)blockIdx.y;
  const int a_row = g.a_row0 + (int)blockIdx.x * BM;
  const int b_row = slot * 3 * U;
  const int kb0 = rank == 0 ? 0 : g.nka0, nkb = rank == 0 ? g.nka0 : g.nka1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.tmA);
    tma_prefetch_desc(&g.tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  cluster_sync_all();     // the peer is resident before anybody writes into its shared memory
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (g.a_bytes ? (uint32_t)g.a_bytes : (uint32_t)A_STAGE_BYTES) + (uint32_t)(3 * U) * BK * 2;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1u);
        uint8_t* sa = smem + s * SL::SUB;
        const int ka = rank == 0 ? g.ka0 + i : g.ka1 + i;
        mbar_expect_tx(&full[s], tx);
        tma_load_2d(sa, &g.tmA, ka * BK, a_row, &full[s]);
        tma_load_2d(sa + A_STAGE_BYTES, &g.tmB, (kb0 + i) * BK, b_row, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(3 * U);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        mbar_wait(&full[s], (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SL::SUB);
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(tfull);
    }
    __syncwarp();
  }

  float acc[3][16];         // this CTA's partial sums of the thread's row and UP units: [r | z | n] (16-column loads; UP are used)
  const int tid = (int)threadIdx.x - 64;
  const int q = warp & 3, part = (warp - 2) >> 2;
  const int row = q * 32 + lane;
  const int c = part * UP;                           // first of this thread's units inside the tile
  const bool epi = warp >= 2;
  const bool mine = epi && (row >> 6) == rank;       // rows 0..63 are finished by rank 0, rows 64..127 by rank 1
  if (epi) {
    EpiGru<U>::stage(ep, g, slot, cst, tid, (int)blockIdx.x * BM);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    if constexpr (UP == 8) {
      tmem_ld8_nowait(tlane + (uint32_t)c, acc[0]);
      tmem_ld8_nowait(tlane + (uint32_t)(U + c), acc[1]);
      tmem_ld8_nowait(tlane + (uint32_t)(2 * U + c), acc[2]);
      tmem_ld_wait();
    } else {               // 12 units: 16-column loads, the 4 extra columns (the next units / gate, inside the allocation) are ignored
      tmem_ld16(tlane + (uint32_t)c, acc[0]);
      tmem_ld16(tlane + (uint32_t)(U + c), acc[1]);
      tmem_ld16(tlane + (uint32_t)(2 * U + c), acc[2]);
    }
    if (!mine) {             // hand the partials of this row to the peer that finishes it
      const uint32_t base = smem_u32(xbuf + (row & 63) * 3 * U + c);
#pragma unroll
      for (int gte = 0; gte < 3; ++gte) {
#pragma unroll
        for (int h = 0; h < UP / 4; ++h) {
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(base + (uint32_t)((gte * U + 4 * h) * 4)), "r"(rank ^ 1));
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(remote), "f"(acc[gte][4 * h]), "f"(acc[gte][4 * h + 1]),
                       "f"(acc[gte][4 * h + 2]), "f"(acc[gte][4 * h + 3])
                       : "memory");
        }
      }
    }
  }
  __syncwarp();
  cluster_arrive_release();  // every thread of both CTAs, exactly once
  cluster_wait_acquire();
  if (mine) {
    const int m = (int)blockIdx.x * BM + row;
    const int u0 = slot * U + c;
    if (m < g.M) {
      const float* px = xbuf + (row & 63) * 3 * U + c;      // the peer's partials of this row
      const int nvalid = min(UP, ep.D - u0);
      float hp[UP], hn[UP];
#pragma unroll
      for (int j = 0; j < UP; ++j) hp[j] = j < nvalid ? __ldg(ep.h_prev + (long)m * ep.ld_hprev + u0 + j) : 0.f;
#pragma unroll
      for (int j = 0; j < UP; ++j) {
        const float pr = px[j], pz = px[U + j], pn = px[2 * U + j];
        const float nx = rank == 0 ? acc[2][j] : pn, nh = rank == 0 ? pn : acc[2][j];
        const float rr = sigmoidf_(acc[0][j] + pr + cst[c + j]);
        const float zz = sigmoidf_(acc[1][j] + pz + cst[U + c + j]);
        const float nn = tanhf_(nx + cst[2 * U + c + j] + rr * (nh + cst[3 * U + c + j]));
        hn[j] = (1.0f - zz) * nn + zz * hp[j];
      }
      float* o = ep.h_out + (long)m * ep.ld_hout + u0;
      __nv_bfloat16* ob = ep.s_h + (long)m * ep.ld_s + u0;
      if (nvalid == UP && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(ob) & 7u) == 0)) {
#pragma unroll
        for (int j = 0; j < UP; j += 4) {
          *reinterpret_cast<float4*>(o + j) = make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]);
          *reinterpret_cast<uint2*>(ob + j) = make_uint2(pack_bf16x2(hn[j], hn[j + 1]), pack_bf16x2(hn[j + 2], hn[j + 3]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < UP; ++j)
          if (j < nvalid) { o[j] = hn[j]; ob[j] = __float2bfloat16_rn(hn[j]); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TCOLS);
}

}  // namespace drm
