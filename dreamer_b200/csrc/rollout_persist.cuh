// The imagination rollout (Dreamer.dream_episodes, Dreamer.py:143-175) as ONE persistent kernel: the time loop, the actor,
// the GRU step, the prior head + categorical sample and the reward / continue heads of every step run inside a single
// cooperative launch of one CTA per SM.  Included from rssm.cu.
//
// Why it is shaped like this (DESIGN.md section 4 has the numbers):
//   * At 1024 start states a step is 9 GFLOP = 6.5 us of tensor work but SEVEN dependent GEMM stages; launched one by one each
//     stage pays a prologue (barrier init, TMEM allocation, descriptor fetch), a launch boundary and a cold pipeline.  Here the
//     CTAs stay resident: barriers, TMEM and the tile constants live across steps, and stages hand over through per-m-tile
//     counters in global memory (release / acquire) instead of kernel boundaries -- an m-tile moves on as soon as ITS
//     producers are done, there is no grid-wide barrier anywhere.
//   * CTAs are specialised.  "GRU" CTAs own one (m-tile, n-tile) of the GRU for the whole horizon; "prior", "sample" and
//     "head" CTAs own one m-tile's MLP chain.  The action enters the GRU through 3 (A <= 4) input columns only, so the GRU
//     CTAs run the whole [z | h] contraction (26 of the 27 k-blocks at the reference sizes) on the tensor cores WHILE the actor
//     chain of the same step is still running, keep the accumulator in TMEM, and add the action term in the epilogue on the
//     CUDA cores once the action is there.  The per-step critical path is actor chain -> GRU epilogue -> prior chain; the GRU
//     main loop, the reward and the continue head are off it.
//   * Weights stay L2-resident and are streamed by TMA every step.  Holding them in shared memory across steps (north-star item
//     1 read literally) needs an N-partition in which every CTA re-reads ALL state rows each step: >= 38 partitions x 3.4 MB
//     = 129 MB of L2 -> SM traffic per step against 117 MB for the output-stationary tiling used here -- the weights (6.2 MB
//     GRU + 3.9 MB heads) and the state (3.4 MB) are the same order of magnitude at this batch size, so residency buys nothing.
//
// Ordering / deadlock freedom: every CTA walks its items in the total order (state j; prior L1/L2 < sample < heads < GRU);
// every wait is on a task that is earlier in that order, so the globally earliest unfinished task can always run.  A GRU CTA
// blocks early (it starts its main loop before the actor of the same step is done), which is safe because GRU CTAs hold GRU
// items only.  All spins are bounded: after ~2 s a waiter records who / what it was waiting for in a host-mapped debug buffer
// and traps, so a scheduling bug fails the launch instead of hanging the GPU.
#pragma once

namespace drm {

constexpr int PS_STAGES = 4;
constexpr int PS_STAGE_BYTES = A_STAGE_BYTES + 256 * BK * 2;       // 48 KB: A tile + up to 256 weight rows per k-block
constexpr int PS_RING_BYTES = PS_STAGES * PS_STAGE_BYTES;          // 192 KB
constexpr int PS_GRU_STAGE_BYTES = A_STAGE_BYTES + 192 * BK * 2;   // 40 KB: GRU CTAs (<= 192 weight rows per k-block)
constexpr int PS_HP_OFF = PS_STAGES * PS_GRU_STAGE_BYTES;          // GRU CTAs: h_prev tile, 128 x 68 fp32, behind their ring
constexpr int PS_WA_OFF = PS_HP_OFF + BM * 68 * 4;                 // GRU CTAs: action-term weights [3U] float4
constexpr int PS_BAR_OFF = PS_RING_BYTES + 16384;
constexpr int PS_EPI_OFF = PS_BAR_OFF + 256;
constexpr int PS_SCHED_OFF = PS_EPI_OFF + 16384;
constexpr int PS_MAX_ITEMS = 20;
constexpr int PS_SCHED_STRIDE = 1 + 3 * PS_MAX_ITEMS;              // ints per CTA: n, then (kind, m, x) per item
constexpr int PS_TOTAL = PS_SCHED_OFF + 256 + 1024;
static_assert(PS_WA_OFF + 3 * 64 * 16 <= PS_BAR_OFF, "GRU aux region overflows");
static_assert(PS_SCHED_STRIDE * 4 <= 256, "schedule record too large");
static_assert(PS_TOTAL <= 232448, "persistent rollout kernel exceeds the 227 KB shared-memory limit");

enum { PS_P12 = 0, PS_P3 = 1, PS_HEAD = 2, PS_GRU = 3 };
enum { PF_H = 0, PF_Z = 1, PF_A = 2, PF_HL1 = 3, PF_P2 = 4, PF_COUNT = 5 };   // per-m-tile counters (one 128-byte line each)
constexpr int PS_DBG_WORDS = 8 * 160 + 8;

struct PersistParams {
  CUtensorMap tmS[2], tmY1, tmY2, tmWgru, tmWp1, tmWp2, tmWp3, tmWh1, tmWh2, tmWh3;
  int B, H, D, DP, ZP, R, A, NB, KS, Mp, mt;
  int U, nt, bn_cat, nq;
  int bnp1, bnp2, bnh1, bnh2, hp1, hp2, hh1, hh2;
  const float *b_ih, *b_hh, *p1_b, *p1_g, *p1_be, *p2_b, *p2_g, *p2_be, *p3_b;
  const float *h1_b, *h1_g, *h1_be, *h2_b, *h2_g, *h2_be, *h3_b, *bk_rew;
  const __nv_bfloat16* Wgru;
  __nv_bfloat16 *S[2], *Y1, *Y2;
  const float *uniforms, *normals;
  float *latent, *hidden, *actions, *rewards, *continues, *mu, *sigma;
  uint8_t* idx;
  unsigned* flags;
  unsigned* dbg;
  const int* sched;
};

// ------------------------------------------------------------------------------------------
// bounded waits
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void ps_timeout(unsigned* dbg, unsigned code, unsigned seen, unsigned want) {
  unsigned* r = dbg + 8 * (blockIdx.x % 160);
  r[0] = code; r[1] = seen; r[2] = want; r[3] = threadIdx.x; r[4] = blockIdx.x;
  __threadfence_system();
  atomicAdd(dbg + 8 * 160, 1u);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ unsigned long long ps_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long PS_TIMEOUT_NS = 2000000000ull;
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
// spin until *f >= want (monotonic counter written with red.release by the producers of the data)
__device__ __forceinline__ void ps_flag_wait(const unsigned* f, unsigned want, unsigned* dbg, unsigned code) {
  if (f == nullptr || want == 0) return;
  unsigned n = 0;
  unsigned long long t0 = 0;
  for (;;) {
    const unsigned v = ld_acquire(f);
    if (v >= want) return;
    if ((++n & 255u) == 0) {
      const unsigned long long t = ps_now();
      if (t0 == 0) t0 = t;
      else if (t - t0 > PS_TIMEOUT_NS) ps_timeout(dbg, code, v, want);
    }
  }
}
__device__ __forceinline__ void ps_mbar_wait(uint64_t* bar, uint32_t parity, unsigned* dbg, unsigned code) {
  unsigned n = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 1023u) == 0) {
      const unsigned long long t = ps_now();
      if (t0 == 0) t0 = t;
      else if (t - t0 > PS_TIMEOUT_NS) ps_timeout(dbg, code, parity, 0xFFFFFFFFu);
    }
  }
}

// ------------------------------------------------------------------------------------------
// GRU epilogue of the persistent kernel: EpiGru's gate math + the action term
//   gi = W_ih [z, a] + b_ih: the z columns were contracted on the tensor cores, the A (<= 4) action columns are added here.
// The action and its weights are rounded to bf16 exactly as the state buffer / packed weights of the launch-per-stage path
// round them, so both paths agree to fp32 summation order.
// ------------------------------------------------------------------------------------------
template <int U>
struct EpiGruP {
  static constexpr int UP = 8;
  static constexpr int PASSES = U / (EPI_PARTS * UP);
  static constexpr int PITCH = U + 4;
  struct Params {
    const float *b_ih, *b_hh;
    const float* h_prev;      // fp32 [M, ld_h]
    float* h_out;             // fp32 [M, ld_h]
    __nv_bfloat16* s_h;       // bf16 h columns of the next state buffer [M, ld_s]
    long ld_h;
    int ld_s, D;
    const __nv_bfloat16* w_a; // packed GRU weights, first action column (row pitch ldw)
    int ldw, A;
    const float* actions;     // fp32 [M, ld_act]: this step's action (written by the actor head of the same step)
    long ld_act;
  };
  // constants [b_r | b_z | b_in | b_hn] -> sm;  action weights -> wa[3U] float4;  h_prev tile -> hp (pitch U + 4)
  static __device__ __forceinline__ void stage(const Params& p, int n_tile, int m0, int M, float* sm, float4* wa, float* hp, int tid) {
    const int D = p.D;
    for (int i = tid; i < U; i += EPI_THREADS) {
      const int u = n_tile * U + i;
      const bool ok = u < D;
      sm[i] = ok ? __ldg(p.b_ih + u) + __ldg(p.b_hh + u) : 0.f;
      sm[U + i] = ok ? __ldg(p.b_ih + D + u) + __ldg(p.b_hh + D + u) : 0.f;
      sm[2 * U + i] = ok ? __ldg(p.b_ih + 2 * D + u) : 0.f;
      sm[3 * U + i] = ok ? __ldg(p.b_hh + 2 * D + u) : 0.f;
    }
    for (int i = tid; i < 3 * U; i += EPI_THREADS) {   // packed rows of this tile: [r (U) | z (U) | n (U)]
      const __nv_bfloat16* w = p.w_a + (long)(n_tile * 3 * U + i) * p.ldw;
      float4 v;
      v.x = p.A > 0 ? __bfloat162float(w[0]) : 0.f;
      v.y = p.A > 1 ? __bfloat162float(w[1]) : 0.f;
      v.z = p.A > 2 ? __bfloat162float(w[2]) : 0.f;
      v.w = p.A > 3 ? __bfloat162float(w[3]) : 0.f;
      wa[i] = v;
    }
    const int u0 = n_tile * U;
    const int nvalid = min(U, D - u0);
    for (int i = tid; i < BM * (U / 4); i += EPI_THREADS) {
      const int r = i / (U / 4), cc = (i % (U / 4)) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && cc < nvalid) {
        const float* src = p.h_prev + (long)(m0 + r) * p.ld_h + u0 + cc;
        if (cc + 4 <= nvalid && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) x = __ldcg(reinterpret_cast<const float4*>(src));
        else { x.x = __ldcg(src); if (cc + 1 < nvalid) x.y = __ldcg(src + 1); if (cc + 2 < nvalid) x.z = __ldcg(src + 2); if (cc + 3 < nvalid) x.w = __ldcg(src + 3); }
      }
      *reinterpret_cast<float4*>(hp + r * PITCH + cc) = x;
    }
  }
  static __device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  static __device__ __forceinline__ void run(const Params& p, int n_tile, int M, const float* sm, const float4* wa, const float* hp,
                                             float* tile, uint32_t taddr, int m, int row, int part, int tid) {
    const int u0 = n_tile * U;
    const int m0 = m - row;
    const int nvalid = min(U, p.D - u0);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) {
      const float* ap = p.actions + (long)m * p.ld_act;
      a.x = bf16r(__ldcg(ap));
      if (p.A > 1) a.y = bf16r(__ldcg(ap + 1));
      if (p.A > 2) a.z = bf16r(__ldcg(ap + 2));
      if (p.A > 3) a.w = bf16r(__ldcg(ap + 3));
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
        const float4 wr = wa[c + j], wz = wa[U + c + j], wn = wa[2 * U + c + j];
        const float ar = fmaf(a.x, wr.x, fmaf(a.y, wr.y, fmaf(a.z, wr.z, a.w * wr.w)));
        const float az = fmaf(a.x, wz.x, fmaf(a.y, wz.y, fmaf(a.z, wz.z, a.w * wz.w)));
        const float an = fmaf(a.x, wn.x, fmaf(a.y, wn.y, fmaf(a.z, wn.z, a.w * wn.w)));
        const float rr = sigmoidf_(r_[j] + ar + sm[c + j]);
        const float zz = sigmoidf_(z_[j] + az + sm[U + c + j]);
        const float nn = tanhf_(nx[j] + an + sm[2 * U + c + j] + rr * (nh[j] + sm[3 * U + c + j]));
        hn[j] = (1.0f - zz) * nn + zz * hp[row * PITCH + c + j];
      }
      tile_put<UP>(tile, PITCH, row, c, hn);
    }
    epi_bar_sync();
    tile_copy_out(tile, PITCH, U, nvalid, m0, M, p.h_out + u0, p.ld_h, p.s_h + u0, p.ld_s, tid);
  }
};

// ------------------------------------------------------------------------------------------
// one GEMM tile inside the persistent kernel
// ------------------------------------------------------------------------------------------
struct PsCtx {
  uint8_t* smem;
  uint64_t *full, *empty, *tmem_full;
  uint32_t tmem;
  uint32_t it;        // k-blocks issued so far (ring position / phase), identical in every thread
  uint32_t tile_no;   // tiles finished so far (tmem_full phase)
  unsigned* dbg;
};
struct PsTile {
  const CUtensorMap *tmA, *tmB;
  int a_row, b_row;
  int ka0, nka0, ka1, nka1;   // A k-block ranges [ka0, ka0 + nka0) then [ka1, ka1 + nka1)
  int b_follows_a;            // 1: the B k-block index equals the A k-block index (GRU weights keep the state's column layout)
  int bn;                     // B rows per k-block (GRU: 3U)
  int stage_bytes;            // ring stride of this CTA's role
  const unsigned* w0; unsigned t0;               // before the first A load
  const unsigned *e0, *e1; unsigned et0, et1;    // before the epilogue reads produced data / overwrites consumed data
  unsigned* sig;                                  // += 1 once the tile's outputs are visible
  unsigned code;                                  // (kind << 24) | (j << 8) | m-tile, for the timeout record
};

// GRU_U = 0: plain N = bn accumulator.  Fn epilogue(tid) runs on the EPI_THREADS epilogue threads between the tmem_full wait and
// the publication of the tile; pre(tid) runs on them while the main loop is in flight.
template <int GRU_U, class Pre, class Epi>
__device__ __forceinline__ void ps_run_tile(PsCtx& c, const PsTile& t, Pre&& pre, Epi&& epilogue) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = t.nka0 + t.nka1;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)A_STAGE_BYTES + (uint32_t)t.bn * BK * 2;
      const int npre = min(nk, PS_STAGES);
      // weights first: they do not depend on anything, so the ring is pre-filled with them while the dependency is still open
      for (int kb = 0; kb < npre; ++kb) {
        const uint32_t i = c.it + kb, s = i % PS_STAGES;
        ps_mbar_wait(&c.empty[s], ((i / PS_STAGES) & 1) ^ 1u, c.dbg, t.code | (1u << 20));
        mbar_expect_tx(&c.full[s], tx);
        const int ka = kb < t.nka0 ? t.ka0 + kb : t.ka1 + (kb - t.nka0);
        tma_load_2d(c.smem + s * t.stage_bytes + A_STAGE_BYTES, t.tmB, (t.b_follows_a ? ka : kb) * BK, t.b_row, &c.full[s]);
      }
      ps_flag_wait(t.w0, t.t0, c.dbg, t.code | (2u << 20));
      fence_proxy_async_all();   // the A operand was written through the generic proxy (possibly by another SM): order it before the TMA reads
      for (int kb = 0; kb < npre; ++kb) {
        const uint32_t s = (c.it + kb) % PS_STAGES;
        const int ka = kb < t.nka0 ? t.ka0 + kb : t.ka1 + (kb - t.nka0);
        tma_load_2d(c.smem + s * t.stage_bytes, t.tmA, ka * BK, t.a_row, &c.full[s]);
      }
      for (int kb = npre; kb < nk; ++kb) {
        const uint32_t i = c.it + kb, s = i % PS_STAGES;
        ps_mbar_wait(&c.empty[s], ((i / PS_STAGES) & 1) ^ 1u, c.dbg, t.code | (3u << 20));
        mbar_expect_tx(&c.full[s], tx);
        const int ka = kb < t.nka0 ? t.ka0 + kb : t.ka1 + (kb - t.nka0);
        tma_load_2d(c.smem + s * t.stage_bytes, t.tmA, ka * BK, t.a_row, &c.full[s]);
        tma_load_2d(c.smem + s * t.stage_bytes + A_STAGE_BYTES, t.tmB, (t.b_follows_a ? ka : kb) * BK, t.b_row, &c.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < nk; ++kb) {
        const uint32_t i = c.it + kb, s = i % PS_STAGES;
        ps_mbar_wait(&c.full[s], (i / PS_STAGES) & 1, c.dbg, t.code | (4u << 20));
        tc_fence_after();
        const uint32_t a_addr = smem_u32(c.smem + s * t.stage_bytes);
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
        if constexpr (GRU_U == 0) {
          const uint32_t idesc = umma_idesc_bf16(t.bn);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(c.tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        } else {
          // TMEM columns [r | z | n_x | n_h]: z k-blocks feed r, z, n_x in one N = 3U MMA; h k-blocks feed r, z (N = 2U) and n_h (N = U)
          constexpr int U = GRU_U;
          if (kb < t.nka0) {
            const uint32_t idesc = umma_idesc_bf16(3 * U);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(c.tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          } else {
            const uint32_t idesc_rz = umma_idesc_bf16(2 * U), idesc_n = umma_idesc_bf16(U);
            const uint64_t bdesc_n = umma_desc_sw128(a_addr + A_STAGE_BYTES + 2 * U * BK * 2);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16(c.tmem, adesc + 2 * k, bdesc + 2 * k, idesc_rz, 1u);
              umma_bf16(c.tmem + 3 * U, adesc + 2 * k, bdesc_n + 2 * k, idesc_n, (kb > t.nka0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&c.empty[s]);
      }
      umma_commit(c.tmem_full);
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    pre(tid);
    if (lane == 0) {
      ps_flag_wait(t.e0, t.et0, c.dbg, t.code | (5u << 20));
      ps_flag_wait(t.e1, t.et1, c.dbg, t.code | (6u << 20));
    }
    __syncwarp();
    epi_bar_sync();
    if (lane == 0) ps_mbar_wait(c.tmem_full, c.tile_no & 1u, c.dbg, t.code | (7u << 20));
    __syncwarp();
    tc_fence_after();
    epilogue(tid);
    // publish: this thread's global stores are made visible device-wide (and to the async proxy of the consumers' TMA loads)
    __threadfence();
    fence_proxy_async_all();
    epi_bar_sync();
    if (tid == 0 && t.sig) red_release_add(t.sig, 1u);
  }
  c.it += (uint32_t)nk;
  c.tile_no += 1;
  // the next tile reuses the ring (the epilogue's transposition buffer) and the TMEM accumulator
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) rollout_persist_kernel(const __grid_constant__ PersistParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PS_BAR_OFF);
  uint64_t* empty = full + PS_STAGES;
  uint64_t* tmem_full = empty + PS_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* epi_sm = reinterpret_cast<float*>(smem + PS_EPI_OFF);
  int* sched = reinterpret_cast<int*>(smem + PS_SCHED_OFF);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.tmS[0]); tma_prefetch_desc(&P.tmS[1]); tma_prefetch_desc(&P.tmY1); tma_prefetch_desc(&P.tmY2);
    tma_prefetch_desc(&P.tmWgru); tma_prefetch_desc(&P.tmWp1); tma_prefetch_desc(&P.tmWp2); tma_prefetch_desc(&P.tmWp3);
    tma_prefetch_desc(&P.tmWh1); tma_prefetch_desc(&P.tmWh2); tma_prefetch_desc(&P.tmWh3);
    for (int s = 0; s < PS_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < PS_SCHED_STRIDE; i += GEMM_THREADS) sched[i] = __ldg(P.sched + (long)blockIdx.x * PS_SCHED_STRIDE + i);
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  PsCtx c;
  c.smem = smem; c.full = full; c.empty = empty; c.tmem_full = tmem_full; c.tmem = *tmem_slot; c.it = 0; c.tile_no = 0; c.dbg = P.dbg;
  const int n_items = sched[0];
  const int B = P.B, H = P.H, D = P.D, ZP = P.ZP, A = P.A, R = P.R, Mp = P.Mp, mt = P.mt;
  const int nkz = ZP / 64, nkh = P.DP / 64, kh0 = nkz + 1;
  const long ldL = (long)(H + 1) * ZP, ldH = (long)(H + 1) * D, ldA = (long)H * A;
  const int q = warp & 3, part = (warp - 2) >> 2, lane = threadIdx.x & 31;
  const int row = q * 32 + lane;
  const uint32_t taddr = c.tmem + ((uint32_t)(q * 32) << 16);
  float* tile = reinterpret_cast<float*>(smem);
  auto flag = [&](int kind, int m) { return P.flags + (long)(kind * mt + m) * 32; };

  for (int j = 0; j <= H; ++j) {
    const int sb = j & 1;
#pragma unroll 1
    for (int it = 0; it < n_items; ++it) {
      const int kind = sched[1 + 3 * it], m_tile = sched[2 + 3 * it], x = sched[3 + 3 * it];
      const int m0 = m_tile * BM;
      const int m = m0 + row;
      PsTile t;
      t.stage_bytes = PS_STAGE_BYTES; t.b_follows_a = 0; t.ka1 = 0; t.nka1 = 0;
      t.w0 = nullptr; t.t0 = 0; t.e0 = nullptr; t.e1 = nullptr; t.et0 = 0; t.et1 = 0; t.sig = nullptr;
      t.code = ((unsigned)kind << 24) | ((unsigned)j << 8) | (unsigned)m_tile;
      if (kind == PS_P12) {
        if (j == 0) continue;
        {   // prior L1: h_j -> Y1[slot 0]   (DynamicsPredictors.py:15-18)
          t.tmA = &P.tmS[sb]; t.tmB = &P.tmWp1; t.a_row = m0; t.b_row = 0; t.ka0 = kh0; t.nka0 = nkh; t.bn = P.bnp1;
          t.w0 = flag(PF_H, m_tile); t.t0 = (unsigned)(P.nt * j);
          const TileG g{B, P.bnp1, 0};
          const EpiLnSilu::Params p{P.p1_b, P.p1_g, P.p1_be, nullptr, 0, P.Y1, 256, 0, 0, P.hp1, 1e-5f, P.bnp1};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, 0, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, 0, tid); });
        }
        {   // prior L2: Y1 -> Y2[slot 0]   (:19-22); same CTA, so program order + the publication fences order the hand-over
          t.tmA = &P.tmY1; t.tmB = &P.tmWp2; t.a_row = m0; t.b_row = 0; t.ka0 = 0; t.nka0 = (P.hp1 + 63) / 64; t.bn = P.bnp2;
          t.w0 = nullptr; t.t0 = 0; t.sig = flag(PF_P2, m_tile);
          t.code |= 1u << 16;
          const TileG g{B, P.bnp2, 0};
          const EpiLnSilu::Params p{P.p2_b, P.p2_g, P.p2_be, nullptr, 0, P.Y2, 256, 0, 0, P.hp2, 1e-5f, P.bnp2};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, 0, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, 0, tid); });
        }
      } else if (kind == PS_P3) {
        if (j == 0) continue;
        // prior logits + sample of latent rows [x * bn_cat / 32, ..): z_j -> S[sb] z columns, latent[:, j], idx[:, j - 1]   (:23, 31-40)
        t.tmA = &P.tmY2; t.tmB = &P.tmWp3; t.a_row = m0; t.b_row = x * P.bn_cat; t.ka0 = 0; t.nka0 = (P.hp2 + 63) / 64; t.bn = P.bn_cat;
        t.w0 = flag(PF_P2, m_tile); t.t0 = (unsigned)j;
        t.sig = flag(PF_Z, m_tile);
        const TileG g{B, P.bn_cat, 0};
        const EpiCat::Params p{P.p3_b, P.uniforms + (long)(j - 1) * B * R, P.latent + (long)j * ZP, nullptr,
                               P.idx ? P.idx + (long)(j - 1) * R : nullptr, P.S[sb], nullptr, ldL, 0, (long)H * R, 0, P.KS, R,
                               RowMap{0, 0, 0, 0}};
        ps_run_tile<0>(c, t, [&](int tid) { EpiCat::stage(p, g, x, epi_sm, tid, m0); },
                       [&](int tid) { EpiCat::run(p, g, epi_sm, tile, taddr, m, row, part, x, tid); });
      } else if (kind == PS_HEAD) {
        const int head = x;
        if (head == HS_ACTOR ? j >= H : j == 0) continue;   // actor on states 0 .. H-1, reward / continue on states 1 .. H
        {   // L1: [z_j | h_j] -> Y1[slot 1 + head]
          t.tmA = &P.tmS[sb]; t.tmB = &P.tmWh1; t.a_row = m0; t.b_row = head * 256;
          t.ka0 = 0; t.nka0 = nkz; t.ka1 = kh0; t.nka1 = nkh; t.bn = P.bnh1;
          t.w0 = flag(PF_Z, m_tile); t.t0 = (unsigned)(P.nq * j);
          t.sig = flag(PF_HL1, m_tile);
          const TileG g{B, P.bnh1, 0};
          const EpiLnSilu::Params p{P.h1_b, P.h1_g, P.h1_be, nullptr, 0, P.Y1, 256, Mp, Mp, P.hh1, 1e-5f, P.bnh1};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
        {   // L2: Y1 -> Y2
          t.tmA = &P.tmY1; t.tmB = &P.tmWh2; t.a_row = Mp + head * Mp + m0; t.b_row = head * 256;
          t.ka0 = 0; t.nka0 = (P.hh1 + 63) / 64; t.ka1 = 0; t.nka1 = 0; t.bn = P.bnh2;
          t.w0 = nullptr; t.t0 = 0; t.sig = nullptr;
          t.code |= 1u << 16;
          const TileG g{B, P.bnh2, 0};
          const EpiLnSilu::Params p{P.h2_b, P.h2_g, P.h2_be, nullptr, 0, P.Y2, 256, Mp, Mp, P.hh2, 1e-5f, P.bnh2};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
        {   // output layer
          t.tmA = &P.tmY2; t.tmB = &P.tmWh3; t.a_row = Mp + head * Mp + m0; t.b_row = head * 256;
          t.ka0 = 0; t.nka0 = (P.hh2 + 63) / 64; t.bn = 256;
          t.sig = head == HS_ACTOR ? flag(PF_A, m_tile) : nullptr;
          t.code |= 2u << 16;
          const TileG g{B, 256, 0};
          EpiHeads::Params hp;
          memset(&hp, 0, sizeof(hp));
          hp.bias = P.h3_b;
          hp.kind[HS_REWARD] = HEAD_BUCKET; hp.kind[HS_CONT] = HEAD_SIGMOID; hp.kind[HS_ACTOR] = HEAD_ACTOR;
          hp.buckets[HS_REWARD] = P.bk_rew;
          hp.NB = P.NB; hp.A = A;
          if (head == HS_ACTOR) {
            hp.normals = P.normals + (long)j * B * A; hp.ld_normals = A;
            hp.mu = P.mu + (long)j * A; hp.sigma = P.sigma + (long)j * A; hp.action = P.actions + (long)j * A; hp.ld_act = ldA;
          } else {
            hp.value[HS_REWARD] = P.rewards + (j - 1); hp.ld_value[HS_REWARD] = H;
            hp.value[HS_CONT] = P.continues + (j - 1); hp.ld_value[HS_CONT] = H;
          }
          ps_run_tile<0>(c, t, [&](int tid) { EpiHeads::stage(hp, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiHeads::run(hp, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
      } else {   // PS_GRU: h_{j+1} = GRU([z_j, a_j], h_j)   (SequenceModel.py:19-24)
        if (j >= H) continue;
        const int n_tile = x;
        t.stage_bytes = PS_GRU_STAGE_BYTES;
        t.tmA = &P.tmS[sb]; t.tmB = &P.tmWgru; t.a_row = m0; t.b_row = n_tile * 3 * P.U;
        t.ka0 = 0; t.nka0 = nkz; t.ka1 = kh0; t.nka1 = nkh; t.b_follows_a = 1; t.bn = 3 * P.U;
        t.w0 = flag(PF_Z, m_tile); t.t0 = (unsigned)(P.nq * j);            // z_j sampled (implies h_j)
        t.e0 = flag(PF_A, m_tile); t.et0 = (unsigned)(j + 1);               // a_j
        t.e1 = flag(PF_HL1, m_tile); t.et1 = j >= 1 ? (unsigned)(1 + 3 * (j - 1)) : 0u;   // every head of state j - 1 has read the buffer h_{j+1} overwrites
        t.sig = flag(PF_H, m_tile);
        float4* wa = reinterpret_cast<float4*>(smem + PS_WA_OFF);
        float* hp_tile = reinterpret_cast<float*>(smem + PS_HP_OFF);
        __nv_bfloat16* s_h = P.S[sb ^ 1] + ZP + 64;
        if (P.U == 32) {
          const EpiGruP<32>::Params p{P.b_ih, P.b_hh, P.hidden + (long)j * D, P.hidden + (long)(j + 1) * D, s_h, ldH, P.KS, D,
                                      P.Wgru + ZP, P.KS, A, P.actions + (long)j * A, ldA};
          ps_run_tile<32>(c, t, [&](int tid) { EpiGruP<32>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<32>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); });
        } else {
          const EpiGruP<64>::Params p{P.b_ih, P.b_hh, P.hidden + (long)j * D, P.hidden + (long)(j + 1) * D, s_h, ldH, P.KS, D,
                                      P.Wgru + ZP, P.KS, A, P.actions + (long)j * A, ldA};
          ps_run_tile<64>(c, t, [&](int tid) { EpiGruP<64>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<64>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); });
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(c.tmem, 256);
}

}  // namespace drm

// ------------------------------------------------------------------------------------------
// host side: static schedule, launch
// ------------------------------------------------------------------------------------------
struct drm_persist {
  int U = 0, bn_cat = 0, nt = 0, nq = 0, n_cta = 0;
  int* sched = nullptr;        // device [n_cta * PS_SCHED_STRIDE]
  unsigned* flags = nullptr;   // device [PF_COUNT * mt * 32]
  unsigned* dbg = nullptr;     // host-mapped [PS_DBG_WORDS]
  size_t flag_bytes = 0;
  bool tried = false, ok = false;
};

namespace drm {

static void persist_free(drm_persist* ps) {
  if (!ps) return;
  if (ps->dbg) cudaFreeHost(ps->dbg);
  delete ps;
}

static int persist_sm_count() {
  static int n = -1;
  if (n < 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
    int per_sm = 0;
    if (sms > 0 && cudaFuncSetAttribute(rollout_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_TOTAL) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_persist_kernel, GEMM_THREADS, PS_TOTAL) == cudaSuccess && per_sm >= 1)
      n = sms;
    else
      n = 0;
    cudaGetLastError();
  }
  return n;
}

// Static schedule: one GRU tile, one prior chain, one sampling tile or one actor chain per CTA (the per-step critical path and
// the GRU accumulators never queue behind other work); the reward / continue chains share whatever CTAs are left.
static bool persist_plan(drm_rollout* r, drm_persist* ps) {
  drm_rssm* m = r->m;
  const int n_sm = persist_sm_count();
  const int mt = r->Mp / BM;
  if (n_sm <= 0 || m->d.A > 4) return false;
  static const int cand[4][2] = {{32, 128}, {64, 128}, {32, 256}, {64, 256}};
  for (int ci = 0; ci < 4; ++ci) {
    const int U = cand[ci][0], bn_cat = cand[ci][1];
    if (m->ZP % bn_cat) continue;
    const int nt = ceil_div(m->d.D, U), nq = m->ZP / bn_cat;
    const int crit = mt * nt + mt + mt * nq + mt;
    const int rest = n_sm - crit;
    if (rest < 1 || ceil_div(2 * mt, rest) > PS_MAX_ITEMS) continue;
    ps->U = U; ps->bn_cat = bn_cat; ps->nt = nt; ps->nq = nq;
    std::vector<int> sc((size_t)n_sm * PS_SCHED_STRIDE, 0);
    int cta = 0;
    auto add = [&](int c, int kind, int mm, int x) {
      int* rec = sc.data() + (size_t)c * PS_SCHED_STRIDE;
      const int n = rec[0]++;
      rec[1 + 3 * n] = kind; rec[2 + 3 * n] = mm; rec[3 + 3 * n] = x;
    };
    // critical-path roles first so that they land on distinct SMs; m-tile-major so that an m-tile's chain is spread over the chip
    for (int mm = 0; mm < mt; ++mm) add(cta++, PS_HEAD, mm, HS_ACTOR);
    for (int mm = 0; mm < mt; ++mm) add(cta++, PS_P12, mm, 0);
    for (int mm = 0; mm < mt; ++mm)
      for (int qq = 0; qq < nq; ++qq) add(cta++, PS_P3, mm, qq);
    for (int mm = 0; mm < mt; ++mm)
      for (int n = 0; n < nt; ++n) add(cta++, PS_GRU, mm, n);
    const int first_rest = cta;
    for (int mm = 0, k = 0; mm < mt; ++mm)
      for (int hd = 0; hd < 2; ++hd, ++k) add(first_rest + k % rest, PS_HEAD, mm, hd == 0 ? HS_REWARD : HS_CONT);
    ps->n_cta = std::min(n_sm, first_rest + std::min(rest, 2 * mt));
    if (dev_alloc(r->allocs, &ps->sched, sc.size()) != DRM_OK) return false;
    if (cudaMemcpy(ps->sched, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    ps->flag_bytes = (size_t)PF_COUNT * mt * 32 * sizeof(unsigned);
    if (dev_alloc(r->allocs, &ps->flags, ps->flag_bytes / sizeof(unsigned)) != DRM_OK) return false;
    if (cudaHostAlloc((void**)&ps->dbg, PS_DBG_WORDS * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return false; }
    memset(ps->dbg, 0, PS_DBG_WORDS * sizeof(unsigned));
    return true;
  }
  return false;
}

static bool persist_eligible(drm_rollout* r) {
  if (!r->ps) r->ps = new drm_persist();
  drm_persist* ps = r->ps;
  if (!ps->tried) {
    ps->tried = true;
    ps->ok = persist_plan(r, ps);
  }
  return ps->ok;
}

static int rollout_persist(drm_rollout* r, const float* z0, const float* h0, const float* uniforms, const float* normals, float* latent,
                           float* hidden, float* actions, float* rewards, float* continues, float* mu, float* sigma, uint8_t* idx,
                           cudaStream_t st) {
  drm_rssm* m = r->m;
  drm_persist* ps = r->ps;
  const int B = r->B, H = r->H, D = m->d.D, ZP = m->ZP;
  const long ldL = (long)(H + 1) * ZP, ldH = (long)(H + 1) * D;
  RC(pack_cols(r->S[0], m->KS, 0, z0, ZP, ZP, B, latent, ldL, st));
  RC(pack_cols(r->S[0], m->KS, ZP + 64, h0, D, D, B, hidden, ldH, st));
  DRM_CUDA(cudaMemsetAsync(ps->flags, 0, ps->flag_bytes, st));
  PersistParams P;
  memset(&P, 0, sizeof(P));
  const int v = ps->U == 64 ? 1 : 0;
  P.tmS[0] = r->tmS[0]; P.tmS[1] = r->tmS[1]; P.tmY1 = r->tmY1; P.tmY2 = r->tmY2;
  P.tmWgru = m->tmWgru2[v]; P.tmWp1 = m->tmWp1; P.tmWp2 = m->tmWp2; P.tmWp3 = ps->bn_cat == 128 ? m->tmWp3h : m->tmWp3;
  P.tmWh1 = m->tmWh1; P.tmWh2 = m->tmWh2; P.tmWh3 = m->tmWh3;
  P.B = B; P.H = H; P.D = D; P.DP = m->DP; P.ZP = ZP; P.R = m->d.R; P.A = m->d.A; P.NB = m->d.NB; P.KS = m->KS; P.Mp = r->Mp; P.mt = r->Mp / BM;
  P.U = ps->U; P.nt = ps->nt; P.bn_cat = ps->bn_cat; P.nq = ps->nq;
  P.bnp1 = m->bnp1; P.bnp2 = m->bnp2; P.bnh1 = m->bnh1; P.bnh2 = m->bnh2;
  P.hp1 = m->d.h_prior[0]; P.hp2 = m->d.h_prior[1]; P.hh1 = m->d.h_head[0]; P.hh2 = m->d.h_head[1];
  P.b_ih = m->b_ih; P.b_hh = m->b_hh;
  P.p1_b = m->p1_b; P.p1_g = m->p1_g; P.p1_be = m->p1_be; P.p2_b = m->p2_b; P.p2_g = m->p2_g; P.p2_be = m->p2_be; P.p3_b = m->p3_b;
  P.h1_b = m->h1_b; P.h1_g = m->h1_g; P.h1_be = m->h1_be; P.h2_b = m->h2_b; P.h2_g = m->h2_g; P.h2_be = m->h2_be; P.h3_b = m->h3_b;
  P.bk_rew = m->bk_rew;
  P.Wgru = m->Wgru2[v];
  P.S[0] = r->S[0]; P.S[1] = r->S[1]; P.Y1 = r->Y1; P.Y2 = r->Y2;
  P.uniforms = uniforms; P.normals = normals;
  P.latent = latent; P.hidden = hidden; P.actions = actions; P.rewards = rewards; P.continues = continues; P.mu = mu; P.sigma = sigma;
  P.idx = idx;
  P.flags = ps->flags; P.dbg = ps->dbg; P.sched = ps->sched;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ps->n_cta);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = PS_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident, or the launch fails: the in-kernel hand-overs need every producer running
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, rollout_persist_kernel, P));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

}  // namespace drm
