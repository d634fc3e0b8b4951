// The imagination rollout (Dreamer.dream_episodes, Dreamer.py:143-175) as ONE persistent kernel: the time loop, the actor,
// the GRU step, the prior head + categorical sample and the reward / continue heads of every step run inside a single
// launch of 4-CTA clusters, one CTA per SM.  Included from rssm.cu.
//
// Why it is shaped like this (DESIGN.md section 4 has the numbers):
//   * At 1024 start states a step is 9 GFLOP = 6.5 us of tensor work but SEVEN dependent GEMM stages; launched one by one each
//     stage pays a prologue (barrier init, TMEM allocation, descriptor fetch), a launch boundary and a cold pipeline.  Here the
//     CTAs stay resident: barriers, TMEM and descriptors live across steps.
//   * CTAs are specialised:
//       "chain" clusters (one 4-CTA cluster per m-tile of 128 start states) run the serial MLP chain of every state: prior L1 ->
//       prior L2 -> logits + sample -> actor L1 -> actor L2 -> actor output.  The four CTAs split every layer's output columns
//       (64 LayerNorm columns / 256 logit columns each); LayerNorm statistics cross the cluster through distributed shared
//       memory, and layers hand over with ONE barrier.cluster (release / acquire) -- no global flag, no kernel boundary.
//       "GRU" CTAs own one (m-tile, n-tile) of the GRU for the whole horizon.  The action enters the GRU through A <= 4 input
//       columns only, so they run the whole [z | h] contraction (26 of the 27 k-blocks at the reference sizes) on the tensor
//       cores WHILE the actor chain of the same step is still running, keep the accumulator in TMEM, and add the action term in
//       the epilogue on the CUDA cores once the action is there.
//       "head" CTAs (whatever SMs are left) run the reward and continue heads of every sampled state; nothing waits for them.
//     Per-step critical path: chain cluster (6 layers) -> GRU epilogue; three global release / acquire hand-overs per step
//     (h ready, z ready, action ready) through per-m-tile counters, everything else inside a cluster.
//   * The state buffer keeps ALL H + 1 states (time-major slabs of bf16 [z | a | h] rows), so no buffer is ever overwritten
//     while somebody may still read it: the only dependencies are true (read-after-write) ones.
//   * Weights stay L2-resident and are streamed by TMA every step.  Holding them in shared memory across steps (north-star item
//     1 read literally) needs an N-partition in which every CTA re-reads ALL state rows each step: >= 38 partitions x 3.4 MB
//     = 129 MB of L2 -> SM traffic per step against 117 MB for the output-stationary tiling used here -- the weights (6.2 MB
//     GRU + 3.9 MB heads) and the state (3.4 MB) are the same order of magnitude at this batch size, so residency buys nothing.
//
// Deadlock freedom: every wait is on work of an EARLIER point of the chain (state j: prior < sample < actor < GRU < state j + 1),
// every CTA walks its own work in that order, and all CTAs are co-resident (the host checks cudaOccupancyMaxActiveClusters).
// All spins are bounded: after ~2 s a waiter records who / what it was waiting for in a host-mapped debug buffer and traps, so
// a scheduling bug fails the launch instead of hanging the GPU.
#pragma once

namespace drm {

constexpr int PS_STAGES = 4;                                       // ring depth at the full 48 KB stage
constexpr int PS_MAX_STAGES = 8;                                   // ring depth with smaller stages (barrier slots)
constexpr int PS_STAGE_BYTES = A_STAGE_BYTES + 256 * BK * 2;       // 48 KB: A tile + up to 256 weight rows per k-block
constexpr int PS_RING_BYTES = PS_STAGES * PS_STAGE_BYTES;          // 192 KB
constexpr int PS_GRU_STAGE_BYTES = A_STAGE_BYTES + 192 * BK * 2;   // 40 KB: GRU CTAs (<= 192 weight rows per k-block)
constexpr int PS_CHAIN_STAGE_BYTES = PS_RING_BYTES / 2;              // chain CTAs: two slots of 96 KB
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

enum { PS_CHAIN = 0, PS_RC = 2, PS_GRU = 3 };   // CTA roles (the record's first item kind)
// per-m-tile counters (one 128-byte line each): GRU tiles finished (h_{j+1} ready), sampling tiles finished (z_j ready); the
// action a_j travels as per-row records that carry their own ready tag (PsActorOut).  Producers of a counter never run a state ahead of one another, so "counter >= c(j)"
// means exactly "everything of states <= j".
enum { PF_H = 0, PF_Z = 1, PF_LP = 2, PF_LA = 3, PF_COUNT = 4 };   // + loads landed: prior L1 / actor L1 of a state (bandwidth gates of the GRU CTAs)
constexpr int PS_DBG_WORDS = 8 * 160 + 8;
constexpr int PS_TRACE_SLOTS = 32;
constexpr unsigned PS_PUB = EPI_THREADS / 32;   // increments of a hand-over counter per published tile (one per epilogue warp)

struct PersistParams {
  CUtensorMap tmS, tmY1, tmY2, tmWgru, tmWp1q, tmWp2q, tmWp3, tmWh1q, tmWh2q, tmWh3a, tmWh1, tmWh2, tmWh3;
  int B, H, D, DP, ZP, R, A, NB, KS, Mp, mt;
  int U, nt, nq;   // GRU tile width, GRU n-tiles, 256-column sampling tiles
  int bnp1, bnp2, bnh1, bnh2, hp1, hp2, hh1, hh2;
  const float *b_ih, *b_hh, *p1_b, *p1_g, *p1_be, *p2_b, *p2_g, *p2_be, *p3_b;
  const float *h1_b, *h1_g, *h1_be, *h2_b, *h2_g, *h2_be, *h3_b, *bk_rew;
  const __nv_bfloat16* Wgru;
  const __nv_bfloat16* Wh3;   // packed output-layer weights [MAX_HEADS * 256, 256]
  __nv_bfloat16 *S, *Y1, *Y2;   // S: time-major state slabs [(H + 1) * B (+ pad), KS]
  const float *uniforms, *normals;
  float *latent, *hidden, *actions, *rewards, *continues, *mu, *sigma;
  uint8_t* idx;
  unsigned* flags;
  unsigned long long* apack;   // [H, Mp] action records (zeroed with the counters before every launch)
  uint8_t* idx_prev;           // [(H + 1), B, R] classes whose one-hot is currently set in S (255 = none)
  unsigned* dbg;
  const int* sched;
  unsigned long long* trace;   // debug: [cta][PS_TRACE_SLOTS][8] timestamps of the tiles of states [trace_j0, trace_j1), or NULL
  int trace_j0, trace_j1;
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

// debug laps inside an epilogue (thread tid == 0 only): the current tile's lap record (8 x u64) is published in shared memory
// 64 bytes below the epilogue scratch; NULL when tracing is off
__device__ __forceinline__ void ps_lap(const float* epi_sm, int tid, int k) {
  if (tid != 0) return;
  unsigned long long* lap = *reinterpret_cast<unsigned long long* const*>(reinterpret_cast<const uint8_t*>(epi_sm) - 64);
  if (lap) lap[k] = ps_now();
}

// ------------------------------------------------------------------------------------------
// GRU epilogue of the persistent kernel: EpiGru's gate math + the action term
//   gi = W_ih [z, a] + b_ih: the z columns were contracted on the tensor cores, the A (<= 3) action columns are added here; the
//   biases b_r, b_z, b_in ride in the fourth lane of the action-term weights (the action vector is padded with a constant 1).
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
    const unsigned long long* apack;   // [M] action records of this state {bf16 a0, a1, a2, tag}: polled until the tag is set; NULL: the action is in the MMA
    unsigned* dbg; unsigned code;
    long ld_prev;             // row pitch of h_prev when it differs from ld_h (0: the same)
  };
  // b_hn -> sm[U];  action weights + biases -> wa[3U] float4 {w_a0, w_a1, w_a2, bias};  h_prev tile -> hp (pitch U + 4)
  static __device__ __forceinline__ void stage(const Params& p, int n_tile, int m0, int M, float* sm, float4* wa, float* hp, int tid) {
    const int D = p.D;
    for (int i = tid; i < U; i += EPI_THREADS) {
      const int u = n_tile * U + i;
      sm[i] = u < D ? __ldg(p.b_hh + 2 * D + u) : 0.f;
    }
    for (int i = tid; i < 3 * U; i += EPI_THREADS) {   // packed rows of this tile: [r (U) | z (U) | n (U)]
      const __nv_bfloat16* w = p.w_a + (long)(n_tile * 3 * U + i) * p.ldw;
      const int gate = i / U, u = n_tile * U + (i - gate * U);
      float4 v;
      v.x = p.A > 0 ? __bfloat162float(w[0]) : 0.f;
      v.y = p.A > 1 ? __bfloat162float(w[1]) : 0.f;
      v.z = p.A > 2 ? __bfloat162float(w[2]) : 0.f;
      v.w = u < D ? (gate < 2 ? __ldg(p.b_ih + gate * D + u) + __ldg(p.b_hh + gate * D + u) : __ldg(p.b_ih + 2 * D + u)) : 0.f;
      wa[i] = v;
    }
    const int u0 = n_tile * U;
    const int nvalid = min(U, D - u0);
    for (int i = tid; i < BM * (U / 4); i += EPI_THREADS) {
      const int r = i / (U / 4), cc = (i % (U / 4)) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && cc < nvalid) {
        const float* src = p.h_prev + (long)(m0 + r) * (p.ld_prev ? p.ld_prev : p.ld_h) + u0 + cc;
        if (cc + 4 <= nvalid && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) x = __ldcg(reinterpret_cast<const float4*>(src));
        else { x.x = __ldcg(src); if (cc + 1 < nvalid) x.y = __ldcg(src + 1); if (cc + 2 < nvalid) x.z = __ldcg(src + 2); if (cc + 3 < nvalid) x.w = __ldcg(src + 3); }
      }
      *reinterpret_cast<float4*>(hp + r * PITCH + cc) = x;
    }
  }
  // The accumulator is complete ~2 us before the action of the same state exists (the actor chain is the longer path), so every
  // TMEM load happens BEFORE the action poll; behind the poll there is only the gate arithmetic, which is MUFU-bound
  // (16 lanes per clock and SM): r and z share ONE reciprocal (1 / ((1 + e_r)(1 + e_z))), 5 instead of 6 MUFU operations per unit.
  // The bf16 h columns of the next state (what the chain's TMA loads read) leave first; the fp32 `hidden` output follows after
  // the tile has been published (post()).
  static __device__ __forceinline__ void run(const Params& p, int n_tile, int M, const float* sm, const float4* wa, const float* hp,
                                             float* tile, uint32_t taddr, int m, int row, int part, int tid) {
    const int u0 = n_tile * U;
    const int m0 = m - row;
    const int nvalid = min(U, p.D - u0);
    float acc[PASSES][4][UP];
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int c = (ps * EPI_PARTS + part) * UP;
      tmem_ld8_nowait(taddr + c, acc[ps][0]);
      tmem_ld8_nowait(taddr + U + c, acc[ps][1]);
      tmem_ld8_nowait(taddr + 2 * U + c, acc[ps][2]);
      tmem_ld8_nowait(taddr + 3 * U + c, acc[ps][3]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int c = (ps * EPI_PARTS + part) * UP;
#pragma unroll
      for (int j = 0; j < UP; ++j) acc[ps][3][j] += sm[c + j];   // n_h + b_hn
    }
    float4 a = make_float4(0.f, 0.f, 0.f, 1.f);
    if (m < M && p.apack != nullptr) {   // the action arrives as one 8-byte record per row whose top half-word is the ready tag: no separate flag, no second round trip
      // (polling with one thread per row + a shared-memory broadcast, or with a __nanosleep back-off, was measured: no faster)
      unsigned long long rec;
      unsigned n = 0;
      unsigned long long t0 = 0;
      for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(rec) : "l"(p.apack + m) : "memory");
        if (rec >> 48) break;
        if ((++n & 255u) == 0) {
          const unsigned long long t = ps_now();
          if (t0 == 0) t0 = t;
          else if (t - t0 > PS_TIMEOUT_NS) ps_timeout(p.dbg, p.code | (5u << 20), 0u, 1u);
        }
      }
      a.x = __uint_as_float(((unsigned)rec & 0xFFFFu) << 16);
      a.y = __uint_as_float((unsigned)rec & 0xFFFF0000u);
      a.z = __uint_as_float(((unsigned)(rec >> 32) & 0xFFFFu) << 16);
    }
    ps_lap(sm, tid, 0);
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int c = (ps * EPI_PARTS + part) * UP;
      float hn[UP];
#pragma unroll
      for (int j = 0; j < UP; ++j) {
        const float4 wr = wa[c + j], wz = wa[U + c + j], wn = wa[2 * U + c + j];
        const float xr = acc[ps][0][j] + fmaf(a.x, wr.x, fmaf(a.y, wr.y, fmaf(a.z, wr.z, wr.w)));   // + action term + bias
        const float xz = acc[ps][1][j] + fmaf(a.x, wz.x, fmaf(a.y, wz.y, fmaf(a.z, wz.z, wz.w)));
        const float xn = acc[ps][2][j] + fmaf(a.x, wn.x, fmaf(a.y, wn.y, fmaf(a.z, wn.z, wn.w)));
        // (the exponent is capped so that the product of the two denominators stays finite; sigmoid(-41) is 0 in fp32 arithmetic anyway)
        const float dr = 1.0f + ex2f_(fminf(-1.4426950408889634f * xr, 60.0f));
        const float dz = 1.0f + ex2f_(fminf(-1.4426950408889634f * xz, 60.0f));
        const float inv = rcpf_(dr * dz);
        const float rr = inv * dz, zz = inv * dr;
        const float nn = tanhf_(fmaf(rr, acc[ps][3][j], xn));
        hn[j] = fmaf(zz, hp[row * PITCH + c + j] - nn, nn);
      }
      tile_put<UP>(tile, PITCH, row, c, hn);
    }
    ps_lap(sm, tid, 1);
    // (stores straight from registers -- 32 + 16 bytes per thread and pass -- were measured: 3.3 - 4.7 us for math + stores against
    // 2.1 + 1.3 us through the transposed tile, so the coalesced copy-out stays)
    epi_bar_sync();
    ps_lap(sm, tid, 2);
    tile_copy_out(tile, PITCH, U, nvalid, m0, M, nullptr, 0, p.s_h + u0, p.ld_s, tid);
    ps_lap(sm, tid, 3);
  }
  static __device__ __forceinline__ void post(const Params& p, int n_tile, int M, const float* tile, int m0, int tid) {
    const int u0 = n_tile * U;
    tile_copy_out(tile, PITCH, U, min(U, p.D - u0), m0, M, p.h_out + u0, p.ld_h, nullptr, 0, tid);
  }
};

// ------------------------------------------------------------------------------------------
// cluster exchanges without barrier.cluster: st.async writes a few bytes into a peer's shared memory through the async proxy and
// counts them on an mbarrier IN THAT PEER (complete_tx), so the receiver just waits on its own mbarrier until every contribution
// has landed -- no all-thread cluster barrier (measured ~0.9 - 1.2 us with its release of the pending remote stores) in the middle
// of an epilogue.  Receivers arm the barrier with arrive.expect_tx; a barrier / buffer pair is reused every second exchange (a CTA
// can run at most ONE exchange ahead of a peer: exchange k + 1 needs that peer's contribution k + 1, sent after it finished k).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_async_v2(float* local_ptr, uint64_t* local_bar, uint32_t cta, float a, float b) {
  uint32_t ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(cta));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(cta));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];\n" ::"r"(ra), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(rb)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(float* local_ptr, uint64_t* local_bar, uint32_t cta, float a, float b, float c, float d) {
  uint32_t ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(cta));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(cta));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(ra),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(rb)
               : "memory");
}

struct PsXchg {
  float* xst;        // [2][4 ranks][128 rows][2] LayerNorm statistics (mean, M2) of every rank's 64 columns
  uint64_t* xbar;    // [2] mbarriers counting the 4 KB of an exchange
  float* xact;       // rank 0: [4 ranks][128 rows][8] actor-output partial sums
  uint64_t* abar;    // rank 0: mbarrier counting their 16 KB
  uint32_t xuse;     // LayerNorm exchanges so far, identical in every thread of the cluster
  uint32_t ause;     // actor exchanges so far
  unsigned* dbg;
};

// EpiLnSiluN4T<false>::compute with the statistics exchanged by st.async + mbarrier
// addv: 16 values added to the pre-activation before the statistics (the posterior's hoisted feature part), or nullptr
__device__ __forceinline__ void ps_ln_compute(const EpiLnSilu::Params& p, const TileG& g, float* sm, PsXchg& x, uint32_t taddr, int m, int row,
                                              int part, int slot, int tid, unsigned code, float (&v)[16], const float* addv = nullptr) {
  const int nv = p.n_valid;
  const int cr = (int)cluster_ctarank();
  const int c0 = part * 16;
  const int gc0 = 64 * cr + c0;
  auto cnt_of = [](int nvv, int first, int width) { return max(0, min(width, nvv - first)); };
  const int cnt = cnt_of(nv, gc0, 16);
  const uint32_t buf = x.xuse & 1u, par = (x.xuse >> 1) & 1u;
  float* xst = x.xst + buf * (4 * 128 * 2);
  if (tid == 0) mbar_expect_tx(&x.xbar[buf], 4u * 128u * 8u);      // this CTA will receive 8 bytes per row from each of the 4 ranks
  tmem_ld16(taddr + c0, v);
  {
    const float4* b4 = reinterpret_cast<const float4*>(sm + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = b4[j];
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
  }
  if (addv) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] += addv[j];
  }
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
  if (part == 0) {   // merge the CTA's four parts, send (mean, M2) of these 64 columns to every CTA of the cluster
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
    for (uint32_t dst = 0; dst < 4; ++dst) st_async_v2(xst + (cr * 128 + row) * 2, &x.xbar[buf], dst, mean_c, M2c);
  }
  if ((tid & 31) == 0) ps_mbar_wait(&x.xbar[buf], par, x.dbg, code | (11u << 20));
  __syncwarp();
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
  const float4* ga = reinterpret_cast<const float4*>(sm + 256 + c0);
  const float4* be = reinterpret_cast<const float4*>(sm + 512 + c0);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 G = ga[j], Bt = be[j];
    v[4 * j] = siluf_(fmaf(fmaf(v[4 * j], rstd, nmr), G.x, Bt.x));
    v[4 * j + 1] = siluf_(fmaf(fmaf(v[4 * j + 1], rstd, nmr), G.y, Bt.y));
    v[4 * j + 2] = siluf_(fmaf(fmaf(v[4 * j + 2], rstd, nmr), G.z, Bt.z));
    v[4 * j + 3] = siluf_(fmaf(fmaf(v[4 * j + 3], rstd, nmr), G.w, Bt.w));
  }
  x.xuse += 1;
}

// ------------------------------------------------------------------------------------------
// prior logits -> sample (EpiCat's arithmetic, DynamicsPredictors.py:31-40) for the persistent kernel.  The straight-through
// latent (onehot + p) - p is exactly 0 off the sampled class, so the fp32 `latent` output is zero-filled once per rollout
// (cudaMemsetAsync) and the epilogue stores ONE float per 32-class row -- no 128 x 256 fp32 tile through shared memory.
// ------------------------------------------------------------------------------------------
struct EpiCatP {
  using Params = EpiCat::Params;
  // the classes whose one-hot is currently set in this tile's z columns -> shared memory, while the main loop runs (the epilogue then
  // contains no global load at all)
  static __device__ __forceinline__ void stage_prev(const Params& p, const uint8_t* idx_prev, const TileG& g, float* sm, int slot, int tid, int m0) {
    uint8_t* old_sm = reinterpret_cast<uint8_t*>(sm + 2304);
    const int G = g.bn >> 5;
    const int ngrp = max(0, min(g.bn, p.R * 32 - slot * g.bn)) >> 5;
    for (int i = tid; i < BM * 8; i += EPI_THREADS) {
      const int r = i >> 3, gi = i & 7;
      old_sm[i] = (p.s_z && m0 + r < g.M && gi < ngrp) ? __ldcg(idx_prev + (long)(m0 + r) * p.R + slot * G + gi) : (uint8_t)255;
    }
  }
  static __device__ __forceinline__ void run(const Params& p, uint8_t* idx_prev, const TileG& g, float* sm, uint32_t taddr,
                                             int m, int row, int part, int slot, int tid) {
    const int m0 = m - row;
    const int G = g.bn >> 5;
    const int col0 = slot * g.bn;
    const int ncols = max(0, min(g.bn, p.R * 32 - col0));
    uint8_t* idx_sm = reinterpret_cast<uint8_t*>(sm + 2048);       // [128 rows][8] (sm[256, 1280) holds the uniforms)
    const int ngrp = ncols >> 5;
    const uint8_t* old_sm = reinterpret_cast<const uint8_t*>(sm + 2304);   // [128 rows][8]: what is set in S (stage_prev, under the main loop)
#pragma unroll 1
    for (int gi = part; gi < G; gi += EPI_PARTS) {
      float v[32];
      tmem_ld32(taddr + gi * 32, v);
      add_const32(v, sm + gi * 32);
      if (p.logits && m < g.M && gi * 32 < ncols) {   // posterior logits of the observe scan: 128 contiguous bytes per thread, whole sectors
        float4* lg = reinterpret_cast<float4*>(p.logits + (long)m * p.ld_logits + col0 + gi * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) lg[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
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
      float cdf = 0.f, phit = 0.f;
      int idx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pj = fmaf(v[j], k, 0.01f * (1.0f / 32.0f));
        phit = (j == 0 || cdf <= u) ? pj : phit;    // cdf (of the classes before j) <= u  <=>  j <= idx: the last hit is p[idx]
        cdf += pj;
        idx += (cdf <= u) ? 1 : 0;
      }
      idx = idx > 31 ? 31 : idx;
      idx_sm[row * 8 + gi] = (uint8_t)idx;
      if (p.latent && m < g.M && gi * 32 < ncols) p.latent[(long)m * p.ld_latent + col0 + gi * 32 + idx] = (1.0f + phit) - phit;
      ps_lap(sm, tid, gi < 4 ? 0 : 1);
    }
    epi_bar_sync();
    ps_lap(sm, tid, 2);
    if (p.idx) {
      for (int i = tid; i < BM * 8; i += EPI_THREADS) {
        const int r = i >> 3, gi = i & 7;
        if (m0 + r < g.M && gi < ngrp) p.idx[(long)(m0 + r) * p.ld_idx + slot * G + gi] = idx_sm[i];
      }
    }
    if (p.s_z) {
      // The z columns of this slab hold the one-hots of the previous rollout (zeros at first): clear the old class, set the new one
      // -- two 2-byte stores per latent row instead of rewriting 64 bytes of zeros (idx_prev remembers what is set; 255 = nothing)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = tid + e * EPI_THREADS, r = i >> 3, gi = i & 7;
        if (m0 + r >= g.M || gi >= ngrp) continue;
        uint8_t* pv = idx_prev + (long)(m0 + r) * p.R + slot * G + gi;
        const int old = old_sm[i], idx = idx_sm[i];
        unsigned short* zrow = reinterpret_cast<unsigned short*>(p.s_z + (long)(m0 + r) * p.ld_s + col0 + gi * 32);
        if (old != idx) {
          if (old != 255) zrow[old] = 0;
          zrow[idx] = 0x3F80;
          *pv = (uint8_t)idx;
        }
      }
    }
    ps_lap(sm, tid, 3);
  }
};

// ------------------------------------------------------------------------------------------
// actor L2 + output layer in one epilogue (Agent.py:182-187, 199-209).  The output layer is 2A <= 6 rows of 200 weights: instead
// of a seventh GEMM stage, every thread multiplies its 16 bf16-rounded activations with the 2A weight rows on the CUDA cores,
// the row's partial sums cross the cluster through distributed shared memory (like the LayerNorm statistics), and rank 0 turns
// them into mu, sigma and the action.  One more barrier.cluster, one hand-over, one TMA round trip and one epilogue less per step.
// ------------------------------------------------------------------------------------------
struct PsActorOut {
  const __nv_bfloat16* w3;   // packed output weights of the actor slot: row k = mu_k, row 16 + k = log-sigma_k; 256 columns
  const float* b3;           // [32] packed the same way
  const float* normals;      // [M, A] of this state
  float *mu, *sigma, *action;   // [M, ld_act]
  unsigned long long* apack;    // [M] this state's action records {bf16 a0, a1, a2, tag 1}: one 8-byte store = data + ready flag for the GRU CTAs
  long ld_act;
  int A, M;
  static constexpr int WOFF = 3584;   // float offset of the staged weight rows [2A][64] inside the epilogue scratch
  __device__ __forceinline__ void stage(float* sm, int tid) const {
    const int cr = (int)cluster_ctarank();
    for (int i = tid; i < 2 * A * 64; i += EPI_THREADS) {
      const int k = i >> 6, c = i & 63;
      const int wrow = k < A ? k : 16 + (k - A);
      sm[WOFF + i] = __bfloat162float(w3[(long)wrow * 256 + 64 * cr + c]);
    }
  }
  // v: this thread's 16 activations (columns 64 * rank + 16 * part ..).  red: 16 KB of the (idle) pipeline ring; the row sums go to
  // rank 0 by st.async (x.xact / x.abar), ranks 1 .. 3 are done as soon as they have sent theirs.
  __device__ __forceinline__ void run(const float (&v)[16], float* sm, float* red, PsXchg& x, int m, int row, int part, int tid, unsigned code) const {
    const int cr = (int)cluster_ctarank();
    const uint32_t par = x.ause & 1u;
    if (cr == 0 && tid == 0) mbar_expect_tx(x.abar, 4u * 128u * 32u);
    float pa[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pa[k] = 0.f;
    float vb[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) vb[j] = __bfloat162float(__float2bfloat16_rn(v[j]));   // what the next GEMM stage would have read
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < 2 * A) {
        const float4* w4 = reinterpret_cast<const float4*>(sm + WOFF + k * 64 + part * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = w4[j];
          pa[k] = fmaf(vb[4 * j], w.x, fmaf(vb[4 * j + 1], w.y, fmaf(vb[4 * j + 2], w.z, fmaf(vb[4 * j + 3], w.w, pa[k]))));
        }
      }
    }
    float4* r4 = reinterpret_cast<float4*>(red + (part * 128 + row) * 8);
    r4[0] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    r4[1] = make_float4(pa[4], pa[5], pa[6], pa[7]);
    ps_lap(sm, tid, 0);
    epi_bar_sync();
    ps_lap(sm, tid, 1);
    float bm[4] = {0.f, 0.f, 0.f, 0.f}, bl[4] = {0.f, 0.f, 0.f, 0.f}, ep[4] = {0.f, 0.f, 0.f, 0.f};
    if (part == 0) {   // the row's sum over this CTA's 64 columns -> rank 0
      float sacc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) sacc[k] = 0.f;
#pragma unroll
      for (int q = 0; q < EPI_PARTS; ++q) {
        const float4 xx = *reinterpret_cast<const float4*>(red + (q * 128 + row) * 8), yy = *reinterpret_cast<const float4*>(red + (q * 128 + row) * 8 + 4);
        sacc[0] += xx.x; sacc[1] += xx.y; sacc[2] += xx.z; sacc[3] += xx.w; sacc[4] += yy.x; sacc[5] += yy.y; sacc[6] += yy.z; sacc[7] += yy.w;
      }
      st_async_v4(x.xact + (cr * 128 + row) * 8, x.abar, 0u, sacc[0], sacc[1], sacc[2], sacc[3]);
      st_async_v4(x.xact + (cr * 128 + row) * 8 + 4, x.abar, 0u, sacc[4], sacc[5], sacc[6], sacc[7]);
      if (cr == 0 && m < M) {   // (fetched while the partial sums travel)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < A) { bm[k] = __ldg(b3 + k); bl[k] = __ldg(b3 + 16 + k); ep[k] = __ldg(normals + (long)m * A + k); }
      }
    }
    ps_lap(sm, tid, 2);
    x.ause += 1;
    if (cr != 0 || part != 0) return;
    if ((tid & 31) == 0) ps_mbar_wait(x.abar, par, x.dbg, code | (12u << 20));
    __syncwarp();
    ps_lap(sm, tid, 3);
    if (m < M) {
      const float* xact = x.xact;
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = (xact[(0 * 128 + row) * 8 + k] + xact[(1 * 128 + row) * 8 + k]) + (xact[(2 * 128 + row) * 8 + k] + xact[(3 * 128 + row) * 8 + k]);
      ps_lap(sm, tid, 4);
      float muv[4], sgv[4], akv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        muv[k] = 0.f; sgv[k] = 0.f; akv[k] = 0.f;
        if (k < A) {
          muv[k] = t[k] + bm[k];
          float ls = t[A + k] + bl[k];
          ls = fminf(fmaxf(ls, -5.0f), 2.0f);
          // softplus on [-5, 2] and tanh through ex2 / lg2 / rcp (relative error ~1e-6): these 128 threads are the only ones between
          // the last partial sum and the GRU CTAs' wake-up, and libm's expf / log1pf / tanhf are ~150 dependent instructions per action
          float l2;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(1.0f + fexpf_(ls)));
          sgv[k] = fmaf(l2, 0.6931471805599453f, 1e-3f);
          akv[k] = tanhf_(muv[k] + sgv[k] * ep[k]);
        }
      }
      ps_lap(sm, tid, 5);
      // the record the GRU CTAs poll goes out FIRST: the 9 row-strided output stores below are 32 sectors per warp instruction and
      // would sit in front of it in the store pipe
      const unsigned long long rec = (unsigned long long)pack_bf16x2(akv[0], akv[1]) | ((unsigned long long)(pack_bf16x2(akv[2], 0.f) & 0xFFFFu) << 32) | (1ull << 48);
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(apack + m), "l"(rec) : "memory");
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < A) {
          mu[(long)m * ld_act + k] = muv[k];
          sigma[(long)m * ld_act + k] = sgv[k];
          action[(long)m * ld_act + k] = akv[k];
        }
      }
      ps_lap(sm, tid, 6);
    }
  }
};

// ------------------------------------------------------------------------------------------
// one GEMM tile inside the persistent kernel
// ------------------------------------------------------------------------------------------
struct PsCtx {
  uint8_t* smem;
  uint64_t *full, *empty, *tmem_full;
  uint32_t tmem;
  uint32_t nsh;       // log2 of the ring depth of this CTA (constant for its whole life: the ring's phase bookkeeping runs across tiles)
  uint32_t it;        // pipeline stages issued so far (ring position / phase), identical in every thread
  uint32_t tile_no;   // tiles finished so far (tmem_full phase)
  unsigned* dbg;
  unsigned long long* tr;   // this tile's trace record (8 x u64) or NULL
  long lap_off;             // distance (u64 words) from a tile's trace record to its lap record
};
struct PsTile {
  const CUtensorMap *tmA, *tmB;
  int a_row, b_row;
  int ka0, nka0, ka1, nka1;   // A k-block ranges [ka0, ka0 + nka0) then [ka1, ka1 + nka1)
  int h_first;                // GRU tiles: 1 = the first range is the h part (r, z, n_h accumulators start there), the second the z part
  int b_follows_a;            // 1: the B k-block index equals the A k-block index (GRU weights keep the state's column layout)
  int bn;                     // B rows per k-block (GRU: 3U); with tmB2: rows of BOTH weight blocks (the MMA's N)
  const CUtensorMap* tmB2;    // second weight block stacked under the first in shared memory (rows bn - n1), or NULL
  int n1, b2_row, b2_koff;    // rows of the first block; row / k-block offset of the second
  int tcol;                   // TMEM column of the accumulator
  int acc0;                   // 1: the accumulator already holds a partial sum (the first MMA accumulates)
  int kps;                    // k-blocks per pipeline stage.  Every full / empty handshake (tcgen05.commit -> mbarrier -> producer -> TMA -> mbarrier)
                              // costs ~0.3 us whatever the ring depth (profiles/scan_knobs.py: 5 k-blocks per stage in ONE slot beat 1 per stage in 8),
                              // so the chain CTAs run TWO slots of 96 KB and put 2 - 4 k-blocks into a stage
  int stage_bytes;            // ring stride of this CTA's role
  int a_bytes;                // shared-memory bytes of the A part of a k-block (A_STAGE_BYTES, or less with a short-box tensor map)
  int a_tx;                   // bytes one A load delivers (A_STAGE_BYTES; less with a short-box tensor map: the tile's tail rows stay stale and are never stored)
  int cbar;                   // cluster barriers the epilogue executes (the producer / MMA warps mirror them)
  int chain;                  // 1: a cluster hand-over follows this tile
  const unsigned* w0; unsigned t0;     // data dependency of the first A k-block range (NULL: resolved by a barrier / program order)
  const unsigned* g0; unsigned gt0;    // bandwidth gate of the first range: wait until the chain's loads of that phase have landed
  const unsigned* w1; unsigned t1;     // the same two for the second range (only honoured with kps == 1)
  const unsigned* g1; unsigned gt1;
  unsigned* lsig;                      // += 1 (relaxed) once every operand of this tile has landed in shared memory
  const unsigned* e0; unsigned et0;    // before the epilogue reads data another CTA produced
  unsigned* sig;                       // += 1 (device-scope release) once the tile's outputs are visible
  unsigned code;                       // (role << 24) | (layer << 16) | (j << 8) | m-tile, for the timeout / trace records
};

__device__ __forceinline__ int ps_ka(const PsTile& t, int kb) { return kb < t.nka0 ? t.ka0 + kb : t.ka1 + (kb - t.nka0); }

// GRU_U = 0: plain N = bn accumulator.  pre(tid) runs on the EPI_THREADS epilogue threads while the main loop is in flight,
// epilogue(tid) between the accumulator-complete wait and the publication of the tile.
struct PsNoPost { __device__ __forceinline__ void operator()(int) const {} };
template <int GRU_U, class Pre, class Epi>
__device__ __forceinline__ void ps_run_tile(PsCtx& c, const PsTile& t, Pre&& pre, Epi&& epilogue);
template <int GRU_U, class Pre, class Epi, class Post>
__device__ __forceinline__ void ps_run_tile(PsCtx& c, const PsTile& t, Pre&& pre, Epi&& epilogue, Post&& post) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = t.nka0 + t.nka1;
  const int kps = t.kps;
  const int n_st = (nk + kps - 1) / kps;
  const int sub_bytes = t.a_bytes + t.bn * BK * 2;   // one k-block: [A | B]
  const uint32_t NSH = c.nsh, NSM = (1u << NSH) - 1u;   // slot = stage & mask, phase = stage >> shift
  if (warp == 0) {
    if (lane == 0) {
      if (c.tr) { c.tr[0] = t.code; c.tr[1] = ps_now(); }
      // issue pipeline stages [st0, st1): with an open dependency / gate the weights go first (they depend on nothing), so the ring is
      // pre-filled while waiting; otherwise A and B of a stage are issued together
      const int tx_sub = sub_bytes - t.a_bytes + t.a_tx;
      auto load_b = [&](uint32_t s, int u, int kb) {
        uint8_t* sb = c.smem + s * t.stage_bytes + u * sub_bytes + t.a_bytes;
        tma_load_2d(sb, t.tmB, (t.b_follows_a ? ps_ka(t, kb) : kb) * BK, t.b_row, &c.full[s]);
        if (t.tmB2) tma_load_2d(sb + t.n1 * BK * 2, t.tmB2, (kb + t.b2_koff) * BK, t.b2_row, &c.full[s]);
      };
      auto issue = [&](int st0, int st1, const unsigned* w, unsigned tw, const unsigned* gte, unsigned tg, bool mark) {
        int st = st0;
        if (w != nullptr || gte != nullptr) {
          const int npre = min(st1 - st0, (int)NSM + 1);
          for (; st < st0 + npre; ++st) {
            const uint32_t i = c.it + st, s = i & NSM;
            ps_mbar_wait(&c.empty[s], ((i >> NSH) & 1) ^ 1u, c.dbg, t.code | (1u << 20));
            const int n_sub = min(kps, nk - st * kps);
            mbar_expect_tx(&c.full[s], (uint32_t)(n_sub * tx_sub));
            for (int u = 0; u < n_sub; ++u) load_b(s, u, st * kps + u);
          }
          ps_flag_wait(gte, tg, c.dbg, t.code | (6u << 20));
          ps_flag_wait(w, tw, c.dbg, t.code | (2u << 20));
          fence_proxy_async_all();   // the A operand was written through the generic proxy by another SM: order it before the TMA reads
          if (mark && c.tr) c.tr[2] = ps_now();
          for (int s2 = st0; s2 < st0 + npre; ++s2) {
            const uint32_t s = (c.it + s2) & NSM;
            const int n_sub = min(kps, nk - s2 * kps);
            for (int u = 0; u < n_sub; ++u)
              tma_load_2d(c.smem + s * t.stage_bytes + u * sub_bytes, t.tmA, ps_ka(t, s2 * kps + u) * BK, t.a_row, &c.full[s]);
          }
        } else if (mark) {
          fence_proxy_async_all();   // (dependency already resolved by a cluster barrier / program order)
          if (c.tr) c.tr[2] = ps_now();
        }
        for (; st < st1; ++st) {
          const uint32_t i = c.it + st, s = i & NSM;
          ps_mbar_wait(&c.empty[s], ((i >> NSH) & 1) ^ 1u, c.dbg, t.code | (3u << 20));
          const int n_sub = min(kps, nk - st * kps);
          mbar_expect_tx(&c.full[s], (uint32_t)(n_sub * tx_sub));
          for (int u = 0; u < n_sub; ++u) {
            const int kb = st * kps + u, ka = ps_ka(t, kb);
            tma_load_2d(c.smem + s * t.stage_bytes + u * sub_bytes, t.tmA, ka * BK, t.a_row, &c.full[s]);
            load_b(s, u, kb);
          }
        }
      };
      if (t.w1 != nullptr || t.g1 != nullptr) {   // two separately released k ranges (the first a whole number of pipeline stages)
        issue(0, t.nka0 / kps, t.w0, t.t0, t.g0, t.gt0, true);
        issue(t.nka0 / kps, n_st, t.w1, t.t1, t.g1, t.gt1, false);
      } else {
        issue(0, n_st, t.w0, t.t0, t.g0, t.gt0, true);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int st = 0; st < n_st; ++st) {
        const uint32_t i = c.it + st, s = i & NSM;
        ps_mbar_wait(&c.full[s], (i >> NSH) & 1, c.dbg, t.code | (4u << 20));
        tc_fence_after();
        if (c.tr && st == 0) c.tr[3] = ps_now();
        const int n_sub = min(kps, nk - st * kps);
        for (int u = 0; u < n_sub; ++u) {
          const int kb = st * kps + u;
          const uint32_t a_addr = smem_u32(c.smem + s * t.stage_bytes + u * sub_bytes);
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(a_addr + t.a_bytes);
          if constexpr (GRU_U == 0) {
            const uint32_t idesc = umma_idesc_bf16(t.bn);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(c.tmem + t.tcol, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k | t.acc0) != 0);
          } else {
            // TMEM columns [r | z | n_x | n_h]: z k-blocks feed r, z, n_x in one N = 3U MMA; h k-blocks feed r, z (N = 2U) and n_h (N = U).
            // The h part comes first (h_j is there long before z_j): the very first z MMA must START n_x while ADDING to r, z, so it is
            // issued as an N = 2U and an N = U instruction.
            constexpr int U = GRU_U;
            const uint32_t idesc_rz = umma_idesc_bf16(2 * U), idesc_n = umma_idesc_bf16(U), idesc_3 = umma_idesc_bf16(3 * U);
            const uint64_t bdesc_n = umma_desc_sw128(a_addr + t.a_bytes + 2 * U * BK * 2);
            if (kb < t.nka0) {   // h part
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_bf16(c.tmem, adesc + 2 * k, bdesc + 2 * k, idesc_rz, (kb | k) != 0);
                umma_bf16(c.tmem + 3 * U, adesc + 2 * k, bdesc_n + 2 * k, idesc_n, (kb | k) != 0);
              }
            } else {             // z part
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if (kb == t.nka0 && k == 0) {
                  umma_bf16(c.tmem, adesc, bdesc, idesc_rz, 1u);
                  umma_bf16(c.tmem + 2 * U, adesc, bdesc_n, idesc_n, 0u);
                } else {
                  umma_bf16(c.tmem, adesc + 2 * k, bdesc + 2 * k, idesc_3, 1u);
                }
              }
            }
          }
        }
        umma_commit(&c.empty[s]);
        if (st == n_st - 1 && t.lsig) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;\n" ::"l"(t.lsig), "r"(1u) : "memory");   // every operand of the tile is in shared memory
      }
      umma_commit(c.tmem_full);
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    if (tid == 0) *reinterpret_cast<unsigned long long**>(c.smem + PS_EPI_OFF - 64) = c.tr ? c.tr + c.lap_off : nullptr;
    pre(tid);
    if (lane == 0) ps_flag_wait(t.e0, t.et0, c.dbg, t.code | (5u << 20));
    __syncwarp();
    epi_bar_sync();
    if (c.tr && tid == 0) c.tr[4] = ps_now();
    if (lane == 0) ps_mbar_wait(c.tmem_full, c.tile_no & 1u, c.dbg, t.code | (7u << 20));
    __syncwarp();
    tc_fence_after();
    if (c.tr && tid == 0) c.tr[5] = ps_now();
    epilogue(tid);
    if (c.tr && tid == 0) c.tr[6] = ps_now();
    if (t.sig) {
      // publish to other CTAs' waiters, WARP BY WARP: every epilogue warp releases its own stores (warp barrier, one device-scope fence
      // + release by lane 0) as soon as it has issued them -- no CTA barrier and no single thread fencing the whole tile's stores; a
      // tile counts PS_PUB (= 16 warps) on the hand-over counter
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        red_release_add(t.sig, 1u);
      }
    }
    post(tid);   // outputs nobody inside the kernel waits for
    if (!t.chain) fence_proxy_async_all();   // this thread's generic-proxy accesses to the ring / its global stores vs the next tile's TMA traffic
                                             // (chain tiles: the cluster hand-over's release / acquire + the consumer's proxy fence order them)
    if (c.tr && tid == 0) c.tr[7] = ps_now();
  }
  if (warp < 2) {   // mirror the cluster barriers of the epilogue (barrier.cluster counts every thread of the cluster)
    __syncwarp();
    for (int b = 0; b < t.cbar; ++b) { cluster_arrive_release(); cluster_wait_acquire(); }
  }
  c.it += (uint32_t)n_st;
  c.tile_no += 1;
  if (c.tr) c.tr += 8;
  // the next tile reuses the ring (the epilogue's transposition buffer) and the TMEM accumulator
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
template <int GRU_U, class Pre, class Epi>
__device__ __forceinline__ void ps_run_tile(PsCtx& c, const PsTile& t, Pre&& pre, Epi&& epilogue) {
  ps_run_tile<GRU_U>(c, t, pre, epilogue, PsNoPost{});
}

// hand-over between two layers of a chain cluster: every thread of the four CTAs; release / acquire at cluster scope orders the
// global stores of the finished layer before the TMA loads of the next one (plus the proxy fence in ps_run_tile)
__device__ __forceinline__ void ps_cluster_handover() {
  cluster_arrive_release();
  cluster_wait_acquire();
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) rollout_persist_kernel(const __grid_constant__ PersistParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PS_BAR_OFF);
  // barrier block (256 bytes = 32 slots): full[0, 8) | empty[8, 16) | tmem_full 16 | TMEM address 17 | xbar 18, 19 | abar 20 | lap pointer 24
  uint64_t* empty = full + PS_MAX_STAGES;
  uint64_t* tmem_full = empty + PS_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* epi_sm = reinterpret_cast<float*>(smem + PS_EPI_OFF);
  int* sched = reinterpret_cast<int*>(smem + PS_SCHED_OFF);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.tmS); tma_prefetch_desc(&P.tmY1); tma_prefetch_desc(&P.tmY2);
    for (int s = 0; s < PS_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    for (int i = 18; i < 21; ++i) mbar_init(&full[i], 1);   // xbar[0], xbar[1], abar
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < PS_SCHED_STRIDE; i += GEMM_THREADS) sched[i] = __ldg(P.sched + (long)blockIdx.x * PS_SCHED_STRIDE + i);
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // every CTA's mbarriers exist before a peer's st.async can target them
  tc_fence_after();

  PsCtx c;
  c.smem = smem; c.full = full; c.empty = empty; c.tmem_full = tmem_full; c.tmem = *tmem_slot; c.nsh = 2; c.it = 0; c.tile_no = 0; c.dbg = P.dbg; c.tr = nullptr; c.lap_off = (long)gridDim.x * PS_TRACE_SLOTS * 8;
  const int n_items = sched[0];
  const int role = n_items > 0 ? sched[1] : -1;
  const int B = P.B, H = P.H, D = P.D, ZP = P.ZP, A = P.A, R = P.R, Mp = P.Mp, mt = P.mt;
  const int nkz = ZP / 64, nkh = P.DP / 64, kh0 = nkz + 1;
  const long ldL = (long)(H + 1) * ZP, ldH = (long)(H + 1) * D, ldA = (long)H * A;
  const int q = warp & 3, part = (warp - 2) >> 2, lane = threadIdx.x & 31;
  const int row = q * 32 + lane;
  const uint32_t taddr = c.tmem + ((uint32_t)(q * 32) << 16);
  float* tile = reinterpret_cast<float*>(smem);
  auto flag = [&](int kind, int m) { return P.flags + (long)(kind * mt + m) * 32; };
  auto tile_init = [&](PsTile& t, int layer, int j, int m_tile) {
    t.stage_bytes = PS_STAGE_BYTES; t.a_bytes = A_STAGE_BYTES; t.a_tx = A_STAGE_BYTES; t.b_follows_a = 0; t.ka0 = 0; t.nka0 = 0; t.ka1 = 0; t.nka1 = 0; t.kps = 1; t.cbar = 0; t.chain = role == PS_CHAIN ? 1 : 0;
    t.w0 = nullptr; t.t0 = 0; t.g0 = nullptr; t.gt0 = 0; t.w1 = nullptr; t.t1 = 0; t.g1 = nullptr; t.gt1 = 0; t.lsig = nullptr;
    t.h_first = 0; t.tmB2 = nullptr; t.n1 = 0; t.b2_row = 0; t.b2_koff = 0; t.tcol = 0; t.acc0 = 0; t.e0 = nullptr; t.et0 = 0; t.sig = nullptr;
    t.code = ((unsigned)role << 24) | ((unsigned)layer << 16) | ((unsigned)j << 8) | (unsigned)m_tile;
  };
  auto trace_window = [&](int j) {   // (at most PS_TRACE_SLOTS tiles per CTA are recorded)
    if (P.trace) {
      if (j == P.trace_j0) c.tr = P.trace + (long)blockIdx.x * PS_TRACE_SLOTS * 8;
      if (j == P.trace_j1) c.tr = nullptr;
    }
  };
  // the [h, z] heads' output layer (EpiHeads) is parameterised once; only the per-state output pointers move
  EpiHeads::Params hp;
  memset(&hp, 0, sizeof(hp));
  hp.bias = P.h3_b;
  hp.kind[HS_REWARD] = HEAD_BUCKET; hp.kind[HS_CONT] = HEAD_SIGMOID; hp.kind[HS_ACTOR] = HEAD_ACTOR;
  hp.buckets[HS_REWARD] = P.bk_rew;
  hp.NB = P.NB; hp.A = A;
  hp.ld_normals = A; hp.ld_act = ldA; hp.ld_value[HS_REWARD] = H; hp.ld_value[HS_CONT] = H;

  if (role == PS_CHAIN) {
    // ---- one 4-CTA cluster per m-tile: prior L1 -> L2 -> logits + sample -> actor L1 -> L2 -> output, for every state ----
    const int m_tile = sched[2];
    const int rank = (int)cluster_ctarank();
    const int m0 = m_tile * BM, m = m0 + row;
    const int ya = (1 + HS_ACTOR) * Mp;   // the actor's slot of the hidden-activation buffers (slot 0 = prior)
    PsXchg xg;
    xg.xst = reinterpret_cast<float*>(smem + PS_RING_BYTES);                 // (the 16 KB behind the ring, unused by chain CTAs otherwise)
    xg.xbar = full + 18; xg.abar = full + 20;
    xg.xact = reinterpret_cast<float*>(smem + PS_RING_BYTES - 16384);        // rank 0: last 16 KB of the ring (idle while the epilogue runs)
    xg.xuse = 0; xg.ause = 0; xg.dbg = P.dbg;
    c.nsh = 1;                                    // two ring slots of up to 96 KB (PS_CHAIN_STAGE_BYTES)
    for (int j = 0; j <= H; ++j) {
      trace_window(j);
      const int s_row = j * B + m0;       // this m-tile's rows of state j in the time-major state buffer
      PsTile t;
      if (j >= 1) {
        {   // prior L1: h_j -> Y1[slot 0], this CTA's 64 columns   (DynamicsPredictors.py:15-18)
          tile_init(t, 0, j, m_tile);
          t.tmA = &P.tmS; t.tmB = &P.tmWp1q; t.a_row = s_row; t.b_row = 64 * rank; t.ka0 = kh0; t.nka0 = nkh; t.bn = 64; t.kps = 4;   // (4 x 24 KB)
          t.stage_bytes = PS_CHAIN_STAGE_BYTES;
          t.tcol = 256;
          if (j < H) {
            // the actor's first layer reads the same h k-blocks: its 64 weight rows ride along under the prior's (one N = 128 MMA per
            // k-step), so the h part of actor L1 is already in TMEM columns [320, 384) when z_j arrives -- 10 of its 26 k-blocks
            // leave the critical path for 8 KB more per k-block here
            t.tmB2 = &P.tmWh1q; t.n1 = 64; t.b2_row = HS_ACTOR * 256 + 64 * rank; t.b2_koff = nkz; t.bn = 128; t.kps = 3;   // (3 x 32 KB)
          }
          t.w0 = flag(PF_H, m_tile); t.t0 = (unsigned)(PS_PUB * P.nt * j);
          t.lsig = flag(PF_LP, m_tile);
          const TileG g{B, 64, 0};
          const EpiLnSilu::Params p{P.p1_b, P.p1_g, P.p1_be, nullptr, 0, P.Y1, 256, 0, 0, P.hp1, 1e-5f, P.bnp1};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSiluN4::stage(p, g, 0, epi_sm, tid, m0); },
                         [&](int tid) {
                           float v[16];
                           ps_ln_compute(p, g, epi_sm, xg, taddr + 256, m, row, part, 0, tid, t.code, v);
                           EpiLnSiluN4::store(p, g, tile, m, row, part, 0, tid, v);
                         });
          ps_cluster_handover();
        }
        {   // prior L2: Y1 -> Y2[slot 0]   (:19-22)
          tile_init(t, 1, j, m_tile);
          t.tmA = &P.tmY1; t.tmB = &P.tmWp2q; t.a_row = m0; t.b_row = 64 * rank; t.ka0 = 0; t.nka0 = (P.hp1 + 63) / 64; t.bn = 64; t.kps = 4; t.stage_bytes = PS_CHAIN_STAGE_BYTES;   // all 4 k-blocks in one stage
          const TileG g{B, 64, 0};
          const EpiLnSilu::Params p{P.p2_b, P.p2_g, P.p2_be, nullptr, 0, P.Y2, 256, 0, 0, P.hp2, 1e-5f, P.bnp2};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSiluN4::stage(p, g, 0, epi_sm, tid, m0); },
                         [&](int tid) {
                           float v[16];
                           ps_ln_compute(p, g, epi_sm, xg, taddr, m, row, part, 0, tid, t.code, v);
                           EpiLnSiluN4::store(p, g, tile, m, row, part, 0, tid, v);
                         });
          ps_cluster_handover();
        }
        // prior logits + sample: 256 logit columns (8 latent rows) per tile, tiles rank, rank + 4, ...   (:23, 31-40)
        // -> z_j into the state buffer, latent[:, j], idx[:, j - 1]
        for (int x = rank; x < P.nq; x += 4) {
          tile_init(t, 2, j, m_tile);
          t.tmA = &P.tmY2; t.tmB = &P.tmWp3; t.a_row = m0; t.b_row = x * 256; t.ka0 = 0; t.nka0 = (P.hp2 + 63) / 64; t.bn = 256; t.kps = 2; t.stage_bytes = PS_CHAIN_STAGE_BYTES;   // 2 x 48 KB
          t.sig = flag(PF_Z, m_tile);
          const TileG g{B, 256, 0};
          const EpiCat::Params p{P.p3_b, P.uniforms + (long)(j - 1) * B * R, P.latent + (long)j * ZP, nullptr,
                                 P.idx ? P.idx + (long)(j - 1) * R : nullptr, P.S + (long)j * B * P.KS, nullptr, ldL, 0, (long)H * R, 0, P.KS, R,
                                 RowMap{0, 0, 0, 0}};
          ps_run_tile<0>(c, t, [&](int tid) { EpiCat::stage(p, g, x, epi_sm, tid, m0); EpiCatP::stage_prev(p, P.idx_prev + (long)j * B * R, g, epi_sm, x, tid, m0); },
                         [&](int tid) { EpiCatP::run(p, P.idx_prev + (long)j * B * R, g, epi_sm, taddr, m, row, part, x, tid); });
        }
        ps_cluster_handover();
      }
      if (j < H) {
        {   // actor L1: [z_j | h_j] -> Y1[actor slot]   (Agent.py:178-181)
          tile_init(t, 0, j, m_tile);
          t.code |= 1u << 19;
          t.tmA = &P.tmS; t.tmB = &P.tmWh1q; t.a_row = s_row; t.b_row = HS_ACTOR * 256 + 64 * rank;
          t.ka0 = 0; t.nka0 = nkz; t.ka1 = kh0; t.nka1 = nkh; t.bn = 64; t.kps = 4; t.stage_bytes = PS_CHAIN_STAGE_BYTES;
          if (j >= 1) {   // the h part is already accumulated (prior L1 above): z k-blocks only
            t.nka1 = 0; t.tcol = 320; t.acc0 = 1;
          }
          t.lsig = flag(PF_LA, m_tile);
          const TileG g{B, 64, 0};
          const EpiLnSilu::Params p{P.h1_b, P.h1_g, P.h1_be, nullptr, 0, P.Y1, 256, Mp, Mp, P.hh1, 1e-5f, P.bnh1};
          const uint32_t ta = taddr + (uint32_t)t.tcol;
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSiluN4::stage(p, g, HS_ACTOR, epi_sm, tid, m0); },
                         [&](int tid) {
                           float v[16];
                           ps_ln_compute(p, g, epi_sm, xg, ta, m, row, part, HS_ACTOR, tid, t.code, v);
                           EpiLnSiluN4::store(p, g, tile, m, row, part, HS_ACTOR, tid, v);
                         });
          ps_cluster_handover();
        }
        {   // actor L2 + output layer -> a_j = tanh(mu + sigma * eps)   (Agent.py:182-187, 199-209)
          tile_init(t, 1, j, m_tile);
          t.code |= 1u << 19;
          t.tmA = &P.tmY1; t.tmB = &P.tmWh2q; t.a_row = ya + m0; t.b_row = HS_ACTOR * 256 + 64 * rank;
          t.ka0 = 0; t.nka0 = (P.hh1 + 63) / 64; t.bn = 64; t.kps = 4; t.stage_bytes = PS_CHAIN_STAGE_BYTES;
          t.chain = 0;   // (no hand-over follows, and its epilogue used the ring as scratch: keep the proxy fence before the next tile's TMA loads)
          const TileG g{B, 64, 0};
          const EpiLnSilu::Params p{P.h2_b, P.h2_g, P.h2_be, nullptr, 0, P.Y2, 256, Mp, Mp, P.hh2, 1e-5f, P.bnh2};
          const PsActorOut ao{P.Wh3 + (long)HS_ACTOR * 256 * 256, P.h3_b + HS_ACTOR * 256, P.normals + (long)j * B * A,
                              P.mu + (long)j * A, P.sigma + (long)j * A, P.actions + (long)j * A, P.apack + (long)j * Mp, ldA, A, B};
          float* red = reinterpret_cast<float*>(smem);
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSiluN4::stage(p, g, HS_ACTOR, epi_sm, tid, m0); ao.stage(epi_sm, tid); },
                         [&](int tid) {
                           float v[16];
                           ps_ln_compute(p, g, epi_sm, xg, taddr, m, row, part, HS_ACTOR, tid, t.code, v);
                           ao.run(v, epi_sm, red, xg, m, row, part, tid, t.code);
                         });
        }
      }
    }
    ps_cluster_handover();   // nobody leaves while a peer may still write into its shared memory
  } else if (role == PS_GRU) {
    // ---- h_{j+1} = GRU([z_j, a_j], h_j)   (SequenceModel.py:19-24): one (m-tile, n-tile) per item ----
    float4* wa = reinterpret_cast<float4*>(smem + PS_WA_OFF);
    float* hp_tile = reinterpret_cast<float*>(smem + PS_HP_OFF);
    for (int j = 0; j < H; ++j) {
      trace_window(j);
#pragma unroll 1
      for (int it = 0; it < n_items; ++it) {
        const int m_tile = sched[2 + 3 * it], n_tile = sched[3 + 3 * it];
        const int m0 = m_tile * BM, m = m0 + row;
        PsTile t;
        tile_init(t, 0, j, m_tile);
        t.stage_bytes = PS_GRU_STAGE_BYTES;
        t.tmA = &P.tmS; t.tmB = &P.tmWgru; t.a_row = j * B + m0; t.b_row = n_tile * 3 * P.U;
        t.ka0 = kh0; t.nka0 = nkh; t.ka1 = 0; t.nka1 = nkz; t.b_follows_a = 1; t.bn = 3 * P.U; t.h_first = 1;
        // The GRU CTAs move 83 MB per step -- alone they would saturate the L2 -> SM fabric for 6 us and starve the chain's small
        // loads exactly when those are on the critical path.  So each half waits until the chain layer that reads the same data has
        // its operands: the h part (needs h_j) behind prior L1 of state j, the z part (needs z_j) behind actor L1 of state j.
        t.w0 = flag(PF_H, m_tile); t.t0 = (unsigned)(PS_PUB * P.nt * j);
        t.g0 = flag(PF_LP, m_tile); t.gt0 = (unsigned)(4 * j);
        t.w1 = flag(PF_Z, m_tile); t.t1 = (unsigned)(PS_PUB * P.nq * j);
        t.g1 = flag(PF_LA, m_tile); t.gt1 = (unsigned)(4 * (j + 1));
        t.sig = flag(PF_H, m_tile);     // (a_j is awaited inside the epilogue, row by row: EpiGruP::run)
        __nv_bfloat16* s_h = P.S + (long)(j + 1) * B * P.KS + ZP + 64;
        if (P.U == 32) {
          const EpiGruP<32>::Params p{P.b_ih, P.b_hh, P.hidden + (long)j * D, P.hidden + (long)(j + 1) * D, s_h, ldH, P.KS, D,
                                      P.Wgru + ZP, P.KS, A, P.apack + (long)j * Mp, P.dbg, t.code};
          ps_run_tile<32>(c, t, [&](int tid) { EpiGruP<32>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<32>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); },
                          [&](int tid) { EpiGruP<32>::post(p, n_tile, B, tile, m0, tid); });
        } else {
          const EpiGruP<64>::Params p{P.b_ih, P.b_hh, P.hidden + (long)j * D, P.hidden + (long)(j + 1) * D, s_h, ldH, P.KS, D,
                                      P.Wgru + ZP, P.KS, A, P.apack + (long)j * Mp, P.dbg, t.code};
          ps_run_tile<64>(c, t, [&](int tid) { EpiGruP<64>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<64>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); },
                          [&](int tid) { EpiGruP<64>::post(p, n_tile, B, tile, m0, tid); });
        }
      }
    }
  } else if (role == PS_RC) {
    // ---- reward / continue heads of the sampled states 1 .. H (DynamicsPredictors.py:64-74, 95-105): nothing waits for these ----
    for (int j = 1; j <= H; ++j) {
      trace_window(j);
#pragma unroll 1
      for (int it = 0; it < n_items; ++it) {
        const int m_tile = sched[2 + 3 * it], head = sched[3 + 3 * it];
        const int m0 = m_tile * BM, m = m0 + row;
        const int yr = (1 + head) * Mp + m0;
        PsTile t;
        {   // L1: [z_j | h_j] -> Y1[slot 1 + head]
          tile_init(t, 0, j, m_tile);
          t.tmA = &P.tmS; t.tmB = &P.tmWh1; t.a_row = j * B + m0; t.b_row = head * 256;
          t.ka0 = 0; t.nka0 = nkz; t.ka1 = kh0; t.nka1 = nkh; t.bn = P.bnh1;
          t.w0 = flag(PF_Z, m_tile); t.t0 = (unsigned)(PS_PUB * P.nq * j);
          const TileG g{B, P.bnh1, 0};
          const EpiLnSilu::Params p{P.h1_b, P.h1_g, P.h1_be, nullptr, 0, P.Y1, 256, Mp, Mp, P.hh1, 1e-5f, P.bnh1};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
        {   // L2: Y1 -> Y2 (same CTA: program order + the fences at the end of a tile order the hand-over)
          tile_init(t, 1, j, m_tile);
          t.tmA = &P.tmY1; t.tmB = &P.tmWh2; t.a_row = yr; t.b_row = head * 256; t.ka0 = 0; t.nka0 = (P.hh1 + 63) / 64; t.bn = P.bnh2;
          const TileG g{B, P.bnh2, 0};
          const EpiLnSilu::Params p{P.h2_b, P.h2_g, P.h2_be, nullptr, 0, P.Y2, 256, Mp, Mp, P.hh2, 1e-5f, P.bnh2};
          ps_run_tile<0>(c, t, [&](int tid) { EpiLnSilu::stage(p, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiLnSilu::run(p, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
        {   // output layer: bucket softmax -> symexp (reward), sigmoid (continue)
          tile_init(t, 2, j, m_tile);
          t.tmA = &P.tmY2; t.tmB = &P.tmWh3; t.a_row = yr; t.b_row = head * 256; t.ka0 = 0; t.nka0 = (P.hh2 + 63) / 64; t.bn = 256;
          const TileG g{B, 256, 0};
          hp.value[HS_REWARD] = P.rewards + (j - 1);
          hp.value[HS_CONT] = P.continues + (j - 1);
          ps_run_tile<0>(c, t, [&](int tid) { EpiHeads::stage(hp, g, head, epi_sm, tid, m0); },
                         [&](int tid) { EpiHeads::run(hp, g, epi_sm, tile, taddr, m, row, part, head, tid); });
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(c.tmem, 512);
}

}  // namespace drm

// ------------------------------------------------------------------------------------------
// host side: static schedule, launch
// ------------------------------------------------------------------------------------------
struct drm_persist {
  int U = 0, nt = 0, nq = 0, n_cta = 0;
  int* sched = nullptr;        // device [n_cta * PS_SCHED_STRIDE]
  unsigned* flags = nullptr;   // device [PF_COUNT * mt * 32] counters, then [H * Mp] 8-byte action records (zeroed together before every launch)
  uint8_t* idx_prev = nullptr; // device [(H + 1) * B * R]
  unsigned* dbg = nullptr;     // host-mapped [PS_DBG_WORDS]
  size_t flag_bytes = 0;
  __nv_bfloat16* S = nullptr;  // time-major state slabs [(H + 1) * B + 128, KS]
  CUtensorMap tmS;
  unsigned long long* trace = nullptr;   // device [n_cta * PS_TRACE_SLOTS * 8] while tracing
  int trace_j0 = 0, trace_j1 = 0;
  bool tried = false, ok = false;
};

namespace drm {

static void persist_free(drm_persist* ps) {
  if (!ps) return;
  if (ps->dbg) cudaFreeHost(ps->dbg);
  delete ps;
}

static void persist_launch_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int n_cta, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(n_cta);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = PS_TOTAL;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

// clusters of 4 CTAs of this kernel the device can hold at once (0: the kernel cannot run here)
static int persist_cluster_capacity() {
  static int n = -1;
  if (n < 0) {
    n = 0;
    if (cudaFuncSetAttribute(rollout_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_TOTAL) == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[1];
      int sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      persist_launch_config(cfg, attr, (sms / 4) * 4, nullptr);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, rollout_persist_kernel, &cfg) == cudaSuccess) n = nc;
    }
    cudaGetLastError();
  }
  return n;
}

// Static schedule.  Cluster c < mt: the MLP chain of m-tile c.  Then the GRU tiles, one per CTA.  Every CTA left over takes
// reward / continue chains (m-tile, head) round-robin.
static bool persist_plan(drm_rollout* r, drm_persist* ps) {
  drm_rssm* m = r->m;
  const int cap = persist_cluster_capacity();
  const int mt = r->Mp / BM;
  if (cap <= 0 || m->d.A > 3 || m->ZP % 256) return false;
  const size_t s_rows = (size_t)(r->H + 1) * r->B + BM;
  if (s_rows * m->KS * sizeof(__nv_bfloat16) > ((size_t)2 << 30)) return false;
  for (int U = 32; U <= 64; U *= 2) {
    const int nt = ceil_div(m->d.D, U);
    const int gru_cl = ceil_div(mt * nt, 4);
    const int spare = (cap - mt - gru_cl) * 4 + (gru_cl * 4 - mt * nt);   // CTAs left for the reward / continue chains
    if (cap < mt + gru_cl || spare < 1 || ceil_div(2 * mt, spare) > PS_MAX_ITEMS) continue;
    ps->U = U; ps->nt = nt; ps->nq = m->ZP / 256;
    const int n_rc = std::min(spare, 2 * mt);
    const int n_cta = round_up(4 * mt + mt * nt + n_rc, 4);
    std::vector<int> sc((size_t)n_cta * PS_SCHED_STRIDE, 0);
    auto add = [&](int c, int kind, int mm, int x) {
      int* rec = sc.data() + (size_t)c * PS_SCHED_STRIDE;
      const int n = rec[0]++;
      rec[1 + 3 * n] = kind; rec[2 + 3 * n] = mm; rec[3 + 3 * n] = x;
    };
    for (int mm = 0; mm < mt; ++mm)
      for (int k = 0; k < 4; ++k) add(4 * mm + k, PS_CHAIN, mm, 0);
    int cta = 4 * mt;
    for (int n = 0; n < nt; ++n)          // n-major: the GRU tiles of one m-tile are spread over the chip's clusters
      for (int mm = 0; mm < mt; ++mm) add(cta++, PS_GRU, mm, n);
    for (int mm = 0, k = 0; mm < mt; ++mm)
      for (int hd = 0; hd < 2; ++hd, ++k) add(cta + k % n_rc, PS_RC, mm, hd == 0 ? HS_REWARD : HS_CONT);
    ps->n_cta = n_cta;
    if (dev_alloc(r->allocs, &ps->sched, sc.size()) != DRM_OK) return false;
    if (cudaMemcpy(ps->sched, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    ps->flag_bytes = (size_t)PF_COUNT * mt * 32 * sizeof(unsigned) + (size_t)r->H * r->Mp * sizeof(unsigned long long);
    if (dev_alloc(r->allocs, &ps->flags, ps->flag_bytes / sizeof(unsigned)) != DRM_OK) return false;
    const size_t n_prev = (size_t)(r->H + 1) * r->B * m->d.R;
    if (dev_alloc(r->allocs, &ps->idx_prev, n_prev) != DRM_OK) return false;
    if (cudaMemset(ps->idx_prev, 255, n_prev) != cudaSuccess) return false;
    if (dev_alloc(r->allocs, &ps->S, s_rows * m->KS) != DRM_OK) return false;
    if (make_tmap_bf16_2d(&ps->tmS, ps->S, s_rows, m->KS, m->KS, BM) != DRM_OK) return false;
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
  DRM_CUDA(cudaMemsetAsync(latent, 0, (size_t)B * ldL * sizeof(float), st));   // the sampling epilogue stores only the sampled class of every latent row
  RC(pack_cols(ps->S, m->KS, 0, z0, ZP, ZP, B, latent, ldL, st));
  RC(pack_cols(ps->S, m->KS, ZP + 64, h0, D, D, B, hidden, ldH, st));
  DRM_CUDA(cudaMemsetAsync(ps->flags, 0, ps->flag_bytes, st));
  PersistParams P;
  memset(&P, 0, sizeof(P));
  const int v = ps->U == 64 ? 1 : 0;
  P.tmS = ps->tmS; P.tmY1 = r->tmY1; P.tmY2 = r->tmY2;
  P.tmWgru = m->tmWgru2[v]; P.tmWp1q = m->tmWp1q; P.tmWp2q = m->tmWp2q; P.tmWp3 = m->tmWp3;
  P.tmWh1q = m->tmWh1q; P.tmWh2q = m->tmWh2q; P.tmWh3a = m->tmWh3a; P.tmWh1 = m->tmWh1; P.tmWh2 = m->tmWh2; P.tmWh3 = m->tmWh3;
  P.B = B; P.H = H; P.D = D; P.DP = m->DP; P.ZP = ZP; P.R = m->d.R; P.A = m->d.A; P.NB = m->d.NB; P.KS = m->KS; P.Mp = r->Mp; P.mt = r->Mp / BM;
  P.U = ps->U; P.nt = ps->nt; P.nq = ps->nq;
  P.bnp1 = m->bnp1; P.bnp2 = m->bnp2; P.bnh1 = m->bnh1; P.bnh2 = m->bnh2;
  P.hp1 = m->d.h_prior[0]; P.hp2 = m->d.h_prior[1]; P.hh1 = m->d.h_head[0]; P.hh2 = m->d.h_head[1];
  P.b_ih = m->b_ih; P.b_hh = m->b_hh;
  P.p1_b = m->p1_b; P.p1_g = m->p1_g; P.p1_be = m->p1_be; P.p2_b = m->p2_b; P.p2_g = m->p2_g; P.p2_be = m->p2_be; P.p3_b = m->p3_b;
  P.h1_b = m->h1_b; P.h1_g = m->h1_g; P.h1_be = m->h1_be; P.h2_b = m->h2_b; P.h2_g = m->h2_g; P.h2_be = m->h2_be; P.h3_b = m->h3_b;
  P.bk_rew = m->bk_rew;
  P.Wgru = m->Wgru2[v];
  P.Wh3 = m->Wh3;
  P.S = ps->S; P.Y1 = r->Y1; P.Y2 = r->Y2;
  P.uniforms = uniforms; P.normals = normals;
  P.latent = latent; P.hidden = hidden; P.actions = actions; P.rewards = rewards; P.continues = continues; P.mu = mu; P.sigma = sigma;
  P.idx = idx;
  P.flags = ps->flags; P.dbg = ps->dbg; P.sched = ps->sched;
  P.apack = reinterpret_cast<unsigned long long*>(ps->flags + (size_t)PF_COUNT * P.mt * 32);
  P.idx_prev = ps->idx_prev;
  P.trace = ps->trace; P.trace_j0 = ps->trace_j0; P.trace_j1 = ps->trace_j1;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  persist_launch_config(cfg, attr, ps->n_cta, st);   // (co-residency of all clusters was checked against the device's capacity in persist_plan)
  profile_begin(DRM_STAGE_ROLLOUT, st);
  DRM_CUDA(cudaLaunchKernelEx(&cfg, rollout_persist_kernel, P));
  profile_end(DRM_STAGE_ROLLOUT, st);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

}  // namespace drm
