// The posterior (observe) scan of WorldModel.unroll_model / Dreamer.warm_start_generator (WorldModel.py:92-107, Dreamer.py:252-261)
// as ONE persistent kernel: the time loop, the GRU step and the posterior head (latent_mapper MLP on [features, h_t] -> logits ->
// sample) of every step run inside a single launch, on the building blocks of rollout_persist.cuh.  Included from rssm.cu after
// vae.cuh's declarations.
//
// Launched stage by stage, a scan step is three dependent launches (GRU, Linear + LayerNorm + SiLU, logits + sample) of 9 - 17 us
// each at 16 sequences -- 37 us per step, 2.4 ms for the 64 steps of BASELINE configs[2], all of it launch / prologue / cold-pipeline
// latency.  Here the CTAs stay resident:
//   "chain" clusters (one 4-CTA cluster per m-tile of 128 sequences): posterior L1 on h_t (+ the hoisted feature part, prefetched
//       into registers under the main loop) with the LayerNorm columns split four ways, then the logits + sample tiles;
//   "GRU" CTAs, one (m-tile, n-tile) each for the whole sequence: the h part of step t + 1 is contracted as soon as h_t exists,
//       the [z_t | a_t] part (the actions are data here, so they ride in the MMA) as soon as z_t is sampled.
// State: the observe workspace's time-major slabs (slab 0 = zero state, slab t + 1 = (z_t | a_t | h_t)); hand-overs are the same
// per-m-tile release / acquire counters as in the rollout kernel.  With <= 32 sequences the A operand is loaded through the
// short-box tensor maps (4 KB instead of 16 KB per k-block; the tile's tail rows are stale and never stored).
#pragma once

namespace drm {

struct ObsPersistParams {
  CUtensorMap tmS, tmY1, tmWgru, tmWehq, tmWe3;
  int B, T, D, DP, ZP, R, KS, mt;
  int U, nt, nq;          // GRU tile width, GRU n-tiles, 256-column sampling tiles
  int h_enc, bn_he;       // posterior hidden width and its padded pitch
  int a_tx;               // bytes of one A load (short-box maps for small batches)
  int h_skip;             // 1: warm start (h_0 = 0, no GRU step before the first posterior), else 0
  const float *b_ih, *b_hh, *e1_b, *e1_g, *e1_be, *e3_b;
  const __nv_bfloat16* Wgru;
  __nv_bfloat16 *S, *Y1;
  const float *featpart, *uniforms, *zero_h;
  float *latent, *hidden, *logits;
  uint8_t *idx, *idx_prev;
  unsigned *flags, *dbg;
  const int* sched;
  unsigned long long* trace;   // debug: [cta][PS_TRACE_SLOTS][8] timestamps of the tiles of steps [trace_j0, trace_j1), or NULL
  int trace_j0, trace_j1;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1) observe_persist_kernel(const __grid_constant__ ObsPersistParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PS_BAR_OFF);
  uint64_t* empty = full + PS_MAX_STAGES;      // (barrier block layout: rollout_persist_kernel)
  uint64_t* tmem_full = empty + PS_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* epi_sm = reinterpret_cast<float*>(smem + PS_EPI_OFF);
  int* sched = reinterpret_cast<int*>(smem + PS_SCHED_OFF);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&P.tmS); tma_prefetch_desc(&P.tmY1);
    for (int s = 0; s < PS_MAX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    for (int i = 18; i < 21; ++i) mbar_init(&full[i], 1);   // xbar[0], xbar[1], (abar)
    mbar_fence_init();
    *reinterpret_cast<unsigned long long**>(smem + PS_EPI_OFF - 64) = nullptr;   // no lap records
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
  const int B = P.B, T = P.T, D = P.D, ZP = P.ZP, R = P.R, mt = P.mt;
  const int nkz = ZP / 64, nkh = P.DP / 64, kh0 = nkz + 1;
  const long ldL = (long)T * ZP, ldH = (long)T * D;
  const int q = warp & 3, part = (warp - 2) >> 2, lane = threadIdx.x & 31;
  const int row = q * 32 + lane;
  const uint32_t taddr = c.tmem + ((uint32_t)(q * 32) << 16);
  float* tile = reinterpret_cast<float*>(smem);
  auto flag = [&](int kind, int m) { return P.flags + (long)(kind * mt + m) * 32; };
  auto trace_window = [&](int j) {
    if (P.trace) {
      if (j == P.trace_j0) c.tr = P.trace + (long)blockIdx.x * PS_TRACE_SLOTS * 8;
      if (j == P.trace_j1) c.tr = nullptr;
    }
  };
  auto tile_init = [&](PsTile& t, int layer, int j, int m_tile) {
    t.stage_bytes = PS_STAGE_BYTES; t.a_bytes = A_STAGE_BYTES; t.a_tx = P.a_tx; t.b_follows_a = 0; t.ka0 = 0; t.nka0 = 0; t.ka1 = 0; t.nka1 = 0; t.kps = 1; t.cbar = 0; t.chain = role == PS_CHAIN ? 1 : 0;
    t.w0 = nullptr; t.t0 = 0; t.g0 = nullptr; t.gt0 = 0; t.w1 = nullptr; t.t1 = 0; t.g1 = nullptr; t.gt1 = 0; t.lsig = nullptr;
    t.h_first = 0; t.tmB2 = nullptr; t.n1 = 0; t.b2_row = 0; t.b2_koff = 0; t.tcol = 0; t.acc0 = 0; t.e0 = nullptr; t.et0 = 0; t.sig = nullptr;
    t.code = ((unsigned)(role + 4) << 24) | ((unsigned)layer << 16) | ((unsigned)(j & 255) << 8) | (unsigned)m_tile;
  };

  if (role == PS_CHAIN) {
    // ---- posterior head of every step: slab j = t + 1 holds h_t (written by the GRU CTAs) and receives z_t ----
    const int m_tile = sched[2];
    const int rank = (int)cluster_ctarank();
    const int m0 = m_tile * BM, m = m0 + row;
    PsXchg xg;
    xg.xst = reinterpret_cast<float*>(smem + PS_RING_BYTES);
    xg.xbar = full + 18; xg.abar = full + 20;
    xg.xact = nullptr;
    xg.xuse = 0; xg.ause = 0; xg.dbg = P.dbg;
    c.nsh = 1;   // two ring slots of 96 KB, 2 - 4 k-blocks per stage (rollout_persist.cuh: a handshake costs ~0.3 us whatever the depth)
    for (int j = 1; j <= T; ++j) {
      trace_window(j);
      const int t_ = j - 1;
      const int s_row = j * B + m0;
      PsTile t;
      {   // latent_mapper.0 on [features, h_t] (VariationalAutoEncoder.py:45-48, 84-86): the feature part was hoisted (featpart), the h part here
        tile_init(t, 0, j, m_tile);
        t.tmA = &P.tmS; t.tmB = &P.tmWehq; t.a_row = s_row; t.b_row = 64 * rank; t.ka0 = kh0; t.nka0 = nkh; t.bn = 64; t.kps = 4; t.stage_bytes = PS_CHAIN_STAGE_BYTES;
        t.w0 = flag(PF_H, m_tile); t.t0 = (unsigned)(PS_PUB * P.nt * (j - P.h_skip));
        const TileG g{B, 64, 0};
        const EpiLnSilu::Params p{P.e1_b, P.e1_g, P.e1_be, nullptr, 0, P.Y1, 256, 0, 0, P.h_enc, 1e-5f, P.bn_he};
        float addv[16];
        ps_run_tile<0>(c, t,
                       [&](int tid) {
                         EpiLnSiluN4::stage(p, g, 0, epi_sm, tid, m0);
                         const int gc0 = 64 * rank + part * 16;
#pragma unroll
                         for (int k = 0; k < 16; ++k) addv[k] = 0.f;
                         if (m < B) {
                           const float* add = P.featpart + ((long)t_ * B + m) * P.bn_he + gc0;
#pragma unroll
                           for (int k = 0; k < 16; ++k)
                             if (gc0 + k < P.h_enc) addv[k] = __ldg(add + k);
                         }
                       },
                       [&](int tid) {
                         float v[16];
                         ps_ln_compute(p, g, epi_sm, xg, taddr, m, row, part, 0, tid, t.code, v, addv);
                         EpiLnSiluN4::store(p, g, tile, m, row, part, 0, tid, v);
                       });
        ps_cluster_handover();
      }
      // latent_mapper.3 -> posterior logits -> sample (VariationalAutoEncoder.py:49, 88-99): 256 logit columns per tile, tiles rank, rank + 4, ...
      for (int x = rank; x < P.nq; x += 4) {
        tile_init(t, 2, j, m_tile);
        t.tmA = &P.tmY1; t.tmB = &P.tmWe3; t.a_row = m0; t.b_row = x * 256; t.ka0 = 0; t.nka0 = (P.h_enc + 63) / 64; t.bn = 256; t.kps = 2; t.stage_bytes = PS_CHAIN_STAGE_BYTES;
        t.sig = flag(PF_Z, m_tile);
        const TileG g{B, 256, 0};
        const EpiCat::Params p{P.e3_b, P.uniforms + (long)t_ * B * R, P.latent + (long)t_ * ZP, P.logits ? P.logits + (long)t_ * ZP : nullptr,
                               P.idx ? P.idx + (long)t_ * R : nullptr, P.S + (long)j * B * P.KS, nullptr, ldL, ldL, (long)T * R, 0, P.KS, R,
                               RowMap{0, 0, 0, 0}};
        ps_run_tile<0>(c, t, [&](int tid) { EpiCat::stage(p, g, x, epi_sm, tid, m0); EpiCatP::stage_prev(p, P.idx_prev + (long)j * B * R, g, epi_sm, x, tid, m0); },
                       [&](int tid) { EpiCatP::run(p, P.idx_prev + (long)j * B * R, g, epi_sm, taddr, m, row, part, x, tid); });
      }
      ps_cluster_handover();   // Y1 is rewritten by the next step's first layer: every rank's loads of it are done
    }
    ps_cluster_handover();   // nobody leaves while a peer may still write into its shared memory
  } else if (role == PS_GRU) {
    // ---- h_t = GRU([z_{t-1}, a_{t-1}], h_{t-1})   (SequenceModel.py:19-24; WorldModel.py:97-99): reads slab t, writes the h columns of slab t + 1 ----
    float4* wa = reinterpret_cast<float4*>(smem + PS_WA_OFF);
    float* hp_tile = reinterpret_cast<float*>(smem + PS_HP_OFF);
    // a step's 27 k-blocks are latency-bound at 16 - 50 sequences: compact stages ([A 4 or 16 KB | 3U weight rows]) and a ring as deep
    // as fits in front of the h_prev tile
    const int gru_sub = P.a_tx + 3 * P.U * BK * 2;
    // k-blocks per handshake: as many as divide the h range (it must be whole stages) with two stages still fitting in front of the
    // h_prev tile -- measured at 16 x 64: 1 per stage / 8 slots 2.23 ms per scan() call, 2 / 8 2.08, 5 / 2 2.01, 5 / 1 2.08
    int gru_kps = 1;
    for (int k = 5; k >= 2; --k)
      if (nkh % k == 0 && 2 * k * gru_sub <= PS_HP_OFF) { gru_kps = k; break; }
    const int gru_stage = gru_kps * gru_sub;
    c.nsh = 0;
    while (c.nsh < 3 && (2 << c.nsh) * gru_stage <= PS_HP_OFF) ++c.nsh;
    for (int s = P.h_skip; s < T; ++s) {
      trace_window(s);
#pragma unroll 1
      for (int it = 0; it < n_items; ++it) {
        const int m_tile = sched[2 + 3 * it], n_tile = sched[3 + 3 * it];
        const int m0 = m_tile * BM, m = m0 + row;
        PsTile t;
        tile_init(t, 0, s, m_tile);
        t.stage_bytes = gru_stage; t.a_bytes = P.a_tx; t.kps = gru_kps;
        t.tmA = &P.tmS; t.tmB = &P.tmWgru; t.a_row = s * B + m0; t.b_row = n_tile * 3 * P.U;
        t.ka0 = kh0; t.nka0 = nkh; t.ka1 = 0; t.nka1 = nkz + 1; t.b_follows_a = 1; t.bn = 3 * P.U; t.h_first = 1;
        t.w0 = flag(PF_H, m_tile); t.t0 = (unsigned)(PS_PUB * P.nt * (s - P.h_skip));   // h_{s-1} (slab s) complete
        t.w1 = flag(PF_Z, m_tile); t.t1 = (unsigned)(PS_PUB * P.nq * s);                // z_{s-1} (slab s) sampled
        t.sig = flag(PF_H, m_tile);
        __nv_bfloat16* s_h = P.S + (long)(s + 1) * B * P.KS + ZP + 64;
        const float* h_prev = s == 0 ? P.zero_h : P.hidden + (long)(s - 1) * D;
        const long ld_prev = s == 0 ? (long)D : ldH;
        if (P.U == 32) {
          const EpiGruP<32>::Params p{P.b_ih, P.b_hh, h_prev, P.hidden + (long)s * D, s_h, ldH, P.KS, D, P.Wgru + ZP, P.KS, 0, nullptr, P.dbg, t.code, ld_prev};
          ps_run_tile<32>(c, t, [&](int tid) { EpiGruP<32>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<32>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); },
                          [&](int tid) { EpiGruP<32>::post(p, n_tile, B, tile, m0, tid); });
        } else {
          const EpiGruP<64>::Params p{P.b_ih, P.b_hh, h_prev, P.hidden + (long)s * D, s_h, ldH, P.KS, D, P.Wgru + ZP, P.KS, 0, nullptr, P.dbg, t.code, ld_prev};
          ps_run_tile<64>(c, t, [&](int tid) { EpiGruP<64>::stage(p, n_tile, m0, B, epi_sm, wa, hp_tile, tid); },
                          [&](int tid) { EpiGruP<64>::run(p, n_tile, B, epi_sm, wa, hp_tile, tile, taddr, m, row, part, tid); },
                          [&](int tid) { EpiGruP<64>::post(p, n_tile, B, tile, m0, tid); });
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
// host side
// ------------------------------------------------------------------------------------------
struct drm_obs_persist {
  int U = 0, nt = 0, nq = 0, n_cta = 0;
  int* sched = nullptr;
  unsigned* flags = nullptr;
  size_t flag_bytes = 0;
  uint8_t* idx_prev = nullptr;
  size_t n_prev = 0;
  unsigned* dbg = nullptr;   // host-mapped
  unsigned long long* trace = nullptr;
  int trace_j0 = 0, trace_j1 = 0;
  bool tried = false, ok = false;
};

namespace drm {

static void obs_persist_free(drm_obs_persist* ps) {
  if (!ps) return;
  if (ps->dbg) cudaFreeHost(ps->dbg);
  delete ps;
}

static int obs_persist_capacity() {
  static int n = -1;
  if (n < 0) {
    n = 0;
    if (cudaFuncSetAttribute(observe_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_TOTAL) == cudaSuccess) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr[1];
      int sms = 0, dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      persist_launch_config(cfg, attr, (sms / 4) * 4, nullptr);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, observe_persist_kernel, &cfg) == cudaSuccess) n = nc;
    }
    cudaGetLastError();
  }
  return n;
}

// cluster c < mt: the posterior chain of m-tile c; then the GRU tiles, one per CTA (narrow tiles first: more SMs share a step)
static bool obs_persist_plan(drm_observe* o, drm_obs_persist* ps) {
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  const int cap = obs_persist_capacity();
  const int mt = ceil_div(o->B, BM);
  if (cap <= 0 || m->ZP % 256 || v->d.h_enc > 256) return false;
  for (int U = 32; U <= 64; U *= 2) {
    const int nt = ceil_div(m->d.D, U);
    if (mt + ceil_div(mt * nt, 4) > cap) continue;
    ps->U = U; ps->nt = nt; ps->nq = m->ZP / 256;
    const int n_cta = round_up(4 * mt + mt * nt, 4);
    std::vector<int> sc((size_t)n_cta * PS_SCHED_STRIDE, 0);
    auto add = [&](int c, int kind, int mm, int x) {
      int* rec = sc.data() + (size_t)c * PS_SCHED_STRIDE;
      const int n = rec[0]++;
      rec[1 + 3 * n] = kind; rec[2 + 3 * n] = mm; rec[3 + 3 * n] = x;
    };
    for (int mm = 0; mm < mt; ++mm)
      for (int k = 0; k < 4; ++k) add(4 * mm + k, PS_CHAIN, mm, 0);
    int cta = 4 * mt;
    for (int n = 0; n < nt; ++n)
      for (int mm = 0; mm < mt; ++mm) add(cta++, PS_GRU, mm, n);
    ps->n_cta = n_cta;
    if (dev_alloc(o->allocs, &ps->sched, sc.size()) != DRM_OK) return false;
    if (cudaMemcpy(ps->sched, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    ps->flag_bytes = (size_t)PF_COUNT * mt * 32 * sizeof(unsigned);
    if (dev_alloc(o->allocs, &ps->flags, ps->flag_bytes / sizeof(unsigned)) != DRM_OK) return false;
    ps->n_prev = (size_t)(o->T + 1) * o->B * m->d.R;
    if (dev_alloc(o->allocs, &ps->idx_prev, ps->n_prev) != DRM_OK) return false;
    if (cudaHostAlloc((void**)&ps->dbg, PS_DBG_WORDS * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return false; }
    memset(ps->dbg, 0, PS_DBG_WORDS * sizeof(unsigned));
    return true;
  }
  return false;
}

static bool obs_persist_eligible(drm_observe* o) {
  if (!o->ps) o->ps = new drm_obs_persist();
  drm_obs_persist* ps = o->ps;
  if (!ps->tried) {
    ps->tried = true;
    ps->ok = obs_persist_plan(o, ps);
  }
  return ps->ok;
}

// the recurrence of drm_observe_scan (after the hoisted conv features / feature part / actions are in place)
static int observe_persist(drm_observe* o, const float* uniforms, int mode, float* latent, float* hidden, float* post_logits, uint8_t* idx,
                           cudaStream_t st) {
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  drm_obs_persist* ps = o->ps;
  const int B = o->B, T = o->T, D = m->d.D, ZP = m->ZP;
  // the sampling epilogue stores one float per latent row into a zero-filled latent and flips single one-hot entries of the z columns
  DRM_CUDA(cudaMemsetAsync(latent, 0, (size_t)B * T * ZP * sizeof(float), st));
  DRM_CUDA(cudaMemsetAsync(ps->idx_prev, 255, ps->n_prev, st));
  DRM_CUDA(cudaMemsetAsync(ps->flags, 0, ps->flag_bytes, st));
  if (mode == 1) DRM_CUDA(cudaMemset2DAsync(hidden, (size_t)T * D * sizeof(float), 0, (size_t)D * sizeof(float), B, st));   // h_0 = 0
  ObsPersistParams P;
  memset(&P, 0, sizeof(P));
  const int vi = ps->U == 64 ? 1 : 0;
  const bool small = B <= SMALL_A_ROWS;
  P.tmS = small ? o->tmS_s : o->tmS; P.tmY1 = small ? o->tmY1_s : o->tmY1;
  P.tmWgru = m->tmWgru2[vi]; P.tmWehq = v->tmWehq; P.tmWe3 = v->tmWe3;
  P.B = B; P.T = T; P.D = D; P.DP = m->DP; P.ZP = ZP; P.R = m->d.R; P.KS = m->KS; P.mt = ceil_div(B, BM);
  P.U = ps->U; P.nt = ps->nt; P.nq = ps->nq;
  P.h_enc = v->d.h_enc; P.bn_he = v->bn_he;
  P.a_tx = small ? SMALL_A_ROWS * 128 : A_STAGE_BYTES;
  P.h_skip = mode == 1 ? 1 : 0;
  P.b_ih = m->b_ih; P.b_hh = m->b_hh; P.e1_b = v->e1_b; P.e1_g = v->e1_g; P.e1_be = v->e1_be; P.e3_b = v->e3_b;
  P.Wgru = m->Wgru2[vi];
  P.S = o->S; P.Y1 = o->Y1;
  P.featpart = o->featpart; P.uniforms = uniforms; P.zero_h = o->zero_h;
  P.latent = latent; P.hidden = hidden; P.logits = post_logits; P.idx = idx; P.idx_prev = ps->idx_prev;
  P.flags = ps->flags; P.dbg = ps->dbg; P.sched = ps->sched;
  P.trace = ps->trace; P.trace_j0 = ps->trace_j0; P.trace_j1 = ps->trace_j1;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  persist_launch_config(cfg, attr, ps->n_cta, st);
  cfg.dynamicSmemBytes = PS_TOTAL;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, observe_persist_kernel, P));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

}  // namespace drm

// debug: per-tile timestamps of the persistent scan (profiles/scan_trace.py); same protocol as drm_rollout_trace
extern "C" int drm_observe_trace(drm_observe* o, int32_t j0, int32_t nj, unsigned long long* out, int64_t n_words) {
  DRM_REQUIRE(o, DRM_ERR_ARG, "drm_observe_trace: NULL workspace");
  RC(check_arch());
  DRM_REQUIRE(drm::obs_persist_eligible(o), DRM_ERR_ARG, "drm_observe_trace: this workspace does not use the persistent kernel");
  drm_obs_persist* ps = o->ps;
  const size_t words = (size_t)ps->n_cta * drm::PS_TRACE_SLOTS * 8 * 2;
  if (!out) {
    if (!ps->trace) RC(drm::dev_alloc(o->allocs, &ps->trace, words));
    DRM_CUDA(cudaMemset(ps->trace, 0, words * sizeof(unsigned long long)));
    ps->trace_j0 = j0; ps->trace_j1 = j0 + nj;
    return DRM_OK;
  }
  DRM_REQUIRE(ps->trace && n_words >= (int64_t)words, DRM_ERR_ARG, "drm_observe_trace: tracing is off or the buffer is too small");
  DRM_CUDA(cudaDeviceSynchronize());
  DRM_CUDA(cudaMemcpy(out, ps->trace, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  ps->trace_j0 = ps->trace_j1 = 0;
  ps->trace = nullptr;
  return DRM_OK;
}
