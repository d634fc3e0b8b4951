// Chained MLP head for SMALL grids: Linear-LN-SiLU -> Linear-LN-SiLU -> output layer of one head, for one 128-row m-tile,
// inside ONE kernel run by a cluster of four CTAs.  Included from rssm.cu.
//
//   layer 0   A = state buffer via TMA,       B = W0 rows [slot*256 + 64*rank, +64)   -> 64 of the 256 hidden columns per CTA
//   layer 1   A = Y (shared memory),          B = W1 rows [slot*256 + 64*rank, +64)
//   layer 2   A = Y (shared memory),          B = output weights: prior 2 x 128 logit columns per CTA, heads 64 columns per CTA
//
// The hidden activations never leave the SMs: after LayerNorm + SiLU every CTA writes its 64-column slice (one k-block of
// the next layer's A operand, bf16, 128-byte swizzled) into its own shared memory and bulk-copies it
// (cp.async.bulk shared::cta -> shared::cluster) into its three peers', so each CTA ends up with the full 128 x 256 operand.  LayerNorm statistics are exchanged the same way.
// Cross-CTA signalling uses mbarriers with remote arrives (release/acquire at cluster scope), never barrier.cluster inside
// the pipeline, so the TMA producer can run ahead freely:
//   stat[l]  16 arrivals (the four part-0 warps of every CTA): "statistics of layer l published everywhere"
//   ybar[l]  1 local arrive + 3 x 16 KB of complete_tx: "my copy of Y_l is complete".  A peer starts writing Y_l only after
//            stat[l], which proves every CTA's MMAs of layer l (the readers of Y_{l-1}) have retired.
// This replaces three dependent launches (two LN stages + the output stage) by one: no kernel boundaries, no L2 round trips.
#pragma once

namespace drm {

constexpr int CH_CN = 4;                 // CTAs per cluster
constexpr int CH_NS = 3;                 // ring stages
constexpr int CH_STAGE = 32768;          // A 16 KB + B up to 16 KB (128 rows)
constexpr int CH_Y_OFF = CH_NS * CH_STAGE;
constexpr int CH_Y_BYTES = 4 * 16384;    // 4 k-blocks of [128 x 64] bf16
constexpr int CH_BAR_OFF = CH_Y_OFF + CH_Y_BYTES;
constexpr int CH_SCR_OFF = CH_BAR_OFF + 256;
// scratch (floats): LN constants of layers 0 / 1, part merge, cross-CTA statistics of layers 0 / 1, output-layer region, bucket merge
constexpr int CS_LN0 = 0, CS_LN1 = 768, CS_XS = 1536, CS_XST0 = 3072, CS_XST1 = 4096, CS_F = 5120, CS_RED = 8192, CS_END = 9728;
constexpr int CH_SMEM = CH_SCR_OFF + CS_END * 4 + 1024;
enum ChainKind { CHAIN_PRIOR = 0, CHAIN_HEADS = 1 };

struct ChainParams {
  CUtensorMap tmA;         // layer-0 A (state buffer), box {64, 128}
  CUtensorMap tmW0, tmW1;  // LN layer weights, box {64, 64}
  CUtensorMap tmW2;        // output weights, box {64, bn2}
  int M, a_row0;
  int ka0, nka0, ka1, nka1;   // layer-0 A k-block ranges
  int nk1, nk2;               // k-blocks of layers 1 and 2
  int nv0, nv1;               // LN feature counts
  int bn2, passes;            // output layer: MMA N per pass, passes per CTA
  const float *b0, *g0, *be0, *b1, *g1, *be1;
  int cs0, cs1;               // LN constants per slot
  int n_slots;
  int y_slot[8];
  float eps;
  int kind;
  EpiCat::Params cat;         // CHAIN_PRIOR
  EpiHeads::Params heads;     // CHAIN_HEADS
  unsigned long long* timeline;  // debug probes of cluster (0,0) rank 0, or NULL
  unsigned long long* cta_times; // debug per-CTA records, or NULL
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(mapa_u32(smem_u32(bar), cta)) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// shared memory -> a peer CTA's shared memory, completion (bytes) signalled on the PEER's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}

__device__ __forceinline__ void cprobe(const ChainParams& p, int i) {
  if (p.timeline && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.timeline[2 * i] = t;
    p.timeline[2 * i + 1] = (unsigned long long)clock64();
  }
}

// LayerNorm + SiLU of one layer for this thread's 16 columns; publishes the bf16 slice into every CTA's Y buffer.
__device__ __forceinline__ void chain_ln_layer(const ChainParams& p, int nv, const float* cst, float* xs, float* xst, uint64_t* stat,
                                               uint64_t* ybar, uint8_t* ybuf, uint32_t taddr, int row, int part, int rank, int lane, int pb) {
  const int c0 = part * 16, gc0 = 64 * rank + c0;
  auto cnt_of = [nv](int first, int width) { return max(0, min(width, nv - first)); };
  const int cnt = cnt_of(gc0, 16);
  float v[16];
  tmem_ld16(taddr + c0, v);
  {
    const float4* b4 = reinterpret_cast<const float4*>(cst + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = b4[j];
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
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
  xs[part * 128 + row] = shift + s1 * inv_cnt;
  xs[512 + part * 128 + row] = fmaxf(s2 - s1 * s1 * inv_cnt, 0.f);
  epi_bar_sync();
  if (threadIdx.x == 64) cprobe(p, pb);
  if (part == 0) {
    const int cnt_c = cnt_of(64 * rank, 64);
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < EPI_PARTS; ++q) tot += xs[q * 128 + row] * (float)cnt_of(64 * rank + 16 * q, 16);
    const float mean_c = cnt_c > 0 ? tot / (float)cnt_c : 0.f;
    float M2c = 0.f;
#pragma unroll
    for (int q = 0; q < EPI_PARTS; ++q) {
      const float d = xs[q * 128 + row] - mean_c;
      M2c += xs[512 + q * 128 + row] + d * d * (float)cnt_of(64 * rank + 16 * q, 16);
    }
#pragma unroll
    for (uint32_t dst = 0; dst < CH_CN; ++dst) {
      st_cluster_f32(xst + (rank * 128 + row) * 2, dst, mean_c);
      st_cluster_f32(xst + (rank * 128 + row) * 2 + 1, dst, M2c);
    }
    __syncwarp();   // part 0 is four whole warps: one release-arrive per warp and destination covers its 32 rows
    if (lane == 0) {
#pragma unroll
      for (uint32_t dst = 0; dst < CH_CN; ++dst) mbar_arrive_remote(stat, dst);
    }
  }
  mbar_wait_cluster(stat, 0);
  if (threadIdx.x == 64) cprobe(p, pb + 1);   // every CTA's statistics are here, and every CTA's MMAs of this layer have retired
  float tot = 0.f;
#pragma unroll
  for (int q = 0; q < CH_CN; ++q) tot += xst[(q * 128 + row) * 2] * (float)cnt_of(64 * q, 64);
  const float mean = tot / (float)nv;
  float M2 = 0.f;
#pragma unroll
  for (int q = 0; q < CH_CN; ++q) {
    const float d = xst[(q * 128 + row) * 2] - mean;
    M2 += xst[(q * 128 + row) * 2 + 1] + d * d * (float)cnt_of(64 * q, 64);
  }
  const float rstd = rsqrtf(M2 / (float)nv + p.eps);
  const float nmr = -mean * rstd;
  {
    const float4* ga = reinterpret_cast<const float4*>(cst + 256 + c0);
    const float4* be = reinterpret_cast<const float4*>(cst + 512 + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 G = ga[j], Bt = be[j];
      v[4 * j] = siluf_(fmaf(fmaf(v[4 * j], rstd, nmr), G.x, Bt.x));
      v[4 * j + 1] = siluf_(fmaf(fmaf(v[4 * j + 1], rstd, nmr), G.y, Bt.y));
      v[4 * j + 2] = siluf_(fmaf(fmaf(v[4 * j + 2], rstd, nmr), G.z, Bt.z));
      v[4 * j + 3] = siluf_(fmaf(fmaf(v[4 * j + 3], rstd, nmr), G.w, Bt.w));
    }
  }
  if (threadIdx.x == 64 && pb == 8) cprobe(p, 15);
  // bf16 slice -> k-block `rank` of Y in all four CTAs (SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7))
  uint4 lo, hi;
  lo.x = pack_bf16x2(v[0], v[1]); lo.y = pack_bf16x2(v[2], v[3]); lo.z = pack_bf16x2(v[4], v[5]); lo.w = pack_bf16x2(v[6], v[7]);
  hi.x = pack_bf16x2(v[8], v[9]); hi.y = pack_bf16x2(v[10], v[11]); hi.z = pack_bf16x2(v[12], v[13]); hi.w = pack_bf16x2(v[14], v[15]);
  const uint32_t base = smem_u32(ybuf) + rank * 16384 + row * 128;
  st_shared_v4(base + (((2 * part) ^ (row & 7)) << 4), lo);
  st_shared_v4(base + (((2 * part + 1) ^ (row & 7)) << 4), hi);
  fence_proxy_async();   // generic-proxy writes -> visible to the async proxy (tensor-core operand reads, bulk copies)
  if (threadIdx.x == 64 && pb == 8) cprobe(p, 16);
  epi_bar_sync();
  if (threadIdx.x == 64 && pb == 8) cprobe(p, 17);
  if (threadIdx.x == 64) {
    // this CTA's copy of Y is complete once the three peer slices have landed (complete_tx) and this arrive has happened
    mbar_expect_tx(ybar, (CH_CN - 1) * 16384);
    const uint32_t src = smem_u32(ybuf) + rank * 16384;
#pragma unroll
    for (uint32_t d = 1; d < CH_CN; ++d) {
      const uint32_t dst = (rank + d) & (CH_CN - 1);
      bulk_copy_to_peer(mapa_u32(src, dst), src, 16384, mapa_u32(smem_u32(ybar), dst));
    }
  }
  if (threadIdx.x == 64) cprobe(p, pb + 2);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CH_BAR_OFF);
  uint64_t* empty = full + CH_NS;
  uint64_t* tfull = empty + CH_NS;   // [3]
  uint64_t* stat = tfull + 3;        // [2]
  uint64_t* ybar = stat + 2;         // [2]
  uint64_t* red = ybar + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red + 1);
  float* scr = reinterpret_cast<float*>(smem + CH_SCR_OFF);
  uint8_t* ybuf = smem + CH_Y_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int slot = p.n_slots > 0 ? p.y_slot[blockIdx.y] : (int)blockIdx.y;
  const int a_row = p.a_row0 + (int)blockIdx.x * BM;
  const int nk0 = p.nka0 + p.nka1;
  const int w_row = slot * 256 + rank * 64;                                    // LN layers
  const int w2_row = p.kind == CHAIN_PRIOR ? rank * 256 : slot * 256 + rank * 64; // output layer (first pass)

  if (threadIdx.x == 0) {
    cprobe(p, 0);
    cta_probe(p.cta_times, 0);
    tma_prefetch_desc(&p.tmA); tma_prefetch_desc(&p.tmW0); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2);
    for (int s = 0; s < CH_NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 3; ++i) mbar_init(&tfull[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&stat[i], CH_CN * 4); mbar_init(&ybar[i], 1); }
    mbar_init(red, CH_CN * 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  // (after the TMEM allocation: see fused_gemm_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  cluster_sync_all();   // peers are resident and their barriers initialised before any remote access
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) cprobe(p, 1);
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");
  if (threadIdx.x == 0) cta_probe(p.cta_times, 1);

  const int n_iter = nk0 + p.nk1 + p.passes * p.nk2;
  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % CH_NS;
        const uint32_t ph = (it / CH_NS) & 1;
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* sa = smem + s * CH_STAGE;
        uint8_t* sb = sa + A_STAGE_BYTES;
        if (it < nk0) {
          const int ka = it < p.nka0 ? p.ka0 + it : p.ka1 + (it - p.nka0);
          mbar_expect_tx(&full[s], A_STAGE_BYTES + 64 * BK * 2);
          tma_load_2d(sa, &p.tmA, ka * BK, a_row, &full[s]);
          tma_load_2d(sb, &p.tmW0, it * BK, w_row, &full[s]);
        } else if (it < nk0 + p.nk1) {
          mbar_expect_tx(&full[s], 64 * BK * 2);
          tma_load_2d(sb, &p.tmW1, (it - nk0) * BK, w_row, &full[s]);
        } else {
          const int j = it - nk0 - p.nk1, pass = j / p.nk2, kb = j - pass * p.nk2;
          mbar_expect_tx(&full[s], (uint32_t)p.bn2 * BK * 2);
          tma_load_2d(sb, &p.tmW2, kb * BK, w2_row + pass * p.bn2, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc64 = umma_idesc_bf16(64), idesc2 = umma_idesc_bf16(p.bn2);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % CH_NS;
        const uint32_t ph = (it / CH_NS) & 1;
        int layer, kb, pass = 0;
        if (it < nk0) { layer = 0; kb = it; }
        else if (it < nk0 + p.nk1) { layer = 1; kb = it - nk0; }
        else { layer = 2; const int j = it - nk0 - p.nk1; pass = j / p.nk2; kb = j - pass * p.nk2; }
        if (layer > 0 && kb == 0 && pass == 0) {   // the previous layer's activations are complete in this CTA's Y buffer
          mbar_wait_cluster(&ybar[layer - 1], 0);
          tc_fence_after();
          cprobe(p, layer == 1 ? 18 : 20);
        }
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t st_addr = smem_u32(smem + s * CH_STAGE);
        const uint64_t adesc = umma_desc_sw128(layer == 0 ? st_addr : smem_u32(ybuf) + kb * 16384);
        const uint64_t bdesc = umma_desc_sw128(st_addr + A_STAGE_BYTES);
        const uint32_t dcol = layer == 0 ? 0u : (layer == 1 ? 64u : 128u + (uint32_t)(pass * p.bn2));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem + dcol, adesc + 2 * k, bdesc + 2 * k, layer == 2 ? idesc2 : idesc64, (kb | k) != 0);
        umma_commit(&empty[s]);
        const bool last = layer == 0 ? it == nk0 - 1 : (layer == 1 ? it == nk0 + p.nk1 - 1 : it == n_iter - 1);
        if (last) {
          umma_commit(&tfull[layer]);
          if (layer > 0) cprobe(p, layer == 1 ? 19 : 21);
        }
      }
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    const int q = warp & 3, part = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int m = (int)blockIdx.x * BM + row;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    // ---- constants of all three layers (weights-derived + host inputs only: may precede griddepcontrol.wait) ----
    for (int i = tid; i < 64; i += EPI_THREADS) {
      const int col = 64 * rank + i;
      scr[CS_LN0 + i] = col < p.nv0 ? __ldg(p.b0 + slot * p.cs0 + col) : 0.f;
      scr[CS_LN0 + 256 + i] = col < p.nv0 ? __ldg(p.g0 + slot * p.cs0 + col) : 0.f;
      scr[CS_LN0 + 512 + i] = col < p.nv0 ? __ldg(p.be0 + slot * p.cs0 + col) : 0.f;
      scr[CS_LN1 + i] = col < p.nv1 ? __ldg(p.b1 + slot * p.cs1 + col) : 0.f;
      scr[CS_LN1 + 256 + i] = col < p.nv1 ? __ldg(p.g1 + slot * p.cs1 + col) : 0.f;
      scr[CS_LN1 + 512 + i] = col < p.nv1 ? __ldg(p.be1 + slot * p.cs1 + col) : 0.f;
    }
    GemmCommon gg;                 // the output-layer epilogues only read bn and M
    gg.bn = p.kind == CHAIN_PRIOR ? 256 : 64;
    gg.M = p.M;
    if (p.kind == CHAIN_PRIOR) EpiCat::stage(p.cat, gg, rank, scr + CS_F, tid);
    else {
      const bool bucket = p.heads.kind[slot] == HEAD_BUCKET;
      for (int i = tid; i < 64; i += EPI_THREADS) {
        const int col = 64 * rank + i;
        scr[CS_F + i] = __ldg(p.heads.bias + slot * 256 + col);
        scr[CS_F + 256 + i] = (bucket && col < p.heads.NB) ? __ldg(p.heads.buckets[slot] + col) : 0.f;
      }
    }
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    epi_bar_sync();
    if (threadIdx.x == 64) cprobe(p, 2);
    // ---- layer 0, layer 1: LayerNorm + SiLU, slices broadcast into every CTA's Y buffer ----
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    if (threadIdx.x == 64) cprobe(p, 3);
    chain_ln_layer(p, p.nv0, scr + CS_LN0, scr + CS_XS, scr + CS_XST0, &stat[0], &ybar[0], ybuf, tlane, row, part, rank, lane, 4);
    mbar_wait(&tfull[1], 0);
    tc_fence_after();
    if (threadIdx.x == 64) cprobe(p, 7);
    chain_ln_layer(p, p.nv1, scr + CS_LN1, scr + CS_XS, scr + CS_XST1, &stat[1], &ybar[1], ybuf, tlane + 64, row, part, rank, lane, 8);
    // ---- output layer ----
    mbar_wait(&tfull[2], 0);
    tc_fence_after();
    if (threadIdx.x == 64) cprobe(p, 11);
    float* tile = reinterpret_cast<float*>(smem);   // ring + Y are free now: every MMA of this CTA has retired
    if (p.kind == CHAIN_PRIOR) {
      EpiCat::run(p.cat, gg, scr + CS_F, tile, tlane + 128, m, row, part, rank, tid);
    } else {
      const EpiHeads::Params& hp = p.heads;
      const bool live = m < p.M;
      const int kind = hp.kind[slot];
      const float* cf = scr + CS_F;
      if (kind == HEAD_BUCKET) {
        const int NB = hp.NB;
        const int c0 = part * 16, gc0 = 64 * rank + c0;
        const int cnt = max(0, min(16, NB - gc0));
        float v[16];
        tmem_ld16(tlane + 128 + c0, v);
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] += cf[c0 + j]; mx = fmaxf(mx, j < cnt ? v[j] : -INFINITY); }
        if (hp.logits[slot]) {   // coalesced logits copy-out of this CTA's 64 columns
          tile_put<16>(tile, 68, row, c0, v);
          epi_bar_sync();
          const int mine = max(0, min(64, NB - 64 * rank));
          if (mine > 0) tile_copy_out(tile, 68, 64, mine, (int)blockIdx.x * BM, p.M, hp.logits[slot] + 64 * rank, hp.ld_logits[slot], nullptr, 0, tid, hp.rm);
        }
        float s = 0.f, ws = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e = j < cnt ? fexpf_(v[j] - mx) : 0.f;
          s += e;
          ws = fmaf(e, cf[256 + c0 + j], ws);
        }
        float* xs = scr + CS_XS;
        xs[part * 128 + row] = mx; xs[512 + part * 128 + row] = s; xs[1024 + part * 128 + row] = ws;
        epi_bar_sync();
        float* redbuf = scr + CS_RED;   // rank 0's copy collects [4 ranks][128 rows][3]
        if (part == 0) {
          float Mx = -INFINITY;
#pragma unroll
          for (int qq = 0; qq < EPI_PARTS; ++qq) Mx = fmaxf(Mx, xs[qq * 128 + row]);
          float S = 0.f, WS = 0.f;
#pragma unroll
          for (int qq = 0; qq < EPI_PARTS; ++qq) {
            const float mq = xs[qq * 128 + row];
            const float f = mq == -INFINITY ? 0.f : fexpf_(mq - Mx);
            S = fmaf(xs[512 + qq * 128 + row], f, S);
            WS = fmaf(xs[1024 + qq * 128 + row], f, WS);
          }
          st_cluster_f32(redbuf + (rank * 128 + row) * 3, 0, Mx);
          st_cluster_f32(redbuf + (rank * 128 + row) * 3 + 1, 0, S);
          st_cluster_f32(redbuf + (rank * 128 + row) * 3 + 2, 0, WS);
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(red, 0);
          __syncwarp();
          if (rank == 0) {
            mbar_wait_cluster(red, 0);
            float G = -INFINITY;
#pragma unroll
            for (int c = 0; c < CH_CN; ++c) G = fmaxf(G, redbuf[(c * 128 + row) * 3]);
            float SS = 0.f, WW = 0.f;
#pragma unroll
            for (int c = 0; c < CH_CN; ++c) {
              const float mq = redbuf[(c * 128 + row) * 3];
              const float f = mq == -INFINITY ? 0.f : fexpf_(mq - G);
              SS = fmaf(redbuf[(c * 128 + row) * 3 + 1], f, SS);
              WW = fmaf(redbuf[(c * 128 + row) * 3 + 2], f, WW);
            }
            if (live && hp.value[slot]) hp.value[slot][map_row(hp.rm, m) * hp.ld_value[slot]] = symexpf_(WW / SS);
          }
        }
      } else if (rank == 0 && part == 0) {
        float v[32];
        tmem_ld32(tlane + 128, v);
        if (kind == HEAD_SIGMOID) {
          const float x = v[0] + cf[0];
          if (live) {
            if (hp.value[slot]) hp.value[slot][map_row(hp.rm, m) * hp.ld_value[slot]] = 1.0f / (1.0f + expf(-x));
            if (hp.logits[slot]) hp.logits[slot][map_row(hp.rm, m) * hp.ld_value[slot]] = x;
          }
        } else if (live) {
          const int A = hp.A;
          float act[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            act[j] = 0.f;
            if (j < A) {
              const float mu = v[j] + cf[j];
              float ls = v[16 + j] + cf[16 + j];
              ls = fminf(fmaxf(ls, -5.0f), 2.0f);
              const float sg = softplusf_(ls) + 1e-3f;
              if (hp.mu) hp.mu[(long)m * hp.ld_act + j] = mu;
              if (hp.sigma) hp.sigma[(long)m * hp.ld_act + j] = sg;
              if (hp.normals) {
                act[j] = tanhf(mu + sg * __ldg(hp.normals + (long)m * hp.ld_normals + j));
                if (hp.action) hp.action[(long)m * hp.ld_act + j] = act[j];
              }
            }
          }
          if (hp.s_a && hp.normals) store_bf16_row<16>(hp.s_a + (long)m * hp.ld_s, act, 16);
        }
      }
    }
  }
  if (threadIdx.x == 64) cprobe(p, 12);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) cprobe(p, 13);
  cluster_sync_all();   // no CTA leaves while a peer may still write into its shared memory or arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (threadIdx.x == 0) { cprobe(p, 14); cta_probe(p.cta_times, 2); }
}

}  // namespace drm
