// The fused GEMM main loop shared by every tensor-core stage of the RSSM path.
//
//   out_tile[128 x N] = A[128 x K] (bf16, K-major, TMA-staged) * B[N x K]^T (bf16, K-major, TMA-staged)
//
// accumulated in TMEM by tcgen05.mma (one elected thread), then handed to a per-stage epilogue
// functor that reads the accumulator with tcgen05.ld -- one thread per output row, so row-wise
// reductions (LayerNorm, softmax, the 32-class CDF) need no shuffles.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..17 = epilogue: TMEM lane quadrant = warp_id % 4, and the four warps sharing a quadrant
// split the tile's columns four ways (`part`), exchanging partial row statistics through smem.
// Pipeline: STAGES-deep smem ring (full/empty mbarriers), one tmem_full barrier.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace drm {

constexpr int BM = 128;           // rows per CTA tile == UMMA M == TMEM lanes
constexpr int BK = 64;            // bf16 per k-block == one 128-byte swizzle row
constexpr int EPI_PARTS = 4;        // epilogue threads per accumulator row
constexpr int EPI_THREADS = 128 * EPI_PARTS;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr int A_STAGE_BYTES = BM * BK * 2;

// What an epilogue functor needs to know about the tile it finishes (the launch-per-stage kernels pass their GemmCommon,
// the persistent rollout kernel builds one per task).
struct TileG {
  int M;             // valid rows (epilogues guard their global stores with m < M)
  int bn;            // B rows per tile == UMMA N of the (first) accumulator group
  int hp_pre;        // 32-unit GRU tiles: 1 = the tile's h_prev was fetched into the epilogue scratch under the main loop
  int wide;          // 1: TF32 mode -- operand buffers (state, hidden activations, weights) hold fp32, 32 elements per k-block
};

// Operand buffers are declared as bf16 pointers; in TF32 mode the same fields address fp32 data: element offset -> pointer
__device__ __host__ __forceinline__ __nv_bfloat16* opnd_at(__nv_bfloat16* base, long elems, int wide) { return base + (elems << wide); }

struct GemmCommon : TileG {
  CUtensorMap tmA;   // activations, box {64, 128}
  CUtensorMap tmB;   // packed weights, box {64, bn}
  int b_slot_rows;   // B rows between consecutive slots (0: == bn)
  int a_row0;        // A row of tile (x = 0, slot = 0)
  int a_y_stride;    // extra A rows per slot
  int ka0, nka0;     // A k-block ranges [ka0, ka0 + nka0) then [ka1, ka1 + nka1); B k-blocks are
  int ka1, nka1;     //   consumed sequentially from 0
  int n_slots;       // 0: slot = blockIdx.y;  > 0: slot = y_slot[blockIdx.y]
  int y_slot[8];
  int band;                // CTA-pair GRU kernel: m-tiles per band of the tile order (even; 0 = 16)
  int a_bytes;             // bytes one A k-block load delivers: 0 = the full 128-row tile; SMALL_A_ROWS * 128 when tmA has a short box
  unsigned long long* cta_times;  // debug: per-CTA {entry, wait over, exit, smid} records, or NULL
  unsigned long long* timeline;  // debug: CTA (0,0) writes {globaltimer ns, clock64} pairs at 8 probe points, or NULL
};

// debug: every CTA records {entry ns, dependency-wait-over ns, exit ns, SM id} behind the 8 stage probes
// (region: timeline_base + DRM_STAGE_COUNT * 16 + stage * 1024, up to 256 CTAs per stage)
__device__ __forceinline__ void cta_probe(unsigned long long* stage_timeline, int which) {
  if (!stage_timeline) return;
  const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (cta >= 256) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  unsigned long long* rec = stage_timeline + 4 * cta;
  rec[which] = t;
  if (which == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    rec[3] = smid;
  }
}

__device__ __forceinline__ void probe(const GemmCommon& g, int i) {
  if (g.timeline && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g.timeline[2 * i] = t;
    g.timeline[2 * i + 1] = (unsigned long long)clock64();
  }
}

// k-blocks per pipeline stage: Epi::KPS when the epilogue declares it, else 1
template <class E, class = void>
struct kps_of { static constexpr int value = 1; };
template <class E>
struct kps_of<E, std::void_t<decltype(E::KPS)>> { static constexpr int value = E::KPS; };

template <int B_ROWS_MAX, int STAGES, int KPS = 1>
struct GemmSmem {
  static constexpr int B_STAGE_BYTES = B_ROWS_MAX * BK * 2;
  static constexpr int SUB_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;    // one k-block: [A 16 KB | B]
  static constexpr int STAGE_BYTES = KPS * SUB_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = BAR_OFF + 256;        // 1024 floats of per-tile constants + 3072 of row-stat exchange
  static constexpr int ZI_OFF = EPI_OFF + 16384;       // 128 rows x 32 sampled indices of this tile's rows
  static constexpr int TOTAL = ZI_OFF + 4096 + 1024;   // barriers + epilogue scratch + indices + alignment slack
  static_assert(B_STAGE_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment");
};

// Epi must provide:
//   static constexpr int B_ROWS_MAX, STAGES, TMEM_COLS;  static constexpr int GRU_U (0 = plain)
//   static constexpr int CLUSTER_N (1 or 4).  CLUSTER_N = 4: the four CTAs of a cluster (gridDim.z) each own bn = 64 columns
//       of one 256-column tile; the epilogue exchanges row statistics through distributed shared memory and every thread of
//       the cluster takes part in ONE barrier.cluster between the two epilogue halves.
//   struct Params;
//   static __device__ void stage(const Params&, const TileG&, int slot, float* sm, int tid, int m0);
//       -- the EPI_THREADS epilogue threads copy the tile's constants (biases, LN affine, buckets)
//          into shared memory while the main loop runs
//   static __device__ void run(const Params&, const TileG&, float* sm, float* tile, uint32_t taddr, int m,
//                              int row, int part, int slot, int tid);
//       -- sm + 1024 = exchange scratch [.][128 rows]; `tile` = the pipeline stages' shared memory, free once the
//          accumulator is complete, used to transpose the output tile so that global stores are fully coalesced
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_THREADS) : "memory"); }

template <class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, Epi::MIN_CTAS)
fused_gemm_kernel(const __grid_constant__ GemmCommon g, const __grid_constant__ typename Epi::Params ep) {
  constexpr int KPS = kps_of<Epi>::value;
  using SL = GemmSmem<Epi::B_ROWS_MAX, Epi::STAGES, KPS>;
  constexpr int STAGES = Epi::STAGES;
  constexpr int CN = Epi::CLUSTER_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = g.nka0 + g.nka1;
  const int bke = g.wide ? BK / 2 : BK;   // elements per 128-byte k-block
  const int slot = g.n_slots > 0 ? g.y_slot[blockIdx.y] : (int)blockIdx.y;
  const int m0 = (int)blockIdx.x * BM;
  const int a_row = g.a_row0 + slot * g.a_y_stride + m0;
  const int b_row = slot * (g.b_slot_rows ? g.b_slot_rows : g.bn) + (CN > 1 ? (int)cluster_ctarank() * g.bn : 0);

  // (everything this CTA reads that an earlier stage produced -- activations via TMA, h_prev -- is touched only after
  // griddepcontrol.wait; the trigger that lets the next stage's CTAs start their prologue follows the TMEM allocation below)
  if (threadIdx.x == 0) {
    probe(g, 0);
    cta_probe(g.cta_times, 0);
    tma_prefetch_desc(&g.tmA);
    tma_prefetch_desc(&g.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Epi::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  // Programmatic dependent launch, triggered only now that this CTA HOLDS its TMEM columns: the next stage's CTAs allocate TMEM in
  // their prologue and then block in griddepcontrol.wait until this grid has finished, so a dependent CTA that reached an SM before
  // a CTA of this grid had allocated could starve it of columns for good.
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  if constexpr (CN > 1) cluster_sync_all();   // peers are resident and their barriers initialised before any remote access
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 2) asm volatile("griddepcontrol.wait;\n" ::: "memory");  // producer / MMA warps; epilogue warps wait after staging

  if (warp == 0) {
    if (lane == 0) {
      probe(g, 1);
      cta_probe(g.cta_times, 1);
      const uint32_t tx_b = (uint32_t)g.bn * BK * 2;
      const uint32_t tx_a = g.a_bytes ? (uint32_t)g.a_bytes : (uint32_t)A_STAGE_BYTES;
      if constexpr (KPS > 1) {
        // KPS k-blocks per stage: one full / empty handshake per KPS k-blocks.  The main loop of the small-grid stages is bound by
        // the MMA thread's per-handshake chain (~210 ns), not by operand delivery, so halve the handshakes.
        const int n_st = (nk + KPS - 1) / KPS;
        for (int st = 0; st < n_st; ++st) {
          const int s = st % STAGES;
          mbar_wait(&empty[s], ((st / STAGES) & 1) ^ 1u);
          const int n_sub = min(KPS, nk - st * KPS);
          mbar_expect_tx(&full[s], (uint32_t)n_sub * (tx_a + tx_b));
          for (int j = 0; j < n_sub; ++j) {
            const int kb = st * KPS + j;
            uint8_t* sa = smem + s * SL::STAGE_BYTES + j * SL::SUB_BYTES;
            const int ka = kb < g.nka0 ? g.ka0 + kb : g.ka1 + (kb - g.nka0);
            tma_load_2d(sa, &g.tmA, ka * bke, a_row, &full[s]);
            tma_load_2d(sa + A_STAGE_BYTES, &g.tmB, kb * bke, b_row, &full[s]);
          }
          if (st == 0) probe(g, 2);
        }
      } else {
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* sa = smem + s * SL::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          const int ka = kb < g.nka0 ? g.ka0 + kb : g.ka1 + (kb - g.nka0);
          mbar_expect_tx(&full[s], tx_b + tx_a);
          tma_load_2d(sa, &g.tmA, ka * bke, a_row, &full[s]);
          tma_load_2d(sb, &g.tmB, kb * bke, b_row, &full[s]);
          if (kb == 0) probe(g, 2);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if constexpr (KPS > 1) {
        const uint32_t idesc = umma_idesc(g.bn, g.wide);
        const int n_st = (nk + KPS - 1) / KPS;
        for (int st = 0; st < n_st; ++st) {
          const int s = st % STAGES;
          mbar_wait(&full[s], (st / STAGES) & 1);
          tc_fence_after();
          if (st == 0) probe(g, 3);
          const int n_sub = min(KPS, nk - st * KPS);
          for (int j = 0; j < n_sub; ++j) {
            const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES + j * SL::SUB_BYTES);
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_op(g.wide, tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (st | j | k) != 0);
          }
          umma_commit(&empty[s]);
        }
      } else {
        for (int kb = 0; kb < nk; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (kb == 0) probe(g, 3);
          const uint32_t a_addr = smem_u32(smem + s * SL::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
          if constexpr (Epi::GRU_U == 0) {
            const uint32_t idesc = umma_idesc(g.bn, g.wide);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_op(g.wide, tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          } else {
            // GRU tile: B rows = [r(U) | z(U) | n(U)], TMEM cols = [r | z | n_x | n_h].
            // x-part k-blocks feed r, z, n_x in one N = 3U MMA; h-part k-blocks feed r, z (N = 2U)
            // and n_h (N = U, B rows 2U.., TMEM cols 3U..) because r multiplies only W_hn h.
            constexpr int U = Epi::GRU_U;
            if (kb < g.nka0) {
              const uint32_t idesc = umma_idesc(3 * U, g.wide);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_op(g.wide, tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            } else {
              const uint32_t idesc_rz = umma_idesc(2 * U, g.wide);
              const uint32_t idesc_n = umma_idesc(U, g.wide);
              const uint64_t bdesc_n = umma_desc_sw128(a_addr + A_STAGE_BYTES + 2 * U * BK * 2);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                umma_op(g.wide, tmem, adesc + 2 * k, bdesc + 2 * k, idesc_rz, 1u);
                umma_op(g.wide, tmem + 3 * U, adesc + 2 * k, bdesc_n + 2 * k, idesc_n, (kb > g.nka0 || k > 0) ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[s]);   // frees the smem stage when these MMAs retire
        }
      }
      umma_commit(tmem_full);    // accumulator complete
      probe(g, 4);
    }
  } else {
    float* epi_sm = reinterpret_cast<float*>(smem + SL::EPI_OFF);
    Epi::stage(ep, g, slot, epi_sm, (int)threadIdx.x - 64, m0);   // weights-derived constants + host inputs only
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if constexpr (Epi::GRU_U == 32) {     // h_prev is final once the dependency wait is over: fetch it under the main loop
      if (g.hp_pre) Epi::load_hprev(ep, g, slot * 32, m0, epi_sm + Epi::HP_OFF, (int)threadIdx.x - 64);
    }
    epi_bar_sync();                                     // epilogue warps only
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (threadIdx.x == 64) probe(g, 5);
    const int q = warp & 3;                             // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;                   // which quarter of the columns
    const int row = q * 32 + lane;
    const int m = m0 + row;
    Epi::run(ep, g, epi_sm, reinterpret_cast<float*>(smem), tmem + ((uint32_t)(q * 32) << 16), m, row, part, slot,
             (int)threadIdx.x - 64);
    if (threadIdx.x == 64) probe(g, 6);
  }
  if constexpr (CN > 1) {
    if (warp < 2) {          // the producer / MMA warps take part in the epilogue's single cluster barrier
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, Epi::TMEM_COLS);
  if (threadIdx.x == 32) { probe(g, 7); cta_probe(g.cta_times, 2); }
}

// ------------------------------------------------------------------------------------------
// output-tile transposition.  An epilogue thread owns (part of) one ROW of the tile; writing rows
// straight to global memory makes every warp-level store touch 32 different lines.  Instead each
// thread drops its values into a padded fp32 shared-memory tile (pitch = ncols + 4 floats: the 8
// threads of a store phase hit 8 distinct 16-byte bank groups), and after a barrier all
// EPI_THREADS threads stream the tile out with consecutive threads on consecutive 16 bytes.
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void tile_put(float* tile, int pitch, int row, int col, const float (&v)[N]) {
  float4* dst = reinterpret_cast<float4*>(tile + row * pitch + col);
#pragma unroll
  for (int j = 0; j < N; j += 4) dst[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
}
// Where tile row m (the GEMM row) lands in the output: identity, time-major -> batch-major, or the
// sub-pixel phase scatter of a stride-2 transposed convolution.
struct RowMap {
  int mode;        // 0: row = m;  1: m = t * B + b -> b * T + t;  2: convT phase (NHWC rows);  3: convT phase, NCHW fp32 planes
  int p0, p1, p2;  // mode 1: B, T;  modes 2/3: Hin, Win, phase (py * 2 + px)
};
__device__ __forceinline__ long map_row(const RowMap& rm, int m) {
  if (rm.mode == 0) return m;
  if (rm.mode == 1) { const int t = m / rm.p0, b = m - t * rm.p0; return (long)b * rm.p1 + t; }
  const int hw = rm.p0 * rm.p1;
  const int n = m / hw, rem = m - n * hw, q = rem / rm.p1, r = rem - q * rm.p1;
  const int oy = 2 * q + (rm.p2 >> 1), ox = 2 * r + (rm.p2 & 1);
  if (rm.mode == 2) return (long)n * 4 * hw + (long)oy * 2 * rm.p1 + ox;
  return (long)n * 4 * hw * 3 + (long)oy * 2 * rm.p1 + ox;   // mode 3: plane 0 of frame n (3 planes of 4 * hw each)
}

// rows [0, 128) x cols [0, ncols) of `tile` -> out_f32[row(m0 + r) * ld_f + c] and / or out_bf[row(m0 + r) * ld_b + c]
// for m0 + r < M and c < nvalid (ncols % 4 == 0; full float4 groups take the vector path).
// Warp w streams rows in passes of 32 / lpr rows; its lanes walk a row 16 bytes apiece.
__device__ __forceinline__ void tile_copy_out(const float* tile, int pitch, int ncols, int nvalid, int m0, int M, float* out_f32,
                                              long ld_f, __nv_bfloat16* out_bf, long ld_b, int tid, RowMap rm = RowMap{0, 0, 0, 0},
                                              int wide = 0) {   // wide: out_bf addresses an fp32 (TF32-rounded) operand buffer
  const bool al_f = out_f32 && ((reinterpret_cast<uintptr_t>(out_f32) & 15u) == 0) && ((ld_f & 3) == 0);
  const bool al_b = out_bf && ((reinterpret_cast<uintptr_t>(out_bf) & 7u) == 0) && ((ld_b & 3) == 0);
  const int rows = min(BM, M - m0);
  if (rm.mode == 3) {     // NCHW fp32: column c is a whole plane apart (3 output channels); one row per warp pass
    const int plane = 4 * rm.p0 * rm.p1;
    for (int r = tid >> 5; r < rows; r += EPI_THREADS / 32) {
      const long orow = map_row(rm, m0 + r);
      if ((tid & 31) < nvalid) out_f32[orow + (long)(tid & 31) * plane] = tile[r * pitch + (tid & 31)];
    }
    return;
  }
  // Narrow tiles (<= 64 columns) would leave most lanes idle at one row per warp pass and make every lane redo the row
  // mapping (three integer divisions for the transposed-conv scatter) for each of its warp's rows: pack 32 / lpr rows into a pass.
  const int lpr = ncols >= 128 ? 32 : (ncols >= 64 ? 16 : (ncols >= 32 ? 8 : 4));   // lanes per row, 4 columns per lane
  const int rpp = 32 / lpr;                                                        // rows per warp pass
  const int lane = tid & 31;
  const int sub = lane / lpr, lane4 = (lane - sub * lpr) << 2;
  for (int r = (tid >> 5) * rpp + sub; r < rows; r += (EPI_THREADS / 32) * rpp) {
    const float* trow = tile + r * pitch;
    const long orow = map_row(rm, m0 + r);
    for (int c = lane4; c < nvalid; c += lpr * 4) {
      const float4 x = *reinterpret_cast<const float4*>(trow + c);
      const bool full = c + 4 <= nvalid;
      if (out_f32) {
        float* o = out_f32 + orow * ld_f + c;
        if (full && al_f) *reinterpret_cast<float4*>(o) = x;
        else { o[0] = x.x; if (c + 1 < nvalid) o[1] = x.y; if (c + 2 < nvalid) o[2] = x.z; if (c + 3 < nvalid) o[3] = x.w; }
      }
      if (out_bf && wide) {
        float* o = reinterpret_cast<float*>(out_bf) + orow * ld_b + c;
        if (full && ((reinterpret_cast<uintptr_t>(out_bf) & 15u) == 0) && ((ld_b & 3) == 0)) {
          *reinterpret_cast<float4*>(o) = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
        } else {
          o[0] = tf32_rn(x.x);
          if (c + 1 < nvalid) o[1] = tf32_rn(x.y);
          if (c + 2 < nvalid) o[2] = tf32_rn(x.z);
          if (c + 3 < nvalid) o[3] = tf32_rn(x.w);
        }
      } else if (out_bf) {
        __nv_bfloat16* o = out_bf + orow * ld_b + c;
        if (full && al_b) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
        else {
          o[0] = __float2bfloat16_rn(x.x);
          if (c + 1 < nvalid) o[1] = __float2bfloat16_rn(x.y);
          if (c + 2 < nvalid) o[2] = __float2bfloat16_rn(x.z);
          if (c + 3 < nvalid) o[3] = __float2bfloat16_rn(x.w);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// row-store helpers: each epilogue thread owns one output row
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void store_f32_row(float* dst, const float (&v)[N], int nvalid) {
  if (nvalid >= N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
    for (int j = 0; j < N; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < nvalid) dst[j] = v[j];
  }
}
template <int N>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const float (&v)[N], int nvalid) {
  if (nvalid >= N && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
    for (int j = 0; j < N; j += 8) {
      uint4 u;
      u.x = pack_bf16x2(v[j], v[j + 1]);
      u.y = pack_bf16x2(v[j + 2], v[j + 3]);
      u.z = pack_bf16x2(v[j + 4], v[j + 5]);
      u.w = pack_bf16x2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(dst + j) = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j < nvalid) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

}  // namespace drm
