// drm_gemm_tf32: the contractions of the hand-scheduled backward passes (bptt.py) on this library's own tensor-core kernel.
//
//   C[M, N] (+)= op(A)[M, K] * op(B)[N, K]^T (+ bias[n])        fp32 in / out, TF32 operands, fp32 accumulate
//
// The three shapes of a linear layer's training step all map onto it (reference: torch.nn.Linear inside
// /root/reference/SequenceModel.py:19-24, DreamerUtils MLPs, WorldModel.py:162-200 / Agent.py:141-153 through autograd):
//   forward re-evaluation   Y  = X W^T + b     A = X [rows, in],  B = W [out, in]
//   input gradient          dX = dY W          A = dY [rows, out], B = W given as [K = out, N = in]   (DRM_GEMM_TRANS_B)
//   weight gradient         dW += dY^T X       A = dY given as [K = rows, M = out], B = X given as [K = rows, N = in]
//                                              (DRM_GEMM_TRANS_A | DRM_GEMM_TRANS_B | DRM_GEMM_ACCUMULATE)
//
// One or two launches, no host synchronisation, capturable in a CUDA graph:
//   * By default both operands first go through pack_tf32_kernel (one launch for both): rewritten K-major with an aligned pitch,
//     ROUNDED to nearest TF32 (cvt.rna), transposed through shared memory when K-last.  Rounding matters: the tensor core truncates a
//     plain fp32 operand (mean relative error -2^-11 per operand, always towards zero), and a recurrence walked over 64 steps
//     compounds that bias; rounded operands are unbiased.  An operand the caller has already rounded (weights, once per backward:
//     drm_pack_tf32) is passed with DRM_GEMM_A_DIRECT / _B_DIRECT and read IN PLACE by TMA whichever way it lies -- K-first rows
//     (128-byte-swizzled boxes of 32 k x rows) or K-last (MN-major: boxes of 32 rows x 32 k, SWIZZLE_128B_ATOM_32B, UMMA descriptors
//     with the transpose bit) -- if its base is 16-byte aligned and its pitch a multiple of 4 floats (else it is packed after all).
//   * gemm_tf32_kernel: 128 x bn output tiles (bn <= 128), a ring of 3 stages (2 when swapped) of 2 k-blocks (32 fp32 = one swizzle row) each,
//     tcgen05.mma kind::tf32 accumulating in TMEM, 16 epilogue warps.  Skinny problems (M <= 64 rows: the per-time-step GEMMs of the
//     recurrences) run swapped -- the WEIGHT rows fill the 128 MMA rows, the few gradient rows are the N dimension -- and their
//     few-row operand needs no pack launch: the epilogue warps, idle under the main loop, round this CTA's K slice of it and write
//     it into shared memory in the swizzled operand layout themselves; the 2 / 4 / 8 CTAs that split a skinny tile's K form a
//     thread-block cluster and reduce through distributed shared memory (each CTA owns a slice of the tile's rows, the peers push
//     their partial sums into it, one cluster barrier, a z-ordered sum).  Other problems with fewer than ~100 tiles are split
//     along K over gridDim.z through global memory: every CTA stores its partial tile, takes a ticket, and the LAST CTA of a tile sums
//     the partials in split order (deterministic whatever the arrival order), applies bias / accumulate and stores C.
#include <algorithm>
#include <cmath>
#include <string>

#include "common.cuh"
#include "internal.h"

namespace drm {
namespace {

constexpr int TM = 128;                 // MMA rows per tile (UMMA M == TMEM lanes)
constexpr int TN = 128;                 // widest N tile
constexpr int TK = 32;                  // fp32 per k-block: one 128-byte swizzle row
constexpr int KPS = 2;                  // k-blocks per pipeline stage
constexpr int STAGES = 3;                // ring slots; swapped problems use two and put the converted operand into the third
constexpr int P_BYTES = TM * TK * 4;    // 16 KB
constexpr int Q_BYTES = TN * TK * 4;    // 16 KB
constexpr int SUB_BYTES = P_BYTES + Q_BYTES;
constexpr int STAGE_BYTES = KPS * SUB_BYTES;
constexpr int BAR_OFF = STAGES * STAGE_BYTES;          // 192 KB
constexpr int QREG_OFF = 2 * STAGE_BYTES;              // swapped skinny problems: this CTA's whole slice of the few-row operand,
constexpr int QREG_BYTES = 64 * 1024;                  //   rounded and swizzled by the epilogue warps (no pack launch)
constexpr int XCHG_OFF = BAR_OFF + 1024;               // cluster split-K: the partial sums the peer CTAs push into this CTA's slice
constexpr int XCHG_BYTES = 32 * 1024;                  //   [split][bn][128 / split] fp32 <= bn * 512 B
constexpr int SMEM_TOTAL = XCHG_OFF + XCHG_BYTES + 1024;   // + alignment slack
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int TILE_PITCH = TN + 4;                     // fp32 words per row of the transposition tile (67.6 KB, reuses the ring)
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
static_assert(QREG_BYTES <= STAGE_BYTES, "the conversion buffer takes the third ring slot");
static_assert(TM * TILE_PITCH * 4 <= BAR_OFF, "output tile must fit in the pipeline's shared memory");

struct GemmArgs {
  CUtensorMap tmP;      // operand on the MMA's M side: [PR, K] K-major, box {32, 128}
  CUtensorMap tmQ;      // operand on the N side: [QR, K] K-major, box {32, bn}
  int PR, QR, nk;       // rows of P / Q, k-blocks
  int bn;               // N per MMA (multiple of 16, <= 128)
  int swap;             // 0: C[p][q];  1: C[q][p]
  int p_mn, q_mn;       // operand lies K-last (MN-major): its map is over [K][rows], box {32 rows, 32 k}
  int accumulate;       // C += result
  const float* bias;    // indexed by the C column, or NULL
  float* C; long ldc;
  float* part;          // split > 1: [split][c_rows_pad][c_cols_pad] partial results in C layout
  long part_ld, part_plane;
  unsigned int* tickets;   // split > 1: one counter per tile, zero on entry, zero again on exit
  const float* q_src;      // q_convert: the N-side operand [QR][K] as the caller gave it (any alignment), pitch q_src_ld
  long q_src_ld;
  int q_convert, K;
  int cluster;             // the `split` CTAs of a tile form a thread-block cluster: 1 = reduce through distributed shared memory, 2 = through global planes
};

// MN-major fp32 operand descriptor.  32-bit MN-major operands have one legal shared-memory layout (cute::UMMA
// Layout_MN_SW128_32B_Atom, layout type SWIZZLE_128B_BASE32B = 1): rows of 128 B = 32 fp32 of the M / N dimension, atoms of 4 such
// rows (4 k), 32-byte chunks XOR-swizzled with the row number mod 4 -- what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
// Here one {32 rows, 32 k} box is 32 consecutive 128-byte rows: the next 4 k are SBO = 512 B further, the next 32 rows of the
// M / N dimension (the next box) LBO = 4096 B further; one MMA (8 k) advances the start address by 1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn_f32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(4096 >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;
  return d;
}

// four consecutive columns of C row r starting at column c (+ bias[c..], + the old value): one float4 when aligned and complete
__device__ __forceinline__ void store4(float* C, long ld, int r, int c, float4 x, int c_cols, const float* bias, bool acc, bool al) {
  float* o = C + (long)r * ld + c;
  if (al && c + 3 < c_cols) {
    if (bias) { const float4 b = *reinterpret_cast<const float4*>(bias + c); x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w; }
    if (acc) { const float4 y = *reinterpret_cast<const float4*>(o); x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }
    *reinterpret_cast<float4*>(o) = x;
  } else {
    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c + i < c_cols) {
        float y = xs[i] + (bias ? bias[c + i] : 0.f);
        if (acc) y += o[i];
        o[i] = y;
      }
  }
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1) gemm_tf32_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);      // [0] TMEM base, [1] "last CTA of the tile" flag
  uint64_t* q_ready = tmem_full + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = (int)blockIdx.x * TM, q0 = (int)blockIdx.y * g.bn;
  const int split = (int)gridDim.z, z = (int)blockIdx.z;
  const int kb0 = (int)(((long)g.nk * z) / split), kb1 = (int)(((long)g.nk * (z + 1)) / split);   // split <= nk: never empty
  const int nkl = kb1 - kb0;
  const int n_st = (nkl + KPS - 1) / KPS;
  const int stages = g.swap ? STAGES - 1 : STAGES;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&g.tmP);
    tma_prefetch_desc(&g.tmQ);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(q_ready, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (g.cluster) cluster_arrive_release();   // phase 1 ("this CTA runs"): waited for just before the first remote store
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel in
  // the stream; nothing below touches global memory before that kernel has completed.  The trigger for the NEXT kernel follows at
  // once (its own prologue then overlaps this kernel; it waits here in turn).
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      const int q_chunks = (g.bn + 31) >> 5;
      const uint32_t tx = (uint32_t)P_BYTES + (g.q_convert ? 0u : (g.q_mn ? (uint32_t)q_chunks * 4096u : (uint32_t)g.bn * TK * 4));
      for (int st = 0; st < n_st; ++st) {
        const int s = st % stages;
        mbar_wait(&empty[s], ((st / stages) & 1) ^ 1u);
        const int n_sub = min(KPS, nkl - st * KPS);
        mbar_expect_tx(&full[s], (uint32_t)n_sub * tx);
        for (int j = 0; j < n_sub; ++j) {
          const int kb = kb0 + st * KPS + j;
          uint8_t* sp = smem + s * STAGE_BYTES + j * SUB_BYTES;
          if (g.p_mn) {
#pragma unroll
            for (int c = 0; c < TM / 32; ++c) tma_load_2d(sp + c * 4096, &g.tmP, p0 + c * 32, kb * TK, &full[s]);
          } else {
            tma_load_2d(sp, &g.tmP, kb * TK, p0, &full[s]);
          }
          if (g.q_convert) {
          } else if (g.q_mn) {
            for (int c = 0; c < q_chunks; ++c) tma_load_2d(sp + P_BYTES + c * 4096, &g.tmQ, q0 + c * 32, kb * TK, &full[s]);
          } else {
            tma_load_2d(sp + P_BYTES, &g.tmQ, kb * TK, q0, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(g.bn, 1) | (g.p_mn ? (1u << 15) : 0u) | (g.q_mn ? (1u << 16) : 0u);
      // descriptor advance per MMA (8 k): 32 bytes along a K-major swizzle row, one 1024-byte group of 8 rows when MN-major
      const uint32_t p_step = g.p_mn ? (1024 >> 4) : 2, q_step = g.q_mn ? (1024 >> 4) : 2;
      if (g.q_convert) mbar_wait(q_ready, 0);
      for (int st = 0; st < n_st; ++st) {
        const int s = st % stages;
        mbar_wait(&full[s], (st / stages) & 1);
        tc_fence_after();
        const int n_sub = min(KPS, nkl - st * KPS);
        for (int j = 0; j < n_sub; ++j) {
          const uint32_t p_addr = smem_u32(smem + s * STAGE_BYTES + j * SUB_BYTES);
          const uint64_t pdesc = g.p_mn ? umma_desc_mn_f32(p_addr) : umma_desc_sw128(p_addr);
          const uint64_t qdesc = g.q_convert ? umma_desc_sw128(smem_u32(smem + QREG_OFF) + (uint32_t)((st * KPS + j) * g.bn * 128))
                                 : (g.q_mn ? umma_desc_mn_f32(p_addr + P_BYTES) : umma_desc_sw128(p_addr + P_BYTES));
#pragma unroll
          for (int k = 0; k < TK / 8; ++k) umma_tf32(tmem, pdesc + p_step * k, qdesc + q_step * k, idesc, (st | j | k) != 0);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int tid = (int)threadIdx.x - 64;
    const int q = warp & 3, part = (warp - 2) >> 2;      // TMEM lane quadrant, 32-column group
    const int row = q * 32 + lane;
    const bool active = part * 32 < g.bn;
    // first destination: C itself, or this split's plane of the partial buffer (C layout, padded: no guards / bias / accumulate)
    const bool direct = split == 1 || g.cluster == 1;
    const int C_rows = g.swap ? g.QR : g.PR, C_cols = g.swap ? g.PR : g.QR;
    float* dst = direct ? g.C : g.part + (long)z * g.part_plane;
    const long ld = direct ? g.ldc : g.part_ld;
    const int c_rows = direct ? C_rows : 0x7fffffff;
    const int c_cols = direct ? C_cols : 0x7fffffff;
    const float* bias = direct ? g.bias : nullptr;
    const bool acc = direct && g.accumulate;

    if (g.q_convert) {
      // this CTA's slice of the few-row operand: global (any alignment) -> round to nearest TF32 -> K-major 128-byte-swizzled
      // k-blocks of bn rows (rows >= QR and k >= K are zeros)
      uint8_t* qreg = smem + QREG_OFF;
      const bool vec = ((reinterpret_cast<uintptr_t>(g.q_src) & 15u) == 0) && ((g.q_src_ld & 3) == 0);
      if (vec) {
        // 16-byte chunks: chunk j of row r = floats [4j, 4j + 4) of this CTA's K slice; consecutive threads take consecutive chunks of
        // a row (coalesced), every load is issued before the first store (at most QREG_BYTES / 16 / 512 = 8 chunks per thread)
        const int cpr = nkl * 8;                                               // chunks per row, <= 256
        const int n_chunk = g.bn * cpr;
        const uint32_t inv = (uint32_t)((0x100000000ull + (uint32_t)cpr - 1u) / (uint32_t)cpr);   // idx / cpr == umulhi(idx, inv) (idx < 4096, 8 <= cpr <= 256)
        constexpr int PER = QREG_BYTES / 16 / EPI_THREADS;
        float4 x[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int idx = tid + i * EPI_THREADS;
          x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < n_chunk) {
            const int r = (int)__umulhi((uint32_t)idx, inv), j = idx - r * cpr;
            const int kg = kb0 * TK + j * 4;
            if (r < g.QR && kg < g.K) {
              const float* src = g.q_src + (long)r * g.q_src_ld + kg;
              if (kg + 3 < g.K) x[i] = *reinterpret_cast<const float4*>(src);
              else { x[i].x = src[0]; if (kg + 1 < g.K) x[i].y = src[1]; if (kg + 2 < g.K) x[i].z = src[2]; }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int idx = tid + i * EPI_THREADS;
          if (idx < n_chunk) {
            const int r = (int)__umulhi((uint32_t)idx, inv), j = idx - r * cpr;
            const int kb = j >> 3, c = j & 7;
            *reinterpret_cast<float4*>(qreg + (long)kb * (g.bn * 128) + r * 128 + ((c ^ (r & 7)) << 4)) =
                make_float4(tf32_rn(x[i].x), tf32_rn(x[i].y), tf32_rn(x[i].z), tf32_rn(x[i].w));
          }
        }
      } else {
        // unaligned source: one float at a time, consecutive threads on consecutive k
        const int kcount = nkl * TK, kbase = kb0 * TK;
        for (int idx = tid; idx < g.bn * kcount; idx += EPI_THREADS) {
          const int r = idx / kcount, kk = idx - r * kcount;
          const int kg = kbase + kk;
          const float xv = (r < g.QR && kg < g.K) ? tf32_rn(g.q_src[(long)r * g.q_src_ld + kg]) : 0.f;
          const int kb = kk >> 5, kin = kk & 31;
          *reinterpret_cast<float*>(qreg + (long)kb * (g.bn * 128) + r * 128 + ((((kin >> 2) ^ (r & 7))) << 4) + ((kin & 3) << 2)) = xv;
        }
      }
      fence_proxy_async();
      epi_bar();
      if (tid == 0) mbar_arrive(q_ready);
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float v[32];
    if (active) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 32), v);
    if (g.cluster == 1) {
      // Cluster split-K (swapped tiles): CTA z owns rows [z * slice, (z + 1) * slice) of the tile.  Every CTA pushes its partial sums
      // of a row into the owner's exchange buffer [source z][q][row in slice] (st.shared::cluster; lanes = consecutive rows), one
      // cluster barrier, then the owner adds the `split` partials in z order: no global round trip, no atomics, no fences.
      const int slice = TM / split;
      float* xchg = reinterpret_cast<float*>(smem + XCHG_OFF);
      cluster_wait_acquire();              // every peer CTA has started: its shared memory may be written
      if (active) {
        const int owner = row / slice, rl = row - owner * slice;
        uint32_t rbase;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(rbase) : "r"(smem_u32(xchg)), "r"((uint32_t)owner));
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qq = part * 32 + j;
          if (qq < g.bn)
            asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(rbase + (uint32_t)(((z * g.bn + qq) * slice + rl) << 2)), "f"(v[j]) : "memory");
        }
      }
      cluster_arrive_release();
      cluster_wait_acquire();
      for (int idx = tid; idx < g.bn * slice; idx += EPI_THREADS) {
        const int qq = idx / slice, rl = idx - qq * slice;
        const int p = p0 + z * slice + rl, cq = q0 + qq;
        if (p >= c_cols || cq >= c_rows) continue;
        float sum = xchg[qq * slice + rl];
        for (int zz = 1; zz < split; ++zz) sum += xchg[(zz * g.bn + qq) * slice + rl];
        if (bias) sum += bias[p];
        float* o = dst + (long)cq * ld + p;
        if (acc) sum += *o;
        *o = sum;
      }
    } else if (g.swap) {
      // C[q][p]: the 32 lanes of a warp hold 32 consecutive p for every q -> coalesced stores straight from registers
      const int p = p0 + row;
      if (active && p < c_cols) {
        const float bp = bias ? bias[p] : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qq = q0 + part * 32 + j;
          if (part * 32 + j < g.bn && qq < c_rows) {
            float* o = dst + (long)qq * ld + p;
            float x = v[j] + bp;
            if (acc) x += *o;
            *o = x;
          }
        }
      }
    } else {
      float* tile = reinterpret_cast<float*>(smem);
      if (active) {
        float4* t4 = reinterpret_cast<float4*>(tile + row * TILE_PITCH + part * 32);
#pragma unroll
        for (int j = 0; j < 32; j += 4) t4[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      epi_bar();
      const int c4n = g.bn >> 2;
      const bool al = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && ((ld & 3) == 0);
      for (int idx = tid; idx < TM * c4n; idx += EPI_THREADS) {
        const int r = idx / c4n, c = (idx - r * c4n) << 2;
        if (p0 + r >= c_rows || q0 + c >= c_cols) continue;
        store4(dst, ld, p0 + r, q0 + c, *reinterpret_cast<const float4*>(tile + r * TILE_PITCH + c), c_cols, bias, acc, al);
      }
    }
    if (g.cluster == 2) {
      // Cluster split-K (128-row tiles): every CTA has stored its partial tile into its plane of the global buffer (L2); one cluster
      // barrier (release / acquire orders those stores for the peers), then CTA z sums rows [z * slice, (z + 1) * slice) of all planes
      // in z order -- each CTA reads one tile's worth of partials whatever the split, all loads of a row group in flight together.
      cluster_wait_acquire();              // phase 1 (start-up) is over
      cluster_arrive_release();
      cluster_wait_acquire();
      const int slice = (TM + split - 1) / split, CN4 = g.bn >> 2;        // (any split of 2 ... 8: the last slice may be short)
      const bool al = ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0) && ((g.ldc & 3) == 0);
      for (int idx = tid; idx < slice * CN4; idx += EPI_THREADS) {
        const int r = idx / CN4, c = (idx - r * CN4) << 2;
        const int row = p0 + z * slice + r, col = q0 + c;
        if (z * slice + r >= TM || row >= C_rows || col >= C_cols) continue;
        const float* src = g.part + (long)row * g.part_ld + col;
        float4 y[8];
#pragma unroll
        for (int zz = 0; zz < 8; ++zz)
          if (zz < split) y[zz] = __ldcg(reinterpret_cast<const float4*>(src + (long)zz * g.part_plane));
        float4 sum = y[0];
#pragma unroll
        for (int zz = 1; zz < 8; ++zz)
          if (zz < split) { sum.x += y[zz].x; sum.y += y[zz].y; sum.z += y[zz].z; sum.w += y[zz].w; }
        store4(g.C, g.ldc, row, col, sum, C_cols, g.bias, g.accumulate != 0, al);
      }
    } else if (!direct) {          // (split > 1 without a cluster: swapped tiles whose K slice does not fit the conversion buffer)
      // ticket: the last CTA of this tile to finish sums the `split` partial planes in z order and stores C
      uint32_t* last_flag = tmem_slot + 1;
      const int tile_id = (int)blockIdx.x + (int)gridDim.x * (int)blockIdx.y;
      __threadfence();
      epi_bar();
      if (tid == 0) *last_flag = (atomicAdd(&g.tickets[tile_id], 1u) == (unsigned)(split - 1)) ? 1u : 0u;
      epi_bar();
      if (*last_flag) {
        __threadfence();
        const int R0 = g.swap ? q0 : p0, C0 = g.swap ? p0 : q0;
        const int RN = g.swap ? g.bn : TM, CN4 = (g.swap ? TM : g.bn) >> 2;
        const bool al = ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0) && ((g.ldc & 3) == 0);
        for (int idx = tid; idx < RN * CN4; idx += EPI_THREADS) {
          const int r = idx / CN4, c = (idx - r * CN4) << 2;
          if (R0 + r >= C_rows || C0 + c >= C_cols) continue;
          const float* src = g.part + (long)(R0 + r) * g.part_ld + C0 + c;
          float4 s = __ldcg(reinterpret_cast<const float4*>(src));
#pragma unroll 8
          for (int zz = 1; zz < split; ++zz) {
            const float4 y = __ldcg(reinterpret_cast<const float4*>(src + (long)zz * g.part_plane));
            s.x += y.x; s.y += y.y; s.z += y.z; s.w += y.w;
          }
          store4(g.C, g.ldc, R0 + r, C0 + c, s, C_cols, g.bias, g.accumulate != 0, al);
        }
        if (tid == 0) g.tickets[tile_id] = 0u;       // ready for the next call on this workspace
      }
    }
  }
  if (g.cluster && warp < 2) {       // the producer / MMA warps take part in the epilogue's cluster barriers
    __syncwarp();
    cluster_wait_acquire();
    cluster_arrive_release();
    cluster_wait_acquire();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TN);
}

// out[r][k] = tf32_rn(trans ? in[k][r] : in[r][k]) for r < rows, k < K; zero for K <= k < ldo.  32 x 32 tiles through shared memory.
struct PackJob {
  const float* in; long ld; int rows, K, trans;
  float* out; int ldo;
  int tiles_r, tiles_k;
};
__global__ void __launch_bounds__(256) pack_tf32_kernel(const PackJob ja, const PackJob jb, int tiles_a) {
  __shared__ float t[32][33];
  asm volatile("griddepcontrol.wait;\n" ::: "memory");                 // (programmatic dependent launch, as in gemm_tf32_kernel)
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  const bool second = (int)blockIdx.x >= tiles_a;
  const PackJob& j = second ? jb : ja;
  const int id = (int)blockIdx.x - (second ? tiles_a : 0);
  const int tr = id / j.tiles_k, tk = id - tr * j.tiles_k;
  const int r0 = tr * 32, k0 = tk * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (j.trans) {
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int k = k0 + i, r = r0 + tx;
      t[i][tx] = (k < j.K && r < j.rows) ? j.in[(long)k * j.ld + r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, k = k0 + tx;
      if (r < j.rows && k < j.ldo) j.out[(long)r * j.ldo + k] = tf32_rn(t[tx][i]);
    }
  } else {
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, k = k0 + tx;
      if (r < j.rows && k < j.ldo) j.out[(long)r * j.ldo + k] = k < j.K ? tf32_rn(j.in[(long)r * j.ld + k]) : 0.f;
    }
  }
}

constexpr long TICKET_BYTES = 1024;     // 256 tile counters at the start of the workspace
constexpr long PART_MAX_BYTES = 148L * TM * TN * 4;   // split * tiles <= 148 CTAs of at most one 128 x 128 plane each

struct Plan {
  int swap, PR, QR, bn, nk, split, mt, nt;
  bool pack_a, pack_b, q_convert;
  int cluster;                     // 0: none, 1: swapped tiles reduced through DSMEM, 2: 128-row tiles reduced through global planes
  long kp;                         // packed pitch (elements)
  long off_a, off_b, off_part;     // workspace offsets (bytes)
  long part_ld, part_plane;        // elements
  long total;                      // workspace bytes
};

inline long align256(long x) { return (x + 255) & ~255L; }
inline bool aligned_operand(const float* p, long ld) { return (((uintptr_t)p & 15u) == 0) && ((ld & 3) == 0); }

// a_direct / b_direct: the operand is read in place by TMA (caller's flag and aligned); a_kfirst: A lies [M][K]
Plan make_plan(int M, int N, int K, bool a_direct, bool b_direct, bool a_kfirst, int force_split = 0) {
  Plan p{};
  p.swap = M <= 64 ? 1 : 0;
  p.PR = p.swap ? N : M;
  p.QR = p.swap ? M : N;
  p.bn = p.QR >= TN ? TN : round_up(p.QR, 16);
  p.nk = ceil_div(K, TK);
  p.mt = ceil_div(p.PR, TM);
  p.nt = ceil_div(p.QR, p.bn);
  const int tiles = p.mt * p.nt;
  p.split = 1;
  p.cluster = 0;
  if (p.swap) {
    if (tiles < 100 && p.nk >= 4) {
      // (fallback for swapped tiles that cannot use the cluster path below: ticket reduction by the last CTA of a tile)
      const double per_plane_us = (double)TM * p.bn * 4 / 150e3;
      const int s = (int)(std::sqrt(p.nk * 0.2 / per_plane_us) + 0.5);
      p.split = std::max(1, std::min({s, 148 / tiles, p.nk / 2, 32}));
    }
  } else {
    // 128-row tiles: 1 ... 8 CTAs per tile as a cluster.  Measured (profiles/README.md): ~0.22 us per k-block of main loop, ~2 us
    // for the store / cluster barrier / slice reduction once K is split
    // (a cluster must fit one GPC, ~16 - 18 SMs at one CTA per SM: clusters of 5 - 8 CTAs only when there are few of them)
    auto fits = [tiles](int sc) { return sc * tiles <= 148 && (sc <= 4 ? true : tiles <= 8); };
    double best = p.nk * 0.22;
    for (int sc = 2; sc <= 8; ++sc) {
      if (!fits(sc) || p.nk / sc < 2) continue;
      const double cost = ceil_div(p.nk, sc) * 0.22 + 2.0;
      if (cost < best) { best = cost; p.split = sc; }
    }
    if (force_split > 0) p.split = force_split <= 8 && force_split * tiles <= 148 && force_split <= p.nk ? force_split : p.split;   // profiling only
    if (p.split > 1) p.cluster = 2;
  }
  // swapped problems: the few-row operand (A) is rounded and swizzled inside the kernel when a CTA's K slice of it fits QREG_BYTES
  p.q_convert = false;
  if (p.swap && !a_direct && a_kfirst && tiles <= 148) {
    const int kb_fit = QREG_BYTES / (p.bn * 128);                 // k-blocks of bn rows that fit
    const int s_min = ceil_div(p.nk, kb_fit);
    if (s_min <= std::min(148 / tiles, p.nk)) { p.q_convert = true; p.split = std::max(p.split, s_min); }
  }
  // swapped tiles of up to 64 rows of C: 2 / 4 / 8 CTAs per tile as a cluster, reduction through distributed shared memory
  if (p.swap && p.nt == 1 && p.bn * 512 <= XCHG_BYTES && p.nk >= 4) {
    const int sc = p.nk >= 16 ? 8 : (p.nk >= 8 ? 4 : 2);
    const bool convert_ok = !(p.swap && !a_direct && a_kfirst) || ceil_div(p.nk, sc) * p.bn * 128 <= QREG_BYTES;
    if (sc * tiles <= 148 && convert_ok) {
      p.cluster = 1;
      p.split = sc;
      p.q_convert = !a_direct && a_kfirst;
    }
  }
  p.pack_a = !a_direct && !p.q_convert; p.pack_b = !b_direct;
  p.kp = round_up(K, 4);
  long off = TICKET_BYTES;
  p.off_a = off; if (p.pack_a) off = align256(off + (long)M * p.kp * 4);
  p.off_b = off; if (p.pack_b) off = align256(off + (long)N * p.kp * 4);
  p.off_part = off;
  if (p.split > 1 && p.cluster != 1) {
    const long c_rows_pad = p.swap ? (long)p.nt * p.bn : (long)p.mt * TM;
    const long c_cols_pad = p.swap ? (long)p.mt * TM : (long)p.nt * p.bn;
    p.part_ld = c_cols_pad;
    p.part_plane = c_rows_pad * c_cols_pad;
    off = align256(off + (long)p.split * p.part_plane * 4);
  }
  p.total = off;
  return p;
}

PackJob pack_job(const float* in, long ld, int rows, int K, int trans, float* out, long ldo) {
  PackJob j{in, ld, rows, K, trans, out, (int)ldo, ceil_div(rows, 32), ceil_div((int)ldo, 32)};
  return j;
}

}  // namespace
}  // namespace drm

using namespace drm;

extern "C" int64_t drm_gemm_tf32_workspace_bytes(int32_t M, int32_t N, int32_t K) {
  if (M < 1 || N < 1 || K < 1) return 0;
  const long kp = round_up(K, 4);
  return TICKET_BYTES + align256((long)M * kp * 4) + align256((long)N * kp * 4) + align256(PART_MAX_BYTES);   // whatever the flags
}

extern "C" int drm_pack_tf32(int32_t rows, int32_t K, const float* in, int64_t ld, int32_t trans, float* out, int64_t ld_out,
                             void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(in && out, DRM_ERR_ARG, "drm_pack_tf32: NULL argument");
  DRM_REQUIRE(rows >= 1 && K >= 1 && ld_out >= K && (ld_out & 3) == 0, DRM_ERR_SHAPE, "drm_pack_tf32: bad shape (ld_out must be a multiple of 4 >= K)");
  DRM_REQUIRE(ld >= (trans ? rows : K), DRM_ERR_SHAPE, "drm_pack_tf32: ld too small");
  const PackJob j = pack_job(in, ld, rows, K, trans, out, ld_out);
  pack_tf32_kernel<<<j.tiles_r * j.tiles_k, 256, 0, (cudaStream_t)stream>>>(j, j, j.tiles_r * j.tiles_k);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_gemm_tf32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                             int64_t ldc, const float* bias, int32_t flags, void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(A && B && C, DRM_ERR_ARG, "drm_gemm_tf32: NULL argument");
  DRM_REQUIRE(M >= 1 && N >= 1 && K >= 1, DRM_ERR_SHAPE, "drm_gemm_tf32: bad shape");
  const bool ta = flags & DRM_GEMM_TRANS_A, tb = flags & DRM_GEMM_TRANS_B;
  DRM_REQUIRE(lda >= (ta ? M : K) && ldb >= (tb ? N : K) && ldc >= N, DRM_ERR_SHAPE, "drm_gemm_tf32: leading dimension too small");
  const Plan p = make_plan(M, N, K, (flags & DRM_GEMM_A_DIRECT) && aligned_operand(A, lda),
                           (flags & DRM_GEMM_B_DIRECT) && aligned_operand(B, ldb), !ta, (flags >> 8) & 0xff);
  DRM_REQUIRE(p.total <= drm_gemm_tf32_workspace_bytes(M, N, K), DRM_ERR_SHAPE, "drm_gemm_tf32: internal: workspace bound");
  DRM_REQUIRE(workspace && workspace_bytes >= p.total && ((uintptr_t)workspace & 255u) == 0, DRM_ERR_ARG,
              "drm_gemm_tf32: workspace missing, misaligned (256 bytes) or smaller than drm_gemm_tf32_workspace_bytes");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  // operands as the kernel sees them: in place (K-first rows or K-last), or the packed K-first copy
  const float* a_k = A; long a_ld = lda; bool a_mn = ta;
  const float* b_k = B; long b_ld = ldb; bool b_mn = tb;
  if (p.pack_a || p.pack_b) {
    float* a_out = reinterpret_cast<float*>(ws + p.off_a);
    float* b_out = reinterpret_cast<float*>(ws + p.off_b);
    PackJob ja = pack_job(A, lda, M, K, ta, a_out, p.kp), jb = pack_job(B, ldb, N, K, tb, b_out, p.kp);
    // (an operand that is not packed has tiles = 0 below; its job is never read)
    int tiles_a = ja.tiles_r * ja.tiles_k, tiles_b = jb.tiles_r * jb.tiles_k;
    if (!p.pack_a) { ja = jb; tiles_a = 0; }
    if (!p.pack_b) tiles_b = 0;
    cudaLaunchConfig_t pc = {};
    pc.gridDim = dim3(tiles_a + tiles_b);
    pc.blockDim = dim3(256);
    pc.stream = st;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pa[0].val.programmaticStreamSerializationAllowed = 1;
    pc.attrs = pa;
    pc.numAttrs = (flags & DRM_GEMM_NO_PDL) ? 0 : 1;
    DRM_CUDA(cudaLaunchKernelEx(&pc, pack_tf32_kernel, ja, jb, tiles_a));
    DRM_LAUNCH_CHECK();
    if (p.pack_a) { a_k = a_out; a_ld = p.kp; a_mn = false; }
    if (p.pack_b) { b_k = b_out; b_ld = p.kp; b_mn = false; }
  }

  GemmArgs g{};
  const float* P = p.swap ? b_k : a_k; const long p_ld = p.swap ? b_ld : a_ld; const bool p_mn = p.swap ? b_mn : a_mn;
  const float* Q = p.swap ? a_k : b_k; const long q_ld = p.swap ? a_ld : b_ld; const bool q_mn = p.swap ? a_mn : b_mn;
  // K-major: map over [rows][K], box {32 k, rows};  MN-major: map over [K][rows], box {32 rows, 32 k}
  if (int rc = p_mn ? make_tmap_f32_mn(&g.tmP, P, (uint64_t)K, (uint64_t)p.PR, (uint64_t)p_ld)
                    : make_tmap_op_2d(&g.tmP, P, (uint64_t)p.PR, (uint64_t)K, (uint64_t)p_ld, TM, 1)) return rc;
  if (p.q_convert) {
    g.tmQ = g.tmP;       // unused
    g.q_src = A; g.q_src_ld = lda;
  } else if (int rc = q_mn ? make_tmap_f32_mn(&g.tmQ, Q, (uint64_t)K, (uint64_t)p.QR, (uint64_t)q_ld)
                           : make_tmap_op_2d(&g.tmQ, Q, (uint64_t)p.QR, (uint64_t)K, (uint64_t)q_ld, (uint32_t)p.bn, 1)) return rc;
  g.q_convert = p.q_convert ? 1 : 0; g.K = K;
  g.PR = p.PR; g.QR = p.QR; g.nk = p.nk; g.bn = p.bn; g.swap = p.swap;
  g.p_mn = p_mn; g.q_mn = q_mn;
  g.accumulate = (flags & DRM_GEMM_ACCUMULATE) ? 1 : 0;
  g.bias = bias; g.C = C; g.ldc = ldc;
  g.part = (p.split > 1 && p.cluster != 1) ? reinterpret_cast<float*>(ws + p.off_part) : nullptr;
  g.cluster = p.cluster;
  g.part_ld = p.part_ld; g.part_plane = p.part_plane;
  g.tickets = reinterpret_cast<unsigned int*>(ws);
  DRM_REQUIRE(p.split == 1 || p.cluster || p.mt * p.nt <= (int)(TICKET_BYTES / 4), DRM_ERR_SHAPE, "drm_gemm_tf32: internal: too many split tiles");

  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.mt, p.nt, p.split);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.cluster) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = (unsigned)p.split;
    ++na;
  }
  if (!(flags & DRM_GEMM_NO_PDL)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel, g));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
