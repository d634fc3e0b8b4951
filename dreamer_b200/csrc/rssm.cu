// RSSM on tcgen05: packed bf16 weights, the per-stage fused GEMM launches and the imagination
// rollout driver (Dreamer.dream_episodes, Dreamer.py:143-175).
//
// HBM layout
//   state buffer S[2]  bf16 [Mp, KS]  columns  [ z : R*C | a : 64 (A used) | h : DP ]   (Mp = B up to 128)
//       -> one TMA descriptor serves every stage: the GRU reads k-blocks z|a (x part) then h,
//          the prior reads h, the [h, z] heads read z then h.  Ping-pong between steps.
//   Y1, Y2             bf16 [6 * Mp, 256]    hidden activations; slot 0 = prior, 1.. = heads
//   packed weights     bf16, K-major, columns permuted to the state layout, rows grouped per tile
//       Wgru [tiles * 3U, 1088 + DP]  rows of tile j = [r | z | n] of hidden units j*U .. j*U+U-1 (packed for U = 32 and 64)
//       Wp1 [bn, DP]  Wp2 [bn, 256]  Wp3 [R*C, 256]
//       Wh1 [5 * bn, R*C + DP]  Wh2 [5 * bn, 256]  Wh3 [5 * 256, 256]
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "epilogues.cuh"
#include "internal.h"

namespace drm {

// ------------------------------------------------------------------------------------------
// small support kernels
// ------------------------------------------------------------------------------------------
__global__ void pack_matrix_kernel(__nv_bfloat16* __restrict__ dst, int ld_dst, int dst_row0, int nrows, int dst_col0,
                                   int ncols, const float* __restrict__ src, int ld_src, const int* __restrict__ row_map,
                                   const int* __restrict__ col_map, int wide = 0) {
  const long total = (long)nrows * ncols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols), c = (int)(i % ncols);
    const int sr = row_map[r], sc = col_map[c];
    const float v = (sr >= 0 && sc >= 0) ? src[(long)sr * ld_src + sc] : 0.f;
    const long o = (long)(dst_row0 + r) * ld_dst + dst_col0 + c;
    if (wide) reinterpret_cast<float*>(dst)[o] = tf32_rn(v);   // TF32 mode: the operand buffer holds fp32
    else dst[o] = __float2bfloat16_rn(v);
  }
}
__global__ void pack_vector_kernel(float* __restrict__ dst, int n, const float* __restrict__ src, const int* __restrict__ map,
                                   float fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = map[i] >= 0 ? src[map[i]] : fill;
}

// fp32 rows -> bf16 columns of a state buffer (and optional fp32 copies into strided outputs).
__global__ void pack_state_kernel(__nv_bfloat16* __restrict__ S, int ld_s, int col0, const float* __restrict__ src,
                                  long ld_src, int ncols, int N, float* __restrict__ copy, long ld_copy, int wide = 0) {
  const long total = (long)N * ncols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols), c = (int)(i % ncols);
    const float v = src[(long)r * ld_src + c];
    if (wide) reinterpret_cast<float*>(S)[(long)r * ld_s + col0 + c] = tf32_rn(v);
    else S[(long)r * ld_s + col0 + c] = __float2bfloat16_rn(v);
    if (copy) copy[(long)r * ld_copy + c] = v;
  }
}
__global__ void f32_to_bf16_pad_kernel(__nv_bfloat16* __restrict__ dst, int ld_dst, const float* __restrict__ src, int rows,
                                       int cols) {
  const long total = (long)rows * ld_dst;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst), c = (int)(i % ld_dst);
    dst[i] = __float2bfloat16_rn(c < cols ? src[(long)r * cols + c] : 0.f);
  }
}

// debug probe buffer: [DRM_STAGE_COUNT][16] u64 (drm_debug_timeline)
static unsigned long long* g_timeline = nullptr;

// Runtime options (drm_set_option): alternative kernel paths, all producing the same results (profiles/README.md).
struct Options {
  int ln_cluster = 1;  // LN stages of small grids split over clusters of 4 CTAs
  int gru_ksplit = 1;  // single-m-tile grids split the GRU tile's K range over a 2-CTA cluster
  int conv_persist = 1;  // narrow conv layers on the persistent GEMM
  int conv_kps = 0;      // implicit-GEMM convs: k-blocks per pipeline stage (0 = as many as fit in 48 KB)
  int conv_chunk = 512;  // frames per conv chunk of an observe workspace (read when the workspace is created)
  int conv_implicit = 1; // conv layers as implicit GEMMs fed by im2col-mode TMA loads (no patch matrix)
  int gru_pair = -1;   // GRU stage on CTA pairs (cta_group::2 M = 256 MMAs, half the weight tile per SM): -1 auto (large grids), 0, 1
  int small_a = 1;     // stages with <= 32 rows load 32-row A boxes
  int gru_u = 0;       // 0: automatic GRU tile width, else 32 / 64
  int gru_band = 0;    // CTA-pair GRU kernel: m-tiles per band of the tile order (0 = 16)
  int persist = 1;     // rollouts: one persistent kernel for the whole horizon when the schedule fits the machine (rollout_persist.cuh)
};
static Options& opts() {
  static Options o;
  return o;
}

static int grid_for(long total, int threads = 256) {
  long b = (total + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

template <class Epi>
static int launch_gemm(const GemmCommon& g, const typename Epi::Params& ep, dim3 grid, cudaStream_t st, int stage = DRM_STAGE_OTHER) {
  using SL = GemmSmem<Epi::B_ROWS_MAX, Epi::STAGES, kps_of<Epi>::value>;
  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(fused_gemm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL));
    attr_set = true;
  }
  profile_begin(stage, st);
  if (g_timeline) {
    const_cast<GemmCommon&>(g).timeline = g_timeline + 16 * stage;
    const_cast<GemmCommon&>(g).cta_times = g_timeline + 16 * DRM_STAGE_COUNT + 1024 * stage;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SL::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (Epi::CLUSTER_N > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;   // z: four column quarters of one LN tile
      attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = Epi::CLUSTER_N;
      ++na;
    }
    if (!profile_on()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this prologue with the previous stage's tail
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    DRM_CUDA(cudaLaunchKernelEx(&cfg, fused_gemm_kernel<Epi>, g, ep));
  }
  profile_end(stage, st);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

}  // namespace drm
#include "gru_pair.cuh"
#include "gru_ksplit.cuh"
namespace drm {

template <int U>
static int launch_gru_ksplit(const GemmCommon& g, const typename EpiGru<U>::Params& ep, int mt, int tiles, cudaStream_t st) {
  using SL = GruKsSmem<U>;
  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(gru_ksplit_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL));
    attr_set = true;
  }
  profile_begin(DRM_STAGE_GRU, st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(mt, tiles, 2);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SL::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 2;
  ++na;
  if (!profile_on()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, gru_ksplit_kernel<U>, g, ep));
  profile_end(DRM_STAGE_GRU, st);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

template <int U>
static int launch_gru_pair(const GemmCommon& g, const typename EpiGru<U>::Params& ep, int mt, int tiles, cudaStream_t st) {
  using SL = GruPairSmem<U>;
  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(gru_pair_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::TOTAL));
    attr_set = true;
  }
  profile_begin(DRM_STAGE_GRU, st);
  if (g_timeline) const_cast<GemmCommon&>(g).cta_times = g_timeline + 16 * DRM_STAGE_COUNT + 1024 * DRM_STAGE_GRU;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(round_up(mt, 2), tiles);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SL::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
  ++na;
  if (!profile_on()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, gru_pair_kernel<U>, g, ep));
  profile_end(DRM_STAGE_GRU, st);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

// ------------------------------------------------------------------------------------------
// packed model
// ------------------------------------------------------------------------------------------
enum Src {
  S_GRU_WIH, S_GRU_WHH, S_GRU_BIH, S_GRU_BHH,
  S_MLP0,  // + 10 * mlp + {w0,b0,g0,be0,w1,b1,g1,be1,w2,b2};  mlp: 0 prior 1 reward 2 cont 3 actor 4 critic 5 target
  S_MU_W = S_MLP0 + 60, S_MU_B, S_LS_W, S_LS_B, S_BK_REW, S_BK_CRIT, S_COUNT
};
enum { HS_REWARD = 0, HS_CONT = 1, HS_ACTOR = 2, HS_CRITIC = 3, HS_TARGET = 4 };
enum : unsigned { HAVE_GRU = 1u << 8, HAVE_PRIOR = 1u << 9 };  // bits 0..4 = head slots

struct MatOp { __nv_bfloat16* dst; int ld_dst, row0, nrows, col0, ncols, src, ld_src; int *row_map, *col_map; };
struct VecOp { float* dst; int n, src; int* map; float fill; };

}  // namespace drm

using namespace drm;

struct drm_rssm {
  drm_dims d;
  int wide = 0;                  // 1: TF32 mode -- every operand buffer below (declared bf16) holds fp32, tensor maps have 32-element boxes
  int ZP, DP, KS, KG, KH;        // state layout: z cols, padded h cols, state pitch, GRU K, head-L1 K
  int gru_tiles2[2];             // GRU n-tiles for U = 32 / 64 (both layouts are packed; the launch picks by grid size)
  int bnp1, bnp2, bnh1, bnh2;    // padded hidden widths (multiples of 16)
  __nv_bfloat16 *Wgru2[2], *Wp1, *Wp2, *Wp3, *Wh1, *Wh2, *Wh3;
  float *b_ih, *b_hh;            // [3D]
  float *p1_b, *p1_g, *p1_be, *p2_b, *p2_g, *p2_be, *p3_b;
  float *h1_b, *h1_g, *h1_be, *h2_b, *h2_g, *h2_be, *h3_b;  // [5 * bn]
  float *bk_rew, *bk_crit;       // [NB]
  CUtensorMap tmWp1q, tmWp2q, tmWh1q, tmWh2q;   // box rows 64: the cluster-of-4 LN stage
  CUtensorMap tmWp3h;                           // box rows 128: half-width categorical tiles for small grids
  CUtensorMap tmWh3a;                           // box rows 32: the actor's mu / log-sigma rows (persistent rollout kernel)
  CUtensorMap tmWgruQ[2];                      // box rows U / 2: the CTA-pair GRU kernel stages gate blocks and n halves separately
  CUtensorMap tmWgru2[2], tmWp1, tmWp2, tmWp3, tmWh1, tmWh2, tmWh3;
  std::vector<MatOp> mat_ops;
  std::vector<VecOp> vec_ops;
  std::vector<void*> allocs;
  bool packed, has_critic;
  unsigned have;   // which weight groups have been packed (HAVE_* | 1 << HS_*)
};

constexpr int SMALL_A_ROWS = 32;   // box rows of the short A tensor maps (see small_a)

struct drm_persist;   // schedule + hand-over counters of the persistent rollout kernel (rollout_persist.cuh)

struct drm_rollout {
  drm_rssm* m;
  drm_persist* ps = nullptr;
  int B, H, Mp;
  __nv_bfloat16* S[2];
  __nv_bfloat16 *Y1, *Y2;
  CUtensorMap tmS[2], tmY1, tmY2;
  CUtensorMap tmS_s[2], tmY1_s, tmY2_s;           // short-box twins (SMALL_A_ROWS rows)
  std::vector<void*> allocs;
};

namespace drm {

template <class T>
static int dev_alloc(std::vector<void*>& bag, T** out, size_t count) {
  void* p = nullptr;
  DRM_CUDA(cudaMalloc(&p, count * sizeof(T) + 16));
  DRM_CUDA(cudaMemset(p, 0, count * sizeof(T) + 16));
  bag.push_back(p);
  *out = static_cast<T*>(p);
  return DRM_OK;
}
static int upload_map(std::vector<void*>& bag, const std::vector<int>& h, int** out) {
  int* d = nullptr;
  if (int rc = dev_alloc(bag, &d, h.size() ? h.size() : 1)) return rc;
  if (!h.empty()) DRM_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
  *out = d;
  return DRM_OK;
}
static std::vector<int> iota_lim(int n, int limit, int offset = 0) {  // i -> offset + i if i < limit else -1
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = i < limit ? offset + i : -1;
  return v;
}

static int add_mat(drm_rssm* m, __nv_bfloat16* dst, int ld_dst, int row0, int col0, int src, int ld_src,
                   const std::vector<int>& rows, const std::vector<int>& cols) {
  MatOp op{dst, ld_dst, row0, (int)rows.size(), col0, (int)cols.size(), src, ld_src, nullptr, nullptr};
  if (int rc = upload_map(m->allocs, rows, &op.row_map)) return rc;
  if (int rc = upload_map(m->allocs, cols, &op.col_map)) return rc;
  m->mat_ops.push_back(op);
  return DRM_OK;
}
static int add_vec(drm_rssm* m, float* dst, int src, const std::vector<int>& map, float fill = 0.f) {
  VecOp op{dst, (int)map.size(), src, nullptr, fill};
  if (int rc = upload_map(m->allocs, map, &op.map)) return rc;
  m->vec_ops.push_back(op);
  return DRM_OK;
}

}  // namespace drm

#define RC(x)            \
  do {                   \
    if (int rc__ = (x)) return rc__; \
  } while (0)

extern "C" int drm_rssm_create(const drm_dims* dims, drm_rssm** out) { return drm_rssm_create_ex(dims, DRM_PRECISION_BF16, out); }

extern "C" int drm_rssm_create_ex(const drm_dims* dims, int32_t precision, drm_rssm** out) {
  RC(check_arch());
  DRM_REQUIRE(dims && out, DRM_ERR_ARG, "drm_rssm_create: NULL argument");
  DRM_REQUIRE(precision == DRM_PRECISION_BF16 || precision == DRM_PRECISION_TF32, DRM_ERR_ARG, "drm_rssm_create_ex: unknown precision");
  const drm_dims d = *dims;
  DRM_REQUIRE(d.C == 32, DRM_ERR_SHAPE, "drm_rssm_create: latent classes C must be 32");
  DRM_REQUIRE(d.R > 0 && (d.R * d.C) % 256 == 0, DRM_ERR_SHAPE, "drm_rssm_create: R * C must be a multiple of 256");
  DRM_REQUIRE(d.D >= 16 && d.D <= 16384, DRM_ERR_SHAPE, "drm_rssm_create: D out of range");
  DRM_REQUIRE(d.A >= 1 && d.A <= 16, DRM_ERR_SHAPE, "drm_rssm_create: A must be in [1, 16]");
  DRM_REQUIRE(d.NB >= 2 && d.NB <= 256, DRM_ERR_SHAPE, "drm_rssm_create: NB must be in [2, 256]");
  for (int i = 0; i < 2; ++i)
    DRM_REQUIRE(d.h_prior[i] >= 1 && d.h_prior[i] <= 256 && d.h_head[i] >= 1 && d.h_head[i] <= 256, DRM_ERR_SHAPE,
                "drm_rssm_create: MLP hidden sizes must be in [1, 256]");
  drm_rssm* m = new drm_rssm();
  m->d = d;
  m->wide = precision == DRM_PRECISION_TF32 ? 1 : 0;
  const int wd = m->wide;
  m->packed = false;
  m->has_critic = false;
  m->have = 0;
  m->ZP = d.R * d.C;
  m->DP = round_up(d.D, 64);
  m->KS = m->ZP + 64 + m->DP;
  m->KG = m->KS;              // GRU consumes [z | a | h]
  m->KH = m->ZP + m->DP;      // head layer 1 consumes [z | h]
  m->gru_tiles2[0] = ceil_div(d.D, 32);
  m->gru_tiles2[1] = ceil_div(d.D, 64);
  m->bnp1 = round_up(d.h_prior[0], 32);   // LN tiles: MMA N rounded to 32 so pad columns are exact zeros
  m->bnp2 = round_up(d.h_prior[1], 32);
  m->bnh1 = round_up(d.h_head[0], 32);
  m->bnh2 = round_up(d.h_head[1], 32);
  const int D = d.D, ZP = m->ZP, DP = m->DP, A = d.A, NB = d.NB;
  auto& bag = m->allocs;
  int rc = DRM_OK;
#define TRY(x) if (rc == DRM_OK) rc = (x)
  for (int v = 0; v < 2; ++v) TRY(dev_alloc(bag, &m->Wgru2[v], ((size_t)m->gru_tiles2[v] * 3 * (32 << v) * m->KG) << wd));
  TRY(dev_alloc(bag, &m->Wp1, ((size_t)256 * DP) << wd));                 // LN-stage weights: 256-row slots (zero padded)
  TRY(dev_alloc(bag, &m->Wp2, ((size_t)256 * 256) << wd));
  TRY(dev_alloc(bag, &m->Wp3, ((size_t)ZP * 256) << wd));
  TRY(dev_alloc(bag, &m->Wh1, ((size_t)MAX_HEADS * 256 * m->KH) << wd));
  TRY(dev_alloc(bag, &m->Wh2, ((size_t)MAX_HEADS * 256 * 256) << wd));
  TRY(dev_alloc(bag, &m->Wh3, ((size_t)MAX_HEADS * 256 * 256) << wd));
  TRY(dev_alloc(bag, &m->b_ih, (size_t)3 * D));
  TRY(dev_alloc(bag, &m->b_hh, (size_t)3 * D));
  TRY(dev_alloc(bag, &m->p1_b, (size_t)m->bnp1)); TRY(dev_alloc(bag, &m->p1_g, (size_t)m->bnp1)); TRY(dev_alloc(bag, &m->p1_be, (size_t)m->bnp1));
  TRY(dev_alloc(bag, &m->p2_b, (size_t)m->bnp2)); TRY(dev_alloc(bag, &m->p2_g, (size_t)m->bnp2)); TRY(dev_alloc(bag, &m->p2_be, (size_t)m->bnp2));
  TRY(dev_alloc(bag, &m->p3_b, (size_t)ZP));
  TRY(dev_alloc(bag, &m->h1_b, (size_t)MAX_HEADS * m->bnh1)); TRY(dev_alloc(bag, &m->h1_g, (size_t)MAX_HEADS * m->bnh1)); TRY(dev_alloc(bag, &m->h1_be, (size_t)MAX_HEADS * m->bnh1));
  TRY(dev_alloc(bag, &m->h2_b, (size_t)MAX_HEADS * m->bnh2)); TRY(dev_alloc(bag, &m->h2_g, (size_t)MAX_HEADS * m->bnh2)); TRY(dev_alloc(bag, &m->h2_be, (size_t)MAX_HEADS * m->bnh2));
  TRY(dev_alloc(bag, &m->h3_b, (size_t)MAX_HEADS * 256));
  TRY(dev_alloc(bag, &m->bk_rew, (size_t)NB));
  TRY(dev_alloc(bag, &m->bk_crit, (size_t)NB));

  // ---- GRU: rows grouped per tile [r | z | n], columns [z | a(pad 64) | h(pad DP)]; packed for U = 32 and U = 64
  for (int v = 0; v < 2; ++v) {
    const int U = 32 << v, nt = m->gru_tiles2[v];
    std::vector<int> rows((size_t)nt * 3 * U);
    for (int j = 0; j < nt; ++j)
      for (int gte = 0; gte < 3; ++gte)
        for (int u = 0; u < U; ++u) {
          const int unit = j * U + u;
          rows[(size_t)j * 3 * U + gte * U + u] = unit < D ? gte * D + unit : -1;
        }
    std::vector<int> xcols(ZP + 64);
    for (int c = 0; c < ZP + 64; ++c) xcols[c] = c < ZP + A ? c : -1;   // reference x = [z, a]  (SequenceModel.py:21)
    TRY(add_mat(m, m->Wgru2[v], m->KG, 0, 0, S_GRU_WIH, ZP + A, rows, xcols));
    TRY(add_mat(m, m->Wgru2[v], m->KG, 0, ZP + 64, S_GRU_WHH, D, rows, iota_lim(DP, D)));
  }
  TRY(add_vec(m, m->b_ih, S_GRU_BIH, iota_lim(3 * D, 3 * D)));
  TRY(add_vec(m, m->b_hh, S_GRU_BHH, iota_lim(3 * D, 3 * D)));
  // ---- prior MLP (input h)
  {
    const int h1 = d.h_prior[0], h2 = d.h_prior[1], s = S_MLP0 + 0;
    TRY(add_mat(m, m->Wp1, DP, 0, 0, s + 0, D, iota_lim(m->bnp1, h1), iota_lim(DP, D)));
    TRY(add_vec(m, m->p1_b, s + 1, iota_lim(m->bnp1, h1)));
    TRY(add_vec(m, m->p1_g, s + 2, iota_lim(m->bnp1, h1)));
    TRY(add_vec(m, m->p1_be, s + 3, iota_lim(m->bnp1, h1)));
    TRY(add_mat(m, m->Wp2, 256, 0, 0, s + 4, h1, iota_lim(m->bnp2, h2), iota_lim(256, h1)));
    TRY(add_vec(m, m->p2_b, s + 5, iota_lim(m->bnp2, h2)));
    TRY(add_vec(m, m->p2_g, s + 6, iota_lim(m->bnp2, h2)));
    TRY(add_vec(m, m->p2_be, s + 7, iota_lim(m->bnp2, h2)));
    TRY(add_mat(m, m->Wp3, 256, 0, 0, s + 8, h2, iota_lim(ZP, ZP), iota_lim(256, h2)));
    TRY(add_vec(m, m->p3_b, s + 9, iota_lim(ZP, ZP)));
  }
  // ---- [h, z] heads: slots reward, cont, actor, critic, target critic
  {
    const int h1 = d.h_head[0], h2 = d.h_head[1];
    std::vector<int> l1cols(m->KH);   // packed columns [z | h]; reference input is [h, z] (DynamicsPredictors.py:65-66)
    for (int c = 0; c < m->KH; ++c) l1cols[c] = c < ZP ? D + c : (c - ZP < D ? c - ZP : -1);
    for (int hs = 0; hs < MAX_HEADS; ++hs) {
      const int mlp = hs + 1, s = S_MLP0 + 10 * mlp;
      TRY(add_mat(m, m->Wh1, m->KH, hs * 256, 0, s + 0, D + ZP, iota_lim(m->bnh1, h1), l1cols));
      TRY(add_vec(m, m->h1_b + hs * m->bnh1, s + 1, iota_lim(m->bnh1, h1)));
      TRY(add_vec(m, m->h1_g + hs * m->bnh1, s + 2, iota_lim(m->bnh1, h1)));
      TRY(add_vec(m, m->h1_be + hs * m->bnh1, s + 3, iota_lim(m->bnh1, h1)));
      TRY(add_mat(m, m->Wh2, 256, hs * 256, 0, s + 4, h1, iota_lim(m->bnh2, h2), iota_lim(256, h1)));
      TRY(add_vec(m, m->h2_b + hs * m->bnh2, s + 5, iota_lim(m->bnh2, h2)));
      TRY(add_vec(m, m->h2_g + hs * m->bnh2, s + 6, iota_lim(m->bnh2, h2)));
      TRY(add_vec(m, m->h2_be + hs * m->bnh2, s + 7, iota_lim(m->bnh2, h2)));
      if (hs == HS_ACTOR) {  // mu rows at 0.., log-sigma rows at 16..  (Agent.py:186-187)
        std::vector<int> r_mu(256, -1), r_ls(256, -1);
        for (int a = 0; a < A; ++a) { r_mu[a] = a; r_ls[16 + a] = a; }
        // two ops writing disjoint rows of the same 256-row block: pack mu rows [0,16) and ls rows [16,32) separately
        TRY(add_mat(m, m->Wh3, 256, hs * 256, 0, S_MU_W, h2, std::vector<int>(r_mu.begin(), r_mu.begin() + 16), iota_lim(256, h2)));
        TRY(add_mat(m, m->Wh3, 256, hs * 256 + 16, 0, S_LS_W, h2, std::vector<int>(r_ls.begin() + 16, r_ls.begin() + 32), iota_lim(256, h2)));
        TRY(add_vec(m, m->h3_b + hs * 256, S_MU_B, std::vector<int>(r_mu.begin(), r_mu.begin() + 16)));
        TRY(add_vec(m, m->h3_b + hs * 256 + 16, S_LS_B, std::vector<int>(r_ls.begin() + 16, r_ls.begin() + 32)));
      } else {
        const int nout = hs == HS_CONT ? 1 : NB;
        TRY(add_mat(m, m->Wh3, 256, hs * 256, 0, s + 8, h2, iota_lim(256, nout), iota_lim(256, h2)));
        TRY(add_vec(m, m->h3_b + hs * 256, s + 9, iota_lim(256, nout)));
      }
    }
    TRY(add_vec(m, m->bk_rew, S_BK_REW, iota_lim(NB, NB)));
    TRY(add_vec(m, m->bk_crit, S_BK_CRIT, iota_lim(NB, NB)));
  }
  // ---- TMA descriptors of the packed weights (box = {64, rows per tile})
  for (int v = 0; v < 2; ++v) {
    const int U = 32 << v;
    TRY(make_tmap_op_2d(&m->tmWgru2[v], m->Wgru2[v], (uint64_t)m->gru_tiles2[v] * 3 * U, m->KG, m->KG, 3 * U, wd));
    TRY(make_tmap_op_2d(&m->tmWgruQ[v], m->Wgru2[v], (uint64_t)m->gru_tiles2[v] * 3 * U, m->KG, m->KG, U / 2, wd));
  }
  TRY(make_tmap_op_2d(&m->tmWp1, m->Wp1, 256, DP, DP, m->bnp1, wd));
  TRY(make_tmap_op_2d(&m->tmWp2, m->Wp2, 256, 256, 256, m->bnp2, wd));
  TRY(make_tmap_op_2d(&m->tmWp1q, m->Wp1, 256, DP, DP, 64, wd));
  TRY(make_tmap_op_2d(&m->tmWp2q, m->Wp2, 256, 256, 256, 64, wd));
  TRY(make_tmap_op_2d(&m->tmWh1q, m->Wh1, (uint64_t)MAX_HEADS * 256, m->KH, m->KH, 64, wd));
  TRY(make_tmap_op_2d(&m->tmWh2q, m->Wh2, (uint64_t)MAX_HEADS * 256, 256, 256, 64, wd));
  TRY(make_tmap_op_2d(&m->tmWp3, m->Wp3, ZP, 256, 256, 256, wd));
  TRY(make_tmap_op_2d(&m->tmWp3h, m->Wp3, ZP, 256, 256, 128, wd));
  TRY(make_tmap_op_2d(&m->tmWh1, m->Wh1, (uint64_t)MAX_HEADS * 256, m->KH, m->KH, m->bnh1, wd));
  TRY(make_tmap_op_2d(&m->tmWh2, m->Wh2, (uint64_t)MAX_HEADS * 256, 256, 256, m->bnh2, wd));
  TRY(make_tmap_op_2d(&m->tmWh3, m->Wh3, (uint64_t)MAX_HEADS * 256, 256, 256, 256, wd));
  TRY(make_tmap_op_2d(&m->tmWh3a, m->Wh3, (uint64_t)MAX_HEADS * 256, 256, 256, 32, wd));
#undef TRY
  if (rc != DRM_OK) {
    drm_rssm_destroy(m);
    return rc;
  }
  *out = m;
  return DRM_OK;
}

extern "C" int drm_rssm_destroy(drm_rssm* m) {
  if (!m) return DRM_OK;
  for (void* p : m->allocs) cudaFree(p);
  delete m;
  return DRM_OK;
}

extern "C" int drm_rssm_pack(drm_rssm* m, const drm_rssm_weights* w, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(m && w, DRM_ERR_ARG, "drm_rssm_pack: NULL argument");
  const float* src[S_COUNT] = {};
  src[S_GRU_WIH] = w->gru_w_ih; src[S_GRU_WHH] = w->gru_w_hh; src[S_GRU_BIH] = w->gru_b_ih; src[S_GRU_BHH] = w->gru_b_hh;
  const drm_mlp_w* mlps[6] = {&w->prior, &w->reward, &w->cont, &w->actor, &w->critic, &w->target_critic};
  for (int i = 0; i < 6; ++i) {
    const float* f[10] = {mlps[i]->w0, mlps[i]->b0, mlps[i]->g0, mlps[i]->be0, mlps[i]->w1,
                          mlps[i]->b1, mlps[i]->g1, mlps[i]->be1, mlps[i]->w2, mlps[i]->b2};
    for (int k = 0; k < 10; ++k) src[S_MLP0 + 10 * i + k] = f[k];
  }
  src[S_MU_W] = w->actor_mu_w; src[S_MU_B] = w->actor_mu_b; src[S_LS_W] = w->actor_ls_w; src[S_LS_B] = w->actor_ls_b;
  src[S_BK_REW] = w->buckets_rew; src[S_BK_CRIT] = w->buckets_crit;
  // Any subset of the weight groups may be given (a mirrored module packs only its own slice); each stage
  // checks that the groups it needs were packed.
  cudaStream_t st = (cudaStream_t)stream;
  for (const MatOp& op : m->mat_ops) {
    if (!src[op.src]) continue;  // optional head (critic) absent
    pack_matrix_kernel<<<grid_for((long)op.nrows * op.ncols), 256, 0, st>>>(op.dst, op.ld_dst, op.row0, op.nrows, op.col0, op.ncols,
                                                                         src[op.src], op.ld_src, op.row_map, op.col_map, m->wide);
    DRM_LAUNCH_CHECK();
  }
  for (const VecOp& op : m->vec_ops) {
    if (!src[op.src]) continue;
    pack_vector_kernel<<<ceil_div(op.n, 256), 256, 0, st>>>(op.dst, op.n, src[op.src], op.map, op.fill);
    DRM_LAUNCH_CHECK();
  }
  m->packed = true;
  if (w->gru_w_ih && w->gru_w_hh && w->gru_b_ih && w->gru_b_hh) m->have |= HAVE_GRU;
  if (w->prior.w0 && w->prior.w1 && w->prior.w2) m->have |= HAVE_PRIOR;
  if (w->reward.w0 && w->reward.w2 && w->buckets_rew) m->have |= 1u << HS_REWARD;
  if (w->cont.w0 && w->cont.w2) m->have |= 1u << HS_CONT;
  if (w->actor.w0 && w->actor_mu_w && w->actor_ls_w) m->have |= 1u << HS_ACTOR;
  if (w->critic.w0 && w->critic.w2 && w->buckets_crit) m->have |= 1u << HS_CRITIC;
  if (w->target_critic.w0 && w->target_critic.w2 && w->buckets_crit) m->have |= 1u << HS_TARGET;
  m->has_critic = (m->have >> HS_CRITIC) & 1u;
  return DRM_OK;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
extern "C" int drm_rollout_create(drm_rssm* m, int32_t B, int32_t H, drm_rollout** out) {
  RC(check_arch());
  DRM_REQUIRE(m && out, DRM_ERR_ARG, "drm_rollout_create: NULL argument");
  DRM_REQUIRE(B >= 1 && H >= 1, DRM_ERR_SHAPE, "drm_rollout_create: B and H must be >= 1");
  drm_rollout* r = new drm_rollout();
  r->m = m; r->B = B; r->H = H; r->Mp = round_up(B, BM);
  int rc = DRM_OK;
#define TRY(x) if (rc == DRM_OK) rc = (x)
  for (int i = 0; i < 2; ++i) TRY(dev_alloc(r->allocs, &r->S[i], ((size_t)r->Mp * m->KS) << m->wide));
  TRY(dev_alloc(r->allocs, &r->Y1, ((size_t)(MAX_HEADS + 1) * r->Mp * 256) << m->wide));
  TRY(dev_alloc(r->allocs, &r->Y2, ((size_t)(MAX_HEADS + 1) * r->Mp * 256) << m->wide));
  for (int i = 0; i < 2; ++i) TRY(make_tmap_op_2d(&r->tmS[i], r->S[i], r->Mp, m->KS, m->KS, BM, m->wide));
  TRY(make_tmap_op_2d(&r->tmY1, r->Y1, (uint64_t)(MAX_HEADS + 1) * r->Mp, 256, 256, BM, m->wide));
  TRY(make_tmap_op_2d(&r->tmY2, r->Y2, (uint64_t)(MAX_HEADS + 1) * r->Mp, 256, 256, BM, m->wide));
  for (int i = 0; i < 2; ++i) TRY(make_tmap_op_2d(&r->tmS_s[i], r->S[i], r->Mp, m->KS, m->KS, SMALL_A_ROWS, m->wide));
  TRY(make_tmap_op_2d(&r->tmY1_s, r->Y1, (uint64_t)(MAX_HEADS + 1) * r->Mp, 256, 256, SMALL_A_ROWS, m->wide));
  TRY(make_tmap_op_2d(&r->tmY2_s, r->Y2, (uint64_t)(MAX_HEADS + 1) * r->Mp, 256, 256, SMALL_A_ROWS, m->wide));
#undef TRY
  if (rc != DRM_OK) {
    drm_rollout_destroy(r);
    return rc;
  }
  *out = r;
  return DRM_OK;
}

namespace drm { static void persist_free(drm_persist* ps); }

extern "C" int drm_rollout_destroy(drm_rollout* r) {
  if (!r) return DRM_OK;
  persist_free(r->ps);
  for (void* p : r->allocs) cudaFree(p);
  delete r;
  return DRM_OK;
}

namespace drm {

static GemmCommon common(const CUtensorMap& A, const CUtensorMap& B, int M, int bn) {
  GemmCommon g;
  memset(&g, 0, sizeof(g));
  g.tmA = A; g.tmB = B; g.M = M; g.bn = bn;
  return g;
}

// A view of "some rows of a state buffer + the hidden-activation buffers": the rollout uses S[sb] from row 0,
// the observe path uses one time-major buffer with a row offset per step.
struct WsView {
  const CUtensorMap* tmS;   // state buffer [rows, KS]
  __nv_bfloat16* S;         // its base pointer
  const CUtensorMap *tmY1, *tmY2;
  __nv_bfloat16 *Y1, *Y2;   // [(MAX_HEADS + 1) * slot_rows, 256]; slot 0 = prior / encoder, 1.. = heads
  int slot_rows;            // rows per Y slot
  int row0;                 // first row of this view inside S and inside every Y slot
  const CUtensorMap *tmS_s, *tmY1_s, *tmY2_s;   // the same buffers with a SMALL_A_ROWS-row box (NULL: none)
};
static WsView view_of(drm_rollout* r, int sb) {
  return WsView{&r->tmS[sb], r->S[sb], &r->tmY1, &r->tmY2, r->Y1, r->Y2, r->Mp, 0, &r->tmS_s[sb], &r->tmY1_s, &r->tmY2_s};
}
// Few rows (one acting environment, a 16-sequence scan step): load only SMALL_A_ROWS rows of every A k-block.  The MMA still
// reads a 128-row tile; rows beyond the box hold stale shared memory, give garbage accumulator rows, and every epilogue writes
// only rows < M.  Cuts the A bytes per k-block from 16 KB to 4 KB on stages that are bound by shared-memory ingress.
static void small_a(GemmCommon& g, const CUtensorMap* small) {
  if (small && g.M <= SMALL_A_ROWS && opts().small_a) {
    g.tmA = *small;
    g.a_bytes = SMALL_A_ROWS * BK * 2;
  }
}
// TF32 mode: k-block counts were written in 64-element units; a 128-byte k-block holds 32 fp32 elements, and every block boundary
// of the state / weight layouts is a multiple of 64 elements, so offsets and counts simply double.
static void set_precision(GemmCommon& g, const drm_rssm* m) {
  g.wide = m->wide;
  if (m->wide) { g.ka0 *= 2; g.nka0 *= 2; g.ka1 *= 2; g.nka1 *= 2; }
}
// GRU: src = [z | a | h_t]  ->  h_{t+1} (fp32 h_out, bf16 into dst's h columns)
static int stage_gru(drm_rssm* m, const WsView& src, const WsView& dst, const float* h_prev, long ld_hprev, float* h_out,
                     long ld_hout, int M, cudaStream_t st, bool allow_ksplit = false) {
  // tile width: 64 hidden units per tile once that still fills the machine twice over (more FLOPs per operand byte), else 32
  const int mt = ceil_div(M, BM);
  int v = (mt * m->gru_tiles2[1] >= 2 * 148) ? 1 : 0;
  if (const int force = opts().gru_u) v = force == 64 ? 1 : 0;
  const int U = 32 << v;
  GemmCommon g = common(*src.tmS, m->tmWgru2[v], M, 3 * U);
  g.a_row0 = src.row0;
  g.ka0 = 0; g.nka0 = m->ZP / 64 + 1;                 // x part: z blocks + the action block
  g.ka1 = m->ZP / 64 + 1; g.nka1 = m->DP / 64;        // h part
  small_a(g, src.tmS_s);
  set_precision(g, m);
  const dim3 grid(mt, m->gru_tiles2[v]);
  __nv_bfloat16* s_h = opnd_at(dst.S, (long)dst.row0 * m->KS + m->ZP + 64, m->wide);
  // CTA pairs (cta_group::2): two m-tiles issue one M = 256 MMA, each SM stages half of the weight tile (gru_pair.cuh).
  // Measured: 16 384 rows, D = 4096: 897 -> 1057 TFLOP/s; D = 600: 647 -> 720; 1024 rows (76 pairs of U = 32 tiles): 21.6 -> 24.4 us,
  // so the automatic choice pairs only the wide-tile (large-grid) configuration.
  const bool pair = opts().gru_pair < 0 ? v == 1 : opts().gru_pair != 0;
  if (pair && mt >= 2 && !m->wide) {   // (TF32 mode runs on the one-CTA-per-tile kernel only)
    g.tmB = m->tmWgruQ[v];
    g.a_bytes = 0;
    g.band = opts().gru_band & ~1;
    if (U == 32) {
      EpiGru<32>::Params p{m->b_ih, m->b_hh, h_prev, h_out, s_h, ld_hprev, ld_hout, m->KS, m->d.D};
      return launch_gru_pair<32>(g, p, mt, m->gru_tiles2[v], st);
    }
    EpiGru<64>::Params p{m->b_ih, m->b_hh, h_prev, h_out, s_h, ld_hprev, ld_hout, m->KS, m->d.D};
    return launch_gru_pair<64>(g, p, mt, m->gru_tiles2[v], st);
  }
  // (allow_ksplit: the posterior scan and the step-level entry point; drm_rollout_run keeps kernels whose per-row arithmetic does
  // not depend on the batch size, so that shards of a rollout concatenate bit-exactly to the full batch)
  if (allow_ksplit && opts().gru_ksplit && mt <= 1 && mt * m->gru_tiles2[0] <= 148 && !m->wide) {
    // tiny grid (the 16-sequence posterior scan, the B = 1 acting path, warm starts): x part and h part of every tile on two
    // CTAs of a cluster, rows swapped for the epilogue (gru_ksplit.cuh).  Measured per imagined step at D = 600: 128 rows
    // 62.1 -> 59.4 us, 256 rows 62.9 -> 65.1 us, 512 rows 66.2 -> 67.7 us, 896 rows 66.0 -> 66.8 us; a 48-unit-tile variant that
    // fits 1024 rows (104 tiles x 2 CTAs) was 8 us slower there -- so only single-m-tile grids take this path.
    g.tmB = m->tmWgru2[0]; g.bn = 96;
    EpiGru<32>::Params p{m->b_ih, m->b_hh, h_prev, h_out, s_h, ld_hprev, ld_hout, m->KS, m->d.D};
    return launch_gru_ksplit<32>(g, p, mt, m->gru_tiles2[0], st);
  }
  g.hp_pre = U == 32 ? 1 : 0;   // h_prev tile prefetched into the epilogue scratch (18 of its 20 KB) under the main loop
  if (U == 32) {
    EpiGru<32>::Params p{m->b_ih, m->b_hh, h_prev, h_out, s_h, ld_hprev, ld_hout, m->KS, m->d.D};
    return launch_gemm<EpiGru<32>>(g, p, grid, st, DRM_STAGE_GRU);
  }
  EpiGru<64>::Params p{m->b_ih, m->b_hh, h_prev, h_out, s_h, ld_hprev, ld_hout, m->KS, m->d.D};
  return launch_gemm<EpiGru<64>>(g, p, grid, st, DRM_STAGE_GRU);
}

// One Linear + LayerNorm + SiLU stage.  Small grids (<= 37 tiles) use the cluster-of-4 column split, larger ones one CTA per tile.
template <bool HAS_ADD>
static int launch_ln(GemmCommon g, const CUtensorMap& tmB_full, const CUtensorMap& tmB_q, int bn_full,
                     typename EpiLnSiluT<HAS_ADD>::Params p, int mt, int n_slots, cudaStream_t st, int stage_id) {
  g.b_slot_rows = 256;
  if (opts().ln_cluster && mt * n_slots <= 37 && !g.wide) {
    g.tmB = tmB_q;
    g.bn = 64;
    return launch_gemm<EpiLnSiluN4T<HAS_ADD>>(g, p, dim3(mt, n_slots, 4), st, stage_id);
  }
  g.tmB = tmB_full;
  g.bn = bn_full;
  return launch_gemm<EpiLnSiluT<HAS_ADD>>(g, p, dim3(mt, n_slots, 1), st, stage_id);
}

// prior MLP on the view's h columns -> logits -> (optional) categorical sample
static int stage_prior(drm_rssm* m, const WsView& v, const float* uniforms, float* latent, long ld_latent, float* logits,
                       long ld_logits, uint8_t* idx, long ld_idx, bool write_sz, RowMap rm, int M, cudaStream_t st) {
  const int mt = ceil_div(M, BM);
  {
    GemmCommon g = common(*v.tmS, m->tmWp1, M, m->bnp1);
    small_a(g, v.tmS_s);
    g.a_row0 = v.row0;
    g.ka0 = m->ZP / 64 + 1; g.nka0 = m->DP / 64;
    g.n_slots = 1; g.y_slot[0] = 0;
    set_precision(g, m);
    EpiLnSilu::Params p{m->p1_b, m->p1_g, m->p1_be, nullptr, 0, v.Y1, 256, v.row0, v.slot_rows, m->d.h_prior[0], 1e-5f, m->bnp1};
    RC(launch_ln<false>(g, m->tmWp1, m->tmWp1q, m->bnp1, p, mt, 1, st, DRM_STAGE_PRIOR_L1));
  }
  {
    GemmCommon g = common(*v.tmY1, m->tmWp2, M, m->bnp2);
    small_a(g, v.tmY1_s);
    g.a_row0 = v.row0;
    g.ka0 = 0; g.nka0 = ceil_div(m->d.h_prior[0], 64);
    g.n_slots = 1; g.y_slot[0] = 0;
    set_precision(g, m);
    EpiLnSilu::Params p{m->p2_b, m->p2_g, m->p2_be, nullptr, 0, v.Y2, 256, v.row0, v.slot_rows, m->d.h_prior[1], 1e-5f, m->bnp2};
    RC(launch_ln<false>(g, m->tmWp2, m->tmWp2q, m->bnp2, p, mt, 1, st, DRM_STAGE_PRIOR_L2));
  }
  {
    const int bn = (mt * (m->ZP / 256) <= 74) ? 128 : 256;   // small grids: half-width tiles, one 32-class group per thread
    GemmCommon g = common(*v.tmY2, bn == 128 ? m->tmWp3h : m->tmWp3, M, bn);
    small_a(g, v.tmY2_s);
    g.a_row0 = v.row0;
    g.ka0 = 0; g.nka0 = ceil_div(m->d.h_prior[1], 64);
    set_precision(g, m);
    EpiCat::Params p{m->p3_b, uniforms, latent, logits, idx, write_sz ? opnd_at(v.S, (long)v.row0 * m->KS, m->wide) : nullptr, nullptr,
                     ld_latent, ld_logits, ld_idx, 0, m->KS, m->d.R, rm};
    RC(launch_gemm<EpiCat>(g, p, dim3(mt, m->ZP / bn), st, DRM_STAGE_PRIOR_CAT));
  }
  return DRM_OK;
}

static void fill_heads(drm_rssm* m, EpiHeads::Params& hp) {
  hp.bias = m->h3_b;
  hp.kind[HS_REWARD] = HEAD_BUCKET; hp.kind[HS_CONT] = HEAD_SIGMOID; hp.kind[HS_ACTOR] = HEAD_ACTOR;
  hp.kind[HS_CRITIC] = HEAD_BUCKET; hp.kind[HS_TARGET] = HEAD_BUCKET;
  hp.buckets[HS_REWARD] = m->bk_rew; hp.buckets[HS_CRITIC] = m->bk_crit; hp.buckets[HS_TARGET] = m->bk_crit;
  hp.NB = m->d.NB; hp.A = m->d.A;
}

// [h, z] heads on the view: slots listed in `slots` (HS_*)
static int stage_heads(drm_rssm* m, const WsView& v, const int* slots, int n_slots, EpiHeads::Params hp, int M, cudaStream_t st) {
  const int mt = ceil_div(M, BM);
  if (n_slots <= 0) return DRM_OK;
  {
    GemmCommon g = common(*v.tmS, m->tmWh1, M, m->bnh1);
    g.a_row0 = v.row0;
    g.ka0 = 0; g.nka0 = m->ZP / 64;                   // z blocks
    g.ka1 = m->ZP / 64 + 1; g.nka1 = m->DP / 64;      // h blocks (the action block is skipped)
    g.n_slots = n_slots;
    for (int i = 0; i < n_slots; ++i) g.y_slot[i] = slots[i];
    small_a(g, v.tmS_s);
    set_precision(g, m);
    EpiLnSilu::Params p{m->h1_b, m->h1_g, m->h1_be, nullptr, 0, v.Y1, 256, v.slot_rows + v.row0, v.slot_rows, m->d.h_head[0], 1e-5f, m->bnh1};
    RC(launch_ln<false>(g, m->tmWh1, m->tmWh1q, m->bnh1, p, mt, n_slots, st, DRM_STAGE_HEADS_L1));
  }
  {
    GemmCommon g = common(*v.tmY1, m->tmWh2, M, m->bnh2);
    small_a(g, v.tmY1_s);
    g.a_row0 = v.slot_rows + v.row0; g.a_y_stride = v.slot_rows;
    g.ka0 = 0; g.nka0 = ceil_div(m->d.h_head[0], 64);
    g.n_slots = n_slots;
    for (int i = 0; i < n_slots; ++i) g.y_slot[i] = slots[i];
    set_precision(g, m);
    EpiLnSilu::Params p{m->h2_b, m->h2_g, m->h2_be, nullptr, 0, v.Y2, 256, v.slot_rows + v.row0, v.slot_rows, m->d.h_head[1], 1e-5f, m->bnh2};
    RC(launch_ln<false>(g, m->tmWh2, m->tmWh2q, m->bnh2, p, mt, n_slots, st, DRM_STAGE_HEADS_L2));
  }
  {
    GemmCommon g = common(*v.tmY2, m->tmWh3, M, 256);
    small_a(g, v.tmY2_s);
    g.a_row0 = v.slot_rows + v.row0; g.a_y_stride = v.slot_rows;
    g.ka0 = 0; g.nka0 = ceil_div(m->d.h_head[1], 64);
    g.n_slots = n_slots;
    for (int i = 0; i < n_slots; ++i) g.y_slot[i] = slots[i];
    set_precision(g, m);
    fill_heads(m, hp);
    RC(launch_gemm<EpiHeads>(g, hp, dim3(mt, n_slots), st, DRM_STAGE_HEADS_OUT));
  }
  return DRM_OK;
}

static int pack_cols(__nv_bfloat16* S, int ld_s, int col0, const float* src, long ld_src, int ncols, int N, float* copy,
                     long ld_copy, cudaStream_t st, int wide = 0) {
  pack_state_kernel<<<grid_for((long)N * ncols), 256, 0, st>>>(S, ld_s, col0, src, ld_src, ncols, N, copy, ld_copy, wide);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
static int pack_state(drm_rollout* r, int sb, int col0, const float* src, long ld_src, int ncols, int N, float* copy,
                      long ld_copy, cudaStream_t st) {
  return pack_cols(r->S[sb], r->m->KS, col0, src, ld_src, ncols, N, copy, ld_copy, st, r->m->wide);
}

// One lane of an imagination rollout: start states [b0, b0 + M) of the workspace, every pointer is the FULL tensor's base.
static int rollout_lane(drm_rollout* r, int b0, int M, const float* z0, const float* h0, const float* uniforms, const float* normals,
                        float* latent, float* hidden, float* actions, float* rewards, float* continues, float* mu, float* sigma,
                        uint8_t* idx, cudaStream_t st) {
  drm_rssm* m = r->m;
  const int B = r->B, H = r->H, D = m->d.D, ZP = m->ZP, A = m->d.A, R = m->d.R;
  const long ldL = (long)(H + 1) * ZP, ldH = (long)(H + 1) * D, ldA = (long)H * A;
  z0 += (long)b0 * ZP; h0 += (long)b0 * D; uniforms += (long)b0 * R; normals += (long)b0 * A;
  latent += b0 * ldL; hidden += b0 * ldH; actions += b0 * ldA; mu += b0 * ldA; sigma += b0 * ldA;
  rewards += (long)b0 * H; continues += (long)b0 * H;
  if (idx) idx += (long)b0 * H * R;
  auto view = [&](int sb) { WsView v = view_of(r, sb); v.row0 = b0; return v; };
  auto S_rows = [&](int sb) { return opnd_at(r->S[sb], (long)b0 * m->KS, m->wide); };
  // t = 0 state into S[0]; latent[:, 0] = z0, hidden[:, 0] = h0
  RC(pack_cols(S_rows(0), m->KS, 0, z0, ZP, ZP, M, latent, ldL, st, m->wide));
  RC(pack_cols(S_rows(0), m->KS, ZP + 64, h0, D, D, M, hidden, ldH, st, m->wide));
  const int actor_only[1] = {HS_ACTOR};
  const int all3[3] = {HS_REWARD, HS_CONT, HS_ACTOR};
  {
    EpiHeads::Params hp;
    memset(&hp, 0, sizeof(hp));
    hp.normals = normals; hp.ld_normals = A; hp.mu = mu; hp.sigma = sigma; hp.action = actions; hp.ld_act = ldA;
    hp.s_a = opnd_at(S_rows(0), ZP, m->wide); hp.ld_s = m->KS;
    RC(stage_heads(m, view(0), actor_only, 1, hp, M, st));
  }
  for (int t = 0; t < H; ++t) {
    const int cur = t & 1, nxt = cur ^ 1;
    RC(stage_gru(m, view(cur), view(nxt), hidden + (long)t * D, ldH, hidden + (long)(t + 1) * D, ldH, M, st));
    RC(stage_prior(m, view(nxt), uniforms + (long)t * B * R, latent + (long)(t + 1) * ZP, ldL, nullptr, 0,
                   idx ? idx + (long)t * R : nullptr, (long)H * R, true, RowMap{0, 0, 0, 0}, M, st));
    EpiHeads::Params hp;
    memset(&hp, 0, sizeof(hp));
    hp.value[HS_REWARD] = rewards + t; hp.ld_value[HS_REWARD] = H;
    hp.value[HS_CONT] = continues + t; hp.ld_value[HS_CONT] = H;
    const bool more = t + 1 < H;
    if (more) {
      hp.normals = normals + (long)(t + 1) * B * A;  // [B, A] slab of step t + 1
      hp.ld_normals = A;
      hp.mu = mu + (long)(t + 1) * A; hp.sigma = sigma + (long)(t + 1) * A; hp.action = actions + (long)(t + 1) * A;
      hp.ld_act = ldA;
      hp.s_a = opnd_at(S_rows(nxt), ZP, m->wide); hp.ld_s = m->KS;
    }
    RC(stage_heads(m, view(nxt), all3, more ? 3 : 2, hp, M, st));
  }
  return DRM_OK;
}

}  // namespace drm

#include "rollout_persist.cuh"

extern "C" int drm_rollout_run(drm_rollout* r, const float* z0, const float* h0, const float* uniforms, const float* normals,
                               float* latent, float* hidden, float* actions, float* rewards, float* continues, float* mu,
                               float* sigma, uint8_t* idx, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(r && z0 && h0 && uniforms && normals && latent && hidden && actions && rewards && continues && mu && sigma,
              DRM_ERR_ARG, "drm_rollout_run: NULL argument");
  drm_rssm* m = r->m;
  DRM_REQUIRE(m->packed, DRM_ERR_ARG, "drm_rollout_run: weights were never packed (call drm_rssm_pack)");
  DRM_REQUIRE((m->have & (HAVE_GRU | HAVE_PRIOR | 7u)) == (HAVE_GRU | HAVE_PRIOR | 7u), DRM_ERR_ARG,
              "drm_rollout_run: GRU, prior, reward, continue and actor weights must all be packed");
  cudaStream_t st = (cudaStream_t)stream;
  // One persistent kernel for the whole horizon when its static schedule fits the machine (rollout_persist.cuh); the
  // launch-per-stage chain otherwise (large batches, where every stage fills the machine and is throughput bound), while the
  // launch-per-stage timeline probe is on, or with option "persist" = 0.
  if (opts().persist && !g_timeline && !r->m->wide && persist_eligible(r))
    return rollout_persist(r, z0, h0, uniforms, normals, latent, hidden, actions, rewards, continues, mu, sigma, idx, st);
  return rollout_lane(r, 0, r->B, z0, h0, uniforms, normals, latent, hidden, actions, rewards, continues, mu, sigma, idx, st);
}

// Which path the next drm_rollout_run takes (1 = persistent kernel) and its geometry; the timeout records of the persistent kernel
// (host-mapped, so they survive a trapped launch): out[0] = 1 / 0, out[1..4] = U, sampling tile width, CTAs, timeouts recorded,
// then up to 8 records {code, seen, want, thread, cta}.
extern "C" int drm_rollout_info(drm_rollout* r, uint32_t* out, int32_t n) {
  DRM_REQUIRE(r && out && n >= 5, DRM_ERR_ARG, "drm_rollout_info: bad argument");
  for (int i = 0; i < n; ++i) out[i] = 0;
  if (!(r->ps && r->ps->tried)) RC(check_arch());   // (no CUDA call once the plan exists: this must still answer after a trapped launch)
  if (r->m->wide) return DRM_OK;   // TF32 handles run launch by launch
  const bool ok = opts().persist && persist_eligible(r);
  out[0] = ok ? 1u : 0u;
  if (!r->ps || !r->ps->ok) return DRM_OK;
  out[1] = (uint32_t)r->ps->U; out[2] = 256u; out[3] = (uint32_t)r->ps->n_cta;
  out[4] = r->ps->dbg[8 * 160];
  int k = 5;
  for (int c = 0; c < 160 && k + 5 <= n; ++c) {
    const unsigned* rec = r->ps->dbg + 8 * c;
    if (rec[0] == 0 && rec[2] == 0) continue;
    for (int j = 0; j < 5; ++j) out[k++] = rec[j];
  }
  return DRM_OK;
}

// Debug: per-tile timestamps of the persistent kernel.  out == NULL: record the tiles of states [j0, j0 + nj) of the following
// rollouts (nj small: 32 tiles per CTA are kept).  out != NULL: copy n_cta * 32 records of 8 u64
// {code, start, dependency seen, first operands, epilogue ready, accumulator ready, epilogue done, published} (globaltimer ns;
// code = kind << 24 | layer << 16 | state << 8 | m-tile) to the host buffer (n_words >= n_cta * 256) and stop recording.
extern "C" int drm_rollout_trace(drm_rollout* r, int32_t j0, int32_t nj, unsigned long long* out, int64_t n_words) {
  DRM_REQUIRE(r, DRM_ERR_ARG, "drm_rollout_trace: NULL workspace");
  RC(check_arch());
  DRM_REQUIRE(!r->m->wide && persist_eligible(r), DRM_ERR_ARG, "drm_rollout_trace: this workspace does not use the persistent kernel");
  drm_persist* ps = r->ps;
  const size_t words = (size_t)ps->n_cta * PS_TRACE_SLOTS * 8 * 2;   // tile records, then the epilogue-lap records
  if (!out) {
    if (!ps->trace) RC(dev_alloc(r->allocs, &ps->trace, words));
    DRM_CUDA(cudaMemset(ps->trace, 0, words * sizeof(unsigned long long)));
    ps->trace_j0 = j0; ps->trace_j1 = j0 + nj;
    return DRM_OK;
  }
  DRM_REQUIRE(ps->trace && n_words >= (int64_t)words, DRM_ERR_ARG, "drm_rollout_trace: tracing is off or the buffer is too small");
  DRM_CUDA(cudaDeviceSynchronize());
  DRM_CUDA(cudaMemcpy(out, ps->trace, words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  ps->trace_j0 = ps->trace_j1 = 0;
  ps->trace = nullptr;   // (the buffer stays in the workspace's allocation bag)
  return DRM_OK;
}

// ------------------------------------------------------------------------------------------
// step-level entry points (the drop-in classes' per-call path)
// ------------------------------------------------------------------------------------------
extern "C" int drm_gru_step(drm_rollout* r, const float* z, const float* h, const float* a, float* h_out, int32_t N,
                            void* stream) {
  RC(check_arch());
  DRM_REQUIRE(r && z && h && a && h_out, DRM_ERR_ARG, "drm_gru_step: NULL argument");
  DRM_REQUIRE(N >= 0 && N <= r->B, DRM_ERR_SHAPE, "drm_gru_step: N exceeds the workspace rows");
  DRM_REQUIRE(r->m->packed && (r->m->have & HAVE_GRU), DRM_ERR_ARG, "drm_gru_step: GRU weights were never packed");
  if (N == 0) return DRM_OK;
  drm_rssm* m = r->m;
  cudaStream_t st = (cudaStream_t)stream;
  RC(pack_state(r, 0, 0, z, m->ZP, m->ZP, N, nullptr, 0, st));
  RC(pack_state(r, 0, m->ZP, a, m->d.A, m->d.A, N, nullptr, 0, st));
  RC(pack_state(r, 0, m->ZP + 64, h, m->d.D, m->d.D, N, nullptr, 0, st));
  return stage_gru(m, view_of(r, 0), view_of(r, 1), h, m->d.D, h_out, m->d.D, N, st, true);
}

extern "C" int drm_prior(drm_rollout* r, const float* h, const float* uniforms, float* logits, float* z_st, uint8_t* idx,
                         int32_t N, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(r && h, DRM_ERR_ARG, "drm_prior: NULL argument");
  DRM_REQUIRE(N >= 0 && N <= r->B, DRM_ERR_SHAPE, "drm_prior: N exceeds the workspace rows");
  DRM_REQUIRE(r->m->packed && (r->m->have & HAVE_PRIOR), DRM_ERR_ARG, "drm_prior: prior weights were never packed");
  if (N == 0) return DRM_OK;
  drm_rssm* m = r->m;
  cudaStream_t st = (cudaStream_t)stream;
  RC(pack_state(r, 0, m->ZP + 64, h, m->d.D, m->d.D, N, nullptr, 0, st));
  return stage_prior(m, view_of(r, 0), uniforms, z_st, m->ZP, logits, m->ZP, idx, m->d.R, false, RowMap{0, 0, 0, 0}, N, st);
}

extern "C" int drm_heads(drm_rollout* r, const float* h, const float* z, const float* normals, int32_t heads,
                         const drm_heads_out* out, int32_t N, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(r && h && z && out, DRM_ERR_ARG, "drm_heads: NULL argument");
  DRM_REQUIRE(N >= 0 && N <= r->B, DRM_ERR_SHAPE, "drm_heads: N exceeds the workspace rows");
  DRM_REQUIRE(r->m->packed, DRM_ERR_ARG, "drm_heads: weights were never packed");
  drm_rssm* m = r->m;
  DRM_REQUIRE(((unsigned)heads & ~m->have & 31u) == 0, DRM_ERR_ARG, "drm_heads: a requested head's weights were never packed");
  if (N == 0) return DRM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  RC(pack_state(r, 0, 0, z, m->ZP, m->ZP, N, nullptr, 0, st));
  RC(pack_state(r, 0, m->ZP + 64, h, m->d.D, m->d.D, N, nullptr, 0, st));
  int slots[MAX_HEADS], n = 0;
  EpiHeads::Params hp;
  memset(&hp, 0, sizeof(hp));
  const int NB = m->d.NB;
  if (heads & DRM_HEAD_REWARD) { slots[n++] = HS_REWARD; hp.value[HS_REWARD] = out->reward; hp.ld_value[HS_REWARD] = 1; hp.logits[HS_REWARD] = out->reward_logits; hp.ld_logits[HS_REWARD] = NB; }
  if (heads & DRM_HEAD_CONT) { slots[n++] = HS_CONT; hp.value[HS_CONT] = out->cont_prob; hp.logits[HS_CONT] = out->cont_logit; hp.ld_value[HS_CONT] = 1; }
  if (heads & DRM_HEAD_ACTOR) { slots[n++] = HS_ACTOR; hp.normals = normals; hp.ld_normals = m->d.A; hp.mu = out->mu; hp.sigma = out->sigma; hp.action = out->action; hp.ld_act = m->d.A; }
  if (heads & DRM_HEAD_CRITIC) { slots[n++] = HS_CRITIC; hp.value[HS_CRITIC] = out->value; hp.ld_value[HS_CRITIC] = 1; hp.logits[HS_CRITIC] = out->value_logits; hp.ld_logits[HS_CRITIC] = NB; }
  if (heads & DRM_HEAD_TARGET_CRITIC) { slots[n++] = HS_TARGET; hp.value[HS_TARGET] = out->target_value; hp.ld_value[HS_TARGET] = 1; }
  return stage_heads(m, view_of(r, 0), slots, n, hp, N, st);
}

// ------------------------------------------------------------------------------------------
// test hook: plain GEMM through the same TMA / tcgen05 main loop
// ------------------------------------------------------------------------------------------
extern "C" int drm_test_gemm(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                             void* stream) {
  RC(check_arch());
  DRM_REQUIRE(A && W && out, DRM_ERR_ARG, "drm_test_gemm: NULL argument");
  DRM_REQUIRE(M >= 1 && N >= 1 && K >= 1, DRM_ERR_SHAPE, "drm_test_gemm: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int Kp = round_up(K, 64), Mp = round_up(M, BM), Np = round_up(N, 256);
  std::vector<void*> bag;
  __nv_bfloat16 *a16 = nullptr, *w16 = nullptr;
  int rc = dev_alloc(bag, &a16, (size_t)Mp * Kp);
  if (rc == DRM_OK) rc = dev_alloc(bag, &w16, (size_t)Np * Kp);
  CUtensorMap tA, tW;
  if (rc == DRM_OK) {
    f32_to_bf16_pad_kernel<<<grid_for((long)M * Kp), 256, 0, st>>>(a16, Kp, A, M, K);
    f32_to_bf16_pad_kernel<<<grid_for((long)N * Kp), 256, 0, st>>>(w16, Kp, W, N, K);
    g_launches.fetch_add(2);
    rc = make_tmap_bf16_2d(&tA, a16, Mp, Kp, Kp, BM);
  }
  if (rc == DRM_OK) rc = make_tmap_bf16_2d(&tW, w16, Np, Kp, Kp, 256);
  if (rc == DRM_OK) {
    GemmCommon g = common(tA, tW, M, 256);
    g.ka0 = 0; g.nka0 = Kp / 64;
    EpiPlain::Params p{bias, out, nullptr, (long)N, 0, N, 0, 0, RowMap{0, 0, 0, 0}};
    rc = launch_gemm<EpiPlain>(g, p, dim3(Mp / BM, Np / 256), st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  for (void* q : bag) cudaFree(q);
  if (rc == DRM_OK && e != cudaSuccess) return fail(DRM_ERR_CUDA, std::string("drm_test_gemm: ") + cudaGetErrorString(e));
  return rc;
}

extern "C" int drm_set_option(const char* name, int32_t value) {
  if (!name) return fail(DRM_ERR_ARG, "drm_set_option: NULL name");
  const std::string n(name);
  Options& o = opts();
  if (n == "ln_cluster") o.ln_cluster = value != 0;
  else if (n == "small_a") o.small_a = value != 0;
  else if (n == "conv_persist") o.conv_persist = value != 0;
  else if (n == "conv_implicit") o.conv_implicit = value != 0;
  else if (n == "conv_chunk") o.conv_chunk = value >= 16 ? value : 512;
  else if (n == "conv_kps") o.conv_kps = value;
  else if (n == "gru_ksplit") o.gru_ksplit = value != 0;
  else if (n == "persist") o.persist = value != 0;
  else if (n == "gru_pair") { DRM_REQUIRE(value >= -1 && value <= 1, DRM_ERR_ARG, "drm_set_option: gru_pair must be -1, 0 or 1"); o.gru_pair = value; }
  else if (n == "gru_band") { DRM_REQUIRE(value >= 0 && value <= 64, DRM_ERR_ARG, "drm_set_option: gru_band must be in [0, 64]"); o.gru_band = value; }
  else if (n == "gru_u") { DRM_REQUIRE(value == 0 || value == 32 || value == 64, DRM_ERR_ARG, "drm_set_option: gru_u must be 0, 32 or 64"); o.gru_u = value; }
  else return fail(DRM_ERR_ARG, "drm_set_option: unknown option '" + n + "'");
  return DRM_OK;
}

// Debug: enable (on = 1) the in-kernel probe of CTA (0,0) of every fused stage, or read the last probes back
// (on = 0 with out != NULL: copies DRM_STAGE_COUNT * (16 + 1024) u64: {globaltimer ns, clock64} x 8 points per stage, then
// per stage 256 CTA records {entry ns, dependency wait over ns, exit ns, SM id}).
extern "C" int drm_debug_timeline(int32_t on, unsigned long long* out_host) {
  if (on) {
    if (!g_timeline) {
      DRM_CUDA(cudaMalloc(&g_timeline, DRM_STAGE_COUNT * (16 + 1024) * sizeof(unsigned long long)));
      DRM_CUDA(cudaMemset(g_timeline, 0, DRM_STAGE_COUNT * (16 + 1024) * sizeof(unsigned long long)));
    }
    return DRM_OK;
  }
  if (g_timeline && out_host) {
    DRM_CUDA(cudaDeviceSynchronize());
    DRM_CUDA(cudaMemcpy(out_host, g_timeline, DRM_STAGE_COUNT * (16 + 1024) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  if (g_timeline) { cudaFree(g_timeline); g_timeline = nullptr; }
  return DRM_OK;
}

#include "conv_persist.cuh"
namespace drm {
// narrow conv layers (<= 64 output channels): one persistent CTA per SM walks the tiles (conv_persist.cuh)
static int launch_conv_persist(const GemmCommon& g, const EpiPlain::Params& p, int phases, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(conv_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM));
    attr_set = true;
  }
  ConvPersist c;
  memset(&c, 0, sizeof(c));
  c.tmA = g.tmA; c.tmB = g.tmB;
  c.M = g.M; c.n_mtiles = ceil_div(g.M, BM); c.phases = phases; c.a_phase_rows = g.a_y_stride;
  c.bn = g.bn; c.nk = g.nka0;
  c.bias = p.bias; c.out = p.out_bf16; c.ld = p.ld_bf16; c.n_valid = p.N; c.act = p.act; c.rm = p.rm;
  const int n_tiles = c.n_mtiles * phases;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles < 148 ? n_tiles : 148);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = CP_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (!profile_on()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, conv_persist_kernel, c));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
static bool use_conv_persist(const GemmCommon& g, const EpiPlain::Params& p) {
  return opts().conv_persist && (g.bn == 32 || g.bn == 64) && p.out_bf16 && !p.out_f32 && g.nka1 == 0 && g.ka0 == 0 && g.a_row0 == 0;
}
}  // namespace drm

#include "conv_implicit.cuh"
namespace drm {
// One conv / transposed-conv layer as an implicit GEMM (conv_implicit.cuh).  tmA: `phases` im2col maps; (Wo, Ho): filter positions
// per frame and phase; transposed = the sub-pixel phase decomposition of ConvTranspose2d(k4, s2, p1), else Conv2d(k4, s2, p1).
static int launch_conv_implicit(const CUtensorMap* tmA, int phases, const CUtensorMap& tmB, int M, int Wo, int Ho, int stride, int n_base,
                                int ntap, int cpad, int chunk, int bn, const float* bias, __nv_bfloat16* out, long ld, int n_valid, int act,
                                RowMap rm, bool transposed, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DRM_CUDA(cudaFuncSetAttribute(conv_implicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CI_SMEM));
    attr_set = true;
  }
  ConvImplicit c;
  memset(&c, 0, sizeof(c));
  for (int p = 0; p < phases; ++p) c.tmA[p] = tmA[p];
  c.tmB = tmB;
  c.M = M; c.n_mtiles = ceil_div(M, BM); c.phases = phases;
  c.Wo = Wo; c.Ho = Ho; c.stride = stride; c.n_base = n_base;
  c.ntap = ntap; c.cpt = cpad / chunk; c.chunk = chunk;
  for (int p = 0; p < phases; ++p) {
    if (!transposed) {
      c.lower_w[p] = -1; c.lower_h[p] = -1;
      for (int t = 0; t < 16; ++t) { c.off_w[p][t] = (unsigned char)(t & 3); c.off_h[p][t] = (unsigned char)(t >> 2); }
    } else {
      // phase (py, px): output (2q + py, 2r + px) sums input rows {q, q - 1} (py = 0) or {q, q + 1} (py = 1), weight tap order
      // (ty, tx) with t = 0 the centre (ct_d in vae.cuh): as a 2-tap window the base row is q - 1 (py = 0) or q (py = 1)
      const int py = p >> 1, px = p & 1;
      c.lower_w[p] = px ? 0 : -1; c.lower_h[p] = py ? 0 : -1;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          c.off_h[p][ty * 2 + tx] = (unsigned char)(py ? ty : 1 - ty);
          c.off_w[p][ty * 2 + tx] = (unsigned char)(px ? tx : 1 - tx);
        }
    }
  }
  c.bn = bn;
  // pipeline shape: stages of up to ~48 KB (one full / empty handshake per stage costs ~0.3 us whatever it carries), at least two slots
  c.sub_bytes = round_up(BM * chunk * 2 + bn * chunk * 2, 1024);
  const int nk = ntap * (cpad / chunk);
  c.kps = std::max(1, std::min(std::min(opts().conv_kps > 0 ? opts().conv_kps : 48 * 1024 / c.sub_bytes, nk), CI_RING_BYTES / (2 * c.sub_bytes)));
  c.stage_bytes = c.kps * c.sub_bytes;
  c.n_stages = std::min(CI_MAX_STAGES, CI_RING_BYTES / c.stage_bytes);
  c.tmem_cols = 32;
  while (c.tmem_cols < 2 * bn) c.tmem_cols *= 2;
  c.bias = bias; c.out = out; c.ld = ld; c.n_valid = n_valid; c.act = act; c.rm = rm;
  const int n_tiles = c.n_mtiles * phases;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles < 148 ? n_tiles : 148);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = CI_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (!profile_on()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  DRM_CUDA(cudaLaunchKernelEx(&cfg, conv_implicit_kernel, c));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
}  // namespace drm

#include "vae.cuh"
