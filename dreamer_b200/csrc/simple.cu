// HBM-bound fp32/u8 kernels of the hot path: the fused 32-class categorical head, the KL-balance
// terms, the replay-ring gather/insert, the lambda-return scan, two-hot cross-entropy and the
// bucket-value readout.  One warp per row everywhere (lane = class / bucket stripe), 16-byte
// vector accesses where the layout allows, grids sized in multiples of the SM count.
#include <algorithm>
#include <string>

#include "common.cuh"
#include "internal.h"

namespace drm {

// Programmatic dependent launch for the kernels of the backward walks (7 launches per time step, each a few microseconds): the next
// kernel's blocks are scheduled while this one drains.  A kernel launched through launch_pdl() starts with PDL_ENTRY(): it waits for
// the previous kernel in the stream to complete (nothing before that point touches global memory) and at once lets its own successor
// be scheduled.  Inside a captured graph these become programmatic dependency edges.
#define PDL_ENTRY()                                              \
  do {                                                           \
    asm volatile("griddepcontrol.wait;\n" ::: "memory");          \
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); \
  } while (0)

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl2(void (*kernel)(KArgs...), dim3 grid, unsigned block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// categorical32: softmax -> unimix -> inverse-CDF sample -> one-hot -> straight-through.
// DynamicsPredictors.py:33-39 / VariationalAutoEncoder.py:88-98.
// A warp owns 32 consecutive rows: the 32 x 32 logit tile is loaded coalesced (float4 per lane) into a padded
// shared-memory tile, each LANE then processes one ROW in registers (max, exp, sum, the left-to-right fp32 CDF of the
// contract, the draw -- no shuffles), writes its outputs back into the tile and the warp stores them coalesced.
// Algorithmic bytes per row: 128 B logits + 4 B uniform read; 128 B z_st (+ 1 B idx, + optional 128 B probs,
// 64 B bf16 one-hot) written.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) categorical32_kernel(const float* __restrict__ logits,
                                                            const float* __restrict__ uniforms,
                                                            uint8_t* __restrict__ idx_out, float* __restrict__ z_st,
                                                            float* __restrict__ probs, uint16_t* __restrict__ z_bf16,
                                                            int64_t n_rows, const uint8_t* __restrict__ idx_in) {
  __shared__ float tiles[8][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float(*t)[33] = tiles[w];
  const int64_t warp = (int64_t)blockIdx.x * 8 + w;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const int64_t ntiles = (n_rows + 31) >> 5;
  for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
    const int64_t row0 = tile << 5;
    const int nr = (int)min((int64_t)32, n_rows - row0);
    // coalesced load: pass k brings rows 4k .. 4k+3 (8 lanes x float4 per row)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = 4 * k + (lane >> 3), c = (lane & 7) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nr) x = __ldg(reinterpret_cast<const float4*>(logits + (row0 + r) * 32 + c));
      t[r][c] = x.x; t[r][c + 1] = x.y; t[r][c + 2] = x.z; t[r][c + 3] = x.w;
    }
    const float u = (uniforms && lane < nr) ? __ldg(uniforms + row0 + lane) : 0.f;
    __syncwarp();
    float v[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) { v[j] = t[lane][j]; mx = fmaxf(mx, v[j]); }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { v[j] = expf(v[j] - mx); s += v[j]; }
    float cdf = 0.f;
    int idx = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = 0.99f * (v[j] / s) + 0.01f * (1.0f / 32.0f);
      cdf += v[j];
      idx += (cdf <= u) ? 1 : 0;
    }
    idx = idx > 31 ? 31 : idx;
    if (idx_in && lane < nr) idx = min((int)idx_in[row0 + lane], 31);   // teacher forcing: the class is given, not drawn
    if (idx_out && lane < nr) idx_out[row0 + lane] = (uint8_t)idx;   // 32 consecutive bytes per warp
    if (probs) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[lane][j] = v[j];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = 4 * k + (lane >> 3), c = (lane & 7) * 4;
        if (r < nr) *reinterpret_cast<float4*>(probs + (row0 + r) * 32 + c) = make_float4(t[r][c], t[r][c + 1], t[r][c + 2], t[r][c + 3]);
      }
    }
    if (z_st) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[lane][j] = ((j == idx ? 1.0f : 0.0f) + v[j]) - v[j];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = 4 * k + (lane >> 3), c = (lane & 7) * 4;
        if (r < nr) __stcs(reinterpret_cast<float4*>(z_st + (row0 + r) * 32 + c), make_float4(t[r][c], t[r][c + 1], t[r][c + 2], t[r][c + 3]));
      }
    }
    if (z_bf16) {   // exact one-hot: every lane needs its rows' indices -> broadcast through the tile
      __syncwarp();
      t[lane][0] = __int_as_float(idx);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {      // pass k: rows 8k .. 8k+7, 4 lanes x uint4 (8 bf16) per row
        const int r = 8 * k + (lane >> 2), c = (lane & 3) * 8;
        if (r < nr) {
          const int ri = __float_as_int(t[r][0]);
          uint32_t q[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) q[e] = (ri == c + 2 * e ? 0x3F80u : 0u) | (ri == c + 2 * e + 1 ? 0x3F800000u : 0u);
          *reinterpret_cast<uint4*>(z_bf16 + (row0 + r) * 32 + c) = make_uint4(q[0], q[1], q[2], q[3]);
        }
      }
    }
    __syncwarp();
  }
}

// The same for FEW rows (the sequential BPTT step: B * R = 512 rows per call): one warp per ROW, one class per lane, so all four
// operand rows are fetched by four independent coalesced 128-byte loads and the three row reductions are shuffle trees -- the row-per-
// lane kernel below needs 16 warps and ~10 us of dependent load / expf latency for such a call, this one ~3 us on 64 SMs.
__global__ void __launch_bounds__(256) categorical32_bwd_rowwarp_kernel(const float* __restrict__ logits, const float* __restrict__ dz,
                                                                        const float* __restrict__ dz2, const float* __restrict__ dl_add,
                                                                        float* __restrict__ dlogits, int64_t n_rows) {
  PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t o = row * 32 + lane;
  const float l = __ldg(logits + o);
  float g = __ldg(dz + o);
  const float g2 = dz2 ? __ldg(dz2 + o) : 0.f;
  const float add = dl_add ? __ldg(dl_add + o) : 0.f;
  g += g2;
  float mx = l;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  // (summation order differs from the row-per-lane kernel's left-to-right sums: gradients agree to fp32 rounding, not bitwise)
  const float e = expf(l - mx);
  float sum = e;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float sft = e * (1.0f / sum);
  float dot = sft * g;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
  dlogits[o] = 0.99f * sft * (g - dot) + add;
}

// Backward of the straight-through sample z = onehot + p - stopgrad(p), p = 0.99 softmax(l) + 0.01 / 32
// (DynamicsPredictors.py:33-39, VariationalAutoEncoder.py:88-98):  dl = 0.99 * s * (g - sum_j s_j g_j), s = softmax(l), g = dz.
// Same tiling as the forward: a warp owns 32 rows, coalesced float4 tile loads, one row per lane.
// Optional addends (the sequential BPTT step fuses its two accumulations): g = dz + dz2, dl += dl_add.
__global__ void __launch_bounds__(256) categorical32_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dz,
                                                                const float* __restrict__ dz2, const float* __restrict__ dl_add,
                                                                float* __restrict__ dlogits, int64_t n_rows) {
  PDL_ENTRY();
  __shared__ float tiles[8][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float(*t)[33] = tiles[w];
  const int64_t warp = (int64_t)blockIdx.x * 8 + w;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const int64_t ntiles = (n_rows + 31) >> 5;
  for (int64_t tile = warp; tile < ntiles; tile += nwarps) {
    const int64_t row0 = tile << 5;
    const int nr = (int)min((int64_t)32, n_rows - row0);
    float s[32], g[32];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const float* src = pass ? dz : logits;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = 4 * k + (lane >> 3), c = (lane & 7) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) {
          x = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * 32 + c));
          if (pass && dz2) {
            const float4 y = __ldg(reinterpret_cast<const float4*>(dz2 + (row0 + r) * 32 + c));
            x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
          }
        }
        t[r][c] = x.x; t[r][c + 1] = x.y; t[r][c + 2] = x.z; t[r][c + 3] = x.w;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) (pass ? g : s)[j] = t[lane][j];
      __syncwarp();
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, s[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
    const float inv = 1.0f / sum;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) { s[j] *= inv; dot = fmaf(s[j], g[j], dot); }
#pragma unroll
    for (int j = 0; j < 32; ++j) t[lane][j] = 0.99f * s[j] * (g[j] - dot);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = 4 * k + (lane >> 3), c = (lane & 7) * 4;
      if (r < nr) {
        float4 o = make_float4(t[r][c], t[r][c + 1], t[r][c + 2], t[r][c + 3]);
        if (dl_add) {
          const float4 y = __ldg(reinterpret_cast<const float4*>(dl_add + (row0 + r) * 32 + c));
          o.x += y.x; o.y += y.y; o.z += y.z; o.w += y.w;
        }
        *reinterpret_cast<float4*>(dlogits + (row0 + r) * 32 + c) = o;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// Backward of y = SiLU(LayerNorm(a) * gamma + beta) for rows of n <= 1024 features (one warp per row; statistics are
// recomputed from the pre-activation a, two-pass like torch):  dln = dy * silu'(ln);  dxh = dln * gamma;
//   da = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)).   dln (optional) feeds the batched dgamma / dbeta sums.
// ------------------------------------------------------------------------------------------
template <int J>   // J = ceil(n / 32) values per lane: 8 for the <= 256-wide hidden layers of the reference (short unrolled loops), else 32
__global__ void __launch_bounds__(256) ln_silu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ da, float* __restrict__ dln_out, int64_t rows, int n, float eps,
                                                          float* __restrict__ dlnx_out = nullptr) {
  PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_n = 1.0f / (float)n;
  for (int64_t row = warp; row < rows; row += nwarps) {
    const float* ar = a + row * n;
    const float* dr = dy + row * n;
    float x[J], d[J];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      x[j] = c < n ? __ldg(ar + c) : 0.f;
      d[j] = c < n ? __ldg(dr + c) : 0.f;   // (fetched with the pre-activation: nothing below waits on a second round trip)
      s += x[j];
    }
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      const float t = c < n ? x[j] - mean : 0.f;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      float dxh = 0.f, xh = 0.f;
      if (c < n) {
        xh = (x[j] - mean) * rstd;
        const float g = __ldg(gamma + c);
        const float ln = fmaf(xh, g, __ldg(beta + c));
        const float sg = 1.0f / (1.0f + expf(-ln));
        const float dl = d[j] * (sg * (1.0f + ln * (1.0f - sg)));   // silu'(x) = sig (1 + x (1 - sig))
        if (dln_out) dln_out[row * n + c] = dl;
        if (dlnx_out) dlnx_out[row * n + c] = dl * xh;       // its column sums are d(loss)/d(gamma)
        dxh = dl * g;
      }
      x[j] = xh; d[j] = dxh;
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
    const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      if (c < n) da[row * n + c] = rstd * (d[j] - m1 - x[j] * m2);
    }
  }
}

// Column sums out[c] (+)= sum_r x[r][c] (bias and LayerNorm-affine gradients of the batched backward: d(loss)/d(bias) = colsum of the
// layer's pre-activation gradient).  One block per 32 columns: lane = column (coalesced 128-byte rows), 8 warps stride the rows, a
// fixed-order combination across the warps -- deterministic.
// (blockIdx.y: a chunk of rows_per_y rows whose sums go to out + blockIdx.y * n -- the first stage of a tall matrix, see drm_colsum)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t rows, int n, int64_t ld, float* __restrict__ out,
                                                     int accumulate, int64_t rows_per_y) {
  PDL_ENTRY();
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = (int)blockIdx.x * 32 + lane;
  x += (int64_t)blockIdx.y * rows_per_y * ld;
  rows = min(rows_per_y, rows - (int64_t)blockIdx.y * rows_per_y);
  out += (int64_t)blockIdx.y * n;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < n) {
    int64_t r = w;
    for (; r + 24 < rows; r += 32) {          // four independent loads in flight per thread
      s0 += x[r * ld + c];
      s1 += x[(r + 8) * ld + c];
      s2 += x[(r + 16) * ld + c];
      s3 += x[(r + 24) * ld + c];
    }
    for (; r < rows; r += 8) s0 += x[r * ld + c];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && c < n) {
    float t = part[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += part[k][lane];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// First stage of the column sums of a bf16 matrix (the conv layers' bias gradients: grad_output as [N * H * W, C] channels-last rows):
// one block per 64 columns x 64 rows, lane = two columns (one 4-byte load), 8 rows per warp all in flight, fp32 partial sums to
// out [gridDim.y][n].
__global__ void __launch_bounds__(256) colsum_bf16_stage_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int n, int64_t ld,
                                                                float* __restrict__ out) {
  PDL_ENTRY();
  __shared__ float part[8][65];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = (int)blockIdx.x * 64 + 2 * lane;
  const int64_t r0 = (int64_t)blockIdx.y * 64 + w * 8;
  float a0 = 0.f, a1 = 0.f;
  if (c < n) {
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = make_float2(0.f, 0.f);
      if (r0 + i < rows) v[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (r0 + i) * ld + c));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0 += v[i].x; a1 += v[i].y; }
  }
  part[w][2 * lane] = a0;
  part[w][2 * lane + 1] = a1;
  __syncthreads();
  if (w < 2) {
    const int cc = (int)blockIdx.x * 64 + w * 32 + lane;
    if (cc < n) {
      float t = part[0][w * 32 + lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += part[k][w * 32 + lane];
      out[(int64_t)blockIdx.y * n + cc] = t;
    }
  }
}

// The decoder's image layer for the TRAINING graph: y = tanh(ConvTranspose2d(k 4, s 2, p 1)(x) + b) to <= 3 image channels
// (VariationalAutoEncoder.py:134-137), straight from the module's fp32 weight [C_in][co][4][4] (rounded to bf16 on the way into shared
// memory, as the bf16-autocast library conv rounds it).  With 3 output channels a GEMM formulation is all overhead and the library's
// kernel for it takes ~340 us at 1024 frames; here one thread per INPUT pixel produces the 2 x 2 output pixels it is the centre of from
// its 3 x 3 neighbourhood (16-byte NHWC vectors), weights as warp-wide float4 broadcasts from shared memory -- the scheme of vae.cuh's
// convt_last_direct_kernel.   x: bf16 NHWC [N, Hin, Win, CP];  out: fp32 NCHW [N, co_n, 2 Hin, 2 Win].
template <int CP>
__global__ void __launch_bounds__(256) convt_image_fwd_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ W,
                                                              const float* __restrict__ bias, float* __restrict__ out, long npix, int Hin,
                                                              int Win, int co_n) {
  __shared__ float4 w_s[16 * CP];
  for (int i = threadIdx.x; i < 16 * CP; i += blockDim.x) {
    const int ph = i / (4 * CP), r = i - ph * 4 * CP;
    const int tap = r / CP, c = r - tap * CP;
    const int py = ph >> 1, px = ph & 1, ty = tap >> 1, tx = tap & 1;
    // output row 2 i + py takes input row i through ky = 1 (py = 0) / 2 (py = 1) [ty = 0] and input row i - 1 / i + 1 through ky = 3 / 0 [ty = 1]
    const int ky = py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 2 : 0);
    const int kx = px == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 2 : 0);
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < co_n && co < 3; ++co) w[co] = __bfloat162float(__float2bfloat16_rn(W[((long)(c * co_n + co) * 4 + ky) * 4 + kx]));
    w_s[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const float b0 = co_n > 0 ? bias[0] : 0.f, b1 = co_n > 1 ? bias[1] : 0.f, b2 = co_n > 2 ? bias[2] : 0.f;
  const int Ho = 2 * Hin, Wo = 2 * Win;
  for (long pix = (long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (long)gridDim.x * blockDim.x) {
    const int j = (int)(pix % Win);
    const int i = (int)((pix / Win) % Hin);
    const long f = pix / ((long)Win * Hin);
    float acc[4][3];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) { acc[ph][0] = b0; acc[ph][1] = b1; acc[ph][2] = b2; }
#pragma unroll
    for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
      for (int dxi = 0; dxi < 3; ++dxi) {
        const int dy = dyi - 1, dx = dxi - 1;
        const int iy = i + dy, ix = j + dx;
        const bool ok = iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
        const uint4* src = reinterpret_cast<const uint4*>(in + ((f * Hin + iy) * Win + ix) * CP);
#pragma unroll
        for (int c8 = 0; c8 < CP / 8; ++c8) {
          uint4 raw = make_uint4(0, 0, 0, 0);
          if (ok) raw = __ldg(src + c8);
          float xv[8];
          const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { xv[2 * e] = __uint_as_float(rw[e] << 16); xv[2 * e + 1] = __uint_as_float(rw[e] & 0xFFFF0000u); }
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            if (!(dy == 0 || (dy == -1 && py == 0) || (dy == 1 && py == 1))) continue;
            const int ty = dy == 0 ? 0 : 1;
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              if (!(dx == 0 || (dx == -1 && px == 0) || (dx == 1 && px == 1))) continue;
              const int tx = dx == 0 ? 0 : 1;
              const int ph = py * 2 + px;
              const float4* wp = w_s + (ph * 4 + ty * 2 + tx) * CP + c8 * 8;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 w = wp[e];
                acc[ph][0] = fmaf(xv[e], w.x, acc[ph][0]);
                acc[ph][1] = fmaf(xv[e], w.y, acc[ph][1]);
                acc[ph][2] = fmaf(xv[e], w.z, acc[ph][2]);
              }
            }
          }
        }
      }
    }
    for (int co = 0; co < co_n; ++co) {
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        float* o = out + ((f * co_n + co) * Ho + 2 * i + py) * Wo + 2 * j;
        float v0 = acc[py * 2][0], v1 = acc[py * 2 + 1][0];
        if (co == 1) { v0 = acc[py * 2][1]; v1 = acc[py * 2 + 1][1]; }
        if (co == 2) { v0 = acc[py * 2][2]; v1 = acc[py * 2 + 1][2]; }
        *reinterpret_cast<float2*>(o) = make_float2(tanhf(v0), tanhf(v1));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward of one nn.GRUCell step (SequenceModel.py:13,19-24; gates [r; u; n], n = tanh(gi_n + r * gh_n), h' = (1 - u) n + u h)
// from the forward pre-activations gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh:
//   dgi = [dr_pre, du_pre, dn_pre],  dgh = [dr_pre, du_pre, dn_pre * r],  dh_prev (+)= dh * u      (elementwise part only;
//   the caller adds dgh W_hh to dh_prev and takes dx = dgi W_ih).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gru_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gi,
                                                      const float* __restrict__ gh, const float* __restrict__ h_prev,
                                                      float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh_prev,
                                                      int accumulate, int64_t rows, int D, const float* __restrict__ dh_add = nullptr) {
  PDL_ENTRY();
  const int64_t total = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / D;
    const int k = (int)(i - b * D);
    const int64_t g0 = b * 3 * D + k;
    const float ghn = gh[g0 + 2 * D];
    const float r = 1.0f / (1.0f + expf(-(gi[g0] + gh[g0])));
    const float u = 1.0f / (1.0f + expf(-(gi[g0 + D] + gh[g0 + D])));
    const float n = tanhf(gi[g0 + 2 * D] + r * ghn);
    const float hp = h_prev ? h_prev[i] : 0.f;
    const float g = dh[i] + (dh_add ? dh_add[i] : 0.f);
    const float dn_pre = g * (1.0f - u) * (1.0f - n * n);
    const float du_pre = g * (hp - n) * u * (1.0f - u);
    const float dr_pre = dn_pre * ghn * r * (1.0f - r);
    dgi[g0] = dr_pre; dgi[g0 + D] = du_pre; dgi[g0 + 2 * D] = dn_pre;
    dgh[g0] = dr_pre; dgh[g0 + D] = du_pre; dgh[g0 + 2 * D] = dn_pre * r;
    if (dh_prev) dh_prev[i] = (accumulate ? dh_prev[i] : 0.f) + g * u;
  }
}

// ------------------------------------------------------------------------------------------
// Backward of the actor head a = tanh(mu + sigma eps), sigma = softplus(clamp(ls, -5, 2)) + 1e-3 (Agent.py:199-209) for one
// BPTT step:  given the direct gradients g_mu, g_sigma and da (NULL at the last step):
//   du = da (1 - a^2);  dmu = g_mu + du;  dsigma = g_sigma + du eps;  dls = dsigma * sigmoid(ls) * [-5 < ls < 2]
// writes d_head [rows, 2A] = [dmu | dls].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) actor_head_bwd_kernel(const float* __restrict__ g_mu, const float* __restrict__ g_sg,
                                                             const float* __restrict__ da, const float* __restrict__ a,
                                                             const float* __restrict__ eps, const float* __restrict__ ls,
                                                             float* __restrict__ d_head, int64_t rows, int A) {
  PDL_ENTRY();
  const int64_t total = rows * A;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / A;
    const int j = (int)(i - r * A);
    float dmu = g_mu[i], dsg = g_sg[i];
    if (da) {
      const float av = a[i];
      const float du = da[i] * (1.0f - av * av);
      dmu += du;
      dsg = fmaf(du, eps[i], dsg);
    }
    const float l = ls[i];
    const float lc = fminf(fmaxf(l, -5.0f), 2.0f);
    const float dls = (l > -5.0f && l < 2.0f) ? dsg / (1.0f + expf(-lc)) : 0.f;
    d_head[r * 2 * A + j] = dmu;
    d_head[r * 2 * A + A + j] = dls;
  }
}

// log pi(a | mu, sigma) of a tanh-squashed Normal, summed over the action dimension (Agent.py:110-115), and its gradient.
//   y = atanh(clamp(a, -1 + 1e-6, 1 - 1e-6));  logp = sum_j [ -(y - mu)^2 / (2 sigma^2) - log sigma - log sqrt(2 pi) - 2 (log 2 - y - softplus(-2 y)) ]
//   g_mu = coef[row] * (y - mu) / sigma^2;   g_sigma = coef[row] * ((y - mu)^2 / sigma^3 - 1 / sigma)        (a is a constant)
// One thread per row (A <= 32 actions); logp / g_mu / g_sigma may each be NULL.
__global__ void __launch_bounds__(256) tanh_normal_logp_kernel(const float* __restrict__ a, const float* __restrict__ mu,
                                                               const float* __restrict__ sigma, const float* __restrict__ coef,
                                                               float* __restrict__ logp, float* __restrict__ g_mu, float* __restrict__ g_sg,
                                                               int64_t rows, int A) {
  PDL_ENTRY();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float cf = coef ? coef[r] : 1.0f;
    float acc = 0.f;
    for (int j = 0; j < A; ++j) {
      const int64_t i = r * A + j;
      const float av = fminf(fmaxf(a[i], -1.0f + 1e-6f), 1.0f - 1e-6f);
      const float y = atanhf(av);
      const float m = mu[i], sg = sigma[i];
      const float d = y - m, inv = 1.0f / sg;
      if (logp) {
        const float t = -2.0f * y;
        const float sp = t > 20.0f ? t : log1pf(expf(t));
        acc += -0.5f * d * d * inv * inv - logf(sg) - 0.9189385332046727f - 2.0f * (0.6931471805599453f - y - sp);
      }
      if (g_mu) g_mu[i] = cf * d * inv * inv;
      if (g_sg) g_sg[i] = cf * (d * d * inv * inv * inv - inv);
    }
    if (logp) logp[r] = acc;
  }
}

// Percentiles of Agent.update_S (Agent.py:78-88: torch.quantile's definition, sorted[lo] + (sorted[hi] - sorted[lo]) * frac) without a
// sort: the four order statistics (ranks r[0..3]) are found by an MSB-first radix select on the order-preserving integer image of the
// floats, 8 bits per pass, all four selections sharing each pass over the data.  One CTA (the set is 1.5 K - 123 K values and
// L2-resident; a pass is n / 1024 loads per thread); histogram updates are aggregated per warp (__match_any_sync) because the
// leading digits of a return set fall into a handful of bins.  out[0] = v[r0] + (v[r1] - v[r0]) * f_lo, out[1] = v[r2] + (v[r3] -
// v[r2]) * f_hi, out[2] = 1 if every value is finite else 0 (the reference then leaves S unchanged, Agent.py:80-81).
__device__ __forceinline__ unsigned f32_sortable(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unsortable(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__global__ void __launch_bounds__(1024) quantile4_kernel(const float* __restrict__ x, int64_t n, long long r0, long long r1, long long r2,
                                                         long long r3, float f_lo, float f_hi, float* __restrict__ out) {
  __shared__ unsigned hist[4][256];
  __shared__ unsigned prefix[4];
  __shared__ long long rank[4];
  __shared__ int bad;
  const int tid = threadIdx.x;
  if (tid < 4) { prefix[tid] = 0u; rank[tid] = tid == 0 ? r0 : (tid == 1 ? r1 : (tid == 2 ? r2 : r3)); }
  if (tid == 0) bad = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 4 * 256; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const unsigned p0 = prefix[0], p1 = prefix[1], p2 = prefix[2], p3 = prefix[3];
    for (int64_t base = 0; base < n; base += blockDim.x) {       // whole warps iterate together (match_any needs the full mask)
      const int64_t i = base + tid;
      const bool valid = i < n;
      const float xv = valid ? x[i] : 0.f;
      if (pass == 0 && valid && !isfinite(xv)) bad = 1;
      const unsigned key = f32_sortable(xv);
      const unsigned digit = (key >> shift) & 255u;
      const unsigned hi = pass == 0 ? 0u : (key >> (shift + 8));
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned pt = t == 0 ? p0 : (t == 1 ? p1 : (t == 2 ? p2 : p3));
        const bool in = valid && hi == pt;
        const unsigned group = __match_any_sync(0xffffffffu, in ? digit : 0xffffffffu);     // lanes with the same digit (or all idle lanes)
        if (in && (int)(__ffs(group) - 1) == (tid & 31)) atomicAdd(&hist[t][digit], (unsigned)__popc(group));
      }
    }
    __syncthreads();
    if (tid < 4) {
      long long cum = 0, rk = rank[tid];
      unsigned pick = 255u;
      for (int b = 0; b < 256; ++b) {
        const unsigned c = hist[tid][b];
        if (rk < cum + (long long)c) { pick = (unsigned)b; break; }
        cum += c;
      }
      prefix[tid] = (prefix[tid] << 8) | pick;
      rank[tid] = rk - cum;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float v0 = f32_unsortable(prefix[0]), v1 = f32_unsortable(prefix[1]), v2 = f32_unsortable(prefix[2]), v3 = f32_unsortable(prefix[3]);
    out[0] = __fadd_rn(v0, __fmul_rn(v1 - v0, f_lo));      // (no fused multiply-add: the same roundings as the reference's torch expression)
    out[1] = __fadd_rn(v2, __fmul_rn(v3 - v2, f_hi));
    out[2] = bad ? 0.f : 1.f;
  }
}

// KL(Cat(post) || Cat(prior)) summed over the rows of a group (one warp per group).
// WorldModel.py:175-181.
__global__ void __launch_bounds__(256) categorical32_kl_kernel(const float* __restrict__ post,
                                                               const float* __restrict__ prior,
                                                               float* __restrict__ kl, int64_t n_groups,
                                                               int rows_per_group) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    float acc = 0.f;
    for (int r = 0; r < rows_per_group; ++r) {
      const int64_t off = (g * rows_per_group + r) * 32 + lane;
      const float a = __ldg(post + off), b = __ldg(prior + off);
      const float ma = warp_max(a), mb = warp_max(b);
      const float lsa = logf(warp_sum(expf(a - ma))), lsb = logf(warp_sum(expf(b - mb)));
      const float lp = a - ma - lsa, lq = b - mb - lsb;
      acc += expf(lp) * (lp - lq);
    }
    acc = warp_sum(acc);
    if (lane == 0) kl[g] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// replay ring: gather of B windows of L consecutive transitions (modulo capacity).
// Buffer.py:49-61.  One block per (b, t) frame: 4-byte pixel loads (a warp reads 128 contiguous
// bytes), 16-byte float4 stores (a warp writes 512 contiguous bytes).
// Algorithmic bytes per (b, t): frame_bytes + 4 (A + 2) read, 4 * that written (61 480 B at 64x64x3).
// ------------------------------------------------------------------------------------------
template <bool NORM>
__global__ void __launch_bounds__(256) replay_gather_kernel(const uint8_t* __restrict__ ring_obs,
                                                            const float* __restrict__ ring_act,
                                                            const float* __restrict__ ring_rew,
                                                            const float* __restrict__ ring_con,
                                                            const int64_t* __restrict__ starts,
                                                            float* __restrict__ obs_out, float* __restrict__ act_out,
                                                            float* __restrict__ rew_out, float* __restrict__ con_out,
                                                            int L, int64_t cap, int frame_bytes, int A) {
  const int64_t bt = blockIdx.x;
  const int b = (int)(bt / L), t = (int)(bt % L);
  const int64_t src = (__ldg(starts + b) + t) % cap;
  const uchar4* __restrict__ in = reinterpret_cast<const uchar4*>(ring_obs + src * (int64_t)frame_bytes);
  float4* __restrict__ out = reinterpret_cast<float4*>(obs_out + bt * (int64_t)frame_bytes);
  const int nvec = frame_bytes >> 2;
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 4 * 256) {
    uchar4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * 256;
      if (i < nvec) v[j] = __ldg(in + i);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * 256;
      if (i < nvec) {
        float4 f = make_float4((float)v[j].x, (float)v[j].y, (float)v[j].z, (float)v[j].w);
        if (NORM) {
          f.x = f.x / 255.0f - 0.5f; f.y = f.y / 255.0f - 0.5f; f.z = f.z / 255.0f - 0.5f; f.w = f.w / 255.0f - 0.5f;
        }
        __stcs(out + i, f);
      }
    }
  }
  if (threadIdx.x < A) act_out[bt * A + threadIdx.x] = __ldg(ring_act + src * A + threadIdx.x);
  if (threadIdx.x == 32) rew_out[bt] = __ldg(ring_rew + src);
  if (threadIdx.x == 33) con_out[bt] = __ldg(ring_con + src);
}

// Buffer.add_to_buffer (Buffer.py:19-30) for n transitions: one block per transition.
__global__ void __launch_bounds__(256) replay_insert_kernel(uint8_t* __restrict__ ring_obs, float* __restrict__ ring_act,
                                                            float* __restrict__ ring_rew, float* __restrict__ ring_con,
                                                            const uint8_t* __restrict__ obs, const float* __restrict__ act,
                                                            const float* __restrict__ rew, const float* __restrict__ con,
                                                            int64_t next_idx, int64_t cap, int frame_bytes, int A) {
  const int i = blockIdx.x;
  const int64_t dst = (next_idx + i) % cap;
  const uint4* __restrict__ in = reinterpret_cast<const uint4*>(obs + (int64_t)i * frame_bytes);
  uint4* __restrict__ out = reinterpret_cast<uint4*>(ring_obs + dst * (int64_t)frame_bytes);
  for (int k = threadIdx.x; k < (frame_bytes >> 4); k += blockDim.x) out[k] = __ldg(in + k);
  if (threadIdx.x < A) ring_act[dst * A + threadIdx.x] = act[(int64_t)i * A + threadIdx.x];
  if (threadIdx.x == 32) ring_rew[dst] = symlogf_(rew[i]);
  if (threadIdx.x == 33) ring_con[dst] = con[i];
}

// ------------------------------------------------------------------------------------------
// lambda-return reverse scan, one thread per start state, registers only.  Agent.py:158-171.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lambda_return_kernel(const float* __restrict__ rew, const float* __restrict__ cont,
                                                            const float* __restrict__ value, float* __restrict__ out, int B,
                                                            int H, float gamma, float lam) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* r = rew + (int64_t)b * H;
  const float* c = cont + (int64_t)b * H;
  const float* v = value + (int64_t)b * (H + 1);
  float* o = out + (int64_t)b * H;
  float nxt = r[H - 1] + gamma * c[H - 1] * v[H];
  o[H - 1] = nxt;
  for (int t = H - 2; t >= 0; --t) {
    nxt = r[t] + gamma * c[t] * ((1.0f - lam) * v[t + 1] + lam * nxt);
    o[t] = nxt;
  }
}

// ------------------------------------------------------------------------------------------
// two-hot cross-entropy without materialising the two-hot (one warp per row).
// DreamerUtils.py:39-50 + WorldModel.py:137-138 / Agent.py:129-134.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void twohot_index_weight(const float* __restrict__ buckets, int NB, float v, int& idx, float& w) {
  const float lo0 = buckets[0], hi0 = buckets[NB - 1];
  v = fminf(fmaxf(v, lo0), hi0);
  // searchsorted(right=True) - 1 == (#buckets <= v) - 1
  int lo = 0, hi = NB;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (buckets[mid] <= v) lo = mid + 1; else hi = mid;
  }
  idx = lo - 1;
  idx = idx > NB - 2 ? NB - 2 : idx;
  idx = idx < 0 ? 0 : idx;
  const float bl = buckets[idx], bh = buckets[idx + 1];
  w = (v - bl) / (bh - bl + 1e-8f);
}

__global__ void __launch_bounds__(256) twohot_ce_kernel(const float* __restrict__ logits, const float* __restrict__ value,
                                                        const float* __restrict__ buckets, float* __restrict__ ll, int64_t N,
                                                        int NB, int apply_symlog) {
  __shared__ float sb[256];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) sb[i] = buckets[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < N; row += nwarps) {
    const float* x = logits + row * NB;
    float xv[8];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      xv[j] = c < NB ? __ldg(x + c) : -INFINITY;
      m = fmaxf(m, xv[j]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += (lane + 32 * j < NB) ? expf(xv[j] - m) : 0.f;
    const float lse = m + logf(warp_sum(s));
    if (lane == 0) {
      float v = __ldg(value + row);
      if (apply_symlog) v = symlogf_(v);
      int idx; float w;
      twohot_index_weight(sb, NB, v, idx, w);
      // a NaN target stays NaN (torch.clamp / the two-hot weights propagate it, DreamerUtils.py:41-49): the caller's skip sees it
      ll[row] = (v != v) ? v : (1.0f - w) * (__ldg(x + idx) - lse) + w * (__ldg(x + idx + 1) - lse);
    }
  }
}

// Backward of twohot_ce_kernel: d ll[row] / d logits[row][c] = twohot(v)[c] - softmax(logits[row])[c]; one warp per row.
//   dlogits[row][c] = scale * (scale_dev ? *scale_dev : 1) * (coef ? coef[row] : 1) * (twohot[c] - softmax[c])
// (coef: a per-row mask / weight; scale_dev: a device scalar such as 1 / (global element count), so no host read is needed)
__global__ void __launch_bounds__(256) twohot_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ value,
                                                            const float* __restrict__ buckets, const float* __restrict__ coef,
                                                            const float* __restrict__ scale_dev, float scale, float* __restrict__ dlogits,
                                                            int64_t N, int NB, int apply_symlog) {
  PDL_ENTRY();
  __shared__ float sb[256];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) sb[i] = buckets[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float sc = scale * (scale_dev ? __ldg(scale_dev) : 1.0f);
  for (int64_t row = warp; row < N; row += nwarps) {
    const float* x = logits + row * NB;
    float xv[8];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      xv[j] = c < NB ? __ldg(x + c) : -INFINITY;
      m = fmaxf(m, xv[j]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { xv[j] = (lane + 32 * j < NB) ? expf(xv[j] - m) : 0.f; s += xv[j]; }
    const float inv = 1.0f / warp_sum(s);
    float v = __ldg(value + row);
    if (apply_symlog) v = symlogf_(v);
    int idx; float w;
    twohot_index_weight(sb, NB, v, idx, w);
    const float g = sc * (coef ? __ldg(coef + row) : 1.0f);
    float* o = dlogits + row * NB;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      if (c < NB) {
        const float t = c == idx ? 1.0f - w : (c == idx + 1 ? w : 0.f);
        o[c] = (v != v) ? v : g * (t - xv[j] * inv);
      }
    }
  }
}

// symexp(sum(softmax(logits) * buckets)), one warp per row.  DynamicsPredictors.py:70-74.
__global__ void __launch_bounds__(256) bucket_value_kernel(const float* __restrict__ logits, const float* __restrict__ buckets,
                                                           float* __restrict__ value, int64_t N, int NB) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp; row < N; row += nwarps) {
    const float* x = logits + row * NB;
    float xv[8];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      xv[j] = c < NB ? __ldg(x + c) : -INFINITY;
      m = fmaxf(m, xv[j]);
    }
    m = warp_max(m);
    float s = 0.f, ws = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      if (c < NB) {
        const float e = expf(xv[j] - m);
        s += e;
        ws += e * __ldg(buckets + c);
      }
    }
    s = warp_sum(s);
    ws = warp_sum(ws);
    if (lane == 0) value[row] = symexpf_(ws / s);
  }
}

static int rows_grid(int64_t rows, int warps_per_block) {
  int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace drm

using namespace drm;

extern "C" int drm_categorical32_fwd(const float* logits, const float* uniforms, uint8_t* idx, float* z_st, float* probs,
                                     uint16_t* z_bf16, int64_t n_rows, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n_rows >= 0, DRM_ERR_SHAPE, "drm_categorical32_fwd: n_rows < 0");
  if (n_rows == 0) return DRM_OK;  // empty input: nothing to do (pointers may be NULL)
  DRM_REQUIRE(logits && uniforms, DRM_ERR_ARG, "drm_categorical32_fwd: logits/uniforms are NULL");
  DRM_REQUIRE(((uintptr_t)logits % 16 == 0) && (!z_st || (uintptr_t)z_st % 16 == 0) && (!probs || (uintptr_t)probs % 16 == 0) &&
                  (!z_bf16 || (uintptr_t)z_bf16 % 16 == 0),
              DRM_ERR_ALIGN, "drm_categorical32_fwd: 16-byte alignment required");
  categorical32_kernel<<<rows_grid((n_rows + 31) / 32, 8), 256, 0, (cudaStream_t)stream>>>(logits, uniforms, idx, z_st, probs, z_bf16, n_rows, nullptr);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_categorical32_st(const float* logits, const uint8_t* idx, float* z_st, float* probs, int64_t n_rows, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n_rows >= 0, DRM_ERR_SHAPE, "drm_categorical32_st: n_rows < 0");
  if (n_rows == 0) return DRM_OK;
  DRM_REQUIRE(logits && idx, DRM_ERR_ARG, "drm_categorical32_st: logits/idx are NULL");
  DRM_REQUIRE(((uintptr_t)logits % 16 == 0) && (!z_st || (uintptr_t)z_st % 16 == 0) && (!probs || (uintptr_t)probs % 16 == 0), DRM_ERR_ALIGN,
              "drm_categorical32_st: 16-byte alignment required");
  categorical32_kernel<<<rows_grid((n_rows + 31) / 32, 8), 256, 0, (cudaStream_t)stream>>>(logits, nullptr, nullptr, z_st, probs, nullptr, n_rows, idx);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_categorical32_bwd(const float* logits, const float* dz, const float* dz2, const float* dl_add, float* dlogits,
                                     int64_t n_rows, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n_rows >= 0, DRM_ERR_SHAPE, "drm_categorical32_bwd: n_rows < 0");
  if (n_rows == 0) return DRM_OK;
  DRM_REQUIRE(logits && dz && dlogits, DRM_ERR_ARG, "drm_categorical32_bwd: NULL pointer");
  DRM_REQUIRE(((uintptr_t)logits % 16 == 0) && ((uintptr_t)dz % 16 == 0) && ((uintptr_t)dlogits % 16 == 0) &&
                  (!dz2 || (uintptr_t)dz2 % 16 == 0) && (!dl_add || (uintptr_t)dl_add % 16 == 0),
              DRM_ERR_ALIGN, "drm_categorical32_bwd: 16-byte alignment required");
  if (n_rows <= 8192) {   // few rows: warp per row (latency-bound calls of the BPTT recurrence)
    DRM_CUDA(launch_pdl(categorical32_bwd_rowwarp_kernel, (unsigned)((n_rows + 7) / 8), 256, (cudaStream_t)stream, logits, dz, dz2, dl_add, dlogits, n_rows));
    DRM_LAUNCH_CHECK();
    return DRM_OK;
  }
  DRM_CUDA(launch_pdl(categorical32_bwd_kernel, (unsigned)rows_grid((n_rows + 31) / 32, 8), 256, (cudaStream_t)stream, logits, dz, dz2, dl_add, dlogits, n_rows));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_ln_silu_bwd_affine(const float* dy, const float* a, const float* gamma, const float* beta, float* da, float* dln,
                                      float* dlnx, int64_t rows, int32_t n, float eps, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && n >= 1 && n <= 1024, DRM_ERR_SHAPE, "drm_ln_silu_bwd: n must be in [1, 1024]");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(dy && a && gamma && beta && da, DRM_ERR_ARG, "drm_ln_silu_bwd: NULL pointer");
  if (n <= 256) DRM_CUDA(launch_pdl(ln_silu_bwd_kernel<8>, (unsigned)rows_grid(rows, 8), 256, (cudaStream_t)stream, dy, a, gamma, beta, da, dln, rows, n, eps, dlnx));
  else DRM_CUDA(launch_pdl(ln_silu_bwd_kernel<32>, (unsigned)rows_grid(rows, 8), 256, (cudaStream_t)stream, dy, a, gamma, beta, da, dln, rows, n, eps, dlnx));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_ln_silu_bwd(const float* dy, const float* a, const float* gamma, const float* beta, float* da, float* dln,
                               int64_t rows, int32_t n, float eps, void* stream) {
  return drm_ln_silu_bwd_affine(dy, a, gamma, beta, da, dln, nullptr, rows, n, eps, stream);
}

extern "C" int drm_colsum(const float* x, int64_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* scratch, void* stream);

extern "C" int drm_convt_image_fwd(const void* x_nhwc_bf16, const float* weight, const float* bias, float* out_nchw, int32_t N, int32_t Hin,
                                   int32_t Win, int32_t C_in, int32_t C_out, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(N >= 1 && Hin >= 1 && Win >= 1 && C_out >= 1 && C_out <= 3, DRM_ERR_SHAPE, "drm_convt_image_fwd: 1 - 3 image channels");
  DRM_REQUIRE(C_in == 8 || C_in == 16 || C_in == 32 || C_in == 64, DRM_ERR_SHAPE, "drm_convt_image_fwd: C_in must be 8, 16, 32 or 64");
  DRM_REQUIRE(x_nhwc_bf16 && weight && bias && out_nchw, DRM_ERR_ARG, "drm_convt_image_fwd: NULL pointer");
  DRM_REQUIRE(((uintptr_t)x_nhwc_bf16 & 15u) == 0 && ((uintptr_t)out_nchw & 7u) == 0, DRM_ERR_ALIGN, "drm_convt_image_fwd: alignment");
  const long npix = (long)N * Hin * Win;
  const int grid = (int)std::min<long>((npix + 255) / 256, 148L * 16);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_nhwc_bf16);
  cudaStream_t st = (cudaStream_t)stream;
  if (C_in == 8) convt_image_fwd_kernel<8><<<grid, 256, 0, st>>>(x, weight, bias, out_nchw, npix, Hin, Win, C_out);
  else if (C_in == 16) convt_image_fwd_kernel<16><<<grid, 256, 0, st>>>(x, weight, bias, out_nchw, npix, Hin, Win, C_out);
  else if (C_in == 32) convt_image_fwd_kernel<32><<<grid, 256, 0, st>>>(x, weight, bias, out_nchw, npix, Hin, Win, C_out);
  else convt_image_fwd_kernel<64><<<grid, 256, 0, st>>>(x, weight, bias, out_nchw, npix, Hin, Win, C_out);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int64_t drm_colsum_bf16_scratch_bytes(int64_t rows, int32_t n) {
  return rows >= 1 && n >= 1 ? ((rows + 63) / 64 + 64) * (int64_t)n * (int64_t)sizeof(float) : 0;
}

// bf16 input (n even, x 4-byte aligned, ld even): 64-row chunks to fp32 partials, then drm_colsum over the partials
extern "C" int drm_colsum_bf16(const void* x, int64_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* scratch, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 1 && n >= 2 && (n & 1) == 0 && ld >= n && (ld & 1) == 0 && ((uintptr_t)x & 3u) == 0, DRM_ERR_SHAPE,
              "drm_colsum_bf16: n and ld must be even, x 4-byte aligned");
  DRM_REQUIRE(x && out && scratch, DRM_ERR_ARG, "drm_colsum_bf16: NULL pointer");
  const int64_t chunks = (rows + 63) / 64;
  DRM_REQUIRE(chunks <= 0x7fffffff / 1 && chunks <= 65535LL * 1024, DRM_ERR_SHAPE, "drm_colsum_bf16: too many rows");
  float* part = static_cast<float*>(scratch);
  // (grid y is limited to 65535: fold longer chunk lists into several launches)
  for (int64_t y0 = 0; y0 < chunks; y0 += 65535) {
    const int64_t ny = chunks - y0 < 65535 ? chunks - y0 : 65535;
    DRM_CUDA(launch_pdl2(colsum_bf16_stage_kernel, dim3((unsigned)((n + 63) / 64), (unsigned)ny), 256, (cudaStream_t)stream,
                         static_cast<const __nv_bfloat16*>(x) + y0 * 64 * ld, rows - y0 * 64, (int)n, ld, part + y0 * n));
    DRM_LAUNCH_CHECK();
  }
  return drm_colsum(part, chunks, n, n, out, accumulate, part + chunks * n, stream);
}

extern "C" int64_t drm_colsum_scratch_bytes(int64_t rows, int32_t n) {
  return rows > 64 && n >= 1 ? (int64_t)64 * n * (int64_t)sizeof(float) : 0;
}

extern "C" int drm_colsum(const float* x, int64_t rows, int32_t n, int64_t ld, float* out, int32_t accumulate, void* scratch, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && n >= 1 && ld >= n, DRM_ERR_SHAPE, "drm_colsum: bad shape");
  DRM_REQUIRE(x && out, DRM_ERR_ARG, "drm_colsum: NULL pointer");
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (rows > 64) {
    // up to 64 row chunks are summed by their own blocks into scratch [chunks][n], a second launch adds the chunks in order (one block
    // per 32 columns walking all rows is latency-bound: 15 us at 1024 rows against 6 us in two stages)
    DRM_REQUIRE(scratch, DRM_ERR_ARG, "drm_colsum: more than 64 rows need drm_colsum_scratch_bytes() of scratch");
    const int64_t rpy = (rows + 63) / 64 < 32 ? 32 : (rows + 63) / 64;
    const int chunks = (int)((rows + rpy - 1) / rpy);
    cfg.gridDim = dim3((n + 31) / 32, chunks);
    DRM_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel, x, rows, (int)n, ld, static_cast<float*>(scratch), 0, rpy));
    DRM_LAUNCH_CHECK();
    cfg.gridDim = dim3((n + 31) / 32);
    DRM_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel, (const float*)scratch, (int64_t)chunks, (int)n, (int64_t)n, out, (int)accumulate, (int64_t)chunks));
    DRM_LAUNCH_CHECK();
    return DRM_OK;
  }
  cfg.gridDim = dim3((n + 31) / 32);
  DRM_CUDA(cudaLaunchKernelEx(&cfg, colsum_kernel, x, rows, (int)n, ld, out, (int)accumulate, rows > 0 ? rows : (int64_t)1));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_gru_bwd(const float* dh, const float* gi, const float* gh, const float* h_prev, float* dgi, float* dgh,
                           float* dh_prev, int32_t accumulate, int64_t rows, int32_t D, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && D >= 1, DRM_ERR_SHAPE, "drm_gru_bwd: bad shape");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(dh && gi && gh && dgi && dgh, DRM_ERR_ARG, "drm_gru_bwd: NULL pointer");
  const int64_t total = rows * D;
  const int64_t want = (total + 255) / 256;
  DRM_CUDA(launch_pdl(gru_bwd_kernel, (unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, (cudaStream_t)stream, dh, gi, gh, h_prev, dgi, dgh, dh_prev, accumulate, rows, D, (const float*)nullptr));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

// drm_gru_bwd with dh = dh + dh_add: the recurrent term dgh_{t+1} W_hh of the BPTT walk is produced by a GEMM on a side stream and
// joins here instead of through an extra accumulation pass over dh.
extern "C" int drm_gru_bwd_add(const float* dh, const float* dh_add, const float* gi, const float* gh, const float* h_prev, float* dgi,
                               float* dgh, float* dh_prev, int32_t accumulate, int64_t rows, int32_t D, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && D >= 1, DRM_ERR_SHAPE, "drm_gru_bwd_add: bad shape");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(dh && gi && gh && dgi && dgh, DRM_ERR_ARG, "drm_gru_bwd_add: NULL pointer");
  const int64_t total = rows * D;
  const int64_t want = (total + 255) / 256;
  DRM_CUDA(launch_pdl(gru_bwd_kernel, (unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, (cudaStream_t)stream, dh, gi, gh, h_prev, dgi, dgh, dh_prev, accumulate, rows, D, dh_add));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_actor_head_bwd(const float* g_mu, const float* g_sigma, const float* da, const float* a, const float* eps,
                                  const float* log_sigma, float* d_head, int64_t rows, int32_t A, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && A >= 1, DRM_ERR_SHAPE, "drm_actor_head_bwd: bad shape");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(g_mu && g_sigma && log_sigma && d_head && (!da || (a && eps)), DRM_ERR_ARG, "drm_actor_head_bwd: NULL pointer");
  const int64_t want = (rows * A + 255) / 256;
  DRM_CUDA(launch_pdl(actor_head_bwd_kernel, (unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, (cudaStream_t)stream, g_mu, g_sigma, da, a, eps, log_sigma, d_head, rows, A));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_tanh_normal_logp(const float* a, const float* mu, const float* sigma, const float* coef, float* logp, float* g_mu,
                                    float* g_sigma, int64_t rows, int32_t A, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(rows >= 0 && A >= 1 && A <= 32, DRM_ERR_SHAPE, "drm_tanh_normal_logp: A must be in [1, 32]");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(a && mu && sigma && (logp || g_mu || g_sigma), DRM_ERR_ARG, "drm_tanh_normal_logp: NULL pointer");
  const int64_t want = (rows + 255) / 256;
  DRM_CUDA(launch_pdl(tanh_normal_logp_kernel, (unsigned)(want > 148 * 8 ? 148 * 8 : want), 256, (cudaStream_t)stream, a, mu, sigma, coef, logp, g_mu, g_sigma,
                       rows, (int)A));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_percentile_pair(const float* x, int64_t n, double p_lo, double p_hi, float* out, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n >= 1 && p_lo >= 0.0 && p_lo <= 1.0 && p_hi >= 0.0 && p_hi <= 1.0, DRM_ERR_SHAPE, "drm_percentile_pair: n >= 1 and percentiles in [0, 1] required");
  DRM_REQUIRE(x && out, DRM_ERR_ARG, "drm_percentile_pair: NULL pointer");
  // torch.quantile's "linear" rule, as Agent.update_S spells it: pos = p (n - 1), lo = floor(pos), hi = min(lo + 1, n - 1)
  auto split = [n](double p, long long& lo, long long& hi, float& f) {
    const double pos = p * (double)(n - 1);
    lo = (long long)pos;
    hi = lo + 1 < n ? lo + 1 : n - 1;
    f = (float)(pos - (double)lo);
  };
  long long r0, r1, r2, r3;
  float f_lo, f_hi;
  split(p_lo, r0, r1, f_lo);
  split(p_hi, r2, r3, f_hi);
  quantile4_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, r0, r1, r2, r3, f_lo, f_hi, out);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_categorical32_kl(const float* post_logits, const float* prior_logits, float* kl, int64_t n_groups,
                                    int rows_per_group, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n_groups >= 0 && rows_per_group > 0, DRM_ERR_SHAPE, "drm_categorical32_kl: bad shape");
  if (n_groups == 0) return DRM_OK;
  DRM_REQUIRE(post_logits && prior_logits && kl, DRM_ERR_ARG, "drm_categorical32_kl: NULL pointer");
  categorical32_kl_kernel<<<rows_grid(n_groups, 8), 256, 0, (cudaStream_t)stream>>>(post_logits, prior_logits, kl, n_groups, rows_per_group);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_replay_gather(const uint8_t* ring_obs, const float* ring_act, const float* ring_rew, const float* ring_con,
                                 const int64_t* starts, float* obs_out, float* act_out, float* rew_out, float* con_out,
                                 int32_t B, int32_t L, int64_t cap, int32_t frame_bytes, int32_t A, int32_t normalise,
                                 void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(B >= 0 && L > 0 && cap > 0 && A > 0 && A <= 32, DRM_ERR_SHAPE, "drm_replay_gather: bad shape");
  if (B == 0) return DRM_OK;
  DRM_REQUIRE(ring_obs && ring_act && ring_rew && ring_con && starts && obs_out && act_out && rew_out && con_out, DRM_ERR_ARG,
              "drm_replay_gather: NULL pointer");
  DRM_REQUIRE(frame_bytes > 0 && frame_bytes % 16 == 0, DRM_ERR_ALIGN, "drm_replay_gather: frame_bytes must be a multiple of 16");
  DRM_REQUIRE(((uintptr_t)ring_obs % 16 == 0) && ((uintptr_t)obs_out % 16 == 0), DRM_ERR_ALIGN, "drm_replay_gather: 16-byte alignment required");
  if (B == 0) return DRM_OK;
  const int64_t frames = (int64_t)B * L;
  DRM_REQUIRE(frames < (1ll << 31), DRM_ERR_SHAPE, "drm_replay_gather: B * L too large");
  if (normalise)
    replay_gather_kernel<true><<<(unsigned)frames, 256, 0, (cudaStream_t)stream>>>(ring_obs, ring_act, ring_rew, ring_con, starts, obs_out, act_out, rew_out, con_out, L, cap, frame_bytes, A);
  else
    replay_gather_kernel<false><<<(unsigned)frames, 256, 0, (cudaStream_t)stream>>>(ring_obs, ring_act, ring_rew, ring_con, starts, obs_out, act_out, rew_out, con_out, L, cap, frame_bytes, A);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_replay_insert(uint8_t* ring_obs, float* ring_act, float* ring_rew, float* ring_con, const uint8_t* obs,
                                 const float* act, const float* rew, const float* con, int64_t next_idx, int32_t n, int64_t cap,
                                 int32_t frame_bytes, int32_t A, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(ring_obs && ring_act && ring_rew && ring_con && obs && act && rew && con, DRM_ERR_ARG, "drm_replay_insert: NULL pointer");
  DRM_REQUIRE(n >= 0 && n <= cap && cap > 0 && A > 0 && A <= 32 && next_idx >= 0 && next_idx < cap, DRM_ERR_SHAPE, "drm_replay_insert: bad shape");
  DRM_REQUIRE(frame_bytes > 0 && frame_bytes % 16 == 0, DRM_ERR_ALIGN, "drm_replay_insert: frame_bytes must be a multiple of 16");
  DRM_REQUIRE(((uintptr_t)ring_obs % 16 == 0) && ((uintptr_t)obs % 16 == 0), DRM_ERR_ALIGN, "drm_replay_insert: 16-byte alignment required");
  if (n == 0) return DRM_OK;
  replay_insert_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(ring_obs, ring_act, ring_rew, ring_con, obs, act, rew, con, next_idx, cap, frame_bytes, A);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_lambda_return(const float* rew, const float* cont, const float* value, float* out, int32_t B, int32_t H,
                                 float gamma, float lambda_, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(B >= 0 && H >= 1, DRM_ERR_SHAPE, "drm_lambda_return: bad shape");
  if (B == 0) return DRM_OK;
  DRM_REQUIRE(rew && cont && value && out, DRM_ERR_ARG, "drm_lambda_return: NULL pointer");
  lambda_return_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(rew, cont, value, out, B, H, gamma, lambda_);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_twohot_ce_bwd(const float* logits, const float* value, const float* buckets, const float* coef, const float* scale_dev,
                                 float scale, float* dlogits, int64_t N, int32_t NB, int32_t apply_symlog, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(N >= 0 && NB >= 2 && NB <= 256, DRM_ERR_SHAPE, "drm_twohot_ce_bwd: NB must be in [2, 256]");
  if (N == 0) return DRM_OK;
  DRM_REQUIRE(logits && value && buckets && dlogits, DRM_ERR_ARG, "drm_twohot_ce_bwd: NULL pointer");
  DRM_CUDA(launch_pdl(twohot_ce_bwd_kernel, (unsigned)rows_grid(N, 8), 256, (cudaStream_t)stream, logits, value, buckets, coef, scale_dev, scale,
                       dlogits, N, (int)NB, (int)apply_symlog));
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_twohot_ce(const float* logits, const float* value, const float* buckets, float* ll, int64_t N, int32_t NB,
                             int32_t apply_symlog, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(N >= 0 && NB >= 2 && NB <= 256, DRM_ERR_SHAPE, "drm_twohot_ce: NB must be in [2, 256]");
  if (N == 0) return DRM_OK;
  DRM_REQUIRE(logits && value && buckets && ll, DRM_ERR_ARG, "drm_twohot_ce: NULL pointer");
  twohot_ce_kernel<<<rows_grid(N, 8), 256, 0, (cudaStream_t)stream>>>(logits, value, buckets, ll, N, NB, apply_symlog);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

// one-hot expansion of sampled classes: idx [n_rows] u8 (class < 32) -> out [n_rows, 32] fp32 (one float4 per thread)
__global__ void onehot32_kernel(const uint8_t* __restrict__ idx, float* __restrict__ out, int64_t n_rows) {
  const int64_t total = n_rows * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 3;
    const int c0 = (int)(i & 7) * 4, k = (int)idx[r] - c0;
    reinterpret_cast<float4*>(out)[i] = make_float4(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f, k == 3 ? 1.f : 0.f);
  }
}

extern "C" int drm_onehot32(const uint8_t* idx, float* out, int64_t n_rows, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(n_rows >= 0, DRM_ERR_SHAPE, "drm_onehot32: negative row count");
  if (n_rows == 0) return DRM_OK;
  DRM_REQUIRE(idx && out, DRM_ERR_ARG, "drm_onehot32: NULL pointer");
  DRM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, DRM_ERR_ALIGN, "drm_onehot32: out must be 16-byte aligned");
  int64_t blocks = (n_rows * 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  onehot32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(idx, out, n_rows);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

extern "C" int drm_bucket_value(const float* logits, const float* buckets, float* value, int64_t N, int32_t NB, void* stream) {
  if (int rc = check_arch()) return rc;
  DRM_REQUIRE(N >= 0 && NB >= 1 && NB <= 256, DRM_ERR_SHAPE, "drm_bucket_value: NB must be in [1, 256]");
  if (N == 0) return DRM_OK;
  DRM_REQUIRE(logits && buckets && value, DRM_ERR_ARG, "drm_bucket_value: NULL pointer");
  bucket_value_kernel<<<rows_grid(N, 8), 256, 0, (cudaStream_t)stream>>>(logits, buckets, value, N, NB);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
