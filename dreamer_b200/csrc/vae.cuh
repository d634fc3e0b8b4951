// VAE encoder / decoder and the posterior (observe) scan on the same TMA + tcgen05 GEMM stages.
// Included at the end of rssm.cu (one translation unit shares the kernels and launch helpers).
//
// Convolutions (VariationalAutoEncoder.py:33-42, 128-137) run as   patch gather (bf16, 16-byte
// vectors)  ->  fused GEMM (bias + SiLU / tanh in the epilogue, coalesced NHWC copy-out):
//   Conv2d k4 s2 p1         : one patch row per output pixel, K = 16 taps x C_in
//   ConvTranspose2d k4 s2 p1: four sub-pixel phases, each a 2x2-tap convolution with K = 4 x C_in;
//                             the phases are the GEMM's blockIdx.y and scatter to (2q+py, 2r+px)
// Activations are NHWC bf16 with the channel count padded to a multiple of 16.
//
// HBM layout of the observe workspace (B sequences, T steps):
//   S      bf16 [(T+1)*B (+pad), KS]   time-major slabs; slab 0 = the zero state, slab t+1 = (z_t | a_t | h_t)
//   Y1, Y2 bf16 [6 * rows_p, 256]      hidden activations, same row indexing as S inside each slot
//   feat   bf16 [T*B, Kf]              conv features, time-major; featpart fp32 [T*B, bn] = features x W_feat^T
#pragma once

namespace drm {

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
__global__ void pack_offsets_kernel(__nv_bfloat16* __restrict__ dst, int ld_dst, int row0, int nrows, int col0, int ncols,
                                    const float* __restrict__ src, const int* __restrict__ roff, const int* __restrict__ coff) {
  const long total = (long)nrows * ncols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ncols), c = (int)(i % ncols);
    const int a = roff[r], b = coff[c];
    dst[(long)(row0 + r) * ld_dst + col0 + c] = __float2bfloat16_rn((a >= 0 && b >= 0) ? src[(long)a + b] : 0.f);
  }
}

// First encoder layer: fp32 NCHW frames -> bf16 patch rows [frame, oy, ox][(ky*4+kx)*3 + c], K padded 48 -> 64.
// Patch frame n reads source frame (tmB > 0 ? ((f0+n) % tmB) * tmT + (f0+n) / tmB : f0+n)  (time-major <- batch-major).
__global__ void im2col_first_kernel(const float* __restrict__ obs, __nv_bfloat16* __restrict__ out, long n_frames, int f0, int H,
                                    int W, int tmB, int tmT) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long total = n_frames * Ho * Wo * 8;  // 8 vectors of 8 bf16 per row
  const bool pow2 = !(Wo & (Wo - 1)) && !(Ho & (Ho - 1)) && (total >> 3) < (1l << 31);
  const int s_wo = __ffs(Wo) - 1, s_ho = __ffs(Ho) - 1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i & 7);
    const long row = i >> 3;
    int ox, oy;
    long n;
    if (pow2) {                               // 32-bit shifts / masks instead of 64-bit div / mod (image sides are powers of two)
      const unsigned r32 = (unsigned)row;
      ox = (int)(r32 & (unsigned)(Wo - 1));
      oy = (int)((r32 >> s_wo) & (unsigned)(Ho - 1));
      n = r32 >> (s_wo + s_ho);
    } else {
      ox = (int)(row % Wo);
      oy = (int)((row / Wo) % Ho);
      n = row / ((long)Wo * Ho);
    }
    const unsigned gn = (unsigned)(f0 + n);
    const long sn = tmB > 0 ? (long)(gn % (unsigned)tmB) * tmT + gn / (unsigned)tmB : (long)gn;
    const float* fr = obs + sn * 3 * H * W;
    float e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      float x = 0.f;
      if (k < 48) {
        const int tap = k / 3, c = k - tap * 3;
        const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) x = __ldg(fr + ((long)c * H + iy) * W + ix);
      }
      e[j] = x;
    }
    *reinterpret_cast<uint4*>(out + row * 64 + v * 8) =
        make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
  }
}

struct Taps { int n; int dy[16]; int dx[16]; };
struct Taps4 { Taps t[4]; };
// NHWC bf16 [frames, Hin, Win, cp] -> patch rows [phase][frame, oy, ox][tap * cp + c]  (cp % 8 == 0, 16-byte vectors)
// input pixel = (oy * stride + dy[tap], ox * stride + dx[tap]), zero outside the image.  blockIdx.y = phase.
__global__ void patch_gather_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows, int Hin, int Win,
                                    int cp, int Ho, int Wo, int stride, const __grid_constant__ Taps4 taps, long phase_rows, int ld_out) {
  const Taps& tp = taps.t[blockIdx.y];
  const int vpt = cp >> 3;                       // vectors per tap
  const int vpr = tp.n * vpt;                    // vectors per row
  const long total = rows * vpr;
  __nv_bfloat16* o = out + (long)blockIdx.y * phase_rows * ld_out;
  // The index arithmetic (five 64-bit div / mod per 16-byte vector) used to cost more than the copy.  In every configuration
  // the reference uses, tap counts, channel counts and image sides are powers of two: shifts and masks on 32-bit indices.
  const bool pow2 = !(vpt & (vpt - 1)) && !(vpr & (vpr - 1)) && !(Wo & (Wo - 1)) && !(Ho & (Ho - 1)) && total < (1l << 31);
  if (pow2) {
    const int s_vpr = __ffs(vpr) - 1, s_vpt = __ffs(vpt) - 1, s_wo = __ffs(Wo) - 1, s_ho = __ffs(Ho) - 1;
    const unsigned n_total = (unsigned)total;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += gridDim.x * blockDim.x) {
      const unsigned row = i >> s_vpr, w = i & (unsigned)(vpr - 1);
      const int t = (int)(w >> s_vpt), c8 = (int)(w & (unsigned)(vpt - 1));
      const int ox = (int)(row & (unsigned)(Wo - 1));
      const int oy = (int)((row >> s_wo) & (unsigned)(Ho - 1));
      const long n = row >> (s_wo + s_ho);
      const int iy = oy * stride + tp.dy[t], ix = ox * stride + tp.dx[t];
      uint4 val = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
        val = __ldg(reinterpret_cast<const uint4*>(in + ((n * Hin + iy) * Win + ix) * cp + c8 * 8));
      *reinterpret_cast<uint4*>(o + (long)row * ld_out + t * cp + c8 * 8) = val;
    }
    return;
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / vpr;
    const int w = (int)(i - row * vpr);
    const int t = w / vpt, c8 = w - t * vpt;
    const int ox = (int)(row % Wo);
    const int oy = (int)((row / Wo) % Ho);
    const long n = row / ((long)Wo * Ho);
    const int iy = oy * stride + tp.dy[t], ix = ox * stride + tp.dx[t];
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
      val = __ldg(reinterpret_cast<const uint4*>(in + ((n * Hin + iy) * Win + ix) * cp + c8 * 8));
    *reinterpret_cast<uint4*>(o + row * ld_out + t * cp + c8 * 8) = val;
  }
}

// Last decoder layer (ConvTranspose2d k4 s2 p1 to <= 3 image channels + tanh, VariationalAutoEncoder.py:134-137) WITHOUT the
// patch matrix: with 3 output channels the GEMM formulation is all overhead (K = 4 taps x cp, N = 3: 32 768 tiles of 128 rows at
// config 3, 1 GB of patches written and read back).  One thread per INPUT pixel produces the 2 x 2 output pixels it is the
// centre of (sub-pixel phases py, px): 16 (neighbour, phase) pairs x cp channels x co FMAs from the 3 x 3 neighbourhood, read
// as 16-byte NHWC vectors through L1; weights (the packed bf16 phase matrices of the GEMM path, so both paths round alike) sit
// in shared memory as one float4 per (phase, tap, channel) and are read as warp-wide broadcasts.
//   in  NHWC bf16 [nf, Hin, Win, cp];  Wp bf16 [(phase * co_pad + co)][tap * cp + c] (row pitch ldw);  out fp32 NCHW.
template <int CP>
__global__ void __launch_bounds__(256) convt_last_direct_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ Wp,
                                                                const float* __restrict__ bias, float* __restrict__ out, long npix,
                                                                int Hin, int Win, int co_pad, int co_n, int ldw) {
  __shared__ float4 w_s[16 * CP];
  for (int i = threadIdx.x; i < 16 * CP; i += blockDim.x) {
    const int ph = i / (4 * CP), r = i - ph * 4 * CP;           // r = tap * CP + c
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int co = 0; co < co_n && co < 3; ++co) w[co] = __bfloat162float(Wp[(long)(ph * co_pad + co) * ldw + r]);
    w_s[i] = make_float4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const float b0 = co_n > 0 ? bias[0] : 0.f, b1 = co_n > 1 ? bias[1] : 0.f, b2 = co_n > 2 ? bias[2] : 0.f;   // co_n <= 3 image channels
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
          // phases this neighbour feeds: dy = 0 -> tap ty = 0 of both py; dy = -1 -> py = 0, ty = 1; dy = +1 -> py = 1, ty = 1
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
        *reinterpret_cast<float2*>(o) = make_float2(tanhf_(v0), tanhf_(v1));
      }
    }
  }
}

// act[b, t, :] -> the action columns of slab t + 1
__global__ void pack_actions_kernel(__nv_bfloat16* __restrict__ S, int ld_s, int col0, const float* __restrict__ act, int B, int T, int A) {
  const long total = (long)B * T * A;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % A);
    const int t = (int)((i / A) % T);
    const int b = (int)(i / ((long)A * T));
    S[((long)(t + 1) * B + b) * ld_s + col0 + j] = __float2bfloat16_rn(act[i]);
  }
}

__global__ void neg_sse_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int len) {
  __shared__ float red[32];
  const long row = blockIdx.x;
  const float4* a4 = reinterpret_cast<const float4*>(a + row * len);
  const float4* b4 = reinterpret_cast<const float4*>(b + row * len);
  float s = 0.f;
  for (int i = threadIdx.x; i < (len >> 2); i += blockDim.x) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (int i = (len & ~3) + threadIdx.x; i < len; i += blockDim.x) { const float d = a[row * len + i] - b[row * len + i]; s += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[row] = -s;
  }
}

struct OffOp { __nv_bfloat16* dst; int ld_dst, row0, nrows, col0, ncols, src; int *roff, *coff; };
enum VSrc { V_ECW = 0, V_ECB = 4, V_EL1W = 8, V_EL1B, V_ELNG, V_ELNB, V_EL2W, V_EL2B, V_DL1W, V_DL1B, V_DLNG, V_DLNB, V_DL2W, V_DL2B,
            V_DCW = 20, V_DCB = 24, V_COUNT = 28 };

}  // namespace drm

struct drm_vae {
  drm_rssm* m;
  drm_vae_dims d;
  int hw4;                // (H / 16) * (W / 16)
  int ec[5], ebn[5];      // encoder channels 3, e1, e2, 2 e2, 4 e2 and their padded pitches
  int dc[5], dbn[5];      // decoder channels 4 d2, 2 d2, d2, d1, 3 and their padded pitches
  int EK[4], DK[4];       // GEMM K of encoder conv i / decoder convT j
  int Kf, feat, dfeat;    // padded feature columns, reference feature counts
  int bn_he, bn_hd;
  __nv_bfloat16 *We[4], *Wef, *Weh, *We3, *Wd1, *Wd2, *Wdc[4];
  float *be[4], *e1_b, *e1_g, *e1_be, *e3_b, *d1_b, *d1_g, *d1_be, *d2_b, *bdc[4];
  CUtensorMap tmWe[4], tmWef, tmWeh, tmWehq, tmWe3, tmWe3h, tmWd1, tmWd1q, tmWd2, tmWdc[4];
  CUtensorMap tmWeI[4], tmWdcI[4];   // implicit-GEMM twins of tmWe / tmWdc: box {chunk, bn} (conv_implicit.cuh)
  int echunk[4], dchunk[4];          // channels per k-block of encoder conv i / decoder convT j (0: layer stays on the patch path)
  std::vector<drm::OffOp> mat_ops;
  std::vector<drm::VecOp> vec_ops;
  std::vector<void*> allocs;
  bool packed;
};

struct drm_obs_persist;   // schedule + counters of the persistent scan kernel (observe_persist.cuh)
namespace drm { static void obs_persist_free(drm_obs_persist* ps); }

struct drm_observe {
  drm_obs_persist* ps = nullptr;
  drm_rssm* m;
  drm_vae* v;
  int B, T, rows, rows_p, FC;
  __nv_bfloat16 *S, *Y1, *Y2, *feat, *dec0, *patch, *actA, *actB;
  float *zero_h, *featpart;
  CUtensorMap tmS, tmY1, tmY2, tmFeat, tmPatchE[4], tmPatchD[4];
  CUtensorMap tmI2cE[4];       // im2col maps over the NHWC input of encoder conv i (i = 1 .. 3)
  CUtensorMap tmI2cD[4][4];    // ... of decoder convT j (j = 0 .. 2), one per sub-pixel phase
  bool i2cE[4] = {false, false, false, false}, i2cD[4] = {false, false, false, false};   // the layer has its im2col map(s)
  CUtensorMap tmS_s, tmY1_s, tmY2_s;   // short-box twins (SMALL_A_ROWS rows) for steps with few sequences
  long patch_elems, act_elems;
  std::vector<void*> allocs;
  bool scanned;
};

namespace drm {

static int add_off(drm_vae* v, __nv_bfloat16* dst, int ld_dst, int row0, int col0, int src, const std::vector<int>& roff,
                   const std::vector<int>& coff) {
  OffOp op{dst, ld_dst, row0, (int)roff.size(), col0, (int)coff.size(), src, nullptr, nullptr};
  if (int rc = upload_map(v->allocs, roff, &op.roff)) return rc;
  if (int rc = upload_map(v->allocs, coff, &op.coff)) return rc;
  v->mat_ops.push_back(op);
  return DRM_OK;
}
static int add_vvec(drm_vae* v, float* dst, int src, const std::vector<int>& map) {
  VecOp op{dst, (int)map.size(), src, nullptr, 0.f};
  if (int rc = upload_map(v->allocs, map, &op.map)) return rc;
  v->vec_ops.push_back(op);
  return DRM_OK;
}
static std::vector<int> scaled(int n, int limit, int scale, int offset = 0) {  // i -> offset + i * scale, or -1
  std::vector<int> r(n);
  for (int i = 0; i < n; ++i) r[i] = i < limit ? offset + i * scale : -1;
  return r;
}
static WsView view_of(drm_observe* o, int row0) {
  return WsView{&o->tmS, o->S, &o->tmY1, &o->tmY2, o->Y1, o->Y2, o->rows_p, row0, &o->tmS_s, &o->tmY1_s, &o->tmY2_s};
}
// sub-pixel decomposition of ConvTranspose2d(k4, s2, p1): output o = 2 q + p gets taps
//   p = 0: (input q, kernel 1), (q - 1, kernel 3);   p = 1: (q, kernel 2), (q + 1, kernel 0)
static inline int ct_k(int p, int t) { return p == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 2 : 0); }
static inline int ct_d(int p, int t) { return t == 0 ? 0 : (p == 0 ? -1 : 1); }

}  // namespace drm

extern "C" int drm_vae_create(drm_rssm* m, const drm_vae_dims* dims, drm_vae** out) {
  RC(check_arch());
  DRM_REQUIRE(m && dims && out, DRM_ERR_ARG, "drm_vae_create: NULL argument");
  DRM_REQUIRE(!m->wide, DRM_ERR_ARG, "drm_vae_create: the observe / VAE path runs on DRM_PRECISION_BF16 handles only");
  const drm_vae_dims d = *dims;
  DRM_REQUIRE(d.H >= 16 && d.W >= 16 && d.H % 16 == 0 && d.W % 16 == 0, DRM_ERR_SHAPE, "drm_vae_create: H, W must be multiples of 16");
  DRM_REQUIRE(d.e1 >= 1 && d.e2 >= 1 && d.d1 >= 1 && d.d2 >= 1 && 4 * d.e2 <= 256 && 4 * d.d2 <= 256 && d.e1 <= 256 && d.d1 <= 256,
              DRM_ERR_SHAPE, "drm_vae_create: conv channel counts must be in [1, 256]");
  DRM_REQUIRE(d.h_enc >= 1 && d.h_enc <= 256 && d.h_dec >= 1 && d.h_dec <= 256, DRM_ERR_SHAPE, "drm_vae_create: hidden sizes must be in [1, 256]");
  drm_vae* v = new drm_vae();
  v->m = m; v->d = d; v->packed = false;
  v->hw4 = (d.H / 16) * (d.W / 16);
  const int ec[5] = {3, d.e1, d.e2, 2 * d.e2, 4 * d.e2};
  const int dc[5] = {4 * d.d2, 2 * d.d2, d.d2, d.d1, 3};
  for (int i = 0; i < 5; ++i) { v->ec[i] = ec[i]; v->ebn[i] = round_up(ec[i], 16); v->dc[i] = dc[i]; v->dbn[i] = round_up(dc[i], 16); }
  v->EK[0] = 64;
  for (int i = 1; i < 4; ++i) v->EK[i] = 16 * v->ebn[i];
  for (int j = 0; j < 4; ++j) v->DK[j] = 4 * v->dbn[j];
  v->Kf = v->hw4 * v->ebn[4];
  v->feat = v->hw4 * ec[4];
  v->dfeat = v->hw4 * dc[0];
  v->bn_he = round_up(d.h_enc, 32);
  v->bn_hd = round_up(d.h_dec, 32);
  DRM_REQUIRE(v->Kf % 64 == 0 && (v->hw4 * v->dbn[0]) % 256 == 0, DRM_ERR_SHAPE, "drm_vae_create: unsupported feature-map size");
  const int D = m->d.D, ZP = m->ZP, DP = m->DP, KH = m->KH;
  auto& bag = v->allocs;
  int rc = DRM_OK;
#define TRY(x) if (rc == DRM_OK) rc = (x)
  for (int i = 0; i < 4; ++i) {
    TRY(dev_alloc(bag, &v->We[i], (size_t)v->ebn[i + 1] * v->EK[i]));
    TRY(dev_alloc(bag, &v->be[i], (size_t)v->ebn[i + 1]));
    TRY(dev_alloc(bag, &v->Wdc[i], (size_t)4 * v->dbn[i + 1] * v->DK[i]));
    TRY(dev_alloc(bag, &v->bdc[i], (size_t)v->dbn[i + 1]));
  }
  TRY(dev_alloc(bag, &v->Wef, (size_t)v->bn_he * v->Kf));
  TRY(dev_alloc(bag, &v->Weh, (size_t)256 * DP));
  TRY(dev_alloc(bag, &v->We3, (size_t)ZP * 256));
  TRY(dev_alloc(bag, &v->e1_b, (size_t)v->bn_he)); TRY(dev_alloc(bag, &v->e1_g, (size_t)v->bn_he)); TRY(dev_alloc(bag, &v->e1_be, (size_t)v->bn_he));
  TRY(dev_alloc(bag, &v->e3_b, (size_t)ZP));
  TRY(dev_alloc(bag, &v->Wd1, (size_t)256 * KH));
  TRY(dev_alloc(bag, &v->Wd2, (size_t)v->hw4 * v->dbn[0] * 256));
  TRY(dev_alloc(bag, &v->d1_b, (size_t)v->bn_hd)); TRY(dev_alloc(bag, &v->d1_g, (size_t)v->bn_hd)); TRY(dev_alloc(bag, &v->d1_be, (size_t)v->bn_hd));
  TRY(dev_alloc(bag, &v->d2_b, (size_t)v->hw4 * v->dbn[0]));
  // ---- encoder convs: W [co, ci, 4, 4] -> rows co, columns (tap, ci)
  for (int i = 0; i < 4 && rc == DRM_OK; ++i) {
    const int ci_n = ec[i], co_n = ec[i + 1];
    std::vector<int> coff(v->EK[i], -1);
    if (i == 0) {
      for (int k = 0; k < 48; ++k) coff[k] = (k % 3) * 16 + k / 3;
    } else {
      for (int t = 0; t < 16; ++t)
        for (int c = 0; c < ci_n; ++c) coff[t * v->ebn[i] + c] = c * 16 + t;
    }
    TRY(add_off(v, v->We[i], v->EK[i], 0, 0, V_ECW + i, scaled(v->ebn[i + 1], co_n, ci_n * 16), coff));
    TRY(add_vvec(v, v->be[i], V_ECB + i, iota_lim(v->ebn[i + 1], co_n)));
  }
  // ---- latent_mapper.0 split into the feature part (hoisted) and the h part (in the scan); .3 -> logits
  {
    std::vector<int> cf(v->Kf, -1);
    for (int p = 0; p < v->hw4; ++p)
      for (int c = 0; c < ec[4]; ++c) cf[p * v->ebn[4] + c] = c * v->hw4 + p;   // reference flatten order is (c, y, x)
    const std::vector<int> rr = scaled(v->bn_he, d.h_enc, v->feat + D);
    TRY(add_off(v, v->Wef, v->Kf, 0, 0, V_EL1W, rr, cf));
    TRY(add_off(v, v->Weh, DP, 0, 0, V_EL1W, rr, scaled(DP, D, 1, v->feat)));
    TRY(add_vvec(v, v->e1_b, V_EL1B, iota_lim(v->bn_he, d.h_enc)));
    TRY(add_vvec(v, v->e1_g, V_ELNG, iota_lim(v->bn_he, d.h_enc)));
    TRY(add_vvec(v, v->e1_be, V_ELNB, iota_lim(v->bn_he, d.h_enc)));
    TRY(add_off(v, v->We3, 256, 0, 0, V_EL2W, scaled(ZP, ZP, d.h_enc), iota_lim(256, d.h_enc)));
    TRY(add_vvec(v, v->e3_b, V_EL2B, iota_lim(ZP, ZP)));
  }
  // ---- decoder upscaler: .0 on [h, z] (packed columns [z | h]); .3 rows permuted to NHWC
  {
    std::vector<int> c1(KH);
    for (int c = 0; c < KH; ++c) c1[c] = c < ZP ? D + c : (c - ZP < D ? c - ZP : -1);
    TRY(add_off(v, v->Wd1, KH, 0, 0, V_DL1W, scaled(v->bn_hd, d.h_dec, D + ZP), c1));
    TRY(add_vvec(v, v->d1_b, V_DL1B, iota_lim(v->bn_hd, d.h_dec)));
    TRY(add_vvec(v, v->d1_g, V_DLNG, iota_lim(v->bn_hd, d.h_dec)));
    TRY(add_vvec(v, v->d1_be, V_DLNB, iota_lim(v->bn_hd, d.h_dec)));
    const int nr = v->hw4 * v->dbn[0];
    std::vector<int> r2(nr, -1), bmap(nr, -1);
    for (int p = 0; p < v->hw4; ++p)
      for (int c = 0; c < dc[0]; ++c) { r2[p * v->dbn[0] + c] = (c * v->hw4 + p) * d.h_dec; bmap[p * v->dbn[0] + c] = c * v->hw4 + p; }
    TRY(add_off(v, v->Wd2, 256, 0, 0, V_DL2W, r2, iota_lim(256, d.h_dec)));
    TRY(add_vvec(v, v->d2_b, V_DL2B, bmap));
  }
  // ---- decoder transposed convs: W [ci, co, 4, 4], one packed block of rows per sub-pixel phase
  for (int j = 0; j < 4 && rc == DRM_OK; ++j) {
    const int ci_n = dc[j], co_n = dc[j + 1];
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      std::vector<int> coff(v->DK[j], -1);
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx)
          for (int c = 0; c < ci_n; ++c)
            coff[(ty * 2 + tx) * v->dbn[j] + c] = c * co_n * 16 + ct_k(py, ty) * 4 + ct_k(px, tx);
      TRY(add_off(v, v->Wdc[j], v->DK[j], ph * v->dbn[j + 1], 0, V_DCW + j, scaled(v->dbn[j + 1], co_n, 16), coff));
    }
    TRY(add_vvec(v, v->bdc[j], V_DCB + j, iota_lim(v->dbn[j + 1], co_n)));
  }
  for (int i = 0; i < 4; ++i) {
    TRY(make_tmap_bf16_2d(&v->tmWe[i], v->We[i], v->ebn[i + 1], v->EK[i], v->EK[i], v->ebn[i + 1]));
    TRY(make_tmap_bf16_2d(&v->tmWdc[i], v->Wdc[i], 4 * v->dbn[i + 1], v->DK[i], v->DK[i], v->dbn[i + 1]));
  }
  // implicit-GEMM layers: k-block = one tap x `chunk` channels; the output width must suit conv_implicit_kernel's epilogue
  auto chunk_of = [](int cpad, int bn) {
    if (bn % 16 || bn > 256 || (bn < 64 && bn != 16 && bn != 32)) return 0;
    return cpad % 64 == 0 ? 64 : (cpad % 32 == 0 ? 32 : 16);
  };
  for (int i = 0; i < 4; ++i) {
    v->echunk[i] = i == 0 ? 0 : chunk_of(v->ebn[i], v->ebn[i + 1]);   // (layer 0 reads fp32 NCHW frames: im2col_first_kernel)
    v->dchunk[i] = i == 3 ? 0 : chunk_of(v->dbn[i], v->dbn[i + 1]);   // (image layer: convt_last_direct_kernel)
    if (v->echunk[i]) TRY(make_tmap_bf16_2d_inner(&v->tmWeI[i], v->We[i], v->ebn[i + 1], v->EK[i], v->EK[i], v->ebn[i + 1], v->echunk[i]));
    if (v->dchunk[i]) TRY(make_tmap_bf16_2d_inner(&v->tmWdcI[i], v->Wdc[i], 4 * v->dbn[i + 1], v->DK[i], v->DK[i], v->dbn[i + 1], v->dchunk[i]));
  }
  TRY(make_tmap_bf16_2d(&v->tmWef, v->Wef, v->bn_he, v->Kf, v->Kf, v->bn_he));
  TRY(make_tmap_bf16_2d(&v->tmWeh, v->Weh, 256, DP, DP, v->bn_he));
  TRY(make_tmap_bf16_2d(&v->tmWehq, v->Weh, 256, DP, DP, 64));
  TRY(make_tmap_bf16_2d(&v->tmWe3, v->We3, ZP, 256, 256, 256));
  TRY(make_tmap_bf16_2d(&v->tmWe3h, v->We3, ZP, 256, 256, 128));
  TRY(make_tmap_bf16_2d(&v->tmWd1, v->Wd1, 256, KH, KH, v->bn_hd));
  TRY(make_tmap_bf16_2d(&v->tmWd1q, v->Wd1, 256, KH, KH, 64));
  TRY(make_tmap_bf16_2d(&v->tmWd2, v->Wd2, (uint64_t)v->hw4 * v->dbn[0], 256, 256, 256));
#undef TRY
  if (rc != DRM_OK) { drm_vae_destroy(v); return rc; }
  *out = v;
  return DRM_OK;
}

extern "C" int drm_vae_destroy(drm_vae* v) {
  if (!v) return DRM_OK;
  for (void* p : v->allocs) cudaFree(p);
  delete v;
  return DRM_OK;
}

extern "C" int drm_vae_pack(drm_vae* v, const drm_vae_weights* w, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(v && w, DRM_ERR_ARG, "drm_vae_pack: NULL argument");
  const float* src[V_COUNT] = {};
  for (int i = 0; i < 4; ++i) { src[V_ECW + i] = w->enc_conv_w[i]; src[V_ECB + i] = w->enc_conv_b[i]; src[V_DCW + i] = w->dec_conv_w[i]; src[V_DCB + i] = w->dec_conv_b[i]; }
  src[V_EL1W] = w->enc_l1_w; src[V_EL1B] = w->enc_l1_b; src[V_ELNG] = w->enc_ln_g; src[V_ELNB] = w->enc_ln_b; src[V_EL2W] = w->enc_l2_w; src[V_EL2B] = w->enc_l2_b;
  src[V_DL1W] = w->dec_l1_w; src[V_DL1B] = w->dec_l1_b; src[V_DLNG] = w->dec_ln_g; src[V_DLNB] = w->dec_ln_b; src[V_DL2W] = w->dec_l2_w; src[V_DL2B] = w->dec_l2_b;
  for (int i = 0; i < V_COUNT; ++i) DRM_REQUIRE(src[i], DRM_ERR_ARG, "drm_vae_pack: every encoder / decoder weight is required");
  cudaStream_t st = (cudaStream_t)stream;
  for (const OffOp& op : v->mat_ops) {
    pack_offsets_kernel<<<grid_for((long)op.nrows * op.ncols), 256, 0, st>>>(op.dst, op.ld_dst, op.row0, op.nrows, op.col0, op.ncols, src[op.src], op.roff, op.coff);
    DRM_LAUNCH_CHECK();
  }
  for (const VecOp& op : v->vec_ops) {
    pack_vector_kernel<<<ceil_div(op.n, 256), 256, 0, st>>>(op.dst, op.n, src[op.src], op.map, op.fill);
    DRM_LAUNCH_CHECK();
  }
  v->packed = true;
  return DRM_OK;
}

extern "C" int drm_observe_create(drm_rssm* m, drm_vae* v, int32_t B, int32_t T, drm_observe** out) {
  RC(check_arch());
  DRM_REQUIRE(m && v && out, DRM_ERR_ARG, "drm_observe_create: NULL argument");
  DRM_REQUIRE(v->m == m, DRM_ERR_ARG, "drm_observe_create: the vae handle belongs to another rssm handle");
  DRM_REQUIRE(B >= 1 && T >= 1, DRM_ERR_SHAPE, "drm_observe_create: B and T must be >= 1");
  drm_observe* o = new drm_observe();
  o->m = m; o->v = v; o->B = B; o->T = T; o->scanned = false;
  o->rows = (T + 1) * B;
  o->rows_p = round_up(o->rows, BM) + BM;
  const int NF = T * B;
  {
    const int fc = opts().conv_chunk;   // frames per conv chunk: 512 measured best (4.37 -> 3.94 ms per 1024-frame forward vs 128); scratch ~0.5 GB
    o->FC = NF < fc ? NF : fc;
  }
  const int H = v->d.H, W = v->d.W;
  long pe = 0, ae = 0;
  {
    int hs = H, ws = W;
    for (int i = 0; i < 4; ++i) {
      hs >>= 1; ws >>= 1;
      pe = std::max(pe, (long)o->FC * hs * ws * v->EK[i]);
      ae = std::max(ae, (long)o->FC * hs * ws * v->ebn[i + 1]);
    }
    hs = H / 16; ws = W / 16;
    ae = std::max(ae, (long)o->FC * hs * ws * v->dbn[0]);
    for (int j = 0; j < 4; ++j) {
      pe = std::max(pe, 4l * o->FC * hs * ws * v->DK[j]);
      hs <<= 1; ws <<= 1;
      ae = std::max(ae, (long)o->FC * hs * ws * v->dbn[j + 1]);
    }
  }
  o->patch_elems = pe + 128l * 4096;   // slack: TMA boxes of the last tile may read past the valid rows
  o->act_elems = ae + 130l * 64 * 256;   // slack: an im2col tile may walk up to 128 filter positions past the chunk's last frame
  int rc = DRM_OK;
#define TRY(x) if (rc == DRM_OK) rc = (x)
  TRY(dev_alloc(o->allocs, &o->S, (size_t)o->rows_p * m->KS));
  TRY(dev_alloc(o->allocs, &o->Y1, (size_t)(MAX_HEADS + 1) * o->rows_p * 256));
  TRY(dev_alloc(o->allocs, &o->Y2, (size_t)(MAX_HEADS + 1) * o->rows_p * 256));
  TRY(dev_alloc(o->allocs, &o->feat, (size_t)(round_up(NF, BM) + BM) * v->Kf));
  TRY(dev_alloc(o->allocs, &o->featpart, (size_t)NF * v->bn_he));
  TRY(dev_alloc(o->allocs, &o->dec0, (size_t)(NF + BM) * v->hw4 * v->dbn[0]));
  TRY(dev_alloc(o->allocs, &o->zero_h, (size_t)B * m->d.D));
  TRY(dev_alloc(o->allocs, &o->patch, (size_t)o->patch_elems));
  TRY(dev_alloc(o->allocs, &o->actA, (size_t)o->act_elems));
  TRY(dev_alloc(o->allocs, &o->actB, (size_t)o->act_elems));
  TRY(make_tmap_bf16_2d(&o->tmS, o->S, o->rows_p, m->KS, m->KS, BM));
  TRY(make_tmap_bf16_2d(&o->tmY1, o->Y1, (uint64_t)(MAX_HEADS + 1) * o->rows_p, 256, 256, BM));
  TRY(make_tmap_bf16_2d(&o->tmY2, o->Y2, (uint64_t)(MAX_HEADS + 1) * o->rows_p, 256, 256, BM));
  TRY(make_tmap_bf16_2d(&o->tmS_s, o->S, o->rows_p, m->KS, m->KS, SMALL_A_ROWS));
  TRY(make_tmap_bf16_2d(&o->tmY1_s, o->Y1, (uint64_t)(MAX_HEADS + 1) * o->rows_p, 256, 256, SMALL_A_ROWS));
  TRY(make_tmap_bf16_2d(&o->tmY2_s, o->Y2, (uint64_t)(MAX_HEADS + 1) * o->rows_p, 256, 256, SMALL_A_ROWS));
  TRY(make_tmap_bf16_2d(&o->tmFeat, o->feat, round_up(NF, BM) + BM, v->Kf, v->Kf, BM));
  for (int i = 0; i < 4; ++i) {
    TRY(make_tmap_bf16_2d(&o->tmPatchE[i], o->patch, (uint64_t)(o->patch_elems / v->EK[i]), v->EK[i], v->EK[i], BM));
    TRY(make_tmap_bf16_2d(&o->tmPatchD[i], o->patch, (uint64_t)(o->patch_elems / v->DK[i]), v->DK[i], v->DK[i], BM));
  }
  {
    // im2col maps: frames mapped = chunk + the frames a last partial tile can reach (never stored, only read)
    int hs = H / 2, ws = W / 2;      // input grid of encoder conv 1
    for (int i = 1; i < 4; ++i) {
      const int extra = ceil_div(BM, (hs / 2) * (ws / 2)) + 1;
      if (v->echunk[i] && (long)(o->FC + extra) * hs * ws * v->ebn[i] <= o->act_elems) {
        TRY(make_tmap_im2col_bf16(&o->tmI2cE[i], (i & 1) ? o->actA : o->actB, v->ebn[i], ws, hs, o->FC + extra, -1, -1, -2, -2, v->echunk[i], BM, 2));
        o->i2cE[i] = rc == DRM_OK;
      }
      hs >>= 1; ws >>= 1;
    }
    hs = H / 16; ws = W / 16;        // input grid of decoder convT 0
    for (int j = 0; j < 3; ++j) {
      const int extra = ceil_div(BM, hs * ws) + 1;
      const __nv_bfloat16* base = j == 0 ? o->dec0 : (j == 1 ? o->actA : o->actB);
      const int nfr = j == 0 ? NF + extra : o->FC + extra;
      const bool fits = j == 0 ? extra <= BM : (long)nfr * hs * ws * v->dbn[j] <= o->act_elems;
      for (int ph = 0; ph < 4 && v->dchunk[j] && fits; ++ph) {
        const int lw = (ph & 1) ? 0 : -1, lh = (ph >> 1) ? 0 : -1;
        TRY(make_tmap_im2col_bf16(&o->tmI2cD[j][ph], base, v->dbn[j], ws, hs, nfr, lw, lh, lw, lh, v->dchunk[j], BM, 1));
        o->i2cD[j] = rc == DRM_OK;
      }
      hs <<= 1; ws <<= 1;
    }
  }
#undef TRY
  if (rc != DRM_OK) { drm_observe_destroy(o); return rc; }
  *out = o;
  return DRM_OK;
}

extern "C" int drm_observe_destroy(drm_observe* o) {
  if (!o) return DRM_OK;
  drm::obs_persist_free(o->ps);
  for (void* p : o->allocs) cudaFree(p);
  delete o;
  return DRM_OK;
}

namespace drm {

// patch rows of encoder conv i + 1 from the NHWC output `cur` (grid hs x ws) of conv i -- only for layers that stay on the patch path
static int encoder_patch_gather(drm_observe* o, const __nv_bfloat16* cur, int nf, int hs, int ws, int i, cudaStream_t st) {
  drm_vae* v = o->v;
  Taps4 tp;
  memset(&tp, 0, sizeof(tp));
  tp.t[0].n = 16;
  for (int t = 0; t < 16; ++t) { tp.t[0].dy[t] = (t >> 2) - 1; tp.t[0].dx[t] = (t & 3) - 1; }
  const long rows = (long)nf * (hs / 2) * (ws / 2);
  patch_gather_kernel<<<dim3(grid_for(rows * 16 * (v->ebn[i + 1] / 8)), 1), 256, 0, st>>>(cur, o->patch, rows, hs, ws, v->ebn[i + 1], hs / 2, ws / 2, 2, tp, 0, v->EK[i + 1]);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}

// Encoder convs over frames [f0, f0 + nf) -> o->feat rows [f0, f0 + nf).  tm = 1: frame n of the (time-major)
// feature matrix is obs[b = n % B, t = n / B]; tm = 0: frames in the given order.
static int encoder_conv_chunk(drm_observe* o, const float* obs, int f0, int nf, int tm, cudaStream_t st) {
  drm_vae* v = o->v;
  const int H = v->d.H, W = v->d.W;
  im2col_first_kernel<<<grid_for((long)nf * (H / 2) * (W / 2) * 8), 256, 0, st>>>(obs, o->patch, nf, f0, H, W, tm ? o->B : 0, tm ? o->T : 0);
  DRM_LAUNCH_CHECK();
  int hs = H / 2, ws = W / 2;
  __nv_bfloat16* cur = o->actA;
  __nv_bfloat16* nxt = o->actB;
  for (int i = 0; i < 4; ++i) {
    const int M = nf * hs * ws;
    const bool implicit_next = i < 3 && o->i2cE[i + 1] && opts().conv_implicit;
    if (i > 0 && o->i2cE[i] && opts().conv_implicit) {
      // implicit GEMM: the A operand comes straight from the previous layer's NHWC output (cur was swapped: the input is `nxt`)
      __nv_bfloat16* dst = i == 3 ? o->feat + (long)f0 * v->Kf : cur;
      RC(launch_conv_implicit(&o->tmI2cE[i], 1, v->tmWeI[i], M, ws, hs, 2, 0, 16, v->ebn[i], v->echunk[i], v->ebn[i + 1], v->be[i], dst,
                              (long)v->ebn[i + 1], v->ebn[i + 1], 1, RowMap{0, 0, 0, 0}, false, st));
      if (i == 3) break;
      if (!implicit_next) RC(encoder_patch_gather(o, cur, nf, hs, ws, i, st));
      hs >>= 1; ws >>= 1;
      std::swap(cur, nxt);
      continue;
    }
    GemmCommon g = common(o->tmPatchE[i], v->tmWe[i], M, v->ebn[i + 1]);
    g.ka0 = 0; g.nka0 = v->EK[i] / 64;
    g.n_slots = 1; g.y_slot[0] = 0;
    __nv_bfloat16* dst = i == 3 ? o->feat + (long)f0 * v->Kf : cur;
    EpiPlain::Params p{v->be[i], nullptr, dst, 0, (long)v->ebn[i + 1], v->ebn[i + 1], 1, 0, RowMap{0, 0, 0, 0}};
    if (use_conv_persist(g, p)) RC(launch_conv_persist(g, p, 1, st));
    else if (v->ebn[i + 1] <= 64) RC(launch_gemm<EpiPlainS>(g, p, dim3(ceil_div(M, BM), 1), st));
    else RC(launch_gemm<EpiPlain>(g, p, dim3(ceil_div(M, BM), 1), st));
    if (i == 3) break;
    if (!implicit_next) RC(encoder_patch_gather(o, cur, nf, hs, ws, i, st));
    hs >>= 1; ws >>= 1;
    std::swap(cur, nxt);
  }
  return DRM_OK;
}

// Decoder tail for `nf` frames whose upscaler output (NHWC [nf, hw4 * dbn0], bf16) is in `act0`:
// four transposed convs -> mu fp32 NCHW at mu_out (frame-major).
static int decoder_conv_chunk(drm_observe* o, const __nv_bfloat16* act0, int nf, float* mu_out, cudaStream_t st) {
  drm_vae* v = o->v;
  int hs = v->d.H / 16, ws = v->d.W / 16;
  __nv_bfloat16* bufs[2] = {o->actA, o->actB};
  const __nv_bfloat16* src = act0;
  for (int j = 0; j < 4; ++j) {
    const long rows = (long)nf * hs * ws;
    Taps4 tp;
    memset(&tp, 0, sizeof(tp));
    for (int ph = 0; ph < 4; ++ph) {
      tp.t[ph].n = 4;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) { tp.t[ph].dy[ty * 2 + tx] = ct_d(ph >> 1, ty); tp.t[ph].dx[ty * 2 + tx] = ct_d(ph & 1, tx); }
    }
    if (j == 3 && (v->dbn[3] == 32 || v->dbn[3] == 16 || v->dbn[3] == 8 || v->dbn[3] == 64)) {
      // image layer: direct kernel, no patch matrix (see convt_last_direct_kernel)
      const int grid = grid_for(rows);
#define DRM_CT_LAST(CPV) convt_last_direct_kernel<CPV><<<grid, 256, 0, st>>>(src, v->Wdc[3], v->bdc[3], mu_out, rows, hs, ws, v->dbn[4], 3, v->DK[3])
      if (v->dbn[3] == 64) DRM_CT_LAST(64); else if (v->dbn[3] == 32) DRM_CT_LAST(32); else if (v->dbn[3] == 16) DRM_CT_LAST(16); else DRM_CT_LAST(8);
#undef DRM_CT_LAST
      DRM_LAUNCH_CHECK();
      break;
    }
    __nv_bfloat16* dst = bufs[j & 1];
    if (j < 3 && o->i2cD[j] && opts().conv_implicit) {
      // implicit GEMM: four sub-pixel phases read the NHWC input through their im2col maps, no patch matrix
      const int n_base = j == 0 ? (int)((act0 - o->dec0) / ((long)v->hw4 * v->dbn[0])) : 0;
      RC(launch_conv_implicit(o->tmI2cD[j], 4, v->tmWdcI[j], (int)rows, ws, hs, 1, n_base, 4, v->dbn[j], v->dchunk[j], v->dbn[j + 1], v->bdc[j],
                              dst, (long)v->dbn[j + 1], v->dbn[j + 1], 1, RowMap{2, hs, ws, 0}, true, st));
      hs <<= 1; ws <<= 1;
      src = dst;
      continue;
    }
    patch_gather_kernel<<<dim3(grid_for(rows * 4 * (v->dbn[j] / 8)), 4), 256, 0, st>>>(src, o->patch, rows, hs, ws, v->dbn[j], hs, ws, 1, tp, rows, v->DK[j]);
    DRM_LAUNCH_CHECK();
    GemmCommon g = common(o->tmPatchD[j], v->tmWdc[j], (int)rows, v->dbn[j + 1]);
    g.a_y_stride = (int)rows;
    g.ka0 = 0; g.nka0 = v->DK[j] / 64;
    EpiPlain::Params p{v->bdc[j], nullptr, dst, 0, (long)v->dbn[j + 1], v->dbn[j + 1], 1, 1, RowMap{2, hs, ws, 0}};
    if (j == 3) { p.out_f32 = mu_out; p.out_bf16 = nullptr; p.ld_f32 = 0; p.N = 3; p.act = 2; p.rm = RowMap{3, hs, ws, 0}; }
    if (use_conv_persist(g, p)) RC(launch_conv_persist(g, p, 4, st));
    else if (v->dbn[j + 1] <= 64) RC(launch_gemm<EpiPlainS>(g, p, dim3(ceil_div((int)rows, BM), 4), st));
    else RC(launch_gemm<EpiPlain>(g, p, dim3(ceil_div((int)rows, BM), 4), st));
    hs <<= 1; ws <<= 1;
    src = dst;
  }
  return DRM_OK;
}

// features x W_feat^T for feature rows [0, nf) -> featpart (fp32, no bias)
static int encoder_feat_part(drm_observe* o, int nf, cudaStream_t st) {
  drm_vae* v = o->v;
  GemmCommon g = common(o->tmFeat, v->tmWef, nf, v->bn_he);
  g.ka0 = 0; g.nka0 = v->Kf / 64;
  g.n_slots = 1; g.y_slot[0] = 0;
  EpiPlain::Params p{nullptr, o->featpart, nullptr, (long)v->bn_he, 0, v->bn_he, 0, 0, RowMap{0, 0, 0, 0}};
  return launch_gemm<EpiPlain>(g, p, dim3(ceil_div(nf, BM), 1), st);
}

// posterior head on view `vw` (M rows): LN(h part + featpart rows) -> logits -> sample
static int encoder_head(drm_observe* o, const WsView& vw, const float* addend, const float* uniforms, float* latent, long ld_latent,
                        float* logits, long ld_logits, uint8_t* idx, long ld_idx, bool write_sz, int M, cudaStream_t st) {
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  const int mt = ceil_div(M, BM);
  {
    GemmCommon g = common(*vw.tmS, v->tmWeh, M, v->bn_he);
    small_a(g, vw.tmS_s);
    g.a_row0 = vw.row0;
    g.ka0 = m->ZP / 64 + 1; g.nka0 = m->DP / 64;
    g.n_slots = 1; g.y_slot[0] = 0;
    EpiLnSiluAdd::Params p{v->e1_b, v->e1_g, v->e1_be, addend, (long)v->bn_he, vw.Y1, 256, vw.row0, vw.slot_rows, v->d.h_enc, 1e-5f, v->bn_he};
    RC(launch_ln<true>(g, v->tmWeh, v->tmWehq, v->bn_he, p, mt, 1, st, DRM_STAGE_OTHER));
  }
  {
    const int bn = (mt * (m->ZP / 256) <= 74) ? 128 : 256;
    GemmCommon g = common(*vw.tmY1, bn == 128 ? v->tmWe3h : v->tmWe3, M, bn);
    small_a(g, vw.tmY1_s);
    g.a_row0 = vw.row0;
    g.ka0 = 0; g.nka0 = ceil_div(v->d.h_enc, 64);
    EpiCat::Params p{v->e3_b, uniforms, latent, logits, idx, write_sz ? vw.S + (long)vw.row0 * m->KS : nullptr, nullptr,
                     ld_latent, ld_logits, ld_idx, 0, m->KS, m->d.R, RowMap{0, 0, 0, 0}};
    RC(launch_gemm<EpiCat>(g, p, dim3(mt, m->ZP / bn), st));
  }
  return DRM_OK;
}

// decoder dense part on view `vw` (M rows of [z | h]) -> NHWC bf16 [M, hw4 * dbn0] at act0, rows mapped by rm
static int decoder_dense(drm_observe* o, const WsView& vw, __nv_bfloat16* act0, RowMap rm, int M, cudaStream_t st, bool z_idx = false) {
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  const int mt = ceil_div(M, BM);
  {
    GemmCommon g = common(*vw.tmS, v->tmWd1, M, v->bn_hd);
    g.a_row0 = vw.row0;
    g.ka0 = 0; g.nka0 = m->ZP / 64;
    g.ka1 = m->ZP / 64 + 1; g.nka1 = m->DP / 64;
    g.n_slots = 1; g.y_slot[0] = 0;
    small_a(g, vw.tmS_s);
    EpiLnSilu::Params p{v->d1_b, v->d1_g, v->d1_be, nullptr, 0, vw.Y1, 256, vw.row0, vw.slot_rows, v->d.h_dec, 1e-5f, v->bn_hd};
    RC(launch_ln<false>(g, v->tmWd1, v->tmWd1q, v->bn_hd, p, mt, 1, st, DRM_STAGE_OTHER));
  }
  {
    const int N = v->hw4 * v->dbn[0];
    GemmCommon g = common(*vw.tmY1, v->tmWd2, M, 256);
    small_a(g, vw.tmY1_s);
    g.a_row0 = vw.row0;
    g.ka0 = 0; g.nka0 = ceil_div(v->d.h_dec, 64);
    EpiPlain::Params p{v->d2_b, nullptr, act0, 0, (long)N, N, 1, 0, rm};
    RC(launch_gemm<EpiPlain>(g, p, dim3(mt, N / 256), st));
  }
  return DRM_OK;
}

}  // namespace drm

#include "observe_persist.cuh"

extern "C" int drm_observe_scan(drm_observe* o, const float* obs, const float* act, const float* uniforms, int32_t mode,
                                float* latent, float* hidden, float* post_logits, uint8_t* idx, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(o && obs && act && uniforms && latent && hidden, DRM_ERR_ARG, "drm_observe_scan: NULL argument");
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  DRM_REQUIRE(m->packed && (m->have & HAVE_GRU) && v->packed, DRM_ERR_ARG, "drm_observe_scan: GRU / VAE weights were never packed");
  DRM_REQUIRE(mode == 0 || mode == 1, DRM_ERR_ARG, "drm_observe_scan: mode must be 0 (unroll) or 1 (warm start)");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = o->B, T = o->T, D = m->d.D, ZP = m->ZP, R = m->d.R, A = m->d.A, KS = m->KS;
  const int NF = T * B;
  // (1) everything that does not depend on h, hoisted out of the recurrence: conv features of all frames
  //     (time-major), their product with the feature columns of latent_mapper.0, and the actions
  for (int f0 = 0; f0 < NF; f0 += o->FC) RC(encoder_conv_chunk(o, obs, f0, std::min(o->FC, NF - f0), 1, st));
  RC(encoder_feat_part(o, NF, st));
  // (2) the recurrence: one persistent kernel for all T steps when its clusters fit the device (observe_persist.cuh), else
  //     three launches per step
  const bool persist = opts().persist && obs_persist_eligible(o);
  // slab 0 (zero state) and slab 1's h; the persistent kernel flips single one-hot entries, so it starts from all-zero z columns
  DRM_CUDA(cudaMemsetAsync(o->S, 0, (size_t)(persist ? (T + 1) : 2) * B * KS * sizeof(__nv_bfloat16), st));
  pack_actions_kernel<<<grid_for((long)B * T * A), 256, 0, st>>>(o->S, KS, ZP, act, B, T, A);
  DRM_LAUNCH_CHECK();
  if (persist) {
    RC(observe_persist(o, uniforms, mode, latent, hidden, post_logits, idx, st));
    o->scanned = true;
    return DRM_OK;
  }
  const long ldL = (long)T * ZP, ldH = (long)T * D;
  for (int t = 0; t < T; ++t) {
    const WsView prev = view_of(o, t * B), cur = view_of(o, (t + 1) * B);
    if (t == 0 && mode == 1) {
      DRM_CUDA(cudaMemset2DAsync(hidden, ldH * sizeof(float), 0, (size_t)D * sizeof(float), B, st));   // h_0 = 0, no GRU step
    } else {
      RC(stage_gru(m, prev, cur, t == 0 ? o->zero_h : hidden + (long)(t - 1) * D, t == 0 ? (long)D : ldH, hidden + (long)t * D, ldH, B, st, true));
    }
    RC(encoder_head(o, cur, o->featpart + (long)t * B * v->bn_he, uniforms + (long)t * B * R, latent + (long)t * ZP, ldL,
                    post_logits ? post_logits + (long)t * ZP : nullptr, ldL, idx ? idx + (long)t * R : nullptr, (long)T * R, true, B, st));
  }
  o->scanned = true;
  return DRM_OK;
}

extern "C" int drm_observe_heads(drm_observe* o, float* prior_logits, float* dec_mu, float* reward_logits, float* cont_logit,
                                 void* stream) {
  RC(check_arch());
  DRM_REQUIRE(o, DRM_ERR_ARG, "drm_observe_heads: NULL argument");
  DRM_REQUIRE(o->scanned, DRM_ERR_ARG, "drm_observe_heads: call drm_observe_scan first");
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  cudaStream_t st = (cudaStream_t)stream;
  const int B = o->B, T = o->T, NF = T * B, ZP = m->ZP;
  const WsView all = view_of(o, B);   // slabs 1..T = steps 0..T-1, time-major rows
  if (prior_logits) {
    DRM_REQUIRE(m->have & HAVE_PRIOR, DRM_ERR_ARG, "drm_observe_heads: prior weights were never packed");
    RC(stage_prior(m, all, nullptr, nullptr, 0, prior_logits, ZP, nullptr, 0, false, RowMap{1, B, T, 0}, NF, st));
  }
  if (dec_mu) {
    // frames leave the dense part in batch-major order, so the conv tail writes [B, T, 3, H, W] directly
    const long fsz = 3l * v->d.H * v->d.W;
    for (int f0 = 0; f0 < NF; f0 += o->FC) {
      const int nf = std::min(o->FC, NF - f0);
      // the chunk is a contiguous range of BATCH-major frames n' = b * T + t; its rows in S are scattered, so the dense
      // part runs once over all rows into the batch-major dec0 ...
      if (f0 == 0) RC(decoder_dense(o, all, o->dec0, RowMap{1, B, T, 0}, NF, st, true));
      // ... and the conv tail consumes it chunk by chunk
      RC(decoder_conv_chunk(o, o->dec0 + (long)f0 * v->hw4 * v->dbn[0], nf, dec_mu + (long)f0 * fsz, st));
    }
  }
  if ((reward_logits || cont_logit) && T > 1) {
    const unsigned need = (reward_logits ? 1u << HS_REWARD : 0u) | (cont_logit ? 1u << HS_CONT : 0u);
    DRM_REQUIRE((m->have & need) == need, DRM_ERR_ARG, "drm_observe_heads: reward / continue weights were never packed");
    int slots[2], n = 0;
    EpiHeads::Params hp;
    memset(&hp, 0, sizeof(hp));
    hp.rm = RowMap{1, B, T - 1, 0};
    if (reward_logits) { slots[n++] = HS_REWARD; hp.logits[HS_REWARD] = reward_logits; hp.ld_logits[HS_REWARD] = m->d.NB; }
    if (cont_logit) { slots[n++] = HS_CONT; hp.logits[HS_CONT] = cont_logit; hp.ld_value[HS_CONT] = 1; }
    RC(stage_heads(m, view_of(o, 2 * B), slots, n, hp, (T - 1) * B, st));
  }
  return DRM_OK;
}

extern "C" int drm_encoder_fwd(drm_observe* o, const float* h, const float* obs, const float* uniforms, float* logits, float* z_st,
                               uint8_t* idx, int32_t N, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(o && h && obs, DRM_ERR_ARG, "drm_encoder_fwd: NULL argument");
  DRM_REQUIRE(N >= 0 && N <= o->B * o->T, DRM_ERR_SHAPE, "drm_encoder_fwd: N exceeds the workspace (B * T rows)");
  DRM_REQUIRE(o->v->packed, DRM_ERR_ARG, "drm_encoder_fwd: VAE weights were never packed");
  if (N == 0) return DRM_OK;
  drm_rssm* m = o->m;
  cudaStream_t st = (cudaStream_t)stream;
  for (int f0 = 0; f0 < N; f0 += o->FC) RC(encoder_conv_chunk(o, obs, f0, std::min(o->FC, N - f0), 0, st));
  RC(encoder_feat_part(o, N, st));
  RC(pack_cols(o->S, m->KS, m->ZP + 64, h, m->d.D, m->d.D, N, nullptr, 0, st));
  o->scanned = false;
  return encoder_head(o, view_of(o, 0), o->featpart, uniforms, z_st, m->ZP, logits, m->ZP, idx, m->d.R, false, N, st);
}

extern "C" int drm_decoder_fwd(drm_observe* o, const float* h, const float* z, float* mu, int32_t N, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(o && h && z && mu, DRM_ERR_ARG, "drm_decoder_fwd: NULL argument");
  DRM_REQUIRE(N >= 0 && N <= o->B * o->T, DRM_ERR_SHAPE, "drm_decoder_fwd: N exceeds the workspace (B * T rows)");
  DRM_REQUIRE(o->v->packed, DRM_ERR_ARG, "drm_decoder_fwd: VAE weights were never packed");
  if (N == 0) return DRM_OK;
  drm_rssm* m = o->m;
  drm_vae* v = o->v;
  cudaStream_t st = (cudaStream_t)stream;
  RC(pack_cols(o->S, m->KS, 0, z, m->ZP, m->ZP, N, nullptr, 0, st));
  RC(pack_cols(o->S, m->KS, m->ZP + 64, h, m->d.D, m->d.D, N, nullptr, 0, st));
  o->scanned = false;
  RC(decoder_dense(o, view_of(o, 0), o->dec0, RowMap{0, 0, 0, 0}, N, st));
  const long fsz = 3l * v->d.H * v->d.W;
  for (int f0 = 0; f0 < N; f0 += o->FC)
    RC(decoder_conv_chunk(o, o->dec0 + (long)f0 * v->hw4 * v->dbn[0], std::min(o->FC, N - f0), mu + (long)f0 * fsz, st));
  return DRM_OK;
}

extern "C" int drm_neg_sse_rows(const float* a, const float* b, float* out, int64_t rows, int32_t len, void* stream) {
  RC(check_arch());
  DRM_REQUIRE(rows >= 0 && len >= 1, DRM_ERR_SHAPE, "drm_neg_sse_rows: bad shape");
  if (rows == 0) return DRM_OK;
  DRM_REQUIRE(a && b && out, DRM_ERR_ARG, "drm_neg_sse_rows: NULL argument");
  DRM_REQUIRE(len % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0), DRM_ERR_ALIGN, "drm_neg_sse_rows: rows must be 16-byte aligned float4 multiples");
  neg_sse_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(a, b, out, len);
  DRM_LAUNCH_CHECK();
  return DRM_OK;
}
