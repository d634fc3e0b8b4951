// Epilogue functors for fused_gemm_kernel.  Each thread owns one accumulator row (one start
// state / one frame), so every row-wise operation below is register-only.
#pragma once

#include "gemm.cuh"

namespace drm {

// ------------------------------------------------------------------------------------------
// plain: out = act(acc + bias); used by the test hook, the decoder/encoder dense layers
// ------------------------------------------------------------------------------------------
struct EpiPlain {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0;
  struct Params {
    const float* bias;       // [N] or NULL
    float* out_f32;          // [M, ld_f32] or NULL
    __nv_bfloat16* out_bf16; // [M, ld_bf16] or NULL
    long ld_f32, ld_bf16;
    int N;                   // valid output columns
    int act;                 // 0 none, 1 SiLU
  };
  static __device__ __forceinline__ void run(const Params& p, const GemmCommon& g, uint32_t taddr, int m, int slot) {
    const int n0 = slot * g.bn;
    for (int c = 0; c < g.bn; c += 32) {
      if (n0 + c >= p.N) break;
      float v[32];
      tmem_ld32(taddr + c, v);
      const int nvalid = min(32, p.N - (n0 + c));
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] + ((p.bias && j < nvalid) ? __ldg(p.bias + n0 + c + j) : 0.f);
        v[j] = p.act == 1 ? siluf_(x) : x;
      }
      if (m < g.M) {
        if (p.out_f32) store_f32_row<32>(p.out_f32 + (long)m * p.ld_f32 + n0 + c, v, nvalid);
        if (p.out_bf16) store_bf16_row<32>(p.out_bf16 + (long)m * p.ld_bf16 + n0 + c, v, nvalid);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// Linear + LayerNorm(eps) + SiLU -> bf16 (the hidden layers of every MLP head:
// DynamicsPredictors.py:15-23, 52-60, 85-93; Agent.py:178-185, 219-227;
// VariationalAutoEncoder.py:50-53, 119-122).  One tile holds the whole feature row (<= 256).
// ------------------------------------------------------------------------------------------
struct EpiLnSilu {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0;
  struct Params {
    const float* bias;   // [slots * bn]
    const float* gamma;  // [slots * bn]
    const float* beta;   // [slots * bn]
    const float* addend; // optional fp32 [M, ld_addend] added before LN (pre-computed partial product)
    long ld_addend;
    __nv_bfloat16* out;  // rows (out_row0 + slot * out_y_stride + m), ld_out columns
    int ld_out, out_row0, out_y_stride;
    int n_valid;         // features (<= bn)
    float eps;
  };
  static __device__ __forceinline__ void run(const Params& p, const GemmCommon& g, uint32_t taddr, int m, int slot) {
    const float* bias = p.bias + slot * g.bn;
    const float* gamma = p.gamma + slot * g.bn;
    const float* beta = p.beta + slot * g.bn;
    const float* add = (p.addend && m < g.M) ? p.addend + (long)m * p.ld_addend : nullptr;
    const int nch = (p.n_valid + 15) >> 4;
    float sum = 0.f;
    for (int c = 0; c < nch; ++c) {
      float v[16];
      tmem_ld16(taddr + c * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c * 16 + j;
        if (n < p.n_valid) sum += v[j] + __ldg(bias + n) + (add ? __ldg(add + n) : 0.f);
      }
    }
    const float mean = sum / (float)p.n_valid;
    float ss = 0.f;
    for (int c = 0; c < nch; ++c) {
      float v[16];
      tmem_ld16(taddr + c * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c * 16 + j;
        if (n < p.n_valid) {
          const float d = v[j] + __ldg(bias + n) + (add ? __ldg(add + n) : 0.f) - mean;
          ss += d * d;
        }
      }
    }
    const float rstd = rsqrtf(ss / (float)p.n_valid + p.eps);
    __nv_bfloat16* out = p.out + (long)(p.out_row0 + slot * p.out_y_stride + m) * p.ld_out;
    for (int c = 0; c < nch; ++c) {
      float v[16];
      tmem_ld16(taddr + c * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c * 16 + j;
        float y = 0.f;
        if (n < p.n_valid) {
          const float x = v[j] + __ldg(bias + n) + (add ? __ldg(add + n) : 0.f);
          y = siluf_((x - mean) * rstd * __ldg(gamma + n) + __ldg(beta + n));
        }
        v[j] = y;
      }
      if (m < g.M) store_bf16_row<16>(out + c * 16, v, 16);  // pad columns inside the chunk are written as 0
    }
  }
};

// ------------------------------------------------------------------------------------------
// GRU gates + state update (nn.GRUCell, SequenceModel.py:13-24), U hidden units per tile.
// TMEM columns: [r | z | n_x | n_h], each U wide.
// ------------------------------------------------------------------------------------------
template <int U>
struct EpiGru {
  static constexpr int B_ROWS_MAX = 3 * U, STAGES = (U == 32 ? 6 : 4), TMEM_COLS = 4 * U, GRU_U = U;
  struct Params {
    const float* b_ih;   // [3D] reference layout [r; z; n]
    const float* b_hh;   // [3D]
    const float* h_prev; // fp32 [M, ld_hprev]
    float* h_out;        // fp32 [M, ld_hout]
    __nv_bfloat16* s_h;  // bf16 h columns of the next state buffer, [M, ld_s]
    long ld_hprev, ld_hout;
    int ld_s, D;
  };
  static __device__ __forceinline__ void run(const Params& p, const GemmCommon& g, uint32_t taddr, int m, int slot) {
    const int u0 = slot * U;
    const int D = p.D;
    for (int c = 0; c < U; c += 16) {
      if (u0 + c >= D) break;
      float r[16], z[16], nx[16], nh[16];
      tmem_ld16(taddr + c, r);
      tmem_ld16(taddr + U + c, z);
      tmem_ld16(taddr + 2 * U + c, nx);
      tmem_ld16(taddr + 3 * U + c, nh);
      const int nvalid = min(16, D - (u0 + c));
      if (m < g.M) {
        const float* hp = p.h_prev + (long)m * p.ld_hprev + u0 + c;
        float hn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hn[j] = 0.f;
          if (j < nvalid) {
            const int u = u0 + c + j;
            const float rr = sigmoidf_(r[j] + __ldg(p.b_ih + u) + __ldg(p.b_hh + u));
            const float zz = sigmoidf_(z[j] + __ldg(p.b_ih + D + u) + __ldg(p.b_hh + D + u));
            const float nn = tanhf(nx[j] + __ldg(p.b_ih + 2 * D + u) + rr * (nh[j] + __ldg(p.b_hh + 2 * D + u)));
            hn[j] = (1.0f - zz) * nn + zz * __ldg(hp + j);
          }
        }
        store_f32_row<16>(p.h_out + (long)m * p.ld_hout + u0 + c, hn, nvalid);
        store_bf16_row<16>(p.s_h + (long)m * p.ld_s + u0 + c, hn, nvalid);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// prior / posterior logits -> 32-class categorical (softmax, 1% unimix, inverse-CDF sample from a
// host-supplied uniform, one-hot, straight-through).  DynamicsPredictors.py:31-40;
// VariationalAutoEncoder.py:85-99.  A 256-column tile = 8 latent rows x 32 classes.
// ------------------------------------------------------------------------------------------
struct EpiCat {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0;
  struct Params {
    const float* bias;     // [R * 32]
    const float* uniforms; // [M, R] for this step, or NULL (logits only)
    float* latent;         // straight-through value, fp32 [M, ld_latent] (R * 32 columns) or NULL
    float* logits;         // fp32 [M, ld_logits] or NULL
    uint8_t* idx;          // [M, ld_idx] or NULL
    __nv_bfloat16* s_z;    // bf16 one-hot into the z columns of a state buffer [M, ld_s] or NULL
    long ld_latent, ld_logits, ld_idx;
    int ld_s, R;
  };
  static __device__ __forceinline__ void run(const Params& p, const GemmCommon& g, uint32_t taddr, int m, int slot) {
    const bool live = m < g.M;
    for (int gi = 0; gi < 8; ++gi) {
      const int lrow = slot * 8 + gi;
      if (lrow >= p.R) break;
      float v[32];
      tmem_ld32(taddr + gi * 32, v);
      const float* b = p.bias + lrow * 32;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] += __ldg(b + j);
        mx = fmaxf(mx, v[j]);
      }
      if (live && p.logits) store_f32_row<32>(p.logits + (long)m * p.ld_logits + lrow * 32, v, 32);
      if (p.uniforms == nullptr) continue;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = expf(v[j] - mx);
        s += v[j];
      }
      const float u = live ? __ldg(p.uniforms + (long)m * p.R + lrow) : 0.f;
      float cdf = 0.f;
      int idx = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = 0.99f * (v[j] / s) + 0.01f * (1.0f / 32.0f);
        cdf += v[j];
        idx += (cdf <= u) ? 1 : 0;
      }
      idx = idx > 31 ? 31 : idx;
      if (live) {
        if (p.idx) p.idx[(long)m * p.ld_idx + lrow] = (uint8_t)idx;
        if (p.s_z) {
          uint4* dst = reinterpret_cast<uint4*>(p.s_z + (long)m * p.ld_s + lrow * 32);
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            uint32_t q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = w * 8 + e * 2;
              q[e] = (idx == j ? 0x3F80u : 0u) | (idx == j + 1 ? 0x3F800000u : 0u);
            }
            dst[w] = make_uint4(q[0], q[1], q[2], q[3]);
          }
        }
        if (p.latent) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((j == idx ? 1.0f : 0.0f) + v[j]) - v[j];
          store_f32_row<32>(p.latent + (long)m * p.ld_latent + lrow * 32, v, 32);
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// output layers of the [h, z] heads: bucket readout (reward / critic), Bernoulli logit
// (continue), tanh-Normal actor.  DynamicsPredictors.py:64-74, 95-105; Agent.py:191-210, 237-241.
// slot = head id.
// ------------------------------------------------------------------------------------------
enum HeadKind { HEAD_BUCKET = 0, HEAD_SIGMOID = 1, HEAD_ACTOR = 2 };
constexpr int MAX_HEADS = 5;

struct EpiHeads {
  static constexpr int B_ROWS_MAX = 256, STAGES = 4, TMEM_COLS = 256, GRU_U = 0;
  struct Params {
    const float* bias;  // [MAX_HEADS * 256]
    int kind[MAX_HEADS];
    const float* buckets[MAX_HEADS];
    float* value[MAX_HEADS];   // bucket: symexp(E[bucket]); sigmoid: probability;  element (m * ld_value)
    float* logits[MAX_HEADS];  // bucket: [M, ld_logits] fp32; sigmoid: the logit (m * ld_value); or NULL
    long ld_value[MAX_HEADS], ld_logits[MAX_HEADS];
    int NB, A;
    // actor
    const float* normals;  // [M, ld_normals] or NULL (then no action is produced)
    float *mu, *sigma, *action;  // [M, ld_act]
    long ld_act, ld_normals;
    __nv_bfloat16* s_a;    // bf16 action into the a columns of a state buffer [M, ld_s] or NULL
    int ld_s;
  };
  static __device__ __forceinline__ void run(const Params& p, const GemmCommon& g, uint32_t taddr, int m, int slot) {
    const float* bias = p.bias + slot * 256;
    const bool live = m < g.M;
    const int kind = p.kind[slot];
    if (kind == HEAD_BUCKET) {
      const int NB = p.NB;
      const int nch = (NB + 31) >> 5;
      float mx = -INFINITY;
      for (int c = 0; c < nch; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const int nvalid = min(32, NB - c * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < nvalid) {
            v[j] += __ldg(bias + c * 32 + j);
            mx = fmaxf(mx, v[j]);
          }
        }
        if (live && p.logits[slot]) store_f32_row<32>(p.logits[slot] + (long)m * p.ld_logits[slot] + c * 32, v, nvalid);
      }
      float s = 0.f, ws = 0.f;
      const float* bk = p.buckets[slot];
      for (int c = 0; c < nch; ++c) {
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        const int nvalid = min(32, NB - c * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < nvalid) {
            const float e = expf(v[j] + __ldg(bias + c * 32 + j) - mx);
            s += e;
            ws += e * __ldg(bk + c * 32 + j);
          }
        }
      }
      if (live && p.value[slot]) p.value[slot][(long)m * p.ld_value[slot]] = symexpf_(ws / s);
    } else if (kind == HEAD_SIGMOID) {
      float v[16];
      tmem_ld16(taddr, v);
      const float x = v[0] + __ldg(bias);
      if (live) {
        if (p.value[slot]) p.value[slot][(long)m * p.ld_value[slot]] = sigmoidf_(x);
        if (p.logits[slot]) p.logits[slot][(long)m * p.ld_value[slot]] = x;
      }
    } else {
      float v[32];
      tmem_ld32(taddr, v);
      if (live) {
        const int A = p.A;
        float act[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          act[j] = 0.f;
          if (j < A) {
            const float mu = v[j] + __ldg(bias + j);
            float ls = v[16 + j] + __ldg(bias + 16 + j);  // log-sigma rows are packed at 16..16+A
            ls = fminf(fmaxf(ls, -5.0f), 2.0f);
            const float sg = softplusf_(ls) + 1e-3f;
            if (p.mu) p.mu[(long)m * p.ld_act + j] = mu;
            if (p.sigma) p.sigma[(long)m * p.ld_act + j] = sg;
            if (p.normals) {
              act[j] = tanhf(mu + sg * __ldg(p.normals + (long)m * p.ld_normals + j));
              if (p.action) p.action[(long)m * p.ld_act + j] = act[j];
            }
          }
        }
        if (p.s_a && p.normals) store_bf16_row<16>(p.s_a + (long)m * p.ld_s, act, 16);
      }
    }
  }
};

}  // namespace drm
