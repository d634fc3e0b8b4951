// Host-side helpers shared by the translation units of libdreamer_b200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/dreamer_b200.h"

namespace drm {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<int64_t> g_launches;

// per-stage event timing (drm_profile_enable / drm_profile_read)
bool profile_on();
void profile_begin(int stage, cudaStream_t st);
void profile_end(int stage, cudaStream_t st);

// 0 when the current device is sm_100-class; caches the answer per device.
int check_arch();

#define DRM_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ::drm::fail(DRM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
  } while (0)

#define DRM_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    ::drm::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return ::drm::fail(DRM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

#define DRM_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) return ::drm::fail((code), (msg)); \
  } while (0)

// 2D bf16 tensor map, K-major: dims {cols, rows}, row pitch ld_elems, box {64, box_rows},
// 128-byte swizzle.  Returns 0 or DRM_ERR_CUDA.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows);

int make_tmap_op_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows, int wide);
int make_tmap_f32_mn(CUtensorMap* tm, const void* base, uint64_t k_rows, uint64_t mn_cols, uint64_t ld_elems);
int make_tmap_bf16_2d_inner(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                            uint32_t inner);
// im2col-mode map over NHWC bf16 [N, H, W, C] (see core.cu)
int make_tmap_im2col_bf16(CUtensorMap* tm, const void* base, uint32_t C, uint32_t W, uint32_t H, uint32_t N, int lower_w, int lower_h,
                          int upper_w, int upper_h, uint32_t chunk, uint32_t pixels, uint32_t stride);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace drm
