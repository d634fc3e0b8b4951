// Library plumbing: thread-local error string, architecture gate, launch counter and the TMA
// tensor-map encoder (driver entry point resolved at run time, so the .so does not link libcuda).
#include <cudaTypedefs.h>

#include <mutex>
#include <string>
#include <vector>

#include "internal.h"

namespace drm {

static thread_local std::string t_err;
std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { t_err = msg; }
int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}

int check_arch() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(DRM_ERR_CUDA, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
  static std::mutex mu;
  static int cached_dev = -1, cached_rc = 0;
  std::lock_guard<std::mutex> lk(mu);
  if (cached_dev == dev) {
    if (cached_rc) t_err = "device is not compute capability 10.x (sm_100a required; no fallback path exists)";
    return cached_rc;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  cached_rc = (major == 10) ? 0 : DRM_ERR_ARCH;
  if (cached_rc)
    t_err = "device is compute capability " + std::to_string(major) + "." + std::to_string(minor) +
            " (sm_100a required; no fallback path exists)";
  return cached_rc;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows) {
  auto fn = encode_fn();
  if (!fn) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15u) || ((ld_elems * 2) & 15u))
    return fail(DRM_ERR_ALIGN, "tensor map base / pitch must be 16-byte aligned");
  if (box_rows == 0 || box_rows > 256) return fail(DRM_ERR_SHAPE, "tensor map box rows must be in [1, 256]");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return DRM_OK;
}

// Operand tensor map of either precision mode: bf16 (box {64, rows}) or, wide = 1, fp32 for kind::tf32 MMAs (box {32, rows});
// both are 128-byte swizzled rows.  ld_elems / cols count ELEMENTS.
int make_tmap_op_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows, int wide) {
  if (!wide) return make_tmap_bf16_2d(tm, base, rows, cols, ld_elems, box_rows);
  auto fn = encode_fn();
  if (!fn) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15u) || ((ld_elems * 4) & 15u)) return fail(DRM_ERR_ALIGN, "tensor map base / pitch must be 16-byte aligned");
  if (box_rows == 0 || box_rows > 256) return fail(DRM_ERR_SHAPE, "tensor map box rows must be in [1, 256]");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult " + std::to_string((int)r));
  return DRM_OK;
}

// fp32 operand lying K-last ([K][rows], pitch ld): box {32 rows, 32 k} written as 32 rows of 128 B with 32-byte chunks swizzled by the
// row number mod 4 (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- the one shared-memory layout tcgen05 accepts for MN-major 32-bit operands.
int make_tmap_f32_mn(CUtensorMap* tm, const void* base, uint64_t k_rows, uint64_t mn_cols, uint64_t ld_elems) {
  auto fn = encode_fn();
  if (!fn) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15u) || ((ld_elems * 4) & 15u)) return fail(DRM_ERR_ALIGN, "tensor map base / pitch must be 16-byte aligned");
  cuuint64_t gdim[2] = {mn_cols, k_rows};
  cuuint64_t gstride[1] = {ld_elems * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled (fp32, MN-major) failed with CUresult " + std::to_string((int)r));
  return DRM_OK;
}

// 2D bf16 tensor map with a narrower inner box: box {inner, box_rows}, swizzle span = inner * 2 bytes (inner = 16 / 32 / 64 elements).
// Weights of the implicit-GEMM convolutions whose k-block is one tap of <= 64 channels.
int make_tmap_bf16_2d_inner(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                            uint32_t inner) {
  auto fn = encode_fn();
  if (!fn) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15u) || ((ld_elems * 2) & 15u)) return fail(DRM_ERR_ALIGN, "tensor map base / pitch must be 16-byte aligned");
  if (box_rows == 0 || box_rows > 256) return fail(DRM_ERR_SHAPE, "tensor map box rows must be in [1, 256]");
  if (inner != 16 && inner != 32 && inner != 64) return fail(DRM_ERR_SHAPE, "tensor map inner box must be 16, 32 or 64 elements");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return DRM_OK;
}

static PFN_cuTensorMapEncodeIm2col_v12000 encode_im2col_fn() {
  static PFN_cuTensorMapEncodeIm2col_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(p);
  });
  return fn;
}

// im2col-mode tensor map over an NHWC bf16 activation [N, H, W, C]: one load delivers `pixels` consecutive filter positions (W, then
// H, then N; traversal stride `stride`) x `chunk` channels of ONE filter tap -- the A operand of an implicit-GEMM k-block, straight
// from the activation (no patch matrix).  The base pixel of a filter position ranges over [lower, extent - 1 + upper] per axis;
// positions outside the image read zeros.
int make_tmap_im2col_bf16(CUtensorMap* tm, const void* base, uint32_t C, uint32_t W, uint32_t H, uint32_t N, int lower_w, int lower_h,
                          int upper_w, int upper_h, uint32_t chunk, uint32_t pixels, uint32_t stride) {
  auto fn = encode_im2col_fn();
  if (!fn) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
  if (((uintptr_t)base & 15u) || ((C * 2) & 15u)) return fail(DRM_ERR_ALIGN, "im2col tensor map base / channel pitch must be 16-byte aligned");
  if (chunk != 16 && chunk != 32 && chunk != 64) return fail(DRM_ERR_SHAPE, "im2col tensor map: channels per load must be 16, 32 or 64");
  if (pixels == 0 || pixels > 1024 || stride == 0 || stride > 8) return fail(DRM_ERR_SHAPE, "im2col tensor map: bad pixel count / stride");
  cuuint64_t gdim[4] = {C, W, H, N};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {lower_w, lower_h}, upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, stride, stride, 1};
  const CUtensorMapSwizzle sw = chunk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (chunk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, lower, upper, chunk, pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DRM_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult " + std::to_string((int)r));
  // Drivers up to CUDA 13.1 set a descriptor bit for tensors below 128 KB that im2col loads must not see (the same correction
  // CUTLASS applies in make_im2col_tma_copy_desc).
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (uint64_t)N * H * W * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return DRM_OK;
}

// ---- per-stage event timing ---------------------------------------------------------------
static bool g_profile = false;
static std::mutex g_prof_mu;
struct EvPair { cudaEvent_t a, b; };
static std::vector<EvPair> g_events[DRM_STAGE_COUNT];
static cudaEvent_t g_open[DRM_STAGE_COUNT];

bool profile_on() { return g_profile; }
void profile_begin(int stage, cudaStream_t st) {
  if (!g_profile) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  g_open[stage] = e;
}
void profile_end(int stage, cudaStream_t st) {
  if (!g_profile) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_events[stage].push_back({g_open[stage], e});
}

}  // namespace drm

extern "C" int drm_profile_enable(int32_t on) {
  drm::g_profile = on != 0;
  return DRM_OK;
}
extern "C" int drm_profile_read(int32_t stage, double* total_ms, int64_t* launches) {
  using namespace drm;
  if (stage < 0 || stage >= DRM_STAGE_COUNT || !total_ms || !launches) return fail(DRM_ERR_ARG, "drm_profile_read: bad argument");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0;
  for (auto& p : g_events[stage]) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    tot += ms;
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  *total_ms = tot;
  *launches = (int64_t)g_events[stage].size();
  g_events[stage].clear();
  return DRM_OK;
}

extern "C" int drm_abi_version(void) { return DRM_ABI_VERSION; }
extern "C" const char* drm_last_error(void) { return drm::t_err.c_str(); }
extern "C" int drm_device_check(void) { return drm::check_arch(); }
extern "C" int64_t drm_launch_count(void) { return drm::g_launches.load(); }
