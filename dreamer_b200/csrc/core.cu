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
