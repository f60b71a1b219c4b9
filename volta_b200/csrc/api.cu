// Library-level entry points of the C ABI (include/volta_b200.h): error reporting, device check,
// GEMM back-end dispatch.
#include <stdarg.h>
#include <string.h>

#include "gemm_common.cuh"

namespace vb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static const uint64_t* g_seed_ptr = nullptr;
static int g_gemm_sm_limit = 0;
int gemm_sm_limit() { return g_gemm_sm_limit; }
const uint64_t* seed_pointer() { return g_seed_ptr; }

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

}  // namespace vb

extern "C" {

int vb_version(void) { return 1003; }

int vb_set_gemm_sm_limit(int32_t sms) {
  vb::g_gemm_sm_limit = sms > 0 ? sms : 0;
  return 0;
}

int vb_set_seed_pointer(const void* dev_ptr) {
  vb::g_seed_ptr = reinterpret_cast<const uint64_t*>(dev_ptr);
  return 0;
}

const char* vb_last_error(void) { return vb::g_err; }

int vb_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  VB_REQUIRE(e == cudaSuccess, VB_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  VB_REQUIRE(major == 10 && minor == 0, VB_ERR_DEVICE,
             "volta_b200 is built for sm_100a (B200) only; device %d is sm_%d%d", dev, major, minor);
  return 0;
}

int vb_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

int vb_gemm_backend(const vb_gemm_args* a) {
  if (!a) return VB_ERR_ARG;
  if (a->backend == VB_GEMM_TC || a->backend == VB_GEMM_SIMT) return a->backend;
  // tiny-N heads (ITM N=2, retrieval N=1) are latency-bound mat-vecs: not worth a 128-wide MMA tile
  if (a->N < 16 || a->M < 8 || a->K < 8) return VB_GEMM_SIMT;
  return vb::gemm_tc_supported(*a, nullptr) ? VB_GEMM_TC : VB_GEMM_SIMT;
}

int vb_gemm(const vb_gemm_args* a, void* stream);

int vb_gemm_bench(const vb_gemm_args* a, void* stream, int32_t repeats, float* ms_per_launch) {
  VB_REQUIRE(a != nullptr && ms_per_launch != nullptr && repeats > 0, VB_ERR_ARG, "vb_gemm_bench: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  VB_REQUIRE(cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess, VB_ERR_DEVICE, "vb_gemm_bench: no events");
  int rc = vb_gemm(a, stream);  // warm: tensor maps, attributes, L2
  cudaEventRecord(e0, st);
  for (int i = 0; i < repeats && rc == 0; ++i) rc = vb_gemm(a, stream);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / (float)repeats;
  return rc;
}

int vb_gemm(const vb_gemm_args* a, void* stream) {
  VB_REQUIRE(a != nullptr, VB_ERR_ARG, "vb_gemm: null args");
  VB_REQUIRE(a->ab_dtype == VB_F32 || a->ab_dtype == VB_BF16, VB_ERR_ARG, "vb_gemm: bad ab_dtype %d", a->ab_dtype);
  VB_REQUIRE(a->d_dtype == VB_F32 || a->d_dtype == VB_BF16, VB_ERR_ARG, "vb_gemm: bad d_dtype %d", a->d_dtype);
  const int be = vb_gemm_backend(a);
  if (be == VB_GEMM_TC) return vb::gemm_tc_launch(*a, (cudaStream_t)stream);
  VB_REQUIRE(a->D2 == nullptr || a->k_seg <= 0, VB_ERR_UNSUPPORTED,
             "vb_gemm: two reduction segments (k_seg / D2) need the tcgen05 back-end (bf16 operands, aligned pitches)");
  return vb::gemm_simt_launch(*a, (cudaStream_t)stream);
}

}  // extern "C"
