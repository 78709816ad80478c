// extern "C" surface shared by every op: version, error reporting, device check, GEMM dispatch.
#include <stdarg.h>
#include <atomic>
#include "sst_common.cuh"

namespace sst {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what, int n_kernels) {
  g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SST_E_LAUNCH;
  }
  return SST_OK;
}

static const unsigned long long* g_salt = nullptr;
const unsigned long long* dropout_salt() { return g_salt; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SST_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int launch_gemm_tcgen05(const SstGemmDesc&, const void*, const void*, void*, const void*, const void*, cudaStream_t);
int launch_gemm_simt(const SstGemmDesc&, const void*, const void*, void*, const void*, const void*, cudaStream_t);

}  // namespace sst

extern "C" {

const char* sst_version(void) { return "sst-b200 0.1 (sm_100a)"; }
const char* sst_last_error(void) { return sst::g_err; }
long long sst_launch_count(void) { return sst::g_launches.load(std::memory_order_relaxed); }

int sst_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { sst::set_error("no CUDA device: %s", cudaGetErrorString(e)); return SST_E_ARCH; }
  if (major != 10) { sst::set_error("device is sm_%d0, libsst.so is built for sm_100a only", major); return SST_E_ARCH; }
  return SST_OK;
}

int sst_set_dropout_salt(const uint64_t* salt_dev) {
  sst::g_salt = reinterpret_cast<const unsigned long long*>(salt_dev);
  return SST_OK;
}

int sst_gemm(const SstGemmDesc* d, const void* A, const void* B, void* C, const void* bias, const void* aux, void* stream) {
  SST_REQUIRE(d && A && B && C, SST_E_ARG, "sst_gemm: null argument");
  SST_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, SST_E_ARG, "sst_gemm: empty problem %ldx%ldx%ld", (long)d->M, (long)d->N, (long)d->K);
  SST_REQUIRE(!(d->epilogue & SST_EPI_BIAS) || bias, SST_E_ARG, "sst_gemm: BIAS epilogue without bias pointer");
  SST_REQUIRE(!(d->epilogue & SST_EPI_MULMASK) || aux, SST_E_ARG, "sst_gemm: MULMASK epilogue without aux pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d->dtype == SST_BF16 && !d->force_simt) return sst::launch_gemm_tcgen05(*d, A, B, C, bias, aux, st);
  return sst::launch_gemm_simt(*d, A, B, C, bias, aux, st);
}

}  // extern "C"
