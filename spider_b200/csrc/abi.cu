// libcsa_b200.so: ABI bookkeeping (version, errors, device checks, driver entry points).
#include <cstdlib>

#include "csa_internal.h"

namespace csa {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

static uint32_t* g_dbg_host = nullptr;
static uint32_t* g_dbg_dev = nullptr;

uint32_t* debug_record_devptr() {
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = nullptr;
    if (cudaHostAlloc(&h, 4 * sizeof(uint32_t), cudaHostAllocMapped) == cudaSuccess) {
      memset(h, 0, 4 * sizeof(uint32_t));
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
        g_dbg_host = static_cast<uint32_t*>(h);
        g_dbg_dev = static_cast<uint32_t*>(d);
      }
    }
  }
  return g_dbg_dev;
}

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("CSA_PDL");
    return !(e && atoi(e) == 0);
  }();
  return on;
}

int sm_count(int device) {
  static int cache[64];
  static bool known[64];
  if (device < 0 || device >= 64) return 0;
  if (!known[device]) {
    int major = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    cache[device] = (major == 10) ? sms : 0;
    known[device] = true;
  }
  return cache[device];
}

}  // namespace csa

using namespace csa;

extern "C" int csa_abi_version(void) { return CSA_ABI_VERSION; }
extern "C" const char* csa_last_error(void) { return g_err; }

extern "C" int csa_device_supported(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n)
    return set_error(CSA_E_DEVICE, "csa_device_supported: no CUDA device %d (%s)", device, cudaGetErrorString(e));
  if (sm_count(device) <= 0)
    return set_error(CSA_E_DEVICE, "csa_device_supported: device %d is not compute capability 10.x", device);
  return 0;
}

extern "C" int csa_debug_stuck(uint32_t* out4_host) {
  if (!out4_host) return set_error(CSA_E_BADARG, "csa_debug_stuck: null out");
  cudaDeviceSynchronize();  // may itself report the trap; ignore, we only want the record
  if (!g_dbg_host) return 0;
  for (int i = 0; i < 4; ++i) out4_host[i] = g_dbg_host[i];
  return g_dbg_host[0] != 0 ? 1 : 0;
}
