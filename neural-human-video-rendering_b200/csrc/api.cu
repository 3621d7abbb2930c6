// Library-level C-ABI: version, error strings, architecture gate, launch counter.
#include <cuda_runtime.h>
#include <atomic>
#include <cstring>
#include "../../include/nhvr.h"

namespace nhvr {

static std::atomic<uint64_t> g_launches{0};
static char g_last_err[256] = "";
static int g_arch_state = -1;   // -1 unknown, 0 bad, 1 ok
static int g_operand_f16 = 0;
int operand_f16() { return g_operand_f16; }
static int32_t* g_overflow_flag = nullptr;
int32_t* overflow_flag() { return g_overflow_flag; }

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void note_cuda_error(cudaError_t e) {
  std::strncpy(g_last_err, cudaGetErrorString(e), sizeof(g_last_err) - 1);
  g_last_err[sizeof(g_last_err) - 1] = 0;
}

int arch_ok_cached() {
  if (g_arch_state < 0) {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) { note_cuda_error(e); g_arch_state = 0; }
    else g_arch_state = (major == 10) ? 1 : 0;
  }
  return g_arch_state;
}

}  // namespace nhvr

extern "C" int nhvr_version(void) { return 100; }

extern "C" const char* nhvr_strerror(int status) {
  switch (status) {
    case NHVR_OK: return "ok";
    case NHVR_ERR_ARCH: return "device is not compute capability 10.x (sm_100a kernels only, no fallback)";
    case NHVR_ERR_SHAPE: return "unsupported or inconsistent shape";
    case NHVR_ERR_ALIGN: return "pointer is not 16-byte aligned";
    case NHVR_ERR_NULL: return "required pointer is NULL";
    case NHVR_ERR_CUDA: return "CUDA runtime error (see nhvr_last_cuda_error)";
    case NHVR_ERR_SMEM: return "tile does not fit in shared / tensor memory";
    case NHVR_ERR_UNSUPPORTED: return "operation not supported";
    default: return "unknown status";
  }
}

extern "C" const char* nhvr_last_cuda_error(void) { return nhvr::g_last_err; }
extern "C" int nhvr_arch_ok(void) { return nhvr::arch_ok_cached() == 1 ? 0 : NHVR_ERR_ARCH; }
extern "C" uint64_t nhvr_launch_count(void) { return nhvr::g_launches.load(); }

extern "C" int nhvr_set_operand_dtype(int is_f16) { nhvr::g_operand_f16 = is_f16 ? 1 : 0; return NHVR_OK; }
extern "C" int nhvr_get_operand_dtype(void) { return nhvr::g_operand_f16; }

extern "C" int nhvr_set_overflow_flag(int32_t* flag) { nhvr::g_overflow_flag = flag; return NHVR_OK; }
