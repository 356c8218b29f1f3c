// Library-wide C-ABI plumbing: error text, ABI version, device check, launch counter.
#include "common.cuh"

namespace alg {
static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace alg

extern "C" int alg_abi_version(void) { return ALG_B200_ABI_VERSION; }
extern "C" const char* alg_last_error(void) { return alg::g_error.c_str(); }
extern "C" int64_t alg_launch_count(void) { return alg::g_launches.load(); }

extern "C" int alg_check_device(void) {
  using namespace alg;
  static std::atomic<int> ok_mask[64];  // per-device cache: 1 = verified sm_100
  int dev = 0;
  ALG_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && ok_mask[dev].load(std::memory_order_relaxed) == 1) return 0;
  int major = 0, minor = 0;
  ALG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ALG_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  ALG_REQUIRE(major == 10, "this library is built for sm_100a (B200) only; found sm_" + std::to_string(major) +
                               std::to_string(minor));
  if (dev >= 0 && dev < 64) ok_mask[dev].store(1, std::memory_order_relaxed);
  return 0;
}
