// Error reporting + version for the C ABI (include/slam2d_b200.h).
#include "common.cuh"

namespace slam {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return (int)e;
}
}  // namespace slam

extern "C" const char* slam_last_error(void) { return slam::g_err.c_str(); }
extern "C" int slam_version(void) { return 100; }
