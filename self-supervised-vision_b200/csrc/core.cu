// Library-level entry points: version, error strings, device check.
#include "host_util.h"

extern "C" {

int ssvb_version(void) { return SSVB_VERSION; }

const char* ssvb_strerror(int rc) {
  switch (rc) {
    case SSVB_OK: return "ok";
    case SSVB_ERR_INVALID: return "ssv_b200: invalid argument (null pointer, non-positive size or bad flag)";
    case SSVB_ERR_ALIGNMENT: return "ssv_b200: pointer / leading dimension not 16-byte aligned (need d % 4 == 0)";
    case SSVB_ERR_UNSUPPORTED: return "ssv_b200: shape not supported by the sm_100a kernels";
    case SSVB_ERR_WORKSPACE: return "ssv_b200: workspace / saved buffer too small";
    case SSVB_ERR_ARCH: return "ssv_b200: current device is not a Blackwell sm_100 GPU (no fallback path exists)";
    case SSVB_ERR_DRIVER: return "ssv_b200: cuTensorMapEncodeTiled unavailable or failed";
    default: break;
  }
  if (rc > 0) return cudaGetErrorString(static_cast<cudaError_t>(rc));
  return "ssv_b200: unknown error";
}

int ssvb_device_check(void) { return ssvb::check_device_sm100(); }

}  // extern "C"
