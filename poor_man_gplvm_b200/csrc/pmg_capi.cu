// Library-level entry points of the C ABI (version, error strings, device query).
#include "pmg_common.cuh"

extern "C" int pmg_version(void) { return 100; }

extern "C" const char* pmg_error_string(int code) {
  switch (code) {
    case PMG_OK: return "ok";
    case PMG_ERR_BAD_ARG: return "bad argument";
    case PMG_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
    case PMG_ERR_ALIGNMENT: return "pointer or leading dimension not aligned";
    case PMG_ERR_WORKSPACE: return "workspace missing or too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

extern "C" int pmg_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return sms;
}
