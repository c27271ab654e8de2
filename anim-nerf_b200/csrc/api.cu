// Library-level entry points of the C ABI (version, error strings).
#include "common.cuh"

extern "C" int an_version(void) { return 100; }   // 0.1.0

extern "C" const char* an_error_string(int code)
{
    switch (code) {
        case AN_OK: return "ok";
        case AN_ERR_ARG: return "invalid argument (null pointer or non-positive extent)";
        case AN_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case AN_ERR_ALIGN: return "pointer alignment requirement not met";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
