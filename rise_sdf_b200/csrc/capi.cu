// C-ABI odds and ends: version + error strings.
#include "common.cuh"

extern "C" {

const char *rsdf_version(void) { return "rsdf_b200 0.1 (sm_100a)"; }

const char *rsdf_error_string(int code) {
    if (code == 0) return "ok";
    if (code == RSDF_EBADARG) return "rsdf: bad argument";
    if (code == RSDF_ECAPACITY) return "rsdf: capacity exceeded";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "rsdf: unknown error";
}

}  // extern "C"
