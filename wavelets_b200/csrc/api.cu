// Version and status text of the C ABI (include/wavelets_b200.h).
#include "common.cuh"

extern "C" {

int wb_abi_version(void) { return WB_ABI_VERSION; }

const char *wb_error_string(int status) {
    switch (status) {
        case WB_OK: return "ok";
        case WB_EINVAL_DTYPE: return "unsupported dtype (WB_F32 or WB_F64 expected)";
        case WB_EINVAL_TAPS: return "unsupported scaling function (WB_TRIANGLE or WB_B3SPLINE expected)";
        case WB_EINVAL_SHAPE: return "invalid shape (batch, H, W must be >= 1 and batch <= 65535)";
        case WB_EINVAL_SCALE: return "invalid scale / number of levels";
        case WB_EINVAL_POINTER: return "null or aliasing device pointer";
        case WB_EINVAL_ARG: return "invalid argument (pitch smaller than width, bad count, ...)";
        case WB_ENOT_FUSABLE: return "shape or alignment outside the fused WOW kernel (use the two-pass route)";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown wavelets_b200 status";
}

}  // extern "C"
