// Library-level entry points of libmspl_b200.so.
#include "common.cuh"

extern "C" int mspl_abi_version(void) { return MSPL_ABI_VERSION; }

extern "C" const char* mspl_strerror(int status) {
    switch (status) {
        case MSPL_OK: return "ok";
        case MSPL_ERR_BAD_ARG: return "bad argument (null pointer, negative size, or S/K/C out of range)";
        case MSPL_ERR_ALIGN: return "misaligned pointer";
        case MSPL_ERR_UNSUPPORTED: return "unsupported configuration for this build";
        case MSPL_ERR_CUDA: return "CUDA launch failed";
        case MSPL_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown mspl status";
    }
}
