// Shared declarations for libof3d (sm_100a). Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>

#include "../../include/of3d.h"

namespace of3d {

constexpr int kMaxTaps = 257;   // radius <= 128  (sigma <= 42)
constexpr int kMaxFrames = 129; // temporal taps (tSig <= 21)

// One 1-D filter passed BY VALUE as a kernel parameter: it lands in constant bank 0, so
// statically indexed taps become immediate constant operands of DFMA/FFMA and there is no
// shared mutable __constant__ state between contexts.
template <typename T>
struct Filt {
    int n;      // number of taps (odd)
    int sym;    // +1 symmetric, -1 antisymmetric, 0 neither (scipy's NI_Correlate1D test)
    T w[kMaxTaps];
};

struct FramePtrs {
    const void* p[kMaxFrames];
};

void set_error(const std::string& msg);
const char* cuda_err_name(cudaError_t e);

#define OF3D_CUDA_TRY(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::of3d::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));     \
            return (_e == cudaErrorMemoryAllocation) ? OF3D_ERR_NOMEM : OF3D_ERR_CUDA;        \
        }                                                                                     \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace of3d
