// Shared device/host helpers for the mspl_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <initializer_list>

#include "../../include/mspl_b200.h"

namespace mspl {

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNearTieMargin = 1e-6f; // BASELINE.json north_star: top-2 probability margin of a "marginal" pixel

#define MSPL_DEVINL __device__ __forceinline__

// ---- streaming vector access: P consecutive fp32 pixels of one class plane -------------------------
// Inputs are read exactly once, so they bypass L1 allocation (ld.global.nc.L1::no_allocate).
template <int P> struct PixVec;
template <> struct PixVec<1> {
    static MSPL_DEVINL void load(const float* p, float (&v)[1]) {
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
    }
    static MSPL_DEVINL void store(float* p, const float (&v)[1]) { __stcs(p, v[0]); }
};
template <> struct PixVec<2> {
    static MSPL_DEVINL void load(const float* p, float (&v)[2]) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
    static MSPL_DEVINL void store(float* p, const float (&v)[2]) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    }
};
template <> struct PixVec<4> {
    static MSPL_DEVINL void load(const float* p, float (&v)[4]) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
    }
    static MSPL_DEVINL void store(float* p, const float (&v)[4]) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};

template <int P> MSPL_DEVINL void store_labels(uint8_t* p, const int (&l)[P]);
template <> MSPL_DEVINL void store_labels<1>(uint8_t* p, const int (&l)[1]) { p[0] = (uint8_t)l[0]; }
template <> MSPL_DEVINL void store_labels<2>(uint8_t* p, const int (&l)[2]) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2((uint8_t)l[0], (uint8_t)l[1]);
}
template <> MSPL_DEVINL void store_labels<4>(uint8_t* p, const int (&l)[4]) {
    *reinterpret_cast<uchar4*>(p) = make_uchar4((uint8_t)l[0], (uint8_t)l[1], (uint8_t)l[2], (uint8_t)l[3]);
}

// exp(x - M) with one FFMA + one MUFU.EX2: callers pass x and Ml = M*log2e.
MSPL_DEVINL float exp_shifted(float x, float Ml) { return exp2f(fmaf(x, kLog2e, -Ml)); }

// Order-preserving map fp32 -> u32 (larger float <=> larger key) and back; used by the radix select.
MSPL_DEVINL uint32_t float_to_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);      // negative: flip all bits; else set the sign bit
}
MSPL_DEVINL float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
// digit of `key` examined by radix pass 0/1/2 (11 + 11 + 10 bits) and the prefix above it
MSPL_DEVINL uint32_t radix_digit(uint32_t key, int pass) {
    return pass == 0 ? (key >> 21) : pass == 1 ? ((key >> 10) & 0x7ffu) : (key & 0x3ffu);
}
MSPL_DEVINL uint32_t radix_prefix(uint32_t key, int pass) {
    return pass == 0 ? 0u : pass == 1 ? (key >> 21) : (key >> 10);
}

// Bin of the LINEAR confidence histogram (MSPL_RADIX_BINS equal steps over [0,1], clamped): monotone non-decreasing in f
// for every float (x2048 is exact), so "bin above / below the class's selected bin" brackets the order statistic.
// bin(f) >= b  <=>  f >= b/2048 for 1 <= b <= 2047 (both sides exact in fp32).  NaN lands in bin 0.
MSPL_DEVINL uint32_t conf_bin(float f) {
    return (uint32_t)max(0, min(MSPL_RADIX_BINS - 1, __float2int_rd(f * (float)MSPL_RADIX_BINS)));
}

MSPL_DEVINL double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int launch_status() { return cudaPeekAtLastError() == cudaSuccess ? MSPL_OK : MSPL_ERR_CUDA; }

// Largest pixel vector width usable for planes of `hw` floats starting at the given bases.
inline int pick_vec(int64_t hw, std::initializer_list<const void*> ptrs) {
    int p = (hw % 4 == 0) ? 4 : (hw % 2 == 0) ? 2 : 1;
    for (const void* q : ptrs) {
        if (!q) continue;
        while (p > 1 && !aligned_to(q, 4 * p)) p >>= 1;
    }
    return p;
}

}  // namespace mspl
