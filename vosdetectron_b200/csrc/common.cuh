// Shared helpers for libvosd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/vosd_b200.h"

namespace vosd {

// Diagnostic launch counter (host side, relaxed atomic).
void count_launch(int n = 1);

// VOSD_B200_DEBUG=1 in the environment: the CUDA error string behind a VOSD_ERR_LAUNCH goes to stderr.
inline int check_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return VOSD_OK;
    if (getenv("VOSD_B200_DEBUG")) fprintf(stderr, "[vosd_b200] CUDA error: %s\n", cudaGetErrorString(e));
    return VOSD_ERR_LAUNCH;
}

// The vosd_debug_* entry points (kernel-family selection for the parity tests) only take effect in processes that
// export VOSD_B200_TEST_HOOKS=1 (tests/conftest.py, tools/): a production process has no mutable library state.
inline bool test_hooks_enabled() {
    const char* e = getenv("VOSD_B200_TEST_HOOKS");
    return e && e[0] == '1';
}

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Order-preserving float <-> uint32 (larger float -> larger uint).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// 128-bit streaming store / load (data touched once: do not pollute L1).
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

}  // namespace vosd
