// Shared device/host helpers for libihgnn_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ihgnn_b200.h"

namespace ihg {

// ---- error reporting (thread-local, read back through ihg_last_error) ----------------
void set_error(const char* fmt, ...);

#define IHG_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            ::ihg::set_error(__VA_ARGS__);        \
            return IHG_ERR_INVALID_ARGUMENT;      \
        }                                         \
    } while (0)

#define IHG_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ::ihg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                   \
            return IHG_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// every kernel launch goes through this macro; the counter backs ihg_launch_count()
void count_launch();
#define IHG_LAUNCH_CHECK()            \
    do {                              \
        ::ihg::count_launch();        \
        IHG_CUDA(cudaGetLastError()); \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Workspace {
    char* base;
    int64_t size;
    int64_t used;
    Workspace(void* p, int64_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
    template <typename T>
    T* take(int64_t count) {
        int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
    bool ok() const { return used <= size && (base != nullptr || used == 0); }
};
inline int64_t ws_slice(int64_t count, int64_t elem) { return align_up(count * elem, 256); }

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (read-once) 128-bit load: do not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4(float* p, const float4& v) {
    *reinterpret_cast<float4*>(p) = v;
}
// streaming 128-bit store (write-once data that is consumed by a later kernel)
__device__ __forceinline__ void stg4_stream(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_add(float4& a, const float4& b) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& b) {
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y);
    a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
}
__device__ __forceinline__ float4 f4_scale(float s, const float4& b) {
    return make_float4(s * b.x, s * b.y, s * b.z, s * b.w);
}
__device__ __forceinline__ float4 f4_mul(const float4& a, const float4& b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

}  // namespace ihg
