// Common device/host helpers for the heal_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#define HEAL_OK 0
#define HEAL_ERR_ARG (-1)
#define HEAL_ERR_WORKSPACE (-2)
#define HEAL_ERR_LAUNCH (-3)
#define HEAL_ERR_UNSUPPORTED (-4)
#define HEAL_ERR_DRIVER (-5)

#define HEAL_NUM_SMS 148  // B200: 2 dies x 74 SMs

// process-wide count of kernels this library has launched (bench.py reports it as gpu_launches)
extern "C" void heal_launch_counter_add(int n);

static inline int heal_check_launch(int kernels_launched = 1) {
    cudaError_t e = cudaGetLastError();
    heal_launch_counter_add(kernels_launched);
    return e == cudaSuccess ? HEAL_OK : HEAL_ERR_LAUNCH;
}

static inline size_t heal_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace; no internal allocation anywhere in the library.
struct HealArena {
    char* base;
    size_t cap;
    size_t off;
    __host__ HealArena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
    template <typename T>
    __host__ T* take(size_t n) {
        off = heal_align_up(off, 256);
        T* r = (T*)(base + off);
        off += n * sizeof(T);
        return r;
    }
    __host__ bool ok() const { return base != nullptr && off <= cap; }
};

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
