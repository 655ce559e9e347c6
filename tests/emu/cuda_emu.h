// Host emulation of the small CUDA subset csrc/fieldhead.cu uses -- TEST INFRASTRUCTURE ONLY.
//
// One CTA at a time; every CUDA thread of the CTA is a real pthread, __syncthreads() is a pthread barrier, __shared__
// arrays are function-local statics.  Built two ways by tests/test_field_head.py: as a shared library whose results
// are compared with torch ops, and as an executable under -fsanitize=thread, where a missing or misplaced
// __syncthreads() shows up as a reported data race on the shared arrays.
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <cmath>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
};
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static pthread_barrier_t* emu_barrier = nullptr;
inline void __syncthreads() { pthread_barrier_wait(emu_barrier); }

namespace bh {
constexpr int kNumSMs = 148;
inline float4 ldg_stream(const float4* p) { return *p; }
inline void stg_stream(float4* p, const float4& v) { *p = v; }
}  // namespace bh

// run `body()` as a grid of `grid` CTAs of `block` threads (CTAs one after the other)
inline void emu_launch(int grid, int block, const std::function<void()>& body) {
    struct Arg {
        const std::function<void()>* body;
        int t, b, block, grid;
    };
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, block);
    emu_barrier = &bar;
    for (int b = 0; b < grid; ++b) {
        std::vector<pthread_t> th(block);
        std::vector<Arg> args(block);
        for (int t = 0; t < block; ++t) {
            args[t] = Arg{&body, t, b, block, grid};
            pthread_create(&th[t], nullptr, [](void* p) -> void* {
                Arg* a = static_cast<Arg*>(p);
                threadIdx.x = a->t;
                blockIdx.x = a->b;
                blockDim.x = a->block;
                gridDim.x = a->grid;
                (*a->body)();
                return nullptr;
            }, &args[t]);
        }
        for (int t = 0; t < block; ++t) pthread_join(th[t], nullptr);
    }
    pthread_barrier_destroy(&bar);
    emu_barrier = nullptr;
}
