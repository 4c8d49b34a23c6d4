// A very small CUDA execution-model emulation for CPU tests of simple kernels (test infrastructure only).
//
// Enough to compile kernel SOURCE unchanged with g++ when it uses only threadIdx / blockIdx / blockDim / gridDim (x),
// __shared__ arrays and __syncthreads(): a launch runs the blocks one after the other; a block is either a plain loop over
// its threads (kernels without barriers) or blockDim.x real threads meeting at a pthread barrier.  `__shared__` becomes a
// function-local static (blocks never overlap, so one copy is exactly per-block storage).  No atomics, shuffles or textures.
#pragma once
#include <pthread.h>

#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

namespace cuda_emu {
struct Dim {
    unsigned x = 1, y = 1, z = 1;
};
inline thread_local Dim threadIdx_, blockIdx_;
inline Dim blockDim_, gridDim_;
inline pthread_barrier_t* barrier_ = nullptr;
inline void sync() {
    if (barrier_) pthread_barrier_wait(barrier_);
}
// run `body()` for every thread of every block; with_barrier: the kernel calls __syncthreads()
inline void launch(unsigned grid, unsigned block, bool with_barrier, const std::function<void()>& body) {
    gridDim_.x = grid;
    blockDim_.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        if (!with_barrier) {
            barrier_ = nullptr;
            for (unsigned t = 0; t < block; ++t) {
                blockIdx_.x = b;
                threadIdx_.x = t;
                body();
            }
            continue;
        }
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, block);
        barrier_ = &bar;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, b]() {
                blockIdx_.x = b;
                threadIdx_.x = t;
                body();
            });
        for (auto& x : th) x.join();
        barrier_ = nullptr;
        pthread_barrier_destroy(&bar);
    }
}
}  // namespace cuda_emu

#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __syncthreads() cuda_emu::sync()
#define threadIdx cuda_emu::threadIdx_
#define blockIdx cuda_emu::blockIdx_
#define blockDim cuda_emu::blockDim_
#define gridDim cuda_emu::gridDim_
