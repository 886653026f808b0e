// Common definitions for the B200-native beta-diversity kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace b2d {

// Geometry shared by host and device code.
constexpr int TILE = 128;        // samples per row/column stripe ("sample block")
constexpr int NODE_STRIPE = 128; // nodes per node stripe (16 bytes of presence bits per sample)

// Embedding layouts in HBM (q = node position in the engine's own postorder):
//   fp64 : emb[block][q - q0][TILE]                      one node row of one block = 1 KB
//   bits : bits[block][stripe][TILE][4 x u32]            one (block, stripe) = 2 KB, node-packed
// Both make "all samples of a block x a run of nodes" one contiguous chunk, which is what the
// pair kernels stream into shared memory with single bulk copies.

extern thread_local std::string g_last_error;

inline int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define B2D_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            cudaGetLastError();                                                            \
            return ::b2d::fail(e__ == cudaErrorMemoryAllocation ? 3 : 2,                   \
                               std::string(#call) + ": " + cudaGetErrorString(e__));      \
        }                                                                                  \
    } while (0)

// ---- condensed (SciPy pdist) indexing --------------------------------------------------
// element (i<j) of an n x n matrix lives at i*n + j - (i+2)(i+1)/2
__host__ __device__ inline int64_t cond_row_start(int64_t i, int64_t n) {
    // index of (i, i+1)
    return i * n - (i * (i + 1)) / 2;
}
__host__ __device__ inline int64_t cond_index(int64_t i, int64_t j, int64_t n) {
    return i * n - ((i + 2) * (i + 1)) / 2 + j;
}

// ---- PTX helpers: mbarrier + 1-D bulk (TMA) copies --------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the warp sleeps in hardware (up to `ns`) instead of spinning
// through the issue slots the working warps need
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++spins > (1u << 22)) __trap();
    }
}
// Bounded wait: a protocol bug traps instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// global -> shared bulk copy engine (SASS: UBLKCP); completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif

}  // namespace b2d
