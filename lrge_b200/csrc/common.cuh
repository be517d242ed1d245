// common.cuh -- shared device/host helpers for liblrge_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>

namespace lrge {

constexpr uint64_t U64MAX = ~0ULL;

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        throw CudaError(buf);
    }
}
#define CUDA_CHECK(x) ::lrge::cuda_check((x), #x, __FILE__, __LINE__)

// bytes currently held by DevBufs of this process (cudaMemGetInfo costs ~0.5 ms per call, so budgets use this instead)
inline std::atomic<size_t> &dev_bytes_in_use() { static std::atomic<size_t> v{0}; return v; }

// RAII device buffer (raw bytes); grows on demand, never shrinks while the context lives.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) { cudaFree(p); dev_bytes_in_use() -= cap; }
        p = nullptr; cap = 0;
    }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        release();
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            char buf[256];
            snprintf(buf, sizeof buf, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            throw CudaError(buf);
        }
        cap = want;
        dev_bytes_in_use() += cap;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ---- PTX wrappers: mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) ----------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

}  // namespace lrge
