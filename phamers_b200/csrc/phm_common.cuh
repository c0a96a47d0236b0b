// Shared helpers for the phamers_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/phamers_b200.h"

namespace phm {

void set_error(const char *fmt, ...);

#define PHM_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            phm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PHM_E_CUDA;                                                                 \
        }                                                                                      \
    } while (0)

#define PHM_REQUIRE(cond, msg)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            phm::set_error("invalid argument: %s (%s)", msg, #cond);                           \
            return PHM_E_ARG;                                                                  \
        }                                                                                      \
    } while (0)

int sm_count();                       // multiprocessor count of the CURRENT device (cached per device)
int max_smem_optin();                 // opt-in shared memory per block of the current device (cached per device)

extern unsigned long long kernel_launches;      // kernels this library has launched in this process (phm_kernel_launches)
// after every <<< >>>: counts the launch and turns a launch error into PHM_E_CUDA
#define PHM_LAUNCH_CHECK()                                                                     \
    do {                                                                                       \
        ++phm::kernel_launches;                                                                \
        PHM_CUDA_CHECK(cudaGetLastError());                                                    \
    } while (0)

extern int time_kernels;              // option "time_kernels": hot kernels are bracketed with CUDA events (phm_last_kernel_ms)

// CUDA-event brackets of the launches of one kernel since the last read (bench.py's roofline numerator)
struct EventRing {
    static constexpr int CAP = 64;
    cudaEvent_t ev[CAP][2];
    int made = 0, used = 0;
    bool begin(cudaStream_t st) {                      // true if this launch is being timed
        if (!time_kernels || used >= CAP) return false;
        if (made <= used) {
            if (cudaEventCreate(&ev[used][0]) != cudaSuccess || cudaEventCreate(&ev[used][1]) != cudaSuccess) return false;
            made = used + 1;
        }
        return cudaEventRecord(ev[used][0], st) == cudaSuccess;
    }
    void end(cudaStream_t st) { cudaEventRecord(ev[used][1], st); ++used; }
    int mean_ms(float *ms) {                           // mean over the recorded launches, then the ring is emptied
        if (used == 0) { set_error("no timed launch of this kernel (set option time_kernels = 1 first)"); return PHM_E_ARG; }
        double total = 0.0;
        for (int i = 0; i < used; ++i) {
            float one = 0.f;
            if (cudaEventSynchronize(ev[i][1]) != cudaSuccess || cudaEventElapsedTime(&one, ev[i][0], ev[i][1]) != cudaSuccess) {
                set_error("reading a kernel timing event failed");
                return PHM_E_CUDA;
            }
            total += one;
        }
        *ms = (float)(total / used);
        used = 0;
        return PHM_OK;
    }
};

constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// streaming 128-bit load: read-only path, do not keep in L1 (every base is read exactly once)
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void red_shared_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4_zero(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}

// ---- mbarrier + bulk-copy (TMA) helpers shared by the scorer and the histogram kernel ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();        // a protocol bug must fail loudly, never hang the GPU
    }
}
// 1-D bulk copy global -> shared (16-byte aligned addresses and size), completing `bytes` on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

// RN(a / b) for integer-valued 0 <= a <= b < 2^53 from r = RN(1 / b): q0 = RN(a r) is within 2 ulp, e = a - b q0 is exact (FMA),
// q0 + e / b = a / b, and a quotient of such integers is never within 2^-93 (relative) of a rounding boundary nor on one, so
// RN(q0 + e r) is the IEEE quotient numpy computes in kmer.normalize_counts (scripts/kmer.py:219-220).  b = 0 (empty contig):
// r = inf, 0 * inf = NaN = 0 / 0.  Checked bit for bit against true division by tests/test_gpu_count.py and
// tests/test_gpu_score.py::test_scoring_from_counts_equals_scoring_from_features.
__device__ __forceinline__ double exact_quotient(double a, double b, double r) {
    const double q0 = a * r;
    const double e = fma(-q0, b, a);
    return fma(e, r, q0);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace phm
