// Micro-benchmark: issue rate of the integer instructions kmer_hist_kernel is made of, alone and mixed, to see which pipe each one
// uses (the kernel is co-limited by the ALU pipe and the shared-memory pipe).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pipe_bench tools/microbench/pipe_bench.cu && /tmp/pipe_bench
// Each thread runs 8 independent dependency chains; prints warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_SHR, OP_LOP, OP_MULLO, OP_MULHI, OP_PRMT, OP_SHR_MULHI, OP_LOP_MULLO, OP_LOP_MULHI, OP_FUNNEL, OP_IADD, OP_LOP_SHR, OP_BREV, OP_POPC };

template <int OP>
__global__ void __launch_bounds__(256) bench(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t x[8];
    uint32_t c1, c2;
    asm volatile("mov.u32 %0, %1;" : "=r"(c1) : "r"(seed | 0x00010000u));     // opaque "constants"
    asm volatile("mov.u32 %0, %1;" : "=r"(c2) : "r"(seed * 3u + 5u));
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x * 8 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (OP == OP_SHR) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x[j]));
                if (OP == OP_LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(c1), "r"(c2));
                if (OP == OP_MULLO) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
                if (OP == OP_MULHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
                if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(x[j]) : "r"(c1));
                if (OP == OP_FUNNEL) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[j]) : "r"(c1));
                if (OP == OP_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
                if (OP == OP_BREV) asm volatile("brev.b32 %0, %0;" : "+r"(x[j]));
                if (OP == OP_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[j]));
                if (OP == OP_SHR_MULHI) { if (j & 1) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x[j])); else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1)); }
                if (OP == OP_LOP_MULLO) { if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(c1), "r"(c2)); else asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1)); }
                if (OP == OP_LOP_MULHI) { if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(c1), "r"(c2)); else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1)); }
                if (OP == OP_LOP_SHR) { if (j & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(c1), "r"(c2)); else asm volatile("shr.u32 %0, %0, 1;" : "+r"(x[j])); }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= x[j];
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

template <int OP>
static void run(const char *name) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *sink; cudaMalloc(&sink, 4);
    const int iters = 20000, ctas = 4;                // 4 CTAs x 8 warps = 8 warps per sub-partition
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<OP><<<sms * ctas, 256>>>(100, 1u, sink);
    cudaEventRecord(a);
    bench<OP><<<sms * ctas, 256>>>(iters, 1u, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double instr = (double)ctas * 8 * iters * 32.0;      // warp-instructions per SM
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-34s %8.3f ms   warp-instructions / clk / sub-partition %.3f   (%s)\n", name, ms, instr / clk / 4.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    run<OP_SHR>("shr (SHF)");
    run<OP_FUNNEL>("shf.r.wrap (SHF)");
    run<OP_LOP>("lop3 (LOP3)");
    run<OP_IADD>("add (IADD3)");
    run<OP_PRMT>("prmt (PRMT)");
    run<OP_BREV>("brev (BREV)");
    run<OP_POPC>("popc (POPC)");
    run<OP_MULLO>("mul.lo (IMAD)");
    run<OP_MULHI>("mul.hi (IMAD.HI)");
    run<OP_LOP_SHR>("lop3 + shr, alternating");
    run<OP_SHR_MULHI>("shr + mul.hi, alternating");
    run<OP_LOP_MULLO>("lop3 + mul.lo, alternating");
    run<OP_LOP_MULHI>("lop3 + mul.hi, alternating");
    return 0;
}
