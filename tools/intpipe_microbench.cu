// Integer-pipe microbenchmark for B200 (sm_100a): measures the sustained issue
// rate (warp-instructions per clock per SM, and lane-ops/s chip-wide) of the
// instruction classes the Poseidon2-M31 kernels are made of.  Its output is the
// measured denominator of the integer roofline (profiles/intpipe_r01.json).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o intpipe tools/intpipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

enum Kind { K_IADD3 = 0, K_LOP3, K_SHF, K_LEAHI, K_IMAD, K_IMADWIDE, K_MIX_WIDE_LEA, K_MIX_IMAD_IADD,
            K_MIX_WIDE_IADD2, K_MULRED, K_UMIN, K_MIX3, K_WIDE_CHAIN, K_WIDE_LOP, K_IMAD_LEA, K_IMADHI, K_IADD3_3IN, K_WIDE_LOP2, K_COUNT };
static const char *names[K_COUNT] = {"iadd3", "lop3", "shf", "lea_hi", "imad", "imad_wide", "wide+lea_hi",
                                     "imad+iadd3", "wide+2alu", "m31_mul_lazy(wide+lea)", "umin", "imad+iadd3+lop3", "imad_wide(chain on lo)", "wide+lop3", "imad+lea_hi", "imad_hi", "iadd3(3-input)", "wide+2lop3"};
static const int ops_per_iter[K_COUNT] = {1, 1, 1, 1, 1, 1, 2, 2, 3, 2, 1, 3, 1, 2, 2, 1, 1, 3};

template <int KIND>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t seed, uint64_t *cycles) {
    uint32_t a[CHAINS], b[CHAINS];
    uint64_t w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = seed + threadIdx.x * 7 + c; b[c] = seed ^ (c * 77 + 1); w[c] = a[c]; }
    uint64_t t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (KIND == K_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[c]) : "r"(b[c]));
                if (KIND == K_LEAHI) { uint32_t t; asm volatile("shr.u32 %0, %1, 1;\n\tadd.u32 %1, %0, %2;" : "=&r"(t), "+r"(a[c]) : "r"(b[c])); }
                if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                if (KIND == K_IMADWIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c]));
                if (KIND == K_MIX_WIDE_LEA || KIND == K_MULRED) {
                    uint64_t x; uint32_t lo, hi;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x) : "r"(a[c]), "r"(b[c]));
                    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
                    asm volatile("shr.u32 %0, %0, 1;\n\tadd.u32 %1, %0, %2;" : "+r"(lo), "=r"(a[c]) : "r"(hi));
                }
                if (KIND == K_MIX_IMAD_IADD) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(a[c]));
                }
                if (KIND == K_MIX_WIDE_IADD2) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(a[c]), "r"(b[c]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(seed));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(a[c]), "r"(seed));
                }
                if (KIND == K_WIDE_CHAIN) { uint32_t lo = (uint32_t)w[c]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(lo), "r"(b[c])); }
                if (KIND == K_WIDE_LOP) {
                    uint32_t lo = (uint32_t)w[c]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(lo), "r"(b[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                }
                if (KIND == K_WIDE_LOP2) {
                    uint32_t lo = (uint32_t)w[c]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(lo), "r"(b[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(b[c]) : "r"(a[c]), "r"(seed));
                }
                if (KIND == K_IMAD_LEA) {
                    uint32_t t;
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(a[c]), "r"(seed));
                    asm volatile("shr.u32 %0, %1, 1;\n\tadd.u32 %1, %0, %2;" : "=&r"(t), "+r"(b[c]) : "r"(seed));
                }
                if (KIND == K_IMADHI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                if (KIND == K_IADD3_3IN) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                if (KIND == K_UMIN) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                if (KIND == K_MIX3) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b[c]), "r"(seed));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(b[c]) : "r"(seed));
                    { uint32_t wl = (uint32_t)w[c]; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(wl) : "r"(a[c]), "r"(seed)); w[c] = wl; }
                }
            }
        }
    }
    uint64_t t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= a[c] ^ b[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(uint32_t *out, uint64_t *cyc, int n_sm, FILE *js, bool last) {
    const int blocks = n_sm * 4, threads = 256;   // 32 warps / SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<KIND><<<blocks, threads>>>(out, 1, cyc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        bench<KIND><<<blocks, threads>>>(out, r + 2, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    uint64_t c0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_instr_per_sm = (double)ITERS * CHAINS * ops_per_iter[KIND] * 32.0;   // 32 warps per SM
    double ipc_sm = warp_instr_per_sm / (double)c0;
    double lane_ops = (double)ITERS * CHAINS * ops_per_iter[KIND] * blocks * threads;
    double tlops = lane_ops / (best * 1e-3) / 1e12;
    printf("%-28s  %8.3f ms  %6.2f warp-instr/clk/SM  %7.2f T lane-ops/s  (clk %llu)\n", names[KIND], best, ipc_sm, tlops,
           (unsigned long long)c0);
    fprintf(js, "  \"%s\": {\"ms\": %.4f, \"warp_instr_per_clk_per_sm\": %.3f, \"tera_lane_ops_per_s\": %.3f}%s\n", names[KIND], best,
            ipc_sm, tlops, last ? "" : ",");
}

int main(int argc, char **argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int n_sm = p.multiProcessorCount;
    uint32_t *out; uint64_t *cyc;
    cudaMalloc(&out, (size_t)n_sm * 4 * 256 * 4);
    cudaMalloc(&cyc, (size_t)n_sm * 4 * 8);
    FILE *js = fopen(argc > 1 ? argv[1] : "intpipe.json", "w");
    fprintf(js, "{\n  \"gpu\": \"%s\", \"sms\": %d, \"note\": \"32 warps/SM, %d independent chains/thread\",\n", p.name, n_sm, CHAINS);
    run<K_IADD3>(out, cyc, n_sm, js, false);
    run<K_LOP3>(out, cyc, n_sm, js, false);
    run<K_SHF>(out, cyc, n_sm, js, false);
    run<K_LEAHI>(out, cyc, n_sm, js, false);
    run<K_UMIN>(out, cyc, n_sm, js, false);
    run<K_IMAD>(out, cyc, n_sm, js, false);
    run<K_IMADWIDE>(out, cyc, n_sm, js, false);
    run<K_MIX_WIDE_LEA>(out, cyc, n_sm, js, false);
    run<K_MIX_IMAD_IADD>(out, cyc, n_sm, js, false);
    run<K_MIX_WIDE_IADD2>(out, cyc, n_sm, js, false);
    run<K_MIX3>(out, cyc, n_sm, js, false);
    run<K_WIDE_CHAIN>(out, cyc, n_sm, js, false);
    run<K_WIDE_LOP>(out, cyc, n_sm, js, false);
    run<K_WIDE_LOP2>(out, cyc, n_sm, js, false);
    run<K_IMAD_LEA>(out, cyc, n_sm, js, false);
    run<K_IMADHI>(out, cyc, n_sm, js, false);
    run<K_IADD3_3IN>(out, cyc, n_sm, js, true);
    fprintf(js, "}\n");
    fclose(js);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
