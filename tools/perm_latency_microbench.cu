// Lone-warp latency of one Poseidon2 permutation (the regime of the transcript chain, the tree-walk kernels and the
// one-permutation levels of K6): a chain of dependent permutations per thread, W warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I recursive-stwo_b200/csrc -o build/perm_latency tools/perm_latency_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "poseidon2.cuh"

template <bool UNROLLED>
__global__ void chain(u32 *io, int reps) {
    u32 s[16];
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (int i = 0; i < 16; i++) s[i] = io[16 * t + i];
    for (int r = 0; r < reps; r++) {
        poseidon2::permute<UNROLLED>(s);
        s[8] ^= 0;          // keeps the sponge shape: next input depends on the whole previous output
    }
    for (int i = 0; i < 16; i++) io[16 * t + i] = s[i];
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, reps = 200;
    u32 *d;
    cudaMalloc(&d, (size_t)sms * 2048 * 64);
    cudaMemset(d, 1, (size_t)sms * 2048 * 64);
    printf("{\"sms\": %d, \"reps\": %d, \"rows\": [", sms, reps);
    bool first = true;
    for (int variant = 0; variant < 2; variant++)
        for (int warps : {1, 2, 4, 8, 16, 32}) {
            const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
            const int blocks = sms * (warps * 32 / threads);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int it = 0; it < 2; it++) {
                cudaEventRecord(e0);
                if (variant) chain<true><<<blocks, threads>>>(d, reps); else chain<false><<<blocks, threads>>>(d, reps);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s{\"unrolled\": %d, \"warps_per_sm\": %d, \"us_per_perm_chain\": %.3f, \"gperms_per_s\": %.3f}", first ? "" : ", ", variant, warps,
                   ms * 1e3 / reps, (double)sms * warps * 32 * reps / (ms * 1e-3) / 1e9);
            first = false;
        }
    printf("]}\n");
    return cudaGetLastError() != cudaSuccess;
}
