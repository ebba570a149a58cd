// Probe: HBM write bandwidth for the trace export's store pattern.  Output = n_items x 13 columns x n_rows u32, column-major per item.
// A "tile" = R rows x 32 items: the CTA's warps write, for each of the 416 (item, column) pairs, one run of R*4 bytes.  Persistent CTAs
// walk tiles row-tile fastest (like k_cs_export_vals_stream).  No loads, no compute: what the memory system does with this pattern.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int R>   // rows per tile: 32 -> 128 B runs (STG.32 x 32 lanes), 64 -> 256 B (STG.64), 128 -> 512 B (STG.128)
__global__ void __launch_bounds__(256) k_scatter(uint32_t *out, uint32_t n_rows, uint32_t n_groups) {
    const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t n_row_tiles = n_rows / R;
    const size_t n_tiles = (size_t)n_row_tiles * n_groups;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t grp = tile / n_row_tiles, row0 = (tile % n_row_tiles) * R;
        for (uint32_t pair = warp; pair < 32 * 13; pair += 8) {
            const uint32_t item = grp * 32 + pair / 13, col = pair % 13;
            uint32_t *o = out + ((size_t)item * 13 + col) * n_rows + row0;
            if (R == 32) o[lane] = lane + pair;
            else if (R == 64) reinterpret_cast<uint2 *>(o)[lane] = make_uint2(lane, pair);
            else reinterpret_cast<uint4 *>(o)[lane] = make_uint4(lane, pair, 0, 1);
        }
    }
}
int main() {
    const uint32_t n_rows = 65536, n_items = 4096, n_groups = n_items / 32;
    const size_t bytes = (size_t)n_items * 13 * n_rows * 4;
    uint32_t *out;
    cudaMalloc(&out, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas = 2; ctas <= 8; ctas *= 2)
        for (int R = 32; R <= 128; R *= 2) {
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (R == 32) k_scatter<32><<<148 * ctas, 256>>>(out, n_rows, n_groups);
                else if (R == 64) k_scatter<64><<<148 * ctas, 256>>>(out, n_rows, n_groups);
                else k_scatter<128><<<148 * ctas, 256>>>(out, n_rows, n_groups);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            printf("{\"ctas_per_sm\": %d, \"run_bytes\": %d, \"ms\": %.3f, \"write_gbs\": %.1f}\n", ctas, R * 4, best, bytes / best / 1e6);
        }
    cudaMemset(out, 0, bytes);
    cudaEventRecord(e0); cudaMemset(out, 1, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"memset_gbs\": %.1f, \"err\": \"%s\"}\n", bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
