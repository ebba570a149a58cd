// Shared host-side plumbing of the C ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/stwo_b200.h"

namespace stwo_b200 {
// launch bookkeeping (stwo_b200_launch_count)
void note_launch(uint64_t n = 1);
// lazily grown pinned host / device staging areas for the host-pointer entry points
int32_t stage_reserve(size_t dev_bytes);
uint8_t *stage_dev();
bool device_ready();
cudaStream_t stage_stream();
inline int32_t cuda_status(cudaError_t e) { return e == cudaSuccess ? STWO_B200_OK : -(int32_t)e; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace stwo_b200

#define STWO_CHECK_DEVICE()                                              \
    do {                                                                 \
        if (!stwo_b200::device_ready()) return STWO_B200_E_NO_DEVICE;    \
    } while (0)
#define STWO_CUDA(call)                                                  \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return -(int32_t)e__;                    \
    } while (0)
