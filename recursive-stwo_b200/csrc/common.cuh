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
// K7 check_poseidon_invocations with an occupancy knob (cs_kernels.cu): ctas_per_sm 0 = fill the SMs, k = a thin resident layer
// tape (optional) + v->perm_hint_inputs: entries covered by a complete record of executed permutations are compared with it
int32_t cs_check_poseidon_launch(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const int32_t *mult_poseidon,
                                 const uint32_t *scratch, int64_t *first_bad, cudaStream_t st, int ctas_per_sm,
                                 const stwo_b200_cs_tape *tape = nullptr);
// STWO_B200_RECORD_INPUTS=0 (profiling): the verifier records permutation outputs only and check_poseidon_invocations re-executes
inline bool record_inputs() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("STWO_B200_RECORD_INPUTS"); v = !(e && e[0] == '0'); }
    return v != 0;
}
// per-device resources of the batch drivers, released by stwo_b200_shutdown (verify_kernels.cu / circuit.cu)
void verify_pools_destroy();
void circuit_streams_destroy();
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
