// K6 (tape evaluation) and K7 (the O(n_rows) finalisation loops + trace export) of the Plonk-with-Poseidon constraint
// system, for batches of items that share one recorded wiring.
//   constraint_system/src/plonk_with_poseidon.rs:337-380  check_arithmetics
//   :382-466 populate_logup_arguments   :468-519 check_poseidon_invocations   :521-628 generate_plonk_with_poseidon_circuit
// Bounds: K6 is integer-issue bound (it executes the circuit's Poseidon2 permutations, one lane per item); K7 export is
// HBM bound: per (row, item) three 16-byte variable reads + thirteen 4-byte column writes = 100 B.
// Layout: per-item arrays are lane-interleaved (tape.cuh) so a warp = 32 items touches contiguous runs.
#include "common.cuh"
#include "tape.cuh"
#include <stdlib.h>
#include <string.h>

using namespace stwo_b200;

namespace {
constexpr int kT = 256;
inline unsigned nblk(size_t n, int t = kT) { return (unsigned)((n + t - 1) / t); }

struct Batch {                       // stwo_b200_cs_values + sizes, passed by value to the kernels
    tape::Q4 *vars; u32 *flow_hash; uint8_t *flow_swap;
    u32 n_batch, lanes, n_vars, n_flow;
    const u32 *hints; u32 hint_stride;
    const u32 *hint_ready; u32 hint_need;
    const u32 *hint_inputs;
    __device__ __forceinline__ tape::View view(u32 item, const u32 *input, u32 n_input_words) const {
        const size_t g = item / lanes, l = item % lanes;
        tape::View v;
        v.vars = vars + g * n_vars * lanes + l;
        v.input = input ? input + g * n_input_words * lanes + l : nullptr;
        v.flow_hash = flow_hash ? reinterpret_cast<tape::Q4 *>(flow_hash) + g * n_flow * 8 * lanes + l : nullptr;
        v.flow_swap = flow_swap ? flow_swap + g * n_flow * lanes + l : nullptr;
        v.stride = lanes;
        v.hint = hints && (!hint_ready || hint_ready[item] == hint_need) ? hints + (size_t)item * hint_stride : nullptr;
        return v;
    }
};

// ---- K6 -------------------------------------------------------------------------------------------------------------------
// One CTA per group of `lanes` items.  Thread t works for item t % lanes in instruction slot t / lanes: with lanes = 32 a
// warp executes one instruction for 32 items (uniform control flow, coalesced variable traffic) and the CTA's warps share
// the instructions of a level; a barrier separates levels.  Instruction words are warp-uniform broadcast loads.
constexpr int kEvalThreads = 1024;
// A level is a range of BUNDLES (level_start, in bundle units); bundle k is the instructions [bundle_start[k], bundle_start[k + 1]), which one
// warp executes back to back (each may read what the one before it wrote).  bundle_start == nullptr: every instruction is its own bundle.
__device__ __forceinline__ void eval_bundle(const tape::View &v, const tape::Ins *__restrict__ ins, const u32 *__restrict__ bundle_start, u32 k,
                                            const tape::Perm *__restrict__ perms, const u32 *__restrict__ eperms) {
    const u32 i0 = bundle_start ? __ldg(bundle_start + k) : k, i1 = bundle_start ? __ldg(bundle_start + k + 1) : k + 1;
    for (u32 i = i0; i < i1; i++) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(ins) + i);
        tape::Ins in; in.op = w.x; in.dst = w.y; in.a = w.z; in.b = w.w;
        tape::eval<false>(v, in, perms, eperms);
    }
}
__global__ void __launch_bounds__(kEvalThreads) k_tape_eval(const tape::Ins *__restrict__ ins, const u32 *__restrict__ level_start, u32 n_levels,
                                                            const tape::Perm *__restrict__ perms, Batch b, const u32 *input, u32 n_input_words,
                                                            const u32 *__restrict__ eperms, const u32 *__restrict__ bundle_start) {
    const u32 lane = threadIdx.x % b.lanes, slot = threadIdx.x / b.lanes, n_slots = kEvalThreads / b.lanes;
    const u32 item = blockIdx.x * b.lanes + lane;
    const bool live = item < b.n_batch;
    const tape::View v = b.view(live ? item : 0, input, n_input_words);
    if (live && slot == 0) tape::prologue(v);
    __syncthreads();
    for (u32 l = 0; l < n_levels; l++) {
        const u32 lo = __ldg(level_start + l), hi = __ldg(level_start + l + 1);
        if (live)
            for (u32 k = lo + slot; k < hi; k += n_slots) eval_bundle(v, ins, bundle_start, k, perms, eperms);
        __syncthreads();
    }
}

// Grid-wide variant for lanes = 32: the (group, instruction) pairs of a level are spread over every warp of a co-resident
// grid (one CTA per SM, cooperative launch) and a grid barrier separates levels.  A level that holds one permutation per
// group (the transcript chain) then keeps n_groups warps on n_groups different SMs busy instead of one warp per CTA, and
// a batch smaller than 32 x #SM items still uses the whole GPU.
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned n_ctas, unsigned &phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        phase += 1;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned target = phase * n_ctas;
        unsigned seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory"); } while (seen < target);
    }
    __syncthreads();
}
__device__ unsigned long long *g_level_clock = nullptr;      // diagnostics (stwo_b200_cs_eval_level_clock)
bool g_level_clock_host = false;
__device__ __forceinline__ void stamp_level(u32 l) {
    if (g_level_clock && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_level_clock[l] = t;
    }
}
// Measured and dropped (tools/level_clock.py, 4096 proofs): dealing a level's permutations to fewer warps so that the rounds come
// out even -- no gain (the permutation levels run at 3.4 G perms/s whatever the rounding); keeping the operand loads of two gates
// in flight per warp -- no gain (a 1000-gate level moves 193 MB in 39 us: the gate levels are bound by L2/HBM bandwidth at 16 B per
// variable, not by latency); walking the items backwards in the odd warps so that gates and permutations overlap -- no gain;
// taking gates in chunks from a per-level atomic counter -- same-address atomics serialise at ~2 ns.
constexpr u32 kNarrowItems = 128;     // (instructions x lane groups per CTA) up to which a level runs CTA-locally
template <bool UNROLLED>
__global__ void __launch_bounds__(kEvalThreads) k_tape_eval_grid(const tape::Ins *__restrict__ ins, const u32 *__restrict__ level_start, u32 n_levels,
                                                                 const tape::Perm *__restrict__ perms, Batch b, const u32 *input, u32 n_input_words,
                                                                 unsigned *barrier, const u32 *__restrict__ eperms, const u32 *__restrict__ bundle_start) {
    const u32 lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const u32 n_groups = (b.n_batch + 31) / 32, n_warps = gridDim.x * (kEvalThreads / 32);
    const u32 gw = warp * gridDim.x + blockIdx.x;            // consecutive work items land on different SMs
    unsigned phase = 0;
    for (u32 g = gw; g < n_groups; g += n_warps)
        if (g * 32 + lane < b.n_batch) tape::prologue(b.view(g * 32 + lane, input, n_input_words));
    grid_barrier(barrier, gridDim.x, phase);
    stamp_level(0);
    // A level with only a few instructions costs a grid barrier and a chain of dependent loads, not bandwidth.  Stretches of such
    // levels run CTA-locally: CTA c owns the lane groups c, c + gridDim.x, ... (a group's variables are then produced and consumed
    // on one SM), its warps share the level's instructions and __syncthreads() separates the levels; the grid barrier returns
    // where a wide level follows.
    const u32 groups_here = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const u32 groups_max = (n_groups + gridDim.x - 1) / gridDim.x;
    auto is_narrow = [&](u32 l) { return (__ldg(level_start + l + 1) - __ldg(level_start + l)) * groups_max <= kNarrowItems; };
    for (u32 l = 0; l < n_levels; l++) {
        const u32 lo = __ldg(level_start + l), hi = __ldg(level_start + l + 1);
        const bool narrow = is_narrow(l);
        if (narrow) {
            const u32 n_items = (hi - lo) * groups_here;
            for (u32 t = warp; t < n_items; t += kEvalThreads / 32) {
                const u32 k = lo + t / groups_here, item = (blockIdx.x + (t % groups_here) * gridDim.x) * 32 + lane;
                if (item < b.n_batch) eval_bundle(b.view(item, input, n_input_words), ins, bundle_start, k, perms, eperms);
            }
        } else {
            const u32 n_items = (hi - lo) * n_groups;
            for (u32 t = gw; t < n_items; t += n_warps) {
                // group fastest: a warp's successive items (stride n_warps) walk through different instructions of the level, so
                // permutations and cheap gates mix evenly; the groups of a one-instruction level land on different SMs
                const u32 k = lo + t / n_groups, item = (t % n_groups) * 32 + lane;
                if (item < b.n_batch) eval_bundle(b.view(item, input, n_input_words), ins, bundle_start, k, perms, eperms);
            }
        }
        if (narrow && l + 1 < n_levels && is_narrow(l + 1)) __syncthreads();
        else grid_barrier(barrier, gridDim.x, phase);
        stamp_level(l + 1);
    }
}

// Cluster variant for lanes = 32: ONE THREAD-BLOCK CLUSTER per lane group (32 items).  The cluster's warps share the instructions of a
// level and a cluster barrier (barrier.cluster, ~0.2 us, hardware co-scheduled CTAs) separates levels -- an order of magnitude
// cheaper than the grid barrier above (~3 us x 265 levels), and a plain launch: no cooperative grid, so the evaluations of several
// small batches (the shape groups of a mixed batch) run beside each other on different streams.  A lane group's variables[] are
// produced and consumed by its own cluster only, while it is resident: the operand reads of a level find in L2 what the levels
// before it wrote (the grid-wide form sweeps every group's variables at every level).
// Consecutive instructions of a level go to different CTAs of the cluster (different SMs).
constexpr int kClusterThreads = 512;
// Measured and dropped (round 2): a bundle loop that keeps a one-variable instruction's result in registers for its successor and requests
// the successor's other operand before computing (the chain link then waits for no load): 1.27 against 1.01 ms at 512 proofs under the
// 64-register cap and 1.82 ms with 120 registers and one CTA per SM -- the extra copy of the arithmetic cases and the lost co-residency
// cost more than the round trips saved; hoisting the hint loads of a permutation above its flow stores: spills at 64 registers, slower.
// Measured and dropped: four instructions of a level per warp at once (all operand loads issued before any is consumed): 110 registers,
// one CTA per SM, 5.2 ms against 3.1 ms at 4096 proofs and 1.4 against 1.1 ms at 512 -- the pass is not short of loads in flight.
// CLOCK: cluster 0 stamps the global timer after every level (tools/level_clock.py)
// Second order (ins2 != nullptr): the tape with every recorded permutation split into "outputs from the record" and "flow entry"
// (dsl::RecordedCircuit::recorded): far fewer levels, valid for a lane group whose 32 items all have a complete permutation record.
// Every warp of the cluster holds the same 32 items, so every warp reaches the same verdict without talking to the others.
template <bool CLOCK>
__global__ void __launch_bounds__(kClusterThreads) k_tape_eval_cluster(const tape::Ins *__restrict__ ins, const u32 *__restrict__ level_start, u32 n_levels,
                                                                       const tape::Perm *__restrict__ perms, Batch b, const u32 *input, u32 n_input_words,
                                                                       const u32 *__restrict__ eperms, const u32 *__restrict__ bundle_start,
                                                                       const tape::Ins *__restrict__ ins2, const u32 *__restrict__ level_start2, u32 n_levels2,
                                                                       const u32 *__restrict__ bundle_start2) {
    extern __shared__ u32 s_level[];                     // level_start (bundle units), n_levels + 1 words: nothing of the level loop waits on it
    u32 rank, csize, grp;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(grp));
    const u32 lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const u32 wi = rank + csize * warp, n_w = csize * (kClusterThreads / 32);
    const u32 item = grp * 32 + lane;
    const bool live = item < b.n_batch;
    const tape::View v = b.view(live ? item : 0, input, n_input_words);
    if (ins2 && __all_sync(0xffffffffu, !live || v.hint != nullptr)) {
        ins = ins2; level_start = level_start2; n_levels = n_levels2; bundle_start = bundle_start2;
    }
    auto cluster_sync = [] {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    };
    for (u32 l = threadIdx.x; l <= n_levels; l += kClusterThreads) s_level[l] = __ldg(level_start + l);
    if (live && wi == 0) tape::prologue(v);
    __syncthreads();
    const uint4 *iw = reinterpret_cast<const uint4 *>(ins);
    // The extent of this warp's first bundle of a level is fetched BEFORE the barrier that ends the level before it: one L2 round trip
    // less per level.  Measured and dropped: prefetch.global.L2 of the next instruction's operands two instructions ahead (3.12 vs
    // 3.04 ms at 4096 proofs: with every lane group resident the pass already keeps ~7 MB of loads in flight).
    auto extent = [&](u32 k, u32 &i0, u32 &i1) {
        if (bundle_start) { i0 = __ldg(bundle_start + k); i1 = __ldg(bundle_start + k + 1); } else { i0 = k; i1 = k + 1; }
    };
    u32 i0 = 0, i1 = 0;
    if (s_level[0] + wi < s_level[1]) extent(s_level[0] + wi, i0, i1);
    cluster_sync();
    if (CLOCK) stamp_level(0);
    for (u32 l = 0; l < n_levels; l++) {
        const u32 lo = s_level[l], hi = s_level[l + 1];
        for (u32 k = lo + wi; k < hi; k += n_w) {
            if (k != lo + wi) extent(k, i0, i1);
            if (live)
                for (u32 i = i0; i < i1; i++) {
                    const uint4 w = __ldg(iw + i);
                    tape::Ins in; in.op = w.x; in.dst = w.y; in.a = w.z; in.b = w.w;
                    tape::eval<false>(v, in, perms, eperms);
                }
        }
        if (l + 1 < n_levels && hi + wi < s_level[l + 2]) extent(hi + wi, i0, i1);
        cluster_sync();
        if (CLOCK) stamp_level(l + 1);
    }
}

// ---- K7: check_arithmetics ---------------------------------------------------------------------------------------------------
// thread = (row, item) with the item fastest: a warp checks one row for 32 items
__global__ void __launch_bounds__(kT) k_cs_check_arith(stwo_b200_cs_wiring w, Batch b, unsigned long long *first_bad) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    const u32 padded = (b.n_batch + b.lanes - 1) / b.lanes * b.lanes;
    if (g >= (size_t)w.n_rows * padded) return;
    // order: group-major, then row, then lane -> consecutive threads share a row and a group
    const u32 lane = (u32)(g % b.lanes);
    const size_t t = g / b.lanes;
    const u32 row = (u32)(t % w.n_rows), grp = (u32)(t / w.n_rows), item = grp * b.lanes + lane;
    if (item >= b.n_batch) return;
    const tape::View v = b.view(item, nullptr, 0);
    const bool follows = w.op_follows_c && w.op_follows_c[row];
    bool ok;
    if (w.kind == 1) {
        const qm31_t vc = tape::ldv(v, __ldg(w.c_wire + row));
        ok = tape::gate_ok_without(tape::ldv(v, __ldg(w.a_wire + row)), tape::ldv(v, __ldg(w.b_wire + row)), vc, follows ? vc.v[0] : __ldg(w.op + row),
                                   __ldg(w.op2 + row), __ldg(w.op3 + row), __ldg(w.op4 + row));
    } else
        ok = tape::row_ok(v, __ldg(w.a_wire + row), __ldg(w.b_wire + row), __ldg(w.c_wire + row), __ldg(w.op + row), __ldg(w.enforce_c_m31 + row), follows);
    if (!ok) atomicMin(first_bad + item, (unsigned long long)row);
}

// ---- K7: populate_logup_arguments (wiring only) ------------------------------------------------------------------------------
// scratch layout: counts[n_vars] | first_key[n_vars] | first_prow[n_vars] | mult_poseidon_vars[n_vars]
__global__ void __launch_bounds__(kT) k_cs_count(stwo_b200_cs_wiring w, u32 *counts, u32 *first_key, u32 *first_prow, u32 *mpv) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    if (g < w.n_rows) {
        const u32 i = (u32)g;
        const u32 a = w.a_wire[i], b = w.b_wire[i], c = w.c_wire[i];
        atomicAdd(counts + a, 1u); atomicAdd(counts + b, 1u); atomicAdd(counts + c, 1u);
        if (w.kind == 1) atomicMin(first_key + c, i);             // plonk_without_poseidon.rs:617-628: first occurrence among the c wires
        else {
            atomicMin(first_key + a, 3 * i); atomicMin(first_key + b, 3 * i + 1); atomicMin(first_key + c, 3 * i + 2);
            atomicMin(first_prow + w.poseidon_wire[i], i);
        }
    }
    if (g < w.num_input) atomicAdd(counts + g + 1, 1u);
    if (g < w.n_flow) {
        atomicAdd(counts + w.flow_swap_addr[g], 1u);
        for (int k = 0; k < 4; k++) { const u32 v = w.flow_wire[4 * g + k]; if (v) atomicAdd(mpv + v, 1u); }
    }
}
__global__ void __launch_bounds__(kT) k_cs_mult(stwo_b200_cs_wiring w, const u32 *counts, const u32 *first_key, const u32 *first_prow,
                                                const u32 *mpv, int32_t *mult_a, int32_t *mult_b, int32_t *mult_c,
                                                int32_t *mult_poseidon, u32 *status) {
    const u32 i = blockIdx.x * kT + threadIdx.x;
    if (i >= w.n_rows) return;
    if (w.kind == 1) { mult_c[i] = first_key[w.c_wire[i]] == i ? 1 - (int32_t)counts[w.c_wire[i]] : 1; return; }
    const u32 a = w.a_wire[i], b = w.b_wire[i], c = w.c_wire[i], pw = w.poseidon_wire[i];
    // first occurrence in the row-major scan a_0,b_0,c_0,a_1,... gets 1 - count, later ones 1.  A wire repeated inside
    // one row (a == b) is a first occurrence only in its earliest slot, which is what the minimum key encodes.
    mult_a[i] = first_key[a] == 3 * i ? 1 - (int32_t)counts[a] : 1;
    mult_b[i] = first_key[b] == 3 * i + 1 ? 1 - (int32_t)counts[b] : 1;
    mult_c[i] = first_key[c] == 3 * i + 2 ? 1 - (int32_t)counts[c] : 1;
    int32_t mp = 0;
    if (pw != 0 && first_prow[pw] == i && mpv[pw] != 0) {
        mp = (int32_t)mpv[pw];
        if (counts[pw] != 1) atomicOr(status, 1u);      // the reference asserts counts[poseidon_wire] == 1
    }
    mult_poseidon[i] = mp;
}

// ---- K7: check_poseidon_invocations ----------------------------------------------------------------------------------------
// thread = (flow entry, item), item fastest. Persistent: the grid is a fixed number of CTAs per SM looping over the flow, so that it
// can run as a thin, fully resident layer (2-3 CTAs per SM) under the HBM-bound export on another stream, or fill the SMs alone.
// RECORDED: a flow entry whose permutation the native verifier executed (perms[e].hint names its record slot) is compared with that
// execution -- entry input == recorded input, entry output == recorded output -- instead of being executed once more; the record holds
// out = permute(in) by construction.  The slot is a property of the entry (warp-uniform); an item whose record is incomplete, and an
// entry without a slot (none in the verifier circuits), is re-executed.
template <bool RECORDED>
__global__ void __launch_bounds__(128) k_cs_check_poseidon(stwo_b200_cs_wiring w, Batch b, const int32_t *mult_poseidon,
                                                           const u32 *first_prow, unsigned long long *first_bad, const tape::Perm *__restrict__ perms) {
    const u32 padded = (b.n_batch + b.lanes - 1) / b.lanes * b.lanes;
    const size_t total = (size_t)w.n_flow * padded;
    for (size_t g = blockIdx.x * (size_t)128 + threadIdx.x; g < total; g += (size_t)gridDim.x * 128) {
        const u32 lane = (u32)(g % b.lanes);
        const size_t t = g / b.lanes;
        const u32 e = (u32)(t % w.n_flow), grp = (u32)(t / w.n_flow), item = grp * b.lanes + lane;
        if (item >= b.n_batch) continue;
        const tape::View v = b.view(item, nullptr, 0);
        const tape::Q4 *h = v.flow_hash + (size_t)e * 8 * v.stride;
        u32 hh[32];
#pragma unroll
        for (int q = 0; q < 8; q++) { const tape::Q4 t = h[(size_t)q * v.stride]; hh[4 * q] = t.x; hh[4 * q + 1] = t.y; hh[4 * q + 2] = t.z; hh[4 * q + 3] = t.w; }
        bool ok = true;
        for (int k = 0; k < 4; k++) {
            const u32 wire = w.flow_wire[4 * e + k];
            if (!wire) continue;
            const u32 row = first_prow[wire];
            if (row >= w.n_rows || mult_poseidon[row] == 0) { ok = false; continue; }    // map.get(..).unwrap() would panic
            const qm31_t l = tape::ldv(v, w.a_wire[row]), r = tape::ldv(v, w.b_wire[row]);
            for (int j = 0; j < 4; j++) ok &= (l.v[j] == hh[8 * k + j]) & (r.v[j] == hh[8 * k + 4 + j]);
        }
        const bool swap = v.flow_swap[(size_t)e * v.stride] != 0;
        const u32 slot = RECORDED ? __ldg(&perms[e].hint) : 0u;
        if (RECORDED && slot && v.hint) {
            // v.hint is set only for an item whose record is complete (Batch::view); the input record runs parallel to the output record
            const uint4 *ro = reinterpret_cast<const uint4 *>(v.hint + (size_t)(slot - 1) * 16);
            const uint4 *ri = reinterpret_cast<const uint4 *>(v.hint + (b.hint_inputs - b.hints) + (size_t)(slot - 1) * 16);
            u32 d = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 o4 = ro[q], i4 = ri[q];
                // the permutation's input is the two halves in executed order: swapped when the swap bit is set
                const int lo = 4 * q, sw = (4 * q + 8) & 15;
                d |= (o4.x ^ hh[16 + lo]) | (o4.y ^ hh[17 + lo]) | (o4.z ^ hh[18 + lo]) | (o4.w ^ hh[19 + lo]);
                d |= (i4.x ^ (swap ? hh[sw] : hh[lo])) | (i4.y ^ (swap ? hh[sw + 1] : hh[lo + 1])) | (i4.z ^ (swap ? hh[sw + 2] : hh[lo + 2])) |
                     (i4.w ^ (swap ? hh[sw + 3] : hh[lo + 3]));
            }
            ok &= d == 0;
        } else {
            u32 st[16];
#pragma unroll
            for (int j = 0; j < 8; j++) { st[j] = swap ? hh[8 + j] : hh[j]; st[8 + j] = swap ? hh[j] : hh[8 + j]; }    // selects: hh stays in registers
            poseidon2::permute<false>(st);
#pragma unroll
            for (int j = 0; j < 16; j++) ok &= st[j] == hh[16 + j];
        }
        if (!ok) atomicMin(first_bad + item, (unsigned long long)e);
    }
}

// ---- K7: trace export ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 m31_of_i32(int32_t v) { return v < 0 ? M31_P - (u32)(-v) : (u32)v; }

__global__ void __launch_bounds__(kT) k_cs_export_pre(stwo_b200_cs_wiring w, const int32_t *mult_a, const int32_t *mult_b,
                                                      const int32_t *mult_c, const int32_t *mult_poseidon, u32 *pre) {
    const u32 i = blockIdx.x * kT + threadIdx.x;
    if (i >= w.n_rows) return;
    const size_t n = w.n_rows;
    if (w.kind == 1) {
        pre[0 * n + i] = m31_of_i32(mult_c[i]); pre[1 * n + i] = w.a_wire[i]; pre[2 * n + i] = w.b_wire[i]; pre[3 * n + i] = w.c_wire[i];
        pre[4 * n + i] = w.op[i]; pre[5 * n + i] = w.op2[i]; pre[6 * n + i] = w.op3[i]; pre[7 * n + i] = w.op4[i];
        return;
    }
    pre[0 * n + i] = m31_of_i32(mult_a[i]); pre[1 * n + i] = m31_of_i32(mult_b[i]); pre[2 * n + i] = m31_of_i32(mult_c[i]);
    pre[3 * n + i] = w.poseidon_wire[i]; pre[4 * n + i] = (u32)mult_poseidon[i]; pre[5 * n + i] = w.enforce_c_m31[i];
    pre[6 * n + i] = w.a_wire[i]; pre[7 * n + i] = w.b_wire[i]; pre[8 * n + i] = w.c_wire[i]; pre[9 * n + i] = w.op[i];
}

constexpr int kCols = 13;            // a_val_0..3, b_val_0..3, c_val_0..3, op
// lanes = 1: thread per (item, row); random 16-byte gathers, coalesced column stores
__global__ void __launch_bounds__(kT) k_cs_export_vals_plain(stwo_b200_cs_wiring w, Batch b, u32 *vals, unsigned long long *first_bad) {
    const size_t g = blockIdx.x * (size_t)kT + threadIdx.x;
    if (g >= (size_t)w.n_rows * b.n_batch) return;
    const u32 item = (u32)(g / w.n_rows), i = (u32)(g % w.n_rows);
    const tape::View v = b.view(item, nullptr, 0);
    const size_t n = w.n_rows;
    u32 *o = vals + (size_t)item * kCols * n + i;
    const qm31_t a = tape::ldv(v, w.a_wire[i]), bb = tape::ldv(v, w.b_wire[i]), c = tape::ldv(v, w.c_wire[i]);
#pragma unroll
    for (int k = 0; k < 4; k++) { o[k * n] = a.v[k]; o[(4 + k) * n] = bb.v[k]; o[(8 + k) * n] = c.v[k]; }
    const u32 op = (w.op_follows_c && w.op_follows_c[i]) ? c.v[0] : w.op[i];
    o[12 * n] = op;
    if (first_bad) {
        const bool ok = w.kind == 1 ? tape::gate_ok_without(a, bb, c, op, w.op2[i], w.op3[i], w.op4[i]) : tape::gate_ok(a, bb, c, op, w.enforce_c_m31[i]);
        if (!ok) atomicMin(first_bad + item, (unsigned long long)i);
    }
}
// lanes = 32: a CTA transposes a tile of 32 rows x 32 items through shared memory.  Load phase: a warp reads one row's
// three variables for 32 items (3 x 512 contiguous bytes).  Store phase: a warp writes 32 consecutive rows of one
// (item, column) = one 128-byte line.
constexpr int kTileRows = 32;
// first_bad != nullptr fuses check_arithmetics into the pass (the three variables of the row are in registers anyway).
// ITEMS = batch items per tile (32; 16 as an experiment: half the shared memory per CTA and 6 CTAs per SM instead of 4, a warp
// reads two rows at a time, 256 B per row -- slower, 5.15 vs 4.66 ms: the shorter runs per row cost more than the occupancy buys).
template <int ITEMS>
__global__ void __launch_bounds__(kT) k_cs_export_vals_tiled(stwo_b200_cs_wiring w, Batch b, u32 *vals, unsigned long long *first_bad) {
    extern __shared__ u32 tile[];                        // [kCols][ITEMS items][33]
    constexpr u32 ROWS_PER_WARP = 32 / ITEMS;
    const u32 warp = threadIdx.x / 32, lane = threadIdx.x % 32, n_warps = kT / 32;
    const u32 il = lane % ITEMS, rsub = lane / ITEMS;
    const u32 row0 = blockIdx.x * kTileRows, grp = blockIdx.y;
    const u32 item = grp * ITEMS + il;
    if (item < b.n_batch) {
        const tape::View v = b.view(item, nullptr, 0);
        // two rows per trip: the six wire loads, then the six 16-byte variable gathers, are in flight together (the pass is bound
        // by the latency of these dependent loads, and the tile's shared memory, not registers, limits the CTAs per SM)
        constexpr u32 STEP = (kT / 32) * ROWS_PER_WARP;
        static_assert(kTileRows % (2 * STEP) == 0, "rows of a tile come in pairs per warp");
        for (u32 r = warp * ROWS_PER_WARP + rsub; r < kTileRows; r += 2 * STEP) {
            const u32 rr[2] = {r, r + STEP};
            u32 wa[2], wb[2], wc[2];
#pragma unroll
            for (int j = 0; j < 2; j++) { wa[j] = __ldg(w.a_wire + row0 + rr[j]); wb[j] = __ldg(w.b_wire + row0 + rr[j]); wc[j] = __ldg(w.c_wire + row0 + rr[j]); }
            qm31_t va[2], vb[2], vc[2];
#pragma unroll
            for (int j = 0; j < 2; j++) { va[j] = tape::ldv(v, wa[j]); vb[j] = tape::ldv(v, wb[j]); vc[j] = tape::ldv(v, wc[j]); }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const u32 i = row0 + rr[j], r_ = rr[j];
                const qm31_t a = va[j], bb = vb[j], c = vc[j];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    tile[((0 + k) * ITEMS + il) * 33 + r_] = a.v[k];
                    tile[((4 + k) * ITEMS + il) * 33 + r_] = bb.v[k];
                    tile[((8 + k) * ITEMS + il) * 33 + r_] = c.v[k];
                }
                const u32 op = (w.op_follows_c && w.op_follows_c[i]) ? c.v[0] : __ldg(w.op + i);
                tile[(12 * ITEMS + il) * 33 + r_] = op;
                if (first_bad) {
                    const bool ok = w.kind == 1 ? tape::gate_ok_without(a, bb, c, op, __ldg(w.op2 + i), __ldg(w.op3 + i), __ldg(w.op4 + i))
                                                : tape::gate_ok(a, bb, c, op, __ldg(w.enforce_c_m31 + i));
                    if (!ok) atomicMin(first_bad + item, (unsigned long long)i);
                }
            }
        }
    }
    __syncthreads();
    const size_t n = w.n_rows;
    for (u32 pair = warp; pair < ITEMS * kCols; pair += n_warps) {
        const u32 it = pair / kCols, col = pair % kCols;
        if (grp * ITEMS + it < b.n_batch) vals[((size_t)(grp * ITEMS + it) * kCols + col) * n + row0 + lane] = tile[(col * ITEMS + it) * 33 + lane];
    }
}
// lanes = 32, streaming form.  A persistent CTA walks tiles of 64 rows x 16 items with two staging buffers.
//
// Why 64 rows: the pass writes 13.9 GB (4096 proofs) as one run per (item, column, tile).  Measured on this B200
// (tools/scatter_write_probe.cu, stores only): 128-byte runs reach 4.2-4.4 TB/s, 256-byte runs 5.9 TB/s (cudaMemset: 7.4) -- with
// 32-row tiles the stores alone take 3.3 ms, whatever the kernel around them does.  A lane owns two consecutive rows and every
// store instruction writes 256 contiguous bytes of one (item, column).
// Why a variable list per tile: the HOST lists, per 32 rows, the DISTINCT variables the 96 wires name (stwo_b200_cs_export_tiles_build:
// 1.16 per row for the verifier circuit of shape S -- padding rows, the constants 0/1/i/j and gate chains repeat wires), so a
// variable crosses L2 -> SM once per half tile, not once per use (5 GB instead of 12.9 GB per 4096 proofs).
//   issue(t+1): the warps gather the two half tiles' variables with 16-byte asynchronous copies (LDGSTS; a half-warp moves the 256
//               contiguous bytes a variable occupies across the tile's 16 items) straight into shared memory -- no registers are
//               held, the whole tile is in flight while
//   consume(t): check_arithmetics, fused: a warp takes a row, 16 items x the two CM31 halves of the gate (tape::gate_ok_half), so the
//               gate kind is uniform across the warp; then the column stores: a warp takes 2 items, lane = row pair reads its
//               rows' QM31 by their slot in the staging buffer (LDS.128) and writes 13 x 8 bytes.
// Everything a tile needs from the wiring (variable lists, slots, row constants) is fetched three tiles ahead by asynchronous
// copies as well: nothing a later trip depends on is held in a register across trips.  Tiles are dealt row-tile fastest: the
// CTAs of the grid sweep one lane group's variables[] together.
constexpr int kXRows = 64, kXItems = 16, kXPitch = kXItems + 1, kXThreads = 512, kXWarps = kXThreads / 32;
// per 64-row tile (stwo_b200_cs_export_tiles_build), 384 words: two 32-row variable lists of 128 words ([0] count, [1..96] variables),
// then 64 row records of 2 words: {slot_a | slot_b << 8 | slot_c << 16 | flags << 24, op}
constexpr int kXListWords = 128, kXTileWords = 2 * kXListWords + 2 * kXRows, kXMetaStages = 4;
enum : u32 { XF_C1 = 1u << 24, XF_OP3 = 1u << 25, XF_OP4 = 1u << 26, XF_FOLLOWS = 1u << 27, XF_SKIP = 1u << 28 };   // C1: enforce_c_m31 / op2
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__global__ void __launch_bounds__(kXThreads, 2) k_cs_export_vals_stream(stwo_b200_cs_wiring w, Batch b, u32 *vals, unsigned long long *first_bad,
                                                                         u32 n_row_tiles, u32 n_tiles) {
    extern __shared__ uint4 xsmem[];
    u32 *xmeta = reinterpret_cast<u32 *>(xsmem);                                 // [kXMetaStages][kXTileWords]
    uint4 *xstage = xsmem + kXMetaStages * kXTileWords / 4;                      // [2 buffers][2 halves][cap][kXPitch]
    const u32 warp = threadIdx.x / 32, lane = threadIdx.x % 32, cap = w.export_cap;
    const u32 il16 = lane & 15u, part = lane >> 4;
    const size_t n2 = w.n_rows / 2;
    const u32 step = gridDim.x;
    const size_t buf_stride = (size_t)2 * cap * kXPitch;
    auto fetch_meta = [&](u32 rt, u32 stage) {           // rt: 64-row tile; 96 x 16 bytes
        if (warp < 3) cp_async16(xmeta + stage * kXTileWords + 128 * warp + 4 * lane, w.export_tiles + (size_t)rt * kXTileWords + 128 * warp + 4 * lane);
    };
    // tile = (lane group, half of its items, 64-row tile), the row tile fastest
    auto issue = [&](u32 gh, u32 stage, u32 buf) {
        const u32 *m0 = xmeta + stage * kXTileWords, *m1 = m0 + kXListWords;
        const u32 count0 = m0[0], total = count0 + m1[0];
        const u32 item = gh * kXItems + il16;
        const uint4 *src = reinterpret_cast<const uint4 *>(b.vars) + (size_t)(gh >> 1) * b.n_vars * 32 + (gh & 1u) * kXItems + il16;
        uint4 *dst = xstage + buf * buf_stride + il16;
        if (item < b.n_batch)
            for (u32 e = 2 * warp + part; e < total; e += 2 * kXWarps) {
                const bool second = e >= count0;
                const u32 k = second ? e - count0 : e, var = (second ? m1 : m0)[1 + k];
                cp_async16(dst + ((second ? cap : 0u) + k) * kXPitch, src + (size_t)var * 32);
            }
    };
    auto advance = [&](u32 &gh, u32 &rt) { rt += step; while (rt >= n_row_tiles) { rt -= n_row_tiles; gh++; } };
    if (blockIdx.x >= n_tiles) return;
    u32 gh = blockIdx.x / n_row_tiles, rt = blockIdx.x % n_row_tiles;            // current tile
    const u32 n_mine = (n_tiles - blockIdx.x + step - 1) / step;
    u32 mg = gh, mrt = rt;                                                        // tile whose meta is fetched next
    for (u32 k = 0; k < 3; k++) {
        if (k < n_mine) fetch_meta(mrt, k);
        cp_async_commit();
        advance(mg, mrt);
    }
    asm volatile("cp.async.wait_group 2;" ::: "memory");                         // meta 0
    __syncthreads();
    issue(gh, 0, 0);
    cp_async_commit();
    u32 ng = gh, nrt = rt;                                                        // tile t + 1
    advance(ng, nrt);
    for (u32 t = 0; t < n_mine; t++) {
        const u32 buf = t & 1u;
        // groups committed so far, oldest first: .. meta(t+1) data(t-1) meta(t+2) data(t); this trip adds meta(t+3), data(t+1)
        asm volatile("cp.async.wait_group 2;" ::: "memory");                     // meta(t+1) landed (copied by other threads: barrier)
        __syncthreads();                                                          // ... and every warp is done with tile t-1: its staging
                                                                                  // buffer and its meta stage (= that of t+3) are free
        if (t + 3 < n_mine) fetch_meta(mrt, (t + 3) % kXMetaStages);
        cp_async_commit();
        advance(mg, mrt);
        if (t + 1 < n_mine) issue(ng, (t + 1) % kXMetaStages, buf ^ 1u);
        cp_async_commit();
        asm volatile("cp.async.wait_group 2;" ::: "memory");                     // data(t) landed; meta(t+3), data(t+1) may be in flight
        __syncthreads();
        const uint2 *rec = reinterpret_cast<const uint2 *>(xmeta + (t % kXMetaStages) * kXTileWords + 2 * kXListWords);
        const uint4 *st = xstage + buf * buf_stride;
        if (first_bad) {
            const u32 item = gh * kXItems + il16;
            if (item < b.n_batch) {
#pragma unroll 2
                for (u32 j = 0; j < kXRows / kXWarps; j++) {
                    const u32 r = warp + kXWarps * j;
                    const uint2 rr = rec[r];                                      // warp-uniform
                    if (rr.x & XF_SKIP) continue;                                 // same wires and constants as an earlier row of this tile
                    const uint4 *sh = st + (r >> 5) * cap * kXPitch + il16;
                    const uint4 a = sh[(rr.x & 255u) * kXPitch], bb = sh[((rr.x >> 8) & 255u) * kXPitch], c = sh[((rr.x >> 16) & 255u) * kXPitch];
                    const qm31_t qa = qm31::mk(a.x, a.y, a.z, a.w), qb = qm31::mk(bb.x, bb.y, bb.z, bb.w), qc = qm31::mk(c.x, c.y, c.z, c.w);
                    const u32 op = (rr.x & XF_FOLLOWS) ? c.x : rr.y;
                    bool ok;
                    if (w.kind == 1) ok = part || tape::gate_ok_without(qa, qb, qc, op, (rr.x >> 24) & 1u, (rr.x >> 25) & 1u, (rr.x >> 26) & 1u);
                    else ok = tape::gate_ok_half(qa, qb, qc, op, rr.x & XF_C1, part);
                    if (!ok) atomicMin(first_bad + item, (unsigned long long)(rt * kXRows + r));
                }
            }
        }
        {
            // lane = rows 2 lane, 2 lane + 1 of the tile (lanes 0..15: first half tile)
            const uint4 r2 = reinterpret_cast<const uint4 *>(rec)[lane];         // the two row records
            const uint4 *sh = st + part * cap * kXPitch;
            const uint4 *sa0 = sh + (r2.x & 255u) * kXPitch, *sb0 = sh + ((r2.x >> 8) & 255u) * kXPitch, *sc0 = sh + ((r2.x >> 16) & 255u) * kXPitch;
            const uint4 *sa1 = sh + (r2.z & 255u) * kXPitch, *sb1 = sh + ((r2.z >> 8) & 255u) * kXPitch, *sc1 = sh + ((r2.z >> 16) & 255u) * kXPitch;
#pragma unroll
            for (u32 j = 0; j < kXItems / kXWarps; j++) {
                const u32 il = warp * (kXItems / kXWarps) + j, item = gh * kXItems + il;
                if (item >= b.n_batch) continue;
                const uint4 a0 = sa0[il], a1 = sa1[il], b0 = sb0[il], b1 = sb1[il], c0 = sc0[il], c1 = sc1[il];
                uint2 *o = reinterpret_cast<uint2 *>(vals + (size_t)item * kCols * 2 * n2 + (size_t)rt * kXRows) + lane;
                o[0] = make_uint2(a0.x, a1.x); o[n2] = make_uint2(a0.y, a1.y); o[2 * n2] = make_uint2(a0.z, a1.z); o[3 * n2] = make_uint2(a0.w, a1.w);
                o[4 * n2] = make_uint2(b0.x, b1.x); o[5 * n2] = make_uint2(b0.y, b1.y); o[6 * n2] = make_uint2(b0.z, b1.z); o[7 * n2] = make_uint2(b0.w, b1.w);
                o[8 * n2] = make_uint2(c0.x, c1.x); o[9 * n2] = make_uint2(c0.y, c1.y); o[10 * n2] = make_uint2(c0.z, c1.z); o[11 * n2] = make_uint2(c0.w, c1.w);
                o[12 * n2] = make_uint2((r2.x & XF_FOLLOWS) ? c0.x : r2.y, (r2.z & XF_FOLLOWS) ? c1.x : r2.w);
            }
        }
        gh = ng; rt = nrt;
        advance(ng, nrt);
        // no trailing barrier: the next trip's first barrier orders this trip's reads before anything is overwritten
    }
}
// ---- K7: PoseidonFlow export ---------------------------------------------------------------------------------------------------
// The flow the tape evaluation recorded (lane-interleaved 16-byte elements: [group][entry][8 quads][32 lanes][4 words]) as plain per-item arrays
// hash[item][entry][32], swap[item][entry], padded to n_pad entries the way pad() does (plonk_with_poseidon.rs:296-321): entries
// (wire 0, C1), (0, C1), (0, C2), (0, C3), swap = false.  A warp transposes one (group, entry) tile through shared memory: reads
// are 128-byte rows across the lanes, writes the 128 contiguous bytes of one item's entry.
__global__ void __launch_bounds__(256) k_cs_export_flow(Batch b, u32 n_pad, const u32 *__restrict__ pad_hash /* 32 words */, u32 *hash_out, uint8_t *swap_out) {
    __shared__ u32 tile[8][32][33];
    const u32 warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const u32 n_groups = (b.n_batch + 31) / 32;
    const size_t n_tiles = (size_t)n_groups * n_pad;
    for (size_t t = blockIdx.x * (size_t)8 + warp; t < n_tiles; t += (size_t)gridDim.x * 8) {
        const u32 grp = (u32)(t / n_pad), e = (u32)(t % n_pad);
        const u32 item = grp * 32 + lane;
        if (e < b.n_flow) {
            const tape::Q4 *src = reinterpret_cast<const tape::Q4 *>(b.flow_hash) + ((size_t)grp * b.n_flow + e) * 8 * 32;
#pragma unroll
            for (u32 q = 0; q < 8; q++) {
                const tape::Q4 t = src[q * 32 + lane];
                tile[warp][4 * q][lane] = t.x; tile[warp][4 * q + 1][lane] = t.y; tile[warp][4 * q + 2][lane] = t.z; tile[warp][4 * q + 3][lane] = t.w;
            }
            __syncwarp();
            for (u32 it = 0; it < 32; it++)
                if (grp * 32 + it < b.n_batch) hash_out[((size_t)(grp * 32 + it) * n_pad + e) * 32 + lane] = tile[warp][lane][it];
            __syncwarp();
            if (item < b.n_batch) swap_out[(size_t)item * n_pad + e] = b.flow_swap[((size_t)grp * b.n_flow + e) * 32 + lane];
        } else {
            const u32 v = __ldg(pad_hash + lane);
            for (u32 it = 0; it < 32; it++)
                if (grp * 32 + it < b.n_batch) hash_out[((size_t)(grp * 32 + it) * n_pad + e) * 32 + lane] = v;
            if (item < b.n_batch) swap_out[(size_t)item * n_pad + e] = 0;
        }
    }
}
__global__ void k_fill64(unsigned long long *p, size_t n, unsigned long long v) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
bool wiring_ok(const stwo_b200_cs_wiring *w) {
    if (!w || w->n_rows < 16 || (w->n_rows & (w->n_rows - 1)) || w->n_vars < 4 || !w->a_wire || !w->b_wire || !w->c_wire || !w->op || w->kind > 1) return false;
    if (w->kind == 1) return w->op2 && w->op3 && w->op4 && w->n_flow == 0;
    return w->poseidon_wire && w->enforce_c_m31 && (w->n_flow == 0 || (w->flow_wire && w->flow_swap_addr));
}
bool values_ok(const stwo_b200_cs_values *v) { return v && v->n_batch && (v->lanes == 1 || v->lanes == 32) && v->variables; }
Batch batch_of(const stwo_b200_cs_values *v, u32 n_vars, u32 n_flow) {
    Batch b;
    b.vars = reinterpret_cast<tape::Q4 *>(v->variables); b.flow_hash = v->flow_hash; b.flow_swap = v->flow_swap;
    b.n_batch = v->n_batch; b.lanes = v->lanes; b.n_vars = n_vars; b.n_flow = n_flow;
    b.hints = v->perm_hints; b.hint_stride = v->perm_hint_stride;
    b.hint_ready = v->perm_hint_ready; b.hint_need = v->perm_hint_need;
    b.hint_inputs = v->perm_hint_inputs;
    return b;
}
}  // namespace

static_assert(sizeof(tape::Perm) == 48 && sizeof(tape::Ins) == 16, "tape records mirror the C ABI");
static_assert(tape::T_PERM_OUT == STWO_B200_T_PERM_OUT && tape::T_PERM_FLOW == STWO_B200_T_PERM_FLOW, "opcodes mirror the C ABI");
static_assert(tape::T_EPOSEIDON == STWO_B200_T_EPOSEIDON && tape::T_M4 == STWO_B200_T_M4 && tape::T_POSEIDON == STWO_B200_T_POSEIDON && tape::T_ADD == STWO_B200_T_ADD && tape::T_BIT == STWO_B200_T_BIT, "opcodes mirror the C ABI");

extern "C" int32_t stwo_b200_cs_eval_tape_dev(const stwo_b200_cs_tape *t, uint32_t n_vars, const uint32_t *witness, const stwo_b200_cs_values *v,
                                              void *stream) {
    STWO_CHECK_DEVICE();
    if (!t || !values_ok(v) || n_vars < 4 || !t->ins || !t->level_start || (t->n_perms && !t->perms) || (t->n_input_words && !witness))
        return STWO_B200_E_BAD_ARG;
    Batch b = batch_of(v, n_vars, t->n_perms);
    const u32 n_groups = (v->n_batch + v->lanes - 1) / v->lanes;
    cudaStream_t st = (cudaStream_t)stream;
    const tape::Ins *ins = reinterpret_cast<const tape::Ins *>(t->ins);
    const tape::Perm *perms = reinterpret_cast<const tape::Perm *>(t->perms);
    // with bundles the levels are ranges of bundles; without, of instructions (every instruction its own bundle)
    const u32 *bundle_start = t->n_bundles ? t->bundle_start : nullptr;
    const u32 *level_start = t->n_bundles ? t->level_bundle : t->level_start;
    if (t->n_bundles && (!t->bundle_start || !t->level_bundle)) return STWO_B200_E_BAD_ARG;
    u32 n_levels = t->n_levels, n_input_words = t->n_input_words;
    const u32 *eperms = t->eperms;
    if (t->n_eperms && !eperms) return STWO_B200_E_BAD_ARG;
    // grid-wide levels while the batch has fewer lane groups than a few waves of SMs; CTA-local levels beyond that
    static int n_sm = 0, coop = 0, grid_mode = -1, unrolled = 0, cluster = -1;
    if (!n_sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        const char *e = getenv("STWO_B200_EVAL_MODE");        // "cta" / "grid" / "cluster": force one kernel (profiling)
        if (e) grid_mode = e[0] == 'g' ? 1 : e[0] == 'c' && e[1] == 'l' ? 2 : 0;
        e = getenv("STWO_B200_EVAL_CLUSTER");                 // CTAs per cluster (1, 2, 4, 8, 16)
        if (e) cluster = atoi(e);
        e = getenv("STWO_B200_EVAL_UNROLLED");                // "1": fully unrolled permutation inside the grid kernel (profiling)
        if (e) unrolled = e[0] == '1';
    }
    // default for lanes = 32: one cluster per lane group (a plain launch: evaluations on different streams run beside each other, which
    // two cooperative grids must not be asked to do).  Cluster size: many CTAs for a handful of lane groups (the levels are then
    // spread over many SMs), two when the batch alone fills the GPU (4096 proofs: all 128 clusters resident at once).
    if (v->lanes == 32 && (grid_mode == 2 || grid_mode < 0)) {
        int c = cluster > 0 ? cluster : (n_groups <= 2 ? 16 : n_groups <= 24 ? 8 : n_groups <= 64 ? 4 : 2);
        auto kernel = g_level_clock_host ? k_tape_eval_cluster<true> : k_tape_eval_cluster<false>;
        if (c > 8) STWO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        // the second order: usable when the caller has a permutation record at all (the kernel decides per lane group)
        const stwo_b200_cs_tape_order *o2 = (t->recorded_order && v->perm_hints && t->recorded_order->ins && t->recorded_order->n_bundles) ? t->recorded_order : nullptr;
        // (the recorder builds that order only under STWO_B200_RECORDED_ORDER=1: see dsl/recorded.hpp for the measurements)
        const u32 max_levels = o2 && o2->n_levels > n_levels ? o2->n_levels : n_levels;
        const size_t smem = ((size_t)max_levels + 1) * 4;
        if (smem > 48 * 1024) STWO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_groups * (unsigned)c); cfg.blockDim = dim3(kClusterThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        STWO_CUDA(cudaLaunchKernelEx(&cfg, kernel, ins, level_start, n_levels, perms, b, witness, n_input_words, eperms, bundle_start,
                                     o2 ? reinterpret_cast<const tape::Ins *>(o2->ins) : (const tape::Ins *)nullptr, o2 ? o2->level_bundle : (const u32 *)nullptr,
                                     o2 ? o2->n_levels : 0u, o2 ? o2->bundle_start : (const u32 *)nullptr));
        note_launch(1);
        return cuda_status(cudaGetLastError());
    }
    const bool use_grid = v->lanes == 32 && coop && (grid_mode == 1 || (grid_mode < 0 && n_groups <= 8u * (u32)n_sm));
    if (use_grid) {
        static unsigned *barriers = nullptr;
        static unsigned next = 0;
        if (!barriers) STWO_CUDA(cudaMalloc(&barriers, 64 * sizeof(unsigned)));
        unsigned *bar = barriers + (next++ % 64);
        STWO_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
        void *args[] = {(void *)&ins, (void *)&level_start, (void *)&n_levels, (void *)&perms, (void *)&b, (void *)&witness, (void *)&n_input_words, (void *)&bar, (void *)&eperms,
                        (void *)&bundle_start};
        const void *fn = unrolled ? (const void *)k_tape_eval_grid<true> : (const void *)k_tape_eval_grid<false>;
        STWO_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)n_sm), dim3(kEvalThreads), args, 0, st));
    } else {
        k_tape_eval<<<n_groups, kEvalThreads, 0, st>>>(ins, level_start, n_levels, perms, b, witness, n_input_words, eperms, bundle_start);
    }
    note_launch(1);
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_eval_level_clock(uint64_t *level_clock) {
    STWO_CHECK_DEVICE();
    unsigned long long *p = (unsigned long long *)level_clock;
    g_level_clock_host = p != nullptr;
    return cuda_status(cudaMemcpyToSymbol(g_level_clock, &p, sizeof p));
}
extern "C" int32_t stwo_b200_cs_check_arithmetics_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, int64_t *first_bad,
                                                      void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !values_ok(v) || !first_bad) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const Batch b = batch_of(v, w->n_vars, w->n_flow);
    const size_t padded = (size_t)(v->n_batch + v->lanes - 1) / v->lanes * v->lanes;
    k_fill64<<<nblk(v->n_batch), kT, 0, st>>>((unsigned long long *)first_bad, v->n_batch, ~0ull);
    k_cs_check_arith<<<nblk((size_t)w->n_rows * padded), kT, 0, st>>>(*w, b, (unsigned long long *)first_bad);
    note_launch(2);
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_populate_logup_dev(const stwo_b200_cs_wiring *w, int32_t *mult_a, int32_t *mult_b, int32_t *mult_c,
                                                   int32_t *mult_poseidon, uint32_t *scratch, uint32_t *status_out, void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !mult_c || !scratch || !status_out) return STWO_B200_E_BAD_ARG;
    if (w->kind == 0 && (!mult_a || !mult_b || !mult_poseidon)) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nv = w->n_vars;
    u32 *counts = scratch, *first_key = scratch + nv, *first_prow = scratch + 2 * nv, *mpv = scratch + 3 * nv;
    STWO_CUDA(cudaMemsetAsync(counts, 0, nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(first_key, 0xff, 2 * nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(mpv, 0, nv * 4, st));
    STWO_CUDA(cudaMemsetAsync(status_out, 0, 4, st));
    size_t span = w->n_rows;
    if (w->n_flow > span) span = w->n_flow;
    if (w->num_input > span) span = w->num_input;
    k_cs_count<<<nblk(span), kT, 0, st>>>(*w, counts, first_key, first_prow, mpv);
    k_cs_mult<<<nblk(w->n_rows), kT, 0, st>>>(*w, counts, first_key, first_prow, mpv, mult_a, mult_b, mult_c, mult_poseidon, status_out);
    note_launch(2);
    return cuda_status(cudaGetLastError());
}
// ctas_per_sm: 0 = fill the SMs (the kernel alone), k = a resident layer of k CTAs per SM (beside another kernel)
int32_t stwo_b200::cs_check_poseidon_launch(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const int32_t *mult_poseidon,
                                            const uint32_t *scratch, int64_t *first_bad, cudaStream_t st, int ctas_per_sm,
                                            const stwo_b200_cs_tape *tape) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !values_ok(v) || !first_bad || !mult_poseidon || !scratch) return STWO_B200_E_BAD_ARG;
    // the record-based check needs the tape's permutation records (entry -> slot), both records and one record per flow entry
    const bool recorded = tape && tape->perms && tape->n_perms == w->n_flow && v->perm_hints && v->perm_hint_inputs && v->perm_hint_stride;
    k_fill64<<<nblk(v->n_batch), kT, 0, st>>>((unsigned long long *)first_bad, v->n_batch, ~0ull);
    note_launch(1);
    if (w->n_flow) {
        if (!v->flow_hash || !v->flow_swap) return STWO_B200_E_BAD_ARG;
        const Batch b = batch_of(v, w->n_vars, w->n_flow);
        const size_t padded = (size_t)(v->n_batch + v->lanes - 1) / v->lanes * v->lanes;
        static int n_sm = 0;
        if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
        const size_t want = nblk((size_t)w->n_flow * padded, 128), cap = (size_t)n_sm * (ctas_per_sm > 0 ? ctas_per_sm : 16);
        if (recorded)
            k_cs_check_poseidon<true><<<(unsigned)(want < cap ? want : cap), 128, 0, st>>>(*w, b, mult_poseidon, scratch + 2 * (size_t)w->n_vars,
                                                                                         (unsigned long long *)first_bad,
                                                                                         reinterpret_cast<const tape::Perm *>(tape->perms));
        else
            k_cs_check_poseidon<false><<<(unsigned)(want < cap ? want : cap), 128, 0, st>>>(*w, b, mult_poseidon, scratch + 2 * (size_t)w->n_vars,
                                                                                          (unsigned long long *)first_bad, nullptr);
        note_launch(1);
    }
    return cuda_status(cudaGetLastError());
}
extern "C" int32_t stwo_b200_cs_check_poseidon_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v,
                                                   const int32_t *mult_poseidon, const uint32_t *scratch, int64_t *first_bad, void *stream) {
    return stwo_b200::cs_check_poseidon_launch(w, v, mult_poseidon, scratch, first_bad, (cudaStream_t)stream, 0);
}
extern "C" int32_t stwo_b200_cs_check_poseidon_recorded_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const stwo_b200_cs_tape *tape,
                                                            const int32_t *mult_poseidon, const uint32_t *scratch, int64_t *first_bad, void *stream) {
    return stwo_b200::cs_check_poseidon_launch(w, v, mult_poseidon, scratch, first_bad, (cudaStream_t)stream, 0, tape);
}
extern "C" int32_t stwo_b200_cs_export_trace_dev(const stwo_b200_cs_wiring *w, const stwo_b200_cs_values *v, const int32_t *mult_a,
                                                 const int32_t *mult_b, const int32_t *mult_c, const int32_t *mult_poseidon,
                                                 uint32_t *preprocessed, uint32_t *values, int64_t *first_bad, void *stream) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(w) || !values_ok(v) || (!values && !preprocessed)) return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (preprocessed) {
        if (!mult_c || (w->kind == 0 && (!mult_a || !mult_b || !mult_poseidon))) return STWO_B200_E_BAD_ARG;
        k_cs_export_pre<<<nblk(w->n_rows), kT, 0, st>>>(*w, mult_a, mult_b, mult_c, mult_poseidon, preprocessed);
        note_launch(1);
    }
    if (first_bad && !values) return STWO_B200_E_BAD_ARG;
    if (values) {
        const Batch b = batch_of(v, w->n_vars, w->n_flow);
        unsigned long long *fb = (unsigned long long *)first_bad;
        if (fb) { k_fill64<<<nblk(v->n_batch), kT, 0, st>>>(fb, v->n_batch, ~0ull); note_launch(1); }
        if (v->lanes == 1) k_cs_export_vals_plain<<<nblk((size_t)w->n_rows * v->n_batch), kT, 0, st>>>(*w, b, values, fb);
        else {
            static int items = 0, n_sm = 0, ctas_per_sm = 2;
            if (!items) {
                const char *e = getenv("STWO_B200_EXPORT_ITEMS");          // 16 / 32: the tiled form (profiling); default: the streaming form
                items = e && atoi(e) == 16 ? 16 : e && atoi(e) == 32 ? 32 : -1;
                e = getenv("STWO_B200_EXPORT_CTAS");
                if (e && atoi(e) > 0) ctas_per_sm = atoi(e);
                int dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
                STWO_CUDA(cudaFuncSetAttribute(k_cs_export_vals_tiled<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCols * 32 * 33 * 4));
                STWO_CUDA(cudaFuncSetAttribute(k_cs_export_vals_tiled<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCols * 16 * 33 * 4));
                STWO_CUDA(cudaFuncSetAttribute(k_cs_export_vals_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2 * 96 * kXPitch * 16 + kXMetaStages * kXTileWords * 4));
            }
            auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
            if (items < 0 && w->export_tiles && w->export_cap >= 1 && w->export_cap <= 96 && al16(w->export_tiles) && al16(v->variables) && al16(values) &&
                w->n_rows >= (u32)kXRows) {
                const size_t kXSmem = (size_t)2 * 2 * w->export_cap * kXPitch * sizeof(uint4) + kXMetaStages * kXTileWords * 4;
                const u32 n_row_tiles = w->n_rows / kXRows, n_item_groups = (v->n_batch + kXItems - 1) / kXItems;
                const size_t n_tiles = (size_t)n_row_tiles * n_item_groups;
                if (n_tiles > 0xffffffffull) return STWO_B200_E_BAD_ARG;
                const size_t cap = (size_t)n_sm * ctas_per_sm;
                k_cs_export_vals_stream<<<(unsigned)(n_tiles < cap ? n_tiles : cap), kXThreads, kXSmem, st>>>(*w, b, values, fb, n_row_tiles, (u32)n_tiles);
            } else {
                if (items < 0) items = 32;
                const size_t smem = (size_t)kCols * items * 33 * 4;
                dim3 grid(w->n_rows / kTileRows, (v->n_batch + items - 1) / items);
                if (items == 32) k_cs_export_vals_tiled<32><<<grid, kT, smem, st>>>(*w, b, values, fb);
                else k_cs_export_vals_tiled<16><<<grid, kT, smem, st>>>(*w, b, values, fb);
            }
        }
        note_launch(1);
    }
    return cuda_status(cudaGetLastError());
}

extern "C" uint32_t stwo_b200_cs_flow_padded_len(uint32_t n_flow) {
    const uint32_t r = (n_flow + 15) / 16 * 16;          // max(N_LANES * 2, n.div_ceil(16) * 16), plonk_with_poseidon.rs:297
    return r < 32 ? 32 : r;
}
extern "C" int32_t stwo_b200_cs_export_flow_dev(const stwo_b200_cs_values *v, uint32_t n_flow, const uint32_t *pad_constants, uint32_t *pad_scratch,
                                                uint32_t *flow_hash_out, uint8_t *flow_swap_out, void *stream) {
    STWO_CHECK_DEVICE();
    if (!values_ok(v) || v->lanes != 32 || !pad_constants || !pad_scratch || !flow_hash_out || !flow_swap_out || (n_flow && (!v->flow_hash || !v->flow_swap)))
        return STWO_B200_E_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // padding entry = (C1, C1, C2, C3): 32 words on the device
    uint32_t pad[32];
    for (int k = 0; k < 8; k++) { pad[k] = pad[8 + k] = pad_constants[k]; pad[16 + k] = pad_constants[8 + k]; pad[24 + k] = pad_constants[16 + k]; }
    for (int k = 0; k < 32; k++) if (pad[k] >= M31_P) return STWO_B200_E_BAD_ARG;
    STWO_CUDA(cudaMemcpyAsync(pad_scratch, pad, sizeof pad, cudaMemcpyHostToDevice, st));
    STWO_CUDA(cudaStreamSynchronize(st));                 // `pad` is a stack buffer
    Batch b = batch_of(v, 4, n_flow);
    const u32 n_pad = stwo_b200_cs_flow_padded_len(n_flow);
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const size_t n_tiles = (size_t)((v->n_batch + 31) / 32) * n_pad, want = (n_tiles + 7) / 8, cap = (size_t)n_sm * 8;
    k_cs_export_flow<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(b, n_pad, pad_scratch, flow_hash_out, flow_swap_out);
    note_launch(1);
    return cuda_status(cudaGetLastError());
}

extern "C" size_t stwo_b200_cs_export_tiles_words(uint32_t n_rows) { return n_rows < (u32)kXRows ? 0 : (size_t)(n_rows / kXRows) * kXTileWords; }
extern "C" int32_t stwo_b200_cs_export_tiles_build(const stwo_b200_cs_wiring *hw, uint32_t *tiles, uint32_t *cap_out) {
    if (!hw || !tiles || !cap_out || !hw->a_wire || !hw->b_wire || !hw->c_wire || !hw->op || hw->n_rows < (u32)kXRows || (hw->n_rows % kXRows) || hw->kind > 1)
        return STWO_B200_E_BAD_ARG;
    if (hw->kind == 1 ? (!hw->op2 || !hw->op3 || !hw->op4) : !hw->enforce_c_m31) return STWO_B200_E_BAD_ARG;
    u32 cap = 1;
    const u32 *cols[3] = {hw->a_wire, hw->b_wire, hw->c_wire};
    for (u32 rt = 0; rt < hw->n_rows / kXRows; rt++) {
        u32 *m = tiles + (size_t)rt * kXTileWords;
        memset(m, 0, kXTileWords * 4);
        u32 *rec = m + 2 * kXListWords;
        for (u32 half = 0; half < 2; half++) {
            u32 *list = m + half * kXListWords;
            u32 count = 0;
            for (u32 r = 0; r < 32; r++) {
                const u32 row = rt * kXRows + half * 32 + r;
                u32 x = 0;
                for (u32 o = 0; o < 3; o++) {                  // first-appearance order: a gate chain gets consecutive slots
                    const u32 v = cols[o][row];
                    u32 k = 0;
                    while (k < count && list[1 + k] != v) k++;
                    if (k == count) list[1 + count++] = v;
                    x |= k << (8 * o);
                }
                if (hw->kind == 1) {
                    if (hw->op2[row] > 1 || hw->op3[row] > 1 || hw->op4[row] > 1) return STWO_B200_E_BAD_ARG;      // not a selector: no packed form
                    x |= (hw->op2[row] ? XF_C1 : 0u) | (hw->op3[row] ? XF_OP3 : 0u) | (hw->op4[row] ? XF_OP4 : 0u);
                } else if (hw->enforce_c_m31[row]) x |= XF_C1;
                if (hw->op_follows_c && hw->op_follows_c[row]) x |= XF_FOLLOWS;
                rec[2 * (half * 32 + r)] = x;
                rec[2 * (half * 32 + r) + 1] = hw->op[row];
            }
            list[0] = count;
            if (count > cap) cap = count;
        }
        // a row that repeats the wires and constants of an earlier row of the tile (the padding rows above all) checks nothing new
        for (u32 r = 1; r < (u32)kXRows; r++) {
            const u32 row = rt * kXRows + r;
            for (u32 q = (r & 32u); q < r; q++) {              // same half: the slots of the two halves are not comparable
                if (((rec[2 * q] ^ rec[2 * r]) & ~XF_SKIP) == 0 && rec[2 * q + 1] == rec[2 * r + 1] && cols[0][rt * kXRows + q] == cols[0][row] &&
                    cols[1][rt * kXRows + q] == cols[1][row] && cols[2][rt * kXRows + q] == cols[2][row]) { rec[2 * r] |= XF_SKIP; break; }
            }
        }
    }
    *cap_out = cap;
    return STWO_B200_OK;
}

extern "C" int32_t stwo_b200_cs_finalize(const stwo_b200_cs_wiring *hw, const stwo_b200_cs_values *hv, uint32_t *trace,
                                         int64_t *bad_row, int64_t *bad_flow) {
    STWO_CHECK_DEVICE();
    if (!wiring_ok(hw) || hw->kind != 0 || !hv || hv->n_batch != 1 || hv->lanes != 1 || !hv->variables || !trace || !bad_row || !bad_flow) return STWO_B200_E_BAD_ARG;
    const size_t nr = hw->n_rows, nv = hw->n_vars, nf = hw->n_flow;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_wires = take(6 * nr * 4), o_follow = take(nr), o_fw = take(nf * 4 * 4 + 4), o_fa = take(nf * 4 + 4), o_vars = take(nv * 16),
                 o_fh = take(nf * 128 + 4), o_fs = take(nf + 4), o_mult = take(4 * nr * 4), o_scr = take((4 * nv + 4) * 4),
                 o_bad = take(16), o_stat = take(4), o_pre = take(10 * nr * 4), o_vals = take(kCols * nr * 4);
    int32_t rc = stage_reserve(off);
    if (rc) return rc;
    cudaStream_t st = stage_stream();
    uint8_t *d = stage_dev();
    u32 *dw = (u32 *)(d + o_wires);
    const u32 *srcs[6] = {hw->a_wire, hw->b_wire, hw->c_wire, hw->poseidon_wire, hw->enforce_c_m31, hw->op};
    for (int k = 0; k < 6; k++) STWO_CUDA(cudaMemcpyAsync(dw + k * nr, srcs[k], nr * 4, cudaMemcpyHostToDevice, st));
    if (hw->op_follows_c) STWO_CUDA(cudaMemcpyAsync(d + o_follow, hw->op_follows_c, nr, cudaMemcpyHostToDevice, st));
    if (nf) {
        if (!hv->flow_hash || !hv->flow_swap) return STWO_B200_E_BAD_ARG;
        STWO_CUDA(cudaMemcpyAsync(d + o_fw, hw->flow_wire, nf * 16, cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d + o_fa, hw->flow_swap_addr, nf * 4, cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d + o_fh, hv->flow_hash, nf * 128, cudaMemcpyHostToDevice, st));
        STWO_CUDA(cudaMemcpyAsync(d + o_fs, hv->flow_swap, nf, cudaMemcpyHostToDevice, st));
    }
    STWO_CUDA(cudaMemcpyAsync(d + o_vars, hv->variables, nv * 16, cudaMemcpyHostToDevice, st));
    stwo_b200_cs_wiring w = *hw;
    w.a_wire = dw; w.b_wire = dw + nr; w.c_wire = dw + 2 * nr; w.poseidon_wire = dw + 3 * nr; w.enforce_c_m31 = dw + 4 * nr; w.op = dw + 5 * nr;
    w.op_follows_c = hw->op_follows_c ? d + o_follow : nullptr;
    w.flow_wire = (u32 *)(d + o_fw); w.flow_swap_addr = (u32 *)(d + o_fa);
    stwo_b200_cs_values v = {1, 1, (u32 *)(d + o_vars), (u32 *)(d + o_fh), d + o_fs, nullptr, 0, nullptr, 0, nullptr};
    int32_t *m = (int32_t *)(d + o_mult);
    int64_t *bad = (int64_t *)(d + o_bad);
    u32 *scr = (u32 *)(d + o_scr), *stat = (u32 *)(d + o_stat), *pre = (u32 *)(d + o_pre), *vals = (u32 *)(d + o_vals);
    if ((rc = stwo_b200_cs_check_arithmetics_dev(&w, &v, bad, st))) return rc;
    if ((rc = stwo_b200_cs_populate_logup_dev(&w, m, m + nr, m + 2 * nr, m + 3 * nr, scr, stat, st))) return rc;
    if ((rc = stwo_b200_cs_check_poseidon_dev(&w, &v, m + 3 * nr, scr, bad + 1, st))) return rc;
    if ((rc = stwo_b200_cs_export_trace_dev(&w, &v, m, m + nr, m + 2 * nr, m + 3 * nr, pre, vals, nullptr, st))) return rc;
    int64_t hb[2];
    u32 hstat = 0;
    STWO_CUDA(cudaMemcpyAsync(hb, bad, 16, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaMemcpyAsync(&hstat, stat, 4, cudaMemcpyDeviceToHost, st));
    // 22 columns: 9 shared preprocessed + the item's op column, then the 12 value columns
    STWO_CUDA(cudaMemcpyAsync(trace, pre, 9 * nr * 4, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaMemcpyAsync(trace + 9 * nr, vals + 12 * nr, nr * 4, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaMemcpyAsync(trace + 10 * nr, vals, 12 * nr * 4, cudaMemcpyDeviceToHost, st));
    STWO_CUDA(cudaStreamSynchronize(st));
    *bad_row = hb[0]; *bad_flow = hb[1];
    if (hstat) *bad_flow = -2;     // a Poseidon wire is referenced by more than one row (reference: assert_eq!(counts[..], 1))
    return STWO_B200_OK;
}
