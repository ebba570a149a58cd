#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native recursive-stwo verifier hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the CPU oracle port on the host cores)

Workload (BASELINE.json configs[4], the single-GPU share of it): a batch of
--proofs replicas per GPU of the reference's own fixture components/test_data/small_proof.bin (shape S:
pow 20, blow-up 5, last 2, 16 queries, 7 inner FRI layers).  One "step" = one batch through the whole path:
  (1) native verification: transcript + PoW, logup sum, OODS, batched Merkle decommitment re-shaped into
      per-query paths, DEEP answers, circle/line folds, last layer, every per-query authentication path;
  (2) the verifier circuit's trace: witness streams gathered from (1), variables[] evaluated along the recorded
      tape (its 3 481 Poseidon2 permutations per proof are exactly the transcript + per-query path permutations (1) just
      executed: their output states are taken from (1)'s record instead of permuting again), check_arithmetics,
      check_poseidon_invocations (re-executes all 3 481 flow entries), export of the 2^16-row x 13 per-proof trace
      columns (examples/single-proof/src/main.rs:33-90).
Metric: verified proofs/s over all GPUs (a proof counts when its verdict is accept AND its circuit checks pass);
`poseidon31_perms_per_sec` rides along.  Both legs run a stream of batches through VerifyTracePipeline (three device slots; upload,
verification and trace pass of neighbouring batches on their own streams).  `value`: blobs already in HBM, no upload.  `e2e`: every
step uploads one batch of blobs from pinned host memory and copies verdicts + check results back to pinned host memory (traces stay
in HBM for the prover that consumes them; shipping them to the host instead would cost 3.4 MB per proof over PCIe, ~70 ms per
4096-proof batch at 200 GB/s).
Prints ONE JSON line on rank 0.
"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before the CUDA context exists (see recursive-stwo_b200/__init__.py)
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "verified_proofs_per_sec"
UNIT = "proofs/s"
FIXTURE = "small_proof.bin"
# warp instructions one permutation executes in the path kernels (ncu smsp__inst_executed / permutations,
# profiles/r01*_ncu.txt) — the unit of the integer-issue roofline
LANE_OPS_PER_PERM = 4719


def ncu_traffic(kernel, n_proofs):
    """DRAM bytes per launch of `kernel` from a committed ncu capture at exactly this batch size (profiles/ncu_traffic.json), else None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(t[kernel][str(n_proofs)]["bytes"])
    except Exception:
        return None


_JSON_OUT = None


def emit(line):
    f = _JSON_OUT or sys.stdout
    f.write(line + "\n")
    f.flush()


def peaks():
    p = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p = {"hbm_gbs": float(m["hbm_gbs"]), "src": "measured"}
    except Exception:
        pass
    try:
        ip = json.load(open(os.path.join(ROOT, "profiles", "intpipe_r01.json")))
        p["int_tlops"] = float(ip["peak_tera_lane_ops_per_s"])
        p["int_src"] = "measured (profiles/intpipe_r01.json)"
    except Exception:
        p["int_tlops"] = 30.9
        p["int_src"] = "fallback"
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.3)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def load_fixture():
    return open(os.path.join(ROOT, "tests", "golden", "proofs", FIXTURE), "rb").read()


# ----------------------------------------------------------------------------------------------------
class CpuPath:
    """The reference path on the host cores: oracle/orc_verify.c (native verifier) + oracle/orc_tape.c (the circuit's
    value arithmetic, check_arithmetics, check_poseidon_invocations, value-column export replayed from the value log that
    oracle/orc_dsl.py records), both on `cores` pthreads over replicas of the fixture."""

    def __init__(self, n_proofs, cores):
        import oracle_py as O
        sys.path.insert(0, O.ORACLE_DIR)
        import orc_dsl as D
        self.O, self.n, self.cores = O, n_proofs, cores
        lib = self.lib = O.load_oracle()
        lib.orc_verify_batch_mt.restype = ctypes.c_uint64
        lib.orc_circuit_trace_mt.restype = ctypes.c_int64
        buf, n = O.load_proof(FIXTURE)
        self.blobs = np.tile(buf[:n], n_proofs)
        self.off = np.arange(n_proofs + 1, dtype=np.uint64) * n
        self.idx = np.array(O.INPUTS_SMALL[0], dtype=np.uint32)
        self.vals = np.array(O.INPUTS_SMALL[1], dtype=np.uint32)
        self.v, self.s = np.zeros(n_proofs, np.uint8), np.zeros(n_proofs, np.uint8)
        cs, _ = D.verifier_circuit(bytes(buf[:n]), D.INPUTS_SINGLE, 1, O.VerifyOut)
        self.ops, self.perms = D.value_log_arrays(cs)
        self.wiring = np.ascontiguousarray([cs.a_wire, cs.b_wire, cs.c_wire, cs.poseidon_wire, cs.enforce_c_m31, cs.op], dtype=np.uint32)
        self.flow_wire = cs.flow_arrays()[0]
        self.n_vars, self.n_rows, self.circuit_perms = len(cs.variables), len(cs.a_wire), len(cs.flow)

    def step(self):
        """one batch: verify, then the circuit of every proof; returns the permutations executed"""
        O, lib = self.O, self.lib
        perms = lib.orc_verify_batch_mt(O.vp(self.blobs), O.vp(self.off), self.n, O.vp(self.idx), O.vp(self.vals), self.idx.size,
                                        O.vp(self.v), O.vp(self.s), self.cores)
        bad = lib.orc_circuit_trace_mt(O.vp(self.ops), len(self.ops), O.vp(self.perms), len(self.perms), self.n_vars, O.vp(self.wiring),
                                       self.n_rows, O.vp(self.flow_wire), self.n, self.cores)
        assert bad == 0 and not self.v.any()
        return perms + 2 * self.circuit_perms * self.n          # evaluation + check_poseidon_invocations


def cpu_rate(n_proofs, seconds, cores):
    cpu = CpuPath(n_proofs, cores)
    cpu.step()
    reps, perms, t0 = 0, 0, time.perf_counter()
    while True:
        perms += cpu.step()
        reps += 1
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return reps * n_proofs / dt, perms / dt, reps, dt


def run_reference(args):
    """The reference's CPU implementation of the path.  The Rust workspace cannot be built here (no cargo/rustc; its
    stwo git dependency is absent), so this is the C restatement in oracle/ — kind "port" — on all host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    n = max(cores * 4, 64)
    cpu = CpuPath(n, cores)
    perms = 0
    for _ in range(args.warmup):
        cpu.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        perms += cpu.step()
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = "%d replicas of %s per step on %d pthreads: native verifier + circuit value log replay, checks and export " \
             "(the GPU arm's step is %d per GPU)" % (n, FIXTURE, cores, args.proofs)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (M31)", "data": "synthetic",
        "config": {"workload": "verify-batch", "fixture": FIXTURE, "proofs_per_gpu": args.proofs, "mode": "full+trace"},
        "poseidon31_perms_per_sec": perms / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------------
def merkle_sweep(pkg, dev, log_n=20, C=8, Q=128, T=2, reps=3, hbm_peak=None):
    """BASELINE configs[2]: commit T trees of 2^log_n leaves (C columns), decommit Q queries per tree, verify every path.
    Returns the commit and the path legs timed apart: the commit streams C*4 B per leaf (HBM-bound for wide leaves), the
    paths are permutation-bound."""
    import torch
    S = importlib.import_module("recursive-stwo_b200.synth")
    n = 1 << log_n
    cols = torch.from_numpy(np.stack([S.synth_m31(t, C * n).reshape(C, n) for t in range(T)]).view(np.int32)).to(dev)
    idx = torch.from_numpy(np.stack([(S.splitmix64(t ^ 0xABCDEF, Q) & np.uint64(n - 1)).astype(np.uint32) for t in range(T)]).view(np.int32)).to(dev)
    nodes = torch.empty((T, 2 * n - 1, 8), dtype=torch.int32, device=dev)
    rid = torch.arange(T, dtype=torch.int32, device=dev).repeat_interleave(Q).contiguous()
    shape = pkg.PathShape.make(log_n, {log_n: C})

    def commit():
        pkg.merkle_commit(cols, nodes)

    def paths():
        pc, sib = pkg.merkle_decommit(cols, nodes, idx)
        return pkg.merkle_path_verify(shape, idx.reshape(-1), pc, sib, nodes[:, 0, :].contiguous(), rid)

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    for _ in range(2):
        commit()
        v = paths()
    assert bool(v.all().item())
    commit_ms, paths_ms = timed(commit, reps), timed(paths, reps * 4)
    leaf = (C + 7) // 8 + 1
    commit_perms, path_perms = T * (n * leaf + n - 1), T * Q * (leaf + log_n)
    commit_bytes = T * (n * C * 4 + (2 * n - 1) * 32)          # leaves read once, every node written once
    out = {"log_leaves": log_n, "columns": C, "queries": Q, "trees": T, "commit_ms": commit_ms, "paths_ms": paths_ms,
           "commit_perms_per_sec": commit_perms / (commit_ms * 1e-3), "path_perms_per_sec": path_perms / (paths_ms * 1e-3),
           "commit_gb_per_sec": commit_bytes / (commit_ms * 1e-3) / 1e9,
           "perms_per_sec": (commit_perms + path_perms) / ((commit_ms + paths_ms) * 1e-3)}
    if hbm_peak:
        out["commit_frac_of_hbm_peak"] = out["commit_gb_per_sec"] / hbm_peak
    return out


def fixture_names():
    """the 15 Poseidon31 fixtures in the order of BASELINE configs[3]"""
    d = os.path.join(ROOT, "tests", "golden", "proofs")
    lv = sorted((f for f in os.listdir(d) if f.startswith("level") and f != "level14-1.bin"),
                key=lambda f: (int(f[5:].split("-")[0]), f))
    return ["small_proof.bin", "recursive_proof_16_15.bin"] + lv


def multi_proofs_leg(pkg, sharding, rank, world, dev, n_total=256, reps=3):
    """BASELINE configs[3] (examples/multi-proofs): 256 independent proofs = the 15 fixtures cycled, VERIFIED AND TRACED (every proof's
    verifier-circuit trace generated, checked and exported, like the headline step) through MixedBatch: grouped by shape, one recorded
    circuit per shape, sharded over the ranks in contiguous blocks of the shape-sorted order cut by WORK (permutations per proof), not
    by count -- the 80-query level1-5 / level4-5 proofs cost ~8x a small one.  Device time, max over ranks."""
    import torch
    import torch.distributed as dist
    d = os.path.join(ROOT, "tests", "golden", "proofs")
    names = fixture_names()
    blobs = {f: open(os.path.join(d, f), "rb").read() for f in names}
    shape_of = {f: pkg.shape_for(blobs[f]) for f in names}
    order = sorted((names[i % len(names)] for i in range(n_total)), key=lambda f: (tuple(shape_of[f].key()), f))
    cost = [pkg.proof_perms(shape_of[f]) for f in order]
    lo, hi = sharding.shard_by_work(cost, world)[rank]
    mine = order[lo:hi]
    # the single-proof fixture carries one public input, the recursion fixtures three: two resident batches
    parts = [pkg.MixedBatch([blobs[f] for f in mine if f.startswith("small") == sm], inputs=pkg.INPUTS_SINGLE if sm else pkg.INPUTS_RECURSIVE)
             for sm in (True, False) if any(f.startswith("small") == sm for f in mine)]

    # the two resident batches run beside each other too: each on its own stream, joined on the current one
    side = [torch.cuda.Stream(dev) for _ in parts]

    def run_parts():
        cur = torch.cuda.current_stream(dev)
        out = []
        for mb, st in zip(parts, side):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                out.append(mb.run(trace=True, export=True)[0])
        for st in side:
            cur.wait_stream(st)
        return out

    def step():
        return sum(int(v.sum().item()) for v in run_parts())

    assert step() == 0, "every fixture must be accepted and every circuit consistent"
    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run_parts()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps, float(sum(cost[lo:hi])), float(hi - lo)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, t)
    ms = float(tmax[0].item())
    rows = sum(g.circuit.info.n_rows * len(g.ids) for mb in parts for g in mb.groups)
    return {"config": "BASELINE configs[3]: 256 proofs = 15 fixtures cycled; verify + circuit trace (check + export) per proof; one recorded circuit per shape; "
                      "ranks take contiguous blocks of the shape-sorted order cut by work",
            "proofs": n_total, "shapes": len({tuple(shape_of[f].key()) for f in names}), "ms_per_batch": ms, "proofs_per_sec": n_total / (ms * 1e-3),
            "max_rank_work_share": float(tmax[1].item()) / sum(cost), "max_rank_proofs": int(tmax[2].item()), "rank0_trace_rows": rows,
            "per_rank": [{"ms": round(float(x[0].item()), 3), "work_share": round(float(x[1].item()) / sum(cost), 3), "proofs": int(x[2].item())} for x in per_rank]}


def trace_gather_leg(sharding, values, rank, world, dev, n_each=256, reps=3):
    """north_star's second collective: the trace columns of n_each proofs per rank gathered on rank 0 over NCCL (NVLink / NVSwitch).
    values: this rank's [n_local, 13, n_rows] trace columns.  Device time, max over ranks."""
    import torch
    import torch.distributed as dist
    n_each = min(n_each, values.shape[0])
    local = values[:n_each]
    out = sharding.gather_trace_columns(local, n_each * world, dst=0)             # warm-up (NCCL channel setup)
    ok = True
    if rank == 0:
        ok = bool(torch.equal(out[:n_each], local)) and out.shape[0] == n_each * world
    del out
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = sharding.gather_trace_columns(local, n_each * world, dst=0)
        del out
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    recv = (world - 1) * local.numel() * 4
    return {"config": "trace columns of %d proofs per rank gathered on rank 0 (NCCL gather)" % n_each, "ok": ok, "ms": ms,
            "bytes_received_by_rank0": recv, "gb_per_sec_into_rank0": recv / (ms * 1e-3) / 1e9}


def shape_leg(pkg, dev, fixture, n, reps=5):
    """the headline step (verify + circuit trace) on replicas of another fixture: SURVEY.md 8a's shape R (recursive_proof_16_15.bin:
    2^16 / 2^15 rows, 16 queries, 8 inner layers, 5 289 path permutations)"""
    import torch
    blob = open(os.path.join(ROOT, "tests", "golden", "proofs", fixture), "rb").read()
    vb = pkg.VerifyBatch([blob] * n, inputs=pkg.INPUTS_RECURSIVE)
    circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_RECURSIVE)

    def step():
        v, _ = vb.run(full=True)
        return v, circ.trace(vb, check=True, export=True, preprocessed=False)
    for _ in range(2):
        v, r = step()
    assert int(v.sum().item()) == 0 and int((r["bad_row"] != -1).sum().item()) == 0 and int((r["bad_flow"] != -1).sum().item()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"fixture": fixture, "proofs": n, "rows": circ.info.n_rows, "poseidon_flow": circ.info.n_flow, "tape_levels": circ.info.n_levels,
            "ms_per_batch": ms, "proofs_per_sec": n / (ms * 1e-3)}


def synthetic_leg(pkg, dev, shape, n=4096, reps=5):
    """BASELINE configs[4] part i: n DISTINCT synthetic FRI + Merkle instances of the headline shape, generated on
    the device (seed = instance index; generation untimed), through the channel replay, the circle / line folds and the FRI tree rebuilds
    (K3, K5, K2; full mode: per-query roots + permutation record) -- next to n replicas of one instance, which is what a replica batch of
    real proofs looks like to these kernels (same positions, same node sharing, same witness consumption order in every lane)."""
    import torch
    out = {"config": "BASELINE configs[4] part i: %d synthetic FRI + Merkle instances, shape of small_proof.bin without proof of work; verify = channel replay + "
                     "folds + FRI tree rebuilds (8 trees, 2112 path permutations covered per instance); with_folding_tape = verify + the trace pass of the "
                     "folding-stage circuit (fri_answers as witnesses), checks on, value columns exported" % n}
    circ = pkg.VerifierCircuit(shape, folding=True)
    out["folding_circuit"] = {k: getattr(circ.info, k) for k in ("n_rows", "n_vars", "n_flow", "n_ins", "n_levels")}
    for name, distinct in (("distinct", True), ("replicas", False)):
        sb = pkg.SynthBatch(shape, n, seed0=0, distinct=distinct)
        for _ in range(2):
            v, _ = sb.run(full=True)
        assert int(v.sum().item()) == 0
        if distinct:
            for _ in range(2):
                r = circ.trace(sb, check=True, export=True, preprocessed=False)
            assert int((r["bad_row"] != -1).sum().item()) == 0 and int((r["bad_flow"] != -1).sum().item()) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                sb.run(full=True)
                circ.trace(sb, check=True, export=True, preprocessed=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out["with_folding_tape"] = {"ms_per_batch": ms, "instances_per_sec": n / (ms * 1e-3)}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sb.run(full=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        dt = sb.fetch(n // 2, "detail")
        out[name] = {"ms_per_batch": ms, "instances_per_sec": n / (ms * 1e-3), "perms_executed_per_instance": int(dt.n_perms_hints + dt.fs.n_transcript_perms)}
        del sb
        torch.cuda.empty_cache()
    del circ
    torch.cuda.empty_cache()
    out["distinct_over_replicas"] = out["distinct"]["instances_per_sec"] / out["replicas"]["instances_per_sec"]
    return out


def divergence_leg(pkg, dev, n=1024, reps=5):
    """What replicas hide: lanes of a warp hold different proofs in production.  The same shape (16, 15; 10 queries) verified and
    traced as n replicas of level10-1.bin and as n proofs alternating level10-1.bin / level11-1.bin (different query positions
    and decommitment layouts in neighbouring lanes)."""
    import torch
    d = os.path.join(ROOT, "tests", "golden", "proofs")
    a, b = (open(os.path.join(d, f), "rb").read() for f in ("level10-1.bin", "level11-1.bin"))
    out = {"shape": "level10-1 / level11-1 (2^16 Plonk rows, 2^15 Poseidon rows, 10 queries)", "proofs": n}
    circ = None
    for name, blobs in (("replicas", [a] * n), ("alternating", [a, b] * (n // 2))):
        vb = pkg.VerifyBatch(blobs, inputs=pkg.INPUTS_RECURSIVE)
        if circ is None:
            circ = pkg.VerifierCircuit(vb.shape, inputs=pkg.INPUTS_RECURSIVE)

        def step():
            v, _ = vb.run(full=True)
            return v, circ.trace(vb, check=True, export=True, preprocessed=False)
        for _ in range(2):
            v, r = step()
        assert int(v.sum().item()) == 0 and int((r["bad_row"] != -1).sum().item()) == 0 and int((r["bad_flow"] != -1).sum().item()) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            vb.run(full=True)
        e1.record()
        for _ in range(reps):
            step()
        e2 = torch.cuda.Event(enable_timing=True)
        e2.record()
        torch.cuda.synchronize()
        out[name] = {"verify_ms": e0.elapsed_time(e1) / reps, "verify_plus_trace_ms": e1.elapsed_time(e2) / reps,
                     "proofs_per_sec": n / (e1.elapsed_time(e2) / reps * 1e-3)}
        del vb
    out["alternating_over_replicas"] = out["alternating"]["proofs_per_sec"] / out["replicas"]["proofs_per_sec"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--proofs", type=int, default=4096, help="proofs per GPU per step (weak scaling) / in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --proofs per GPU; strong: --proofs in total, ceil(proofs / N) per GPU (BASELINE configs[4]: 4096 across 8 GPUs)")
    ap.add_argument("--lanes", type=int, default=None, help="pipeline lanes: independent verify / trace stream pairs (VerifyTracePipeline; default: its own choice)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the K1 / Merkle-sweep side measurements")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to fd 1 (NCCL's version banner) go to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("recursive-stwo_b200")
    sharding = importlib.import_module("recursive-stwo_b200.sharding")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg.init(local)
    warmup = max(args.warmup, 3)

    blob = load_fixture()
    n_total = args.proofs * world if args.scaling == "weak" else args.proofs
    lo, hi = sharding.shard_range(n_total, rank, world)
    # A stream of batches through the public API: VerifyTracePipeline = three device slots, the upload / verification / trace pass of
    # neighbouring batches on their own streams.  Every step verifies AND traces one whole batch; the device-resident leg skips the
    # upload, the end-to-end leg does everything.
    pipe = pkg.VerifyTracePipeline([blob] * (hi - lo), inputs=pkg.INPUTS_SINGLE, lanes=args.lanes)
    vb, circ = pipe.slots[0], pipe.circuit                              # the circuit is recorded once per shape (host)
    ci = circ.info
    ws_mb = (vb.ws_bytes + circ.workspace_bytes(hi - lo)) >> 20
    blob_mb = vb.h_words.numel() * 4 >> 20
    trace_mb = (hi - lo) * 13 * ci.n_rows * 4 >> 20

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(v, s):
        return sharding.gather_verdicts(v, s, n_total)

    # the 10 preprocessed columns depend on the shape only: written once, outside the steps
    v0, _ = vb.run(full=True)
    r0 = circ.trace(vb, check=True, export=True, preprocessed=True)
    torch.cuda.synchronize()
    assert int(v0.sum().item()) == 0 and int((r0["bad_row"] != -1).sum().item()) == 0 and int((r0["bad_flow"] != -1).sum().item()) == 0
    # spot parity inside the bench: the exported trace of one proof against the committed golden digest
    import hashlib
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "trace_digests.json")))["chain"][0]
    tr = pkg.VerifierCircuit.assemble_trace(r0["preprocessed"], r0["values"][(hi - lo) // 2])
    assert hashlib.sha256(np.ascontiguousarray(tr, dtype="<u4").tobytes()).hexdigest() == gold["trace_sha256"], "trace differs from the golden"
    dt0 = vb.fetch(0, "detail")
    # permutations EXECUTED per proof: transcript + tree rebuilds -- every distinct permutation of a proof once.  The per-query paths are
    # not hashed again (each node's states are handed to the queries whose path runs through it), the tape evaluation takes the circuit's
    # permutations from that record (STWO_B200_TRACE_NATIVE_HINTS), and check_poseidon_invocations compares each of the circuit's
    # ci.n_flow flow entries with the recorded (input, output) pair of the execution instead of executing it a third time.
    perms_per_proof = dt0.n_perms_hints + dt0.fs.n_transcript_perms
    assert dt0.n_perms_paths == 3481, "permutation count the record covers differs from the reference's circuit (SURVEY App. C)"

    def check_last(h):
        hv, hs, hb, hf = pipe.result(h)
        assert int(hv.sum()) == 0 and hv.numel() == n_total, "every replica of the fixture must be accepted"
        assert int((hb != -1).sum()) == 0 and int((hf != -1).sum()) == 0

    for _ in range(warmup):
        h = pipe.step(upload=False, gather=gather)
    pipe.join()
    check_last(h)

    # ---- timed region: device-resident ----------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = pkg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        h = pipe.step(upload=False, gather=gather)
    pipe.fence()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pkg.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    check_last(h)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = n_total * args.steps / (ms * 1e-3)

    # ---- e2e: pinned host blobs -> device -> verdicts and check results on the host, every step ----------
    # the same pipeline with its upload stage: every step copies one batch of blobs from pinned host memory (copy stream, beside the
    # kernels of the batches before it) and ends with the device->host copy of that batch's verdicts and check results into pinned
    # memory; the timed region ends when the last of those copies has landed.
    for _ in range(3):
        h = pipe.step(gather=gather)
    pipe.join()
    check_last(h)
    e2e_steps = max(3, args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h = pipe.step(gather=gather)
    pipe.join()
    dt = time.perf_counter() - t0
    check_last(h)
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(t_e.item())
    h2d = vb.h_words.numel() * 4 + vb.h_off.numel() * 8
    d2h = 2 * n_total + 16 * (hi - lo)

    # ---- stage breakdown + roofline of the dominant kernel (CUDA events between the stage kernels) -------
    pk = peaks()
    acc = {}
    reps = 3
    for _ in range(reps):
        vb.run(full=True, timed=True)
        for k, t in vb.stage_ms().items():
            acc[k] = acc.get(k, 0.0) + t / reps
        circ.trace(vb, check=True, export=True, preprocessed=False, timed=True)
        for k, t in circ.stage_ms().items():
            acc["trace_" + k] = acc.get("trace_" + k, 0.0) + t / reps
    # the same check by re-execution of every flow entry (STWO_B200_TRACE_RECHECK_POSEIDON), for the record
    r_re = circ.trace(vb, check=True, export=True, preprocessed=False, timed=True, recheck=True)
    recheck_ms = circ.stage_ms()["check_poseidon"]
    assert int((r_re["bad_flow"] != -1).sum().item()) == 0
    total_ms = sum(acc.values())
    dom = max(acc, key=acc.get)
    n_local = hi - lo
    sh = vb.shape
    kernel_of = {"trace_eval": "k_tape_eval_cluster", "trace_check_poseidon": "k_cs_check_poseidon", "trace_export": "k_cs_export_vals_stream",
                 "trace_check_arithmetics": "k_cs_check_arith", "trace_gather": "k_gather_witness", "fiat_shamir": "k_transcript16",
                 "single_tree": "k_single_tree_coop", "pair_tree": "k_pair_tree_coop", "folds": "k_folds_coop"}

    # the HBM-bound kernel of the path: trace export.  Algorithmic (compulsory) bytes per proof: variables[] read once (n_vars x 16 B) +
    # 13 value columns written (n_rows x 13 x 4 B); the ~2.6 uses of a variable are served from L2 / shared memory, not from HBM.
    export_bytes = n_local * (ci.n_vars * 16 + ci.n_rows * 13 * 4)
    export_gbs = export_bytes / (acc["trace_export"] * 1e-3) / 1e9
    export_traffic = ncu_traffic("k_cs_export_vals_stream", n_local)
    roofline_export = {"bound": "hbm", "kernel": "k_cs_export_vals_stream", "achieved": export_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                       "frac": export_gbs / pk["hbm_gbs"], "peak_src": pk["src"], "algorithmic_bytes_per_launch": export_bytes,
                       "launch_ms": acc["trace_export"], "share_of_step": acc["trace_export"] / total_ms,
                       "traffic": export_traffic,
                       "dram_frac": (export_traffic / (acc["trace_export"] * 1e-3) / 1e9 / pk["hbm_gbs"]) if export_traffic else None,
                       "note": "check_arithmetics is fused into this pass; algorithmic bytes = variables once + 13 columns per proof; dram_frac = ncu DRAM "
                               "bytes of a capture at this batch size / this run's launch time / peak (null without such a capture)"}

    # the busiest permutation kernel against the integer roofline (permutations each stage executes per proof: SURVEY App. C,
    # counted by the kernels themselves)
    single_paths = sum(pkg.path_perms(pkg.PathShape.make(d, lay)) for d, lay in _tree_shapes(sh)) * sh.n_queries
    pair_paths = pkg.proof_perms(sh) - single_paths
    share = single_paths / float(single_paths + pair_paths)
    perms_of = {"fiat_shamir": dt0.fs.n_transcript_perms, "single_path": single_paths, "pair_path": pair_paths,
                "single_tree": dt0.n_perms_hints * share, "pair_tree": dt0.n_perms_hints * (1 - share)}
    hdom = max(perms_of, key=lambda k: acc.get(k, 0.0))
    h_perms = perms_of[hdom] * n_local
    h_rate = h_perms / (acc[hdom] * 1e-3)
    achieved = h_rate * LANE_OPS_PER_PERM / 1e12
    roofline_hashing = {
        "bound": "int32-issue", "kernel": kernel_of.get(hdom, "k_" + hdom), "achieved": achieved, "peak": pk["int_tlops"], "unit": "T lane-ops/s",
        "frac": achieved / pk["int_tlops"], "peak_src": pk["int_src"], "lane_ops_per_perm": LANE_OPS_PER_PERM,
        "perms_per_launch": h_perms, "perms_per_sec": h_rate, "launch_ms": acc[hdom], "share_of_step": acc[hdom] / total_ms,
        "traffic": ncu_traffic(kernel_of.get(hdom, ""), n_local),
        "check_poseidon_by_reexecution": {"launch_ms": recheck_ms, "perms_per_sec": ci.n_flow * n_local / (recheck_ms * 1e-3),
                                          "frac": ci.n_flow * n_local / (recheck_ms * 1e-3) * LANE_OPS_PER_PERM / 1e12 / pk["int_tlops"]},
        "note": "the tree rebuild is a layer-parallel walk with G lanes per tree (idle lanes at narrow layers, 128 registers, 16 warps per SM) that "
                "also writes the permutation record; the bare permutation kernel K1 reaches 0.63 of the peak (secondary.k1_frac_of_int_peak) and "
                "check_poseidon_invocations by re-execution 0.64 (check_poseidon_by_reexecution; ncu: profiles/r02e_k_cs_check_poseidon_ncu.txt)"}
    # `roofline` = the dominant HBM-bound kernel of the step (the trace export); `roofline_hashing` = the dominant integer-bound kernel
    # (for this path the largest kernels are co-dominant: check_poseidon_invocations, export, tape evaluation within ~10 % of each other)
    roofline = dict(roofline_export)
    roofline["largest_kernel_of_step"] = kernel_of.get(dom, "k_" + dom)
    # the tape evaluation with native hints moves variables only: what one launch reads and writes per proof (operands + results of
    # every tape instruction, 16 B each, + 64 B hint + 192 B flow record per permutation) against its launch time
    eval_bytes = n_local * ((ci.n_ins - ci.n_flow) * 48 + ci.n_flow * (64 + 64 + 192 + 64))
    eval_traffic = ncu_traffic("k_tape_eval_cluster", n_local)
    roofline_eval = {"bound": "hbm", "kernel": "k_tape_eval_cluster", "achieved": eval_bytes / (acc["trace_eval"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                     "unit": "GB/s", "frac": eval_bytes / (acc["trace_eval"] * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes_per_launch": eval_bytes,
                     "launch_ms": acc["trace_eval"], "share_of_step": acc["trace_eval"] / total_ms, "traffic": eval_traffic,
                     "note": "operand traffic counted at 16 B per use: most uses are L2 hits, so frac overstates DRAM use; the pass is latency bound "
                             "(dependent levels), see DESIGN.md"}
    roofline["stage_ms"] = acc

    secondary = {}
    if not args.no_secondary and rank == 0:
        n_states = 1 << 22
        st = torch.randint(0, 2**31 - 1, (n_states, 16), dtype=torch.int32, device=dev)
        for _ in range(3):
            pkg.poseidon2_permute(st)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(5):
            pkg.poseidon2_permute(st)
        k1.record()
        torch.cuda.synchronize()
        k1_rate = n_states * 5 / (k0.elapsed_time(k1) * 1e-3)
        del st
        secondary = {"k1_permute_perms_per_sec": k1_rate, "k1_frac_of_int_peak": k1_rate * 4264 / 1e12 / pk["int_tlops"],
                     "merkle_sweep_config": "BASELINE configs[2]: T trees x 2^20 leaves, leaf width C, Q queries per tree; commit and "
                                            "decommit+path-verify timed apart (SURVEY.md 8d config 3)",
                     "merkle_sweep": [merkle_sweep(pkg, dev, C=C, Q=Q, T=T, hbm_peak=pk["hbm_gbs"])
                                      for C, Q, T in ((4, 16, 2), (8, 128, 2), (50, 64, 2), (60, 16, 2), (60, 128, 2), (8, 32, 1), (4, 32, 64))]}
        secondary["merkle_sweep_perms_per_sec"] = secondary["merkle_sweep"][1]["perms_per_sec"]
        secondary["lane_divergence"] = divergence_leg(pkg, dev)
        secondary["synthetic_4096"] = synthetic_leg(pkg, dev, pkg.shape_from_config(pkg.PcsConfig(0, sh.log_blowup, sh.log_last, sh.n_queries),
                                                                                    sh.log_size_plonk, sh.log_size_poseidon))
        secondary["shape_R"] = shape_leg(pkg, dev, "recursive_proof_16_15.bin", 1024)

    if not args.no_secondary:
        mp = multi_proofs_leg(pkg, sharding, rank, world, dev)            # every rank takes its block of the 256
        if rank == 0:
            secondary["multi_proofs"] = mp
        if world > 1:
            vb.run(full=True)
            rr = circ.trace(vb, check=True, export=True, preprocessed=False)
            tg = trace_gather_leg(sharding, rr["values"], rank, world, dev)
            del rr
            if rank == 0:
                secondary["trace_gather"] = tg
    if not args.no_secondary and rank == 0:
        # BASELINE configs[1] (examples/last-layer): the last-layer circuit's trace for 256 replicas of the Poseidon31 twin of
        # hybrid_hash.bin (Plonk-without-Poseidon system, emulated Poseidon2, 2^17 rows x 20 columns), verification included
        lblob = open(os.path.join(ROOT, "tests", "golden", "proofs", "level13-1.bin"), "rb").read()
        lvb = pkg.VerifyBatch([lblob] * 256, inputs=pkg.INPUTS_RECURSIVE)
        lcirc = pkg.VerifierCircuit(lvb.shape, last_layer=True)

        def last_step(pre=False):
            lv, _ = lvb.run(full=True)
            return lv, lcirc.trace(lvb, check=True, export=True, preprocessed=pre)
        for _ in range(3):
            lv, lr = last_step(pre=True)
        assert int(lv.sum().item()) == 0 and int((lr["bad_row"] != -1).sum().item()) == 0
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(5):
            last_step()
        k1.record()
        torch.cuda.synchronize()
        lms = k0.elapsed_time(k1) / 5
        lcirc.trace(lvb, check=True, export=True, preprocessed=False, timed=True)
        secondary["last_layer"] = {"config": "BASELINE configs[1]: 256 replicas of level13-1.bin, verify + last-layer circuit trace",
                                   "rows": lcirc.info.n_rows, "columns": 20, "tape_levels": lcirc.info.n_levels,
                                   "ms_per_batch": lms, "proofs_per_sec": 256 / (lms * 1e-3), "trace_stage_ms": lcirc.stage_ms()}
        del lvb, lcirc

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32 (M31)", "data": "synthetic",
            "config": {"workload": "verify-batch", "fixture": FIXTURE, "proofs_per_gpu": hi - lo, "proofs_total": n_total, "mode": "full+trace",
                       "circuit": {"rows": ci.n_rows, "rows_unpadded": ci.n_rows_unpadded, "variables": ci.n_vars, "poseidon_flow": ci.n_flow,
                                   "tape_levels": ci.n_levels, "witness_words": ci.n_input_words},
                       "shape": dict(zip(("log_size_plonk", "log_size_poseidon", "pow_bits", "log_blowup", "log_last", "n_queries", "n_inner"), sh.key())),
                       "perms_per_proof": perms_per_proof,
                       "perms_compared_with_record_per_proof": ci.n_flow,
                       "l2": "inputs larger than L2: %d MB of proof blobs + %d MB of workspace + %d MB of trace columns per step" % (blob_mb, ws_mb, trace_mb),
                       "parallelism": "proofs sharded by rank in contiguous blocks; NCCL all-gather of verdict bytes only",
                       "pipeline": "%d device slots, %d lane(s); upload | verification | trace pass of neighbouring steps on their own streams; "
                                   "CUDA_DEVICE_MAX_CONNECTIONS=%s" % (len(pipe.slots), len(pipe.lane_streams), os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))},
            "poseidon31_perms_per_sec": value * perms_per_proof,
            # what the reference's own path executes for the same proofs: hints (every tree node once) + transcript, the DSL's per-query
            # paths + transcript again, check_poseidon_invocations a third time -- the work this rate of proofs stands for
            "poseidon31_perms_reference_equivalent_per_sec": value * (perms_per_proof + 2 * ci.n_flow),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e.item()) / e2e_steps * 1e3},
            "roofline": roofline,
            "roofline_export": roofline_export,
            "roofline_hashing": roofline_hashing,
            "roofline_eval": roofline_eval,
            "secondary": secondary,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rate, prate, reps_c, dtc = cpu_rate(max(cores * 4, 64), args.cpu_seconds, cores)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "poseidon31_perms_per_sec": prate,
                                   "sample": "%d x %d replicas of %s with the oracle port (native verifier + circuit value-log replay, checks, "
                                             "export) on %d pthreads (%.1f s)" % (reps_c, max(cores * 4, 64), FIXTURE, cores, dtc)}
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _tree_shapes(sh):
    """(depth, {log_size: n_cols}) of the four commitment trees of a proof shape"""
    lp, ls, mf = sh.log_size_plonk + sh.log_blowup, sh.log_size_poseidon + sh.log_blowup, sh.max_first
    out = []
    for a, b in ((10, 40), (12, 48), (8, 8)):
        lay = {}
        lay[lp] = lay.get(lp, 0) + a
        lay[ls] = lay.get(ls, 0) + b
        out.append((max(lp, ls), lay))
    out.append((mf, {mf: 8}))
    return out


if __name__ == "__main__":
    main()
