"""The recursive verifier circuit over the C ABI: record once per proof shape (host logic, no device), then generate the
circle-Plonk trace of a whole verified batch on the device.

Host-side mirror of the circuit half of the reference drivers
  examples/single-proof/src/main.rs:48-90   (PlonkWithPoseidonProofVar::new_witness .. FoldingResults::compute, cs.pad(),
                                             check_arithmetics, populate_logup_arguments, check_poseidon_invocations,
                                             generate_plonk_with_poseidon_circuit),
  examples/multi-proofs/src/main.rs:62-139  (the same with `multipliers` verifications inside one constraint system).
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import (CircuitInfo, COLUMNS, CFETCH, TRACE_CHECK_ARITHMETICS, TRACE_CHECK_POSEIDON, TRACE_TIMED, TRACE_NATIVE_HINTS, TRACE_RECHECK_POSEIDON,
                   TRACE_STAGES)
from .verifier import INPUTS_RECURSIVE

COLUMN_NAMES = ("mult_a", "mult_b", "mult_c", "poseidon_wire", "mult_poseidon", "enforce_c_m31", "a_wire", "b_wire", "c_wire", "op",
                "a_val_0", "a_val_1", "a_val_2", "a_val_3", "b_val_0", "b_val_1", "b_val_2", "b_val_3", "c_val_0", "c_val_1", "c_val_2",
                "c_val_3")          # PlonkWithAcceleratorCircuitTrace field order (plonk_with_poseidon.rs:542-619)


COLUMN_NAMES_WITHOUT = ("mult_c", "a_wire", "b_wire", "c_wire", "op1", "op2", "op3", "op4", "a_val_0", "a_val_1", "a_val_2", "a_val_3",
                        "b_val_0", "b_val_1", "b_val_2", "b_val_3", "c_val_0", "c_val_1", "c_val_2", "c_val_3")
# PlonkWithoutAcceleratorCircuitTrace field order (plonk_without_poseidon.rs:645-708)


class VerifierCircuit:
    """A recorded verifier circuit for one proof shape.  last_layer=True records the circuit of examples/last-layer
    (components/last/*, Plonk-without-Poseidon system, emulated Poseidon2) instead of the recursive verifier; folding=True the folding
    stage alone over witnesses (FoldingResults::compute; BASELINE configs[4] part i), which traces a verified SynthBatch or VerifyBatch."""

    def __init__(self, shape, inputs=INPUTS_RECURSIVE, multipliers=1, last_layer=False, folding=False):
        h = ctypes.c_void_p()
        if folding:
            _lib.call("stwo_b200_circuit_record_folding", ctypes.byref(shape), ctypes.byref(h))
        elif last_layer:
            _lib.call("stwo_b200_circuit_record_last_layer", ctypes.byref(shape), ctypes.byref(h))
        else:
            idx = np.ascontiguousarray(inputs[0], dtype=np.uint32)
            vals = np.ascontiguousarray(inputs[1], dtype=np.uint32)
            _lib.call("stwo_b200_circuit_record_verifier", ctypes.byref(shape), idx.ctypes.data_as(ctypes.c_void_p),
                      vals.ctypes.data_as(ctypes.c_void_p), idx.size, multipliers, ctypes.byref(h))
        self._h = h
        self.shape, self.multipliers = shape, multipliers
        info = CircuitInfo()
        _lib.call("stwo_b200_circuit_get_info", self._h, ctypes.byref(info))
        self.info = info
        self._ws = None

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.load().stwo_b200_circuit_free(self._h)
            except Exception:                        # interpreter shutdown: the module globals are already gone
                pass
            self._h = None

    def column(self, name):
        """host copy of one recorded wiring column (uint32)"""
        i = self.info
        n = {"flow_wire": 4 * i.n_flow, "flow_swap_addr": i.n_flow, "level_start": i.n_levels + 1}.get(name, i.n_rows)
        out = np.zeros(n, dtype=np.uint32)
        _lib.call("stwo_b200_circuit_get_column", self._h, COLUMNS[name], out.ctypes.data_as(ctypes.c_void_p), n)
        return out

    def workspace_bytes(self, n_proofs):
        return int(_lib.load().stwo_b200_circuit_workspace_bytes(self._h, n_proofs))

    def trace(self, batch, check=True, export=True, preprocessed=True, timed=False, native_hints=None, recheck=False):
        """Trace generation for a batch that `batch.run()` has verified (VerifyBatch keeps the hints in its workspace).
        Returns dict(values=[n, 13, n_rows] | None, preprocessed=[10, n_rows] | None, bad_row=[n] | None, bad_flow=[n] | None)
        as torch tensors on the batch's device.  `values` is ONE buffer per circuit object (13.3 GB for 4096 proofs of shape S), reused
        by every call: the next trace() overwrites it -- consume or copy it first."""
        import torch
        from .hashing import _dptr, _stream
        dev = batch.d_words.device
        n, nr = batch.n, self.info.n_rows
        need = self.workspace_bytes(n)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._n = n
        out = dict(values=None, preprocessed=None, bad_row=None, bad_flow=None)
        if export:
            if getattr(self, "_values", None) is None or self._values.shape[0] != n or self._values.device != dev:
                self._values = torch.empty((n, 13, nr), dtype=torch.int32, device=dev)
            out["values"] = self._values
        if preprocessed:
            out["preprocessed"] = torch.empty((self.info.n_preprocessed_columns, nr), dtype=torch.int32, device=dev)
        if check:
            out["bad_row"] = torch.empty(n, dtype=torch.int64, device=dev)
            out["bad_flow"] = torch.empty(n, dtype=torch.int64, device=dev)
        flags = ((TRACE_CHECK_ARITHMETICS | TRACE_CHECK_POSEIDON) if check else 0) | (TRACE_TIMED if timed else 0)
        if native_hints if native_hints is not None else getattr(batch, "last_full", False):
            flags |= TRACE_NATIVE_HINTS
        if recheck:                                  # check_poseidon_invocations re-executes every entry instead of comparing with the record
            flags |= TRACE_RECHECK_POSEIDON
        p = lambda t: _dptr(t) if t is not None else None
        _lib.call("stwo_b200_circuit_trace_batch_dev", self._h, _dptr(batch.d_words), _dptr(batch.d_off), n, _dptr(batch.d_ws), _dptr(self._ws),
                  self._ws.numel(), flags, p(out["preprocessed"]), p(out["values"]), p(out["bad_row"]), p(out["bad_flow"]), _stream())
        return out

    def export_flow(self, pad_constants):
        """The prover-facing PoseidonFlow of the batch the last trace() call traced, padded like pad() with the
        caller's CONSTANT_1 / CONSTANT_2 / CONSTANT_3 (pad_constants: [3, 8] M31 words; plonk_with_poseidon.rs:13-15,296-321).
        -> dict(hash=[n, n_flow_padded, 32] int32, swap=[n, n_flow_padded] uint8 on the device,
                wire=[n_flow_padded, 4], swap_addr=[n_flow_padded] numpy: the wiring part, zero for the padding)"""
        import torch
        from .hashing import _dptr, _stream
        n, n_pad = self._n, self.info.n_flow_padded
        dev = self._ws.device
        h = torch.empty((n, n_pad, 32), dtype=torch.int32, device=dev)
        sw = torch.empty((n, n_pad), dtype=torch.uint8, device=dev)
        pc = np.ascontiguousarray(pad_constants, dtype=np.uint32).reshape(24)
        _lib.call("stwo_b200_circuit_export_flow_dev", self._h, n, _dptr(self._ws), self._ws.numel(), pc.ctypes.data_as(ctypes.c_void_p), _dptr(h),
                  _dptr(sw), _stream())
        wire = np.zeros((n_pad, 4), dtype=np.uint32)
        addr = np.zeros(n_pad, dtype=np.uint32)
        wire[: self.info.n_flow] = self.column("flow_wire").reshape(-1, 4)
        addr[: self.info.n_flow] = self.column("flow_swap_addr")
        return dict(hash=h, swap=sw, wire=wire, swap_addr=addr)

    def stage_ms(self):
        ms = (ctypes.c_float * len(TRACE_STAGES))()
        _lib.call("stwo_b200_circuit_stage_ms", ms)
        return dict(zip(TRACE_STAGES, [float(x) for x in ms]))

    def fetch(self, p, what):
        """one proof's variables [n_vars, 4] / flow_hash [n_flow, 32] / flow_swap [n_flow] / witness [n_input_words] (host)"""
        from .hashing import _dptr, _stream
        i = self.info
        shape, dt = {"variables": ((i.n_vars, 4), np.uint32), "flow_hash": ((i.n_flow, 32), np.uint32), "flow_swap": ((i.n_flow,), np.uint8),
                     "witness": ((i.n_input_words,), np.uint32)}[what]
        out = np.zeros(shape, dtype=dt)
        _lib.call("stwo_b200_circuit_fetch", self._h, _dptr(self._ws), self._n, p, CFETCH[what], out.ctypes.data_as(ctypes.c_void_p), out.nbytes,
                  _stream())
        return out

    @staticmethod
    def assemble_trace(preprocessed, values_of_one):
        """22 trace columns of one proof in the reference's order from the shared preprocessed block and its 13 value columns"""
        pre = preprocessed.cpu().numpy().view(np.uint32)
        val = values_of_one.cpu().numpy().view(np.uint32)
        if pre.shape[0] == 8:                    # Plonk-without-Poseidon: mult_c, wires, op1 (per proof), op2..op4, 12 values
            out = np.empty((20, pre.shape[1]), dtype=np.uint32)
            out[:4] = pre[:4]
            out[4] = val[12]
            out[5:8] = pre[5:8]
            out[8:] = val[:12]
            return out
        out = np.empty((22, pre.shape[1]), dtype=np.uint32)
        out[:9] = pre[:9]
        out[9] = val[12]
        out[10:] = val[:12]
        return out


_CIRCUITS = {}


def cached_circuit(shape, inputs=INPUTS_RECURSIVE, multipliers=1, last_layer=False):
    """VerifierCircuit per (shape, public inputs, multipliers): recording is host work that depends on the shape only
    (examples/multi-proofs/src/main.rs:62-139 re-records for every proof; here once per shape and process)."""
    key = (tuple(shape.key()), tuple(inputs[0]), tuple(map(tuple, inputs[1])), multipliers, last_layer)
    if key not in _CIRCUITS:
        _CIRCUITS[key] = VerifierCircuit(shape, inputs=inputs, multipliers=multipliers, last_layer=last_layer)
    return _CIRCUITS[key]


class MixedBatch:
    """A batch of proofs of ANY mix of shapes kept resident on the device, verified and traced shape group by shape group: the
    host-side mirror of examples/multi-proofs/src/main.rs:173-296 run as one batch (BASELINE configs[3]).  Every proof's verdict
    comes back in the caller's order; the traces stay on the device per shape group (`groups[k].trace`).
    config: allow-list of PcsConfigs (a blob claiming another one, or not parsing at all, is rejected at stage parse)."""

    class Group:
        pass

    def __init__(self, blobs, inputs=INPUTS_RECURSIVE, config=None, multipliers=1):
        from .verifier import REFERENCE_CONFIGS, VerifyBatch, proof_shape, shape_for
        config = REFERENCE_CONFIGS if config is None else config
        self.n = len(blobs)
        by_shape, self.unparsed = {}, []
        for i, b in enumerate(blobs):
            try:
                by_shape.setdefault(tuple(shape_for(b, config).key()), []).append(i)
            except ValueError:
                self.unparsed.append(i)
        self.groups = []
        for key, ids in sorted(by_shape.items(), key=lambda kv: -len(kv[1])):
            g = MixedBatch.Group()
            g.ids = np.array(ids, dtype=np.int64)
            g.batch = VerifyBatch([blobs[i] for i in ids], inputs=inputs, config=config)
            g.circuit = cached_circuit(g.batch.shape, inputs=inputs, multipliers=multipliers)
            g.trace = None
            self.groups.append(g)
        # submission order = longest chain first: a group is a chain of latency-bound kernels whose length grows with the shape (queries,
        # tree depth, tape levels), and submitting a group costs the host a fraction of a millisecond -- the heaviest shape must not
        # be the one that starts last
        if os.environ.get("STWO_B200_MIXED_ORDER", "heavy") == "heavy":
            self.groups.sort(key=lambda g: -(g.circuit.info.n_flow + g.circuit.info.n_rows / 16.0))

    def cost(self):
        """relative cost of each group's proofs (permutations of the flow + rows / 16), for work-balanced sharding"""
        return {tuple(g.batch.shape.key()): g.circuit.info.n_flow + g.circuit.info.n_rows / 16.0 for g in self.groups}

    def run(self, trace=True, export=True, concurrent=True):
        """-> (verdict uint8[n], stage uint8[n]) as torch tensors on the device, in the caller's order.  With trace: a proof whose circuit
        checks fail (check_arithmetics / check_poseidon_invocations) counts as rejected.
        concurrent: every shape group runs on its own stream.  A group of a few dozen proofs is a chain of latency-bound kernels (the
        transcript is 100-255 sequential permutations, the tape 300-400 dependent levels) that leaves most of the GPU idle; the groups
        of a mixed batch fill it beside each other, and the batch takes about as long as its slowest group."""
        import torch
        dev = self.groups[0].batch.d_words.device if self.groups else torch.device("cuda", torch.cuda.current_device())
        verdict = torch.ones(self.n, dtype=torch.uint8, device=dev)
        stage = torch.ones(self.n, dtype=torch.uint8, device=dev)          # unparsed: (reject, parse)
        main = torch.cuda.current_stream(dev)
        if concurrent and len(self.groups) > 1:
            if not hasattr(self, "_streams"):
                self._streams = [torch.cuda.Stream(dev) for _ in self.groups]
                self._idx = [torch.from_numpy(g.ids).to(dev) for g in self.groups]
            fork = torch.cuda.Event()
            fork.record(main)
        for k, g in enumerate(self.groups):
            st = self._streams[k] if concurrent and len(self.groups) > 1 else main
            if st is not main:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                v, s = g.batch.run(full=True)
                if trace:
                    g.trace = g.circuit.trace(g.batch, check=True, export=export, preprocessed=False)
                    bad = (g.trace["bad_row"] != -1) | (g.trace["bad_flow"] != -1)
                    v = torch.where(bad & (v == 0), torch.full_like(v, 1), v)
                idx = self._idx[k] if hasattr(self, "_idx") else torch.from_numpy(g.ids).to(dev)
                verdict[idx] = v
                stage[idx] = s
            if st is not main:
                done = torch.cuda.Event()
                done.record(st)
                main.wait_event(done)
        return verdict, stage


class VerifyTracePipeline:
    """A stream of same-shape batches through upload, verification and trace generation, the three stages of neighbouring batches
    running beside each other on their own streams (copy | verify | trace):

        pipe = VerifyTracePipeline(first_blobs, inputs)        # n_slots device slots of that batch size
        for blobs in batches: h = pipe.step(blobs)             # enqueues everything for one batch, returns at once
        pipe.join(); verdict, stage, bad_row, bad_flow = pipe.result(h)     # host (pinned) tensors of that step

    Verification is a set of latency-bound chains and integer-bound tree hashing, the trace pass alternates memory-bound and
    integer-bound kernels: batch k+1's verification fills what batch k's trace pass leaves idle (measured on B200: 15.1 -> 14.1 ms
    per 4096-proof batch, 3.70 -> 2.92 ms per 512-proof batch).  Possible because no kernel of either pass is a cooperative grid.
    The trace columns of a step (pipe.values, on the device) are valid from that step's `traced` event until the next step's trace
    pass starts: a consumer on another stream waits for pipe.traced[h] and records its own event into pipe.consumed before the next
    step() (None: nothing to wait for)."""

    def __init__(self, blobs, inputs=INPUTS_RECURSIVE, config=None, n_slots=None, check=True, export=True, lanes=None):
        """lanes: independent (verify stream, trace stream, circuit workspace + trace buffer) sets; batch k runs on lane k % lanes, so with
        several lanes the latency-bound kernels of several batches' verifications (and trace passes) overlap as well -- what small batches
        need (512 proofs of shape S on B200: 2.64 / 2.34 / 2.18 ms with 1 / 2 / 3 lanes; 4096 proofs: no difference).  Default: 3 lanes up
        to 2048 proofs per batch, 1 above; n_slots = 2 * lanes + 1 keeps every lane busy.  This needs the hardware work queues the package
        asks for on import (CUDA_DEVICE_MAX_CONNECTIONS=32): with the driver's 8, lanes alias onto one queue and gain nothing."""
        import torch
        if lanes is None:
            lanes = 3 if len(blobs) <= 2048 else 1
        if n_slots is None:
            n_slots = 2 * lanes + 1
        from .verifier import REFERENCE_CONFIGS, VerifyBatch
        config = REFERENCE_CONFIGS if config is None else config
        self.slots = [VerifyBatch(blobs, inputs=inputs, config=config) for _ in range(n_slots)]
        self.shape, self.n = self.slots[0].shape, self.slots[0].n
        self.circuits = [VerifierCircuit(self.shape, inputs=inputs) for _ in range(lanes)]
        self.circuit = self.circuits[0]
        self.check, self.export = check, export
        dev = self.slots[0].d_words.device
        self.s_copy = torch.cuda.Stream(dev)
        self.lane_streams = [(torch.cuda.Stream(dev), torch.cuda.Stream(dev)) for _ in range(lanes)]
        self.s_verify, self.s_trace = self.lane_streams[0]
        self.uploaded = [torch.cuda.Event() for _ in range(n_slots)]
        self.verified = [torch.cuda.Event() for _ in range(n_slots)]
        self.traced = [torch.cuda.Event() for _ in range(n_slots)]
        self.consumed = None
        self.host = [dict(verdict=torch.empty(self.n, dtype=torch.uint8).pin_memory(), stage=torch.empty(self.n, dtype=torch.uint8).pin_memory(),
                          bad_row=torch.empty(self.n, dtype=torch.int64).pin_memory(), bad_flow=torch.empty(self.n, dtype=torch.int64).pin_memory())
                     for _ in range(n_slots)]
        self.values = None
        self._k = 0
        cur = torch.cuda.current_stream(dev)
        for st in [self.s_copy] + [x for pair in self.lane_streams for x in pair]:
            st.wait_stream(cur)
        for e in self.traced:
            e.record(self.s_trace)
        self.values_of_lane = [None] * lanes

    def step(self, blobs=None, gather=None, upload=True):
        """enqueue one batch: upload (blobs=None re-sends the slot's pinned host copy; upload=False: the slot's device copy is used as it
        is), verify, trace, results to pinned host memory.
        gather: optional callable (verdict, stage) -> (verdict, stage) run on the trace stream (the NCCL all-gather of a sharded job)."""
        import torch
        i = self._k % len(self.slots)
        slot = self.slots[i]
        lane = self._k % len(self.lane_streams)
        s_verify, s_trace = self.lane_streams[lane]
        circuit = self.circuits[lane]
        if blobs is not None:
            from .verifier import _as_aligned
            if self._k >= len(self.slots):
                self.uploaded[i].synchronize()         # the slot's pinned buffers are the source of its previous upload
            words = np.concatenate([np.frombuffer(_as_aligned(b)[0].tobytes(), dtype=np.uint32) for b in blobs])
            off = np.zeros(len(blobs) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([(len(b) + 3) // 4 for b in blobs])
            if len(blobs) != slot.n or words.size > slot.h_words.numel():
                raise ValueError("a pipeline takes batches of the size and (at most) the byte length it was built with")
            slot.h_words[: words.size].copy_(torch.from_numpy(words.view(np.int32)))
            slot.h_off.copy_(torch.from_numpy(off.view(np.int64)))
        with torch.cuda.stream(self.s_copy):
            self.s_copy.wait_event(self.traced[i])     # the slot's previous trace pass is done with its blobs and workspace
            if upload:
                slot.upload()
            self.uploaded[i].record(self.s_copy)
        with torch.cuda.stream(s_verify):
            s_verify.wait_event(self.uploaded[i])
            v, s = slot.run(full=True)
            self.verified[i].record(s_verify)
        with torch.cuda.stream(s_trace):
            s_trace.wait_event(self.verified[i])
            if self.consumed is not None:
                s_trace.wait_event(self.consumed)
            r = circuit.trace(slot, check=self.check, export=self.export, preprocessed=False)
            self.values = self.values_of_lane[lane] = r["values"]
            h = self.host[i]
            if self.check:
                bad = (r["bad_row"] != -1) | (r["bad_flow"] != -1)
                v = torch.where(bad & (v == 0), torch.full_like(v, 1), v)
                h["bad_row"].copy_(r["bad_row"], non_blocking=True)
                h["bad_flow"].copy_(r["bad_flow"], non_blocking=True)
            if gather is not None:
                v, s = gather(v, s)
                if h["verdict"].numel() != v.numel():
                    h["verdict"], h["stage"] = torch.empty(v.numel(), dtype=torch.uint8).pin_memory(), torch.empty(v.numel(), dtype=torch.uint8).pin_memory()
            h["verdict"].copy_(v, non_blocking=True)
            h["stage"].copy_(s, non_blocking=True)
            self.traced[i].record(s_trace)
        self._k += 1
        return i

    def fence(self, stream=None):
        """make `stream` (default: the current one) wait for everything enqueued so far"""
        import torch
        stream = stream if stream is not None else torch.cuda.current_stream(self.slots[0].d_words.device)
        for st in [self.s_copy] + [x for pair in self.lane_streams for x in pair]:
            stream.wait_stream(st)

    def join(self):
        """wait (host) until everything enqueued so far has finished"""
        for st in [self.s_copy] + [x for pair in self.lane_streams for x in pair]:
            st.synchronize()

    def result(self, handle):
        h = self.host[handle]
        return h["verdict"], h["stage"], h["bad_row"], h["bad_flow"]
