"""The recursive verifier circuit over the C ABI: record once per proof shape (host logic, no device), then generate the
circle-Plonk trace of a whole verified batch on the device.

Host-side mirror of the circuit half of the reference drivers
  examples/single-proof/src/main.rs:48-90   (PlonkWithPoseidonProofVar::new_witness .. FoldingResults::compute, cs.pad(),
                                             check_arithmetics, populate_logup_arguments, check_poseidon_invocations,
                                             generate_plonk_with_poseidon_circuit),
  examples/multi-proofs/src/main.rs:62-139  (the same with `multipliers` verifications inside one constraint system).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (CircuitInfo, COLUMNS, CFETCH, TRACE_CHECK_ARITHMETICS, TRACE_CHECK_POSEIDON, TRACE_TIMED, TRACE_NATIVE_HINTS, TRACE_STAGES)
from .verifier import INPUTS_RECURSIVE

COLUMN_NAMES = ("mult_a", "mult_b", "mult_c", "poseidon_wire", "mult_poseidon", "enforce_c_m31", "a_wire", "b_wire", "c_wire", "op",
                "a_val_0", "a_val_1", "a_val_2", "a_val_3", "b_val_0", "b_val_1", "b_val_2", "b_val_3", "c_val_0", "c_val_1", "c_val_2",
                "c_val_3")          # PlonkWithAcceleratorCircuitTrace field order (plonk_with_poseidon.rs:542-619)


COLUMN_NAMES_WITHOUT = ("mult_c", "a_wire", "b_wire", "c_wire", "op1", "op2", "op3", "op4", "a_val_0", "a_val_1", "a_val_2", "a_val_3",
                        "b_val_0", "b_val_1", "b_val_2", "b_val_3", "c_val_0", "c_val_1", "c_val_2", "c_val_3")
# PlonkWithoutAcceleratorCircuitTrace field order (plonk_without_poseidon.rs:645-708)


class VerifierCircuit:
    """A recorded verifier circuit for one proof shape.  last_layer=True records the circuit of examples/last-layer
    (components/last/*, Plonk-without-Poseidon system, emulated Poseidon2) instead of the recursive verifier."""

    def __init__(self, shape, inputs=INPUTS_RECURSIVE, multipliers=1, last_layer=False):
        h = ctypes.c_void_p()
        if last_layer:
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

    def trace(self, batch, check=True, export=True, preprocessed=True, timed=False, native_hints=None):
        """Trace generation for a batch that `batch.run()` has verified (VerifyBatch keeps the hints in its workspace).
        Returns dict(values=[n, 13, n_rows] | None, preprocessed=[10, n_rows] | None, bad_row=[n] | None, bad_flow=[n] | None)
        as torch tensors on the batch's device."""
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
