"""ctypes binding of libstwo_b200.so (the C ABI in include/stwo_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present,
every product entry point raises.  The library is built in-tree by
`make -C recursive-stwo_b200/csrc` (see __graft_entry__.build).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstwo_b200.so")
MAX_DEPTH = 32

OK = 0
E_NO_DEVICE = -1000
E_BAD_ARG = -1001
E_SHAPE = -1002
E_NO_NCCL = -1003
E_NCCL = -1004


class PathShape(ctypes.Structure):
    """stwo_b200_path_shape"""
    _fields_ = [("depth", ctypes.c_uint32), ("n_cols", ctypes.c_uint32 * (MAX_DEPTH + 1))]

    @classmethod
    def make(cls, depth, cols_by_log_size):
        """cols_by_log_size: {log_size: n_columns}; the leaf layer (log_size == depth) must be present."""
        s = cls()
        s.depth = depth
        for h, n in cols_by_log_size.items():
            if not 0 <= h <= depth:
                raise ValueError("column log size outside the tree")
            s.n_cols[h] = n
        if s.n_cols[depth] == 0:
            raise ValueError("leaf layer has no columns")
        return s

    def cols_per_path(self):
        return sum(self.n_cols[h] for h in range(self.depth + 1))


class StwoB200Error(RuntimeError):
    def __init__(self, fn, status):
        self.status = status
        if status <= -1000:
            what = {E_NO_DEVICE: "no sm_100 CUDA device / library not initialised", E_BAD_ARG: "bad argument",
                    E_SHAPE: "unsupported shape", E_NO_NCCL: "libnccl.so.2 not loadable", E_NCCL: "NCCL call failed"}.get(status, "error")
        else:
            what = "cudaError %d" % (-status)
        super().__init__("%s failed: %s (status %d)" % (fn, what, status))


class ProofShape(ctypes.Structure):
    """stwo_b200_proof_shape"""
    _fields_ = [(n, ctypes.c_uint32) for n in ("log_size_plonk", "log_size_poseidon", "pow_bits", "log_blowup", "log_last",
                                               "n_queries", "n_inner")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)

    @property
    def max_first(self):
        return self.log_last + self.log_blowup + 1 + self.n_inner


class PcsConfig(ctypes.Structure):
    """stwo_b200_pcs_config: the commitment-scheme parameters the CALLER fixes (stwo PcsConfig / FriConfig)"""
    _fields_ = [(n, ctypes.c_uint32) for n in ("pow_bits", "log_blowup", "log_last", "n_queries")]

    def key(self):
        return (self.pow_bits, self.log_blowup, self.log_last, self.n_queries)

    def __repr__(self):
        return "PcsConfig(pow_bits=%d, log_blowup=%d, log_last=%d, n_queries=%d)" % self.key()


_Q = ctypes.c_uint32 * 4


class _FsOut(ctypes.Structure):
    _fields_ = [(n, _Q) for n in ("z", "alpha", "random_coeff", "oods_t", "oods_x", "oods_y", "after_coeff")] + [
        ("fri_alphas", _Q * 33), ("digest_after_nonce", ctypes.c_uint32 * 8), ("raw_queries", ctypes.c_uint32 * 128),
        ("n_transcript_perms", ctypes.c_uint32), ("pow_ok", ctypes.c_uint32)]


class VerifyDetail(ctypes.Structure):
    """stwo_b200_verify_detail"""
    _fields_ = [("fs", _FsOut), ("oods_computed", _Q), ("oods_expected", _Q), ("n_logs", ctypes.c_uint32),
                ("log_sizes", ctypes.c_uint32 * 3), ("fail_mask", ctypes.c_uint32), ("verdict", ctypes.c_uint32),
                ("stage", ctypes.c_uint32), ("n_perms_hints", ctypes.c_uint32), ("n_perms_paths", ctypes.c_uint32)]


VERIFY_FULL = 1
VERIFY_TIMED = 2
VERIFY_PATH_KERNELS = 8
STAGE_KERNELS = ("fiat_shamir", "single_tree", "group", "answer", "folds", "pair_tree", "single_path", "pair_path", "verdict")
FETCH = {"detail": 0, "domain_points": 1, "answers": 2, "circle_folds": 3, "line_folds": 4, "last_evals": 5, "path_roots": 6,
         "path_cols": 7, "path_siblings": 8, "pair_hints": 9, "perm_record": 10, "record_trees": 11, "perm_record_inputs": 12}
STAGES = {0: "ok", 1: "parse", 2: "pow", 3: "logup", 4: "oods", 5: "merkle", 6: "fri_first", 7: "fri_inner", 8: "fri_last",
          9: "unsupported"}

class CsWiring(ctypes.Structure):
    """stwo_b200_cs_wiring"""
    _fields_ = [(n, ctypes.c_uint32) for n in ("n_vars", "n_rows", "n_flow", "num_input")] + [
        (n, ctypes.c_void_p) for n in ("a_wire", "b_wire", "c_wire", "poseidon_wire", "enforce_c_m31", "op", "op_follows_c", "flow_wire",
                                       "flow_swap_addr")] + [("kind", ctypes.c_uint32)] + [(n, ctypes.c_void_p) for n in ("op2", "op3", "op4")] + [
        ("export_tiles", ctypes.c_void_p), ("export_cap", ctypes.c_uint32)]


class CsValues(ctypes.Structure):
    """stwo_b200_cs_values"""
    _fields_ = [("n_batch", ctypes.c_uint32), ("lanes", ctypes.c_uint32), ("variables", ctypes.c_void_p), ("flow_hash", ctypes.c_void_p),
                ("flow_swap", ctypes.c_void_p), ("perm_hints", ctypes.c_void_p), ("perm_hint_stride", ctypes.c_uint32),
                ("perm_hint_ready", ctypes.c_void_p), ("perm_hint_need", ctypes.c_uint32), ("perm_hint_inputs", ctypes.c_void_p)]


class CsTape(ctypes.Structure):
    """stwo_b200_cs_tape"""
    _fields_ = [(n, ctypes.c_uint32) for n in ("n_ins", "n_perms", "n_levels", "n_input_words")] + [
        (n, ctypes.c_void_p) for n in ("ins", "level_start", "perms")] + [("n_eperms", ctypes.c_uint32), ("eperms", ctypes.c_void_p),
                                                                           ("n_bundles", ctypes.c_uint32), ("bundle_start", ctypes.c_void_p), ("level_bundle", ctypes.c_void_p),
                                                                           ("recorded_order", ctypes.c_void_p)]


class CircuitInfo(ctypes.Structure):
    """stwo_b200_circuit_info"""
    _fields_ = [(n, ctypes.c_uint32) for n in ("n_rows", "n_rows_unpadded", "n_vars", "n_flow", "n_flow_padded", "n_input_words", "n_ins",
                                               "n_levels", "num_input", "words_per_instance", "kind", "n_preprocessed_columns")]


TRACE_CHECK_ARITHMETICS, TRACE_CHECK_POSEIDON, TRACE_TIMED, TRACE_NATIVE_HINTS, TRACE_RECHECK_POSEIDON = 1, 2, 4, 8, 16
TRACE_STAGES = ("gather", "eval", "check_arithmetics", "check_poseidon", "export")
COLUMNS = {"a_wire": 0, "b_wire": 1, "c_wire": 2, "poseidon_wire": 3, "enforce_c_m31": 4, "op": 5, "op_follows_c": 6, "flow_wire": 7,
           "flow_swap_addr": 8, "level_start": 10, "op2": 11, "op3": 12, "op4": 13}
CFETCH = {"variables": 0, "flow_hash": 1, "flow_swap": 2, "witness": 3}

_vp, _u32, _i32, _sz, _u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int32, ctypes.c_size_t, ctypes.c_uint64
_WIR_P, _VAL_P, _TAPE_P, _INFO_P = ctypes.POINTER(CsWiring), ctypes.POINTER(CsValues), ctypes.POINTER(CsTape), ctypes.POINTER(CircuitInfo)
_SHAPE_P = ctypes.POINTER(PathShape)
_PSHAPE_P = ctypes.POINTER(ProofShape)
_CFG_P = ctypes.POINTER(PcsConfig)

# name -> (restype, argtypes); the test-suite checks this list against include/stwo_b200.h
SIGNATURES = {
    "stwo_b200_init": (_i32, [_i32]),
    "stwo_b200_shutdown": (_i32, []),
    "stwo_b200_version": (_u32, []),
    "stwo_b200_launch_count": (_u64, []),
    "stwo_b200_poseidon2_permute": (_i32, [_vp, _sz]),
    "stwo_b200_poseidon2_permute_dev": (_i32, [_vp, _sz, _vp]),
    "stwo_b200_poseidon2_permute_dev_variant": (_i32, [_vp, _sz, _i32, _vp]),
    "stwo_b200_hash_node_batch_dev": (_i32, [_vp, _vp, _u32, _sz, _sz, _vp, _vp]),
    "stwo_b200_hash_node_batch": (_i32, [_vp, _vp, _u32, _sz, _sz, _vp]),
    "stwo_b200_merkle_commit_dev": (_i32, [_vp, _u32, _u32, _u32, _vp, _vp]),
    "stwo_b200_merkle_commit": (_i32, [_vp, _u32, _u32, _u32, _vp]),
    "stwo_b200_merkle_decommit_dev": (_i32, [_vp, _u32, _u32, _u32, _vp, _vp, _u32, _vp, _vp, _vp]),
    "stwo_b200_merkle_path_verify_dev": (_i32, [_SHAPE_P, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_merkle_path_verify": (_i32, [_SHAPE_P, _sz, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "stwo_b200_path_perms": (_u32, [_SHAPE_P]),
    "stwo_b200_proof_shape_of": (_i32, [_vp, _sz, _PSHAPE_P]),
    "stwo_b200_shape_from_config": (_i32, [_CFG_P, _u32, _u32, _PSHAPE_P]),
    "stwo_b200_verify_workspace_bytes": (_sz, [_PSHAPE_P, _u32]),
    "stwo_b200_proof_perms": (_u64, [_PSHAPE_P]),
    "stwo_b200_proof_record_slots": (_u32, [_PSHAPE_P]),
    "stwo_b200_verify_proofs_batch_dev": (_i32, [_vp, _vp, _u32, _PSHAPE_P, _vp, _vp, _u32, _u32, _vp, _sz, _vp, _vp, _vp]),
    "stwo_b200_verify_proofs_batch_pinned_dev": (_i32, [_vp, _vp, _vp, _vp, _u32, _PSHAPE_P, _vp, _vp, _u32, _u32, _vp, _sz, _vp, _vp, _vp]),
    "stwo_b200_verify_proofs_batch": (_i32, [_vp, _vp, _u32, _vp, _u32, _vp, _vp, _u32, _u32, _vp, _vp]),
    "stwo_b200_verify_stage_ms": (_i32, [_vp]),
    "stwo_b200_verify_fetch": (_i32, [_vp, _PSHAPE_P, _u32, _u32, _u32, _vp, _sz, _vp]),
    "stwo_b200_verify_fetch_batch": (_i32, [_vp, _PSHAPE_P, _u32, _u32, _vp, _sz, _vp]),
    "stwo_b200_channel_replay_batch": (_i32, [_vp, _vp, _u32, _CFG_P, _vp, _vp, _u32, _vp, _vp, _vp]),
    "stwo_b200_fri_answers_batch": (_i32, [_vp, _vp, _u32, _CFG_P, _vp, _vp, _u32, _vp, _vp, _vp, _vp]),
    "stwo_b200_fri_fold_batch": (_i32, [_vp, _vp, _u32, _CFG_P, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_hash_column_capacity_batch": (_i32, [_vp, _u32, _sz, _vp]),
    "stwo_b200_hash_column_capacity_batch_dev": (_i32, [_vp, _u32, _sz, _vp, _vp]),
    "stwo_b200_shard_range": (_i32, [_u64, _u32, _u32, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]),
    "stwo_b200_comm_unique_id": (_i32, [_vp]),
    "stwo_b200_comm_init": (_i32, [_vp, _u32, _u32, ctypes.POINTER(_vp)]),
    "stwo_b200_comm_destroy": (_i32, [_vp]),
    "stwo_b200_gather_verdicts_scratch_bytes": (_sz, [_u32, _u64]),
    "stwo_b200_gather_verdicts": (_i32, [_vp, _u32, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "stwo_b200_gather_trace_columns": (_i32, [_vp, _u32, _u32, _u32, _u64, _sz, _vp, _vp, _vp]),
    "stwo_b200_cs_flow_padded_len": (_u32, [_u32]),
    "stwo_b200_cs_export_flow_dev": (_i32, [_VAL_P, _u32, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_circuit_export_flow_dev": (_i32, [_vp, _u32, _vp, _sz, _vp, _vp, _vp, _vp]),
    "stwo_b200_synth_blob_words": (_u32, [_PSHAPE_P]),
    "stwo_b200_synth_scratch_bytes": (_sz, [_PSHAPE_P, _u32]),
    "stwo_b200_synth_generate_dev": (_i32, [_PSHAPE_P, _u32, _u64, _vp, _vp, _vp, _sz, _vp, _vp]),
    "stwo_b200_synth_verify_batch_dev": (_i32, [_vp, _vp, _u32, _PSHAPE_P, _u32, _vp, _sz, _vp, _vp, _vp]),
    "stwo_b200_cs_eval_tape_dev": (_i32, [_TAPE_P, _u32, _vp, _VAL_P, _vp]),
    "stwo_b200_cs_eval_level_clock": (_i32, [_vp]),
    "stwo_b200_cs_check_arithmetics_dev": (_i32, [_WIR_P, _VAL_P, _vp, _vp]),
    "stwo_b200_cs_populate_logup_dev": (_i32, [_WIR_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_cs_check_poseidon_dev": (_i32, [_WIR_P, _VAL_P, _vp, _vp, _vp, _vp]),
    "stwo_b200_cs_check_poseidon_recorded_dev": (_i32, [_WIR_P, _VAL_P, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_cs_export_trace_dev": (_i32, [_WIR_P, _VAL_P, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_cs_finalize": (_i32, [_WIR_P, _VAL_P, _vp, _vp, _vp]),
    "stwo_b200_cs_export_tiles_words": (_sz, [_u32]),
    "stwo_b200_cs_export_tiles_build": (_i32, [_WIR_P, _vp, _vp]),
    "stwo_b200_circuit_record_verifier": (_i32, [_PSHAPE_P, _vp, _vp, _u32, _u32, ctypes.POINTER(_vp)]),
    "stwo_b200_circuit_record_last_layer": (_i32, [_PSHAPE_P, ctypes.POINTER(_vp)]),
    "stwo_b200_circuit_record_folding": (_i32, [_PSHAPE_P, ctypes.POINTER(_vp)]),
    "stwo_b200_circuit_free": (None, [_vp]),
    "stwo_b200_circuit_get_info": (_i32, [_vp, _INFO_P]),
    "stwo_b200_circuit_get_column": (_i32, [_vp, _u32, _vp, _sz]),
    "stwo_b200_circuit_workspace_bytes": (_sz, [_vp, _u32]),
    "stwo_b200_circuit_trace_batch_dev": (_i32, [_vp, _vp, _vp, _u32, _vp, _vp, _sz, _u32, _vp, _vp, _vp, _vp, _vp]),
    "stwo_b200_circuit_stage_ms": (_i32, [_vp]),
    "stwo_b200_circuit_fetch": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _sz, _vp]),
}

_lib = None


def load():
    """Load the shared library (no device needed) and set the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libstwo_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C recursive-stwo_b200/csrc`; there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(fn, status):
    if status != OK:
        raise StwoB200Error(fn, status)


def call(name, *args):
    status = getattr(load(), name)(*args)
    check(name, status)
