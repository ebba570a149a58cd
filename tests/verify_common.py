"""Shared helpers for the verifier parity tests: ctypes mirrors of the product structs (Detail, Workspace) and the
field-by-field comparison of a product run (host-sim or GPU) with the oracle."""
import ctypes

import numpy as np

import oracle_py as O

MAX_INNER, MAX_Q, MAX_LOGS = 32, 128, 3
_Q = ctypes.c_uint32 * 4


class FsOut(ctypes.Structure):
    _fields_ = [(n, _Q) for n in ("z", "alpha", "random_coeff", "oods_t", "oods_x", "oods_y", "after_coeff")] + [
        ("fri_alphas", _Q * (MAX_INNER + 1)), ("digest_after_nonce", ctypes.c_uint32 * 8), ("raw_queries", ctypes.c_uint32 * MAX_Q),
        ("n_transcript_perms", ctypes.c_uint32), ("pow_ok", ctypes.c_uint32)]


class Detail(ctypes.Structure):
    """verify::Detail (recursive-stwo_b200/csrc/verify.cuh) == stwo_b200_verify_detail (include/stwo_b200.h)"""
    _fields_ = [("fs", FsOut), ("oods_computed", _Q), ("oods_expected", _Q), ("n_logs", ctypes.c_uint32),
                ("log_sizes", ctypes.c_uint32 * MAX_LOGS), ("fail_mask", ctypes.c_uint32), ("verdict", ctypes.c_uint32),
                ("stage", ctypes.c_uint32), ("n_perms_hints", ctypes.c_uint32), ("n_perms_paths", ctypes.c_uint32)]


def shape_of(buf):
    """(log_size_plonk, log_size_poseidon, pow_bits, log_blowup, log_last, n_queries, n_inner) read from a blob"""
    w = np.frombuffer(buf.tobytes(), dtype=np.uint32)
    out = O.verify_proof(buf, buf.size, O.INPUTS_SMALL)      # only to learn n_inner (parse succeeds before anything can fail)
    return np.array([w[0], w[1], w[10], w[11], w[12], w[13], out.n_inner], dtype=np.uint32)


def pack(blobs):
    """list of (buf, length) -> (words array, offsets in words)"""
    words, off = [], [0]
    for buf, n in blobs:
        assert n % 4 == 0
        w = np.frombuffer(buf[:n].tobytes(), dtype=np.uint32)
        words.append(w)
        off.append(off[-1] + w.size)
    return np.concatenate(words), np.array(off, dtype=np.uint64)


def compare_detail(dt, o, full=True):
    """Detail (product) vs VerifyOut (oracle): every draw, digest, OODS value and counter"""
    assert (dt.verdict, dt.stage) == (o.verdict, o.stage), (dt.verdict, dt.stage, o.verdict, o.stage, bin(dt.fail_mask))
    if o.stage == 1:
        return
    for k in ("z", "alpha", "random_coeff", "oods_t", "oods_x", "oods_y", "after_coeff"):
        assert list(getattr(dt.fs, k)) == list(getattr(o, k)), k
    for i in range(o.n_inner + 1):
        assert list(dt.fs.fri_alphas[i]) == list(o.fri_alphas[i])
    assert list(dt.fs.digest_after_nonce) == list(o.digest_after_nonce)
    assert list(dt.fs.raw_queries)[: o.n_queries] == list(o.raw_queries)[: o.n_queries]
    assert dt.fs.n_transcript_perms == o.n_transcript_perms
    if o.stage in (2, 3):
        return
    assert list(dt.oods_computed) == list(o.oods_computed) and list(dt.oods_expected) == list(o.oods_expected)
    if o.verdict == 0:
        assert dt.n_logs == o.n_logs and list(dt.log_sizes)[: dt.n_logs] == list(o.log_sizes)[: o.n_logs]
        if full:
            assert dt.n_perms_paths == o.n_perms_paths


def compare_arrays(arrs, o, nq):
    """answers / folds / roots of an accepted proof: arrs is a dict of numpy arrays for ONE proof"""
    for g in range(o.n_logs):
        assert np.array_equal(arrs["domain_points"][g, :nq], O.np.ctypeslib.as_array(o.domain_points)[g, :nq])
        assert np.array_equal(arrs["answers"][g, :nq], np.ctypeslib.as_array(o.fri_answers)[g, :nq])
        assert np.array_equal(arrs["circle_folds"][g, :nq], np.ctypeslib.as_array(o.circle_folds)[g, :nq])
    for li in range(o.n_inner):
        assert np.array_equal(arrs["line_folds"][li, :nq], np.ctypeslib.as_array(o.line_folds)[li, :nq])
    assert np.array_equal(arrs["last_evals"][:nq], np.ctypeslib.as_array(o.last_layer_evals)[:nq])
    if "path_roots" in arrs:
        want = np.ctypeslib.as_array(o.path_roots)
        for t in range(5 + o.n_inner):
            assert np.array_equal(arrs["path_roots"][t, :nq], want[t, :nq]), "tree %d" % t
