"""stwo_b200 — B200-native (sm_100a) verifier hot path of recursive-stwo.

Host-side mirror of the reference's primitives for the accelerated path.  Everything here calls
the C ABI in libstwo_b200.so (include/stwo_b200.h); torch is used only to own device memory and
streams.  The directory name has a hyphen: import it with
`importlib.import_module("recursive-stwo_b200")` (see __graft_entry__.py).
"""
import os as _os

# Hardware work queues: the batch drivers use a dozen streams; with the driver's default of 8 queues independent chains alias onto one
# queue and wait for each other.  Read at context creation, so it is set on import, before the first CUDA call (runtime.cu does the same
# in stwo_b200_init for hosts without this package).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: E402
from ._lib import PathShape, StwoB200Error  # noqa: E402
from .hashing import (init, poseidon2_permute, poseidon2_permute_host, hash_node_batch, merkle_commit,
                      merkle_commit_host, merkle_decommit, merkle_path_verify, merkle_path_verify_host,
                      launch_count, path_perms)

from .verifier import (verify_proofs, VerifyBatch, VerifyStream, proof_shape, proof_perms, shape_for, shape_from_config,
                       INPUTS_SINGLE, INPUTS_RECURSIVE, REFERENCE_CONFIGS, CONFIG_SINGLE, CONFIG_STANDARD, CONFIG_FAST_PROVER,
                       CONFIG_FAST_PROVER2, CONFIG_FAST_VERIFIER, CONFIG_FAST_VERIFIER2, CONFIG_FAST_VERIFIER3)
from ._lib import ProofShape, PcsConfig, VerifyDetail, STAGES
from .circuit import VerifierCircuit, MixedBatch, VerifyTracePipeline, cached_circuit
from .synthetic import SynthBatch
from .stages import channel_replay, fri_answers, fri_folds, hash_column_capacity

__all__ = ["SynthBatch", "channel_replay", "fri_answers", "fri_folds", "hash_column_capacity", "VerifierCircuit", "MixedBatch", "VerifyTracePipeline", "cached_circuit", "verify_proofs", "VerifyBatch", "VerifyStream", "proof_shape", "proof_perms", "shape_for", "shape_from_config", "PcsConfig", "REFERENCE_CONFIGS", "INPUTS_SINGLE", "INPUTS_RECURSIVE", "ProofShape",
           "VerifyDetail", "STAGES","init", "poseidon2_permute", "poseidon2_permute_host", "hash_node_batch", "merkle_commit",
           "merkle_commit_host", "merkle_decommit", "merkle_path_verify", "merkle_path_verify_host", "launch_count",
           "path_perms", "PathShape", "StwoB200Error"]
