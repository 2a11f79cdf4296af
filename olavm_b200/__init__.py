"""olavm_b200 -- B200 (sm_100a) STARK proving backend for OlaVM's `circuits` crate.

Host-side mirror (Python over the C ABI in include/ola_gpu.h) of the reference interfaces on the proving
hot path: plonky2 `cfft` transforms, Poseidon hashing / MerkleTree, `PolynomialBatch`.
The arithmetic lives in olavm_b200/csrc (hand-written CUDA for sm_100a); there is no CPU fallback.
"""
from ._lib import OlaError, load  # noqa: F401
from .context import Context  # noqa: F401
from .pcs import MerkleCap, PolynomialBatch  # noqa: F401
from . import cfft, generation, hashing, prover  # noqa: F401
from .prover import BLAKE3, POSEIDON, prove_with_device_traces, prove_with_traces, verify_proof, verify_subsystem_proof  # noqa: F401

GOLDILOCKS_P = 0xFFFFFFFF00000001
COSET_SHIFT = 7
