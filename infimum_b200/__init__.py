"""infimum_b200 — B200-native Poseidon-BN254 hasher and poll-tree merge.

A drop-in for the one data-parallel hot path of rhysbalevicius/infimum
(pallet/src/hash/poseidon.rs + pallet/src/poll/state.rs as driven by
`merge_poll_state`), implemented as hand-written CUDA for sm_100a behind a
C ABI (include/infimum_b200.h).  This package is the host-side mirror of the
reference interface over that ABI.  CUDA only: there is no CPU fallback.
"""
from .errors import DeviceError, MerkleTreeError, PoseidonError
from .context import Context, get_context
from .hasher import HASH_LEN, MAX_X5_LEN, MODULUS, Poseidon, PoseidonParameters
from .leaves import interaction_leaves, registration_leaves, replay_interactions, replay_registrations
from .paths import RetainedTree, compute_merkle_root_from_path, merkle_roots_from_paths, verify_outcome
from .poll import Commitment, Poll, PollConfig
from .tree import (PollStateTree, empty_ballot_roots, get_merkle_zeroes, merge_interactions,
                   merge_registrations, new_interaction_tree, new_registration_tree)

__all__ = [
    "Context", "get_context", "Poseidon", "PoseidonParameters", "PollStateTree", "PoseidonError", "MerkleTreeError",
    "DeviceError", "MODULUS", "HASH_LEN", "MAX_X5_LEN", "get_merkle_zeroes", "empty_ballot_roots",
    "merge_registrations", "merge_interactions", "new_registration_tree", "new_interaction_tree",
    "registration_leaves", "interaction_leaves", "replay_registrations", "replay_interactions", "Poll", "PollConfig", "Commitment",
    "RetainedTree", "compute_merkle_root_from_path", "merkle_roots_from_paths", "verify_outcome",
]
