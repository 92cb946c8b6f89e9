"""ctypes binding of the C ABI declared in include/infimum_b200.h.

This module is plumbing: it loads libinfimum_b200.so (built in-tree by
infimum_b200.build) and declares argument types.  There is no fallback of any
kind: if the library is missing, or no CUDA device is usable, it raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# INFIMUM_B200_LIB: load another build of the same library (kernel-variant experiments)
LIB_PATH = os.environ.get("INFIMUM_B200_LIB") or os.path.join(HERE, "libinfimum_b200.so")

# return codes (include/infimum_b200.h)
OK = 0
ERR_TREE_ALREADY_FULL = 1
ERR_TREE_ALREADY_MERGED = 2
ERR_HASH_FAILED = 3
ERR_MERGE_FAILED = 4
ERR_INVALID_NUMBER_OF_INPUTS = 16
ERR_EMPTY_INPUT = 17
ERR_INVALID_INPUT_LENGTH = 18
ERR_INVALID_WIDTH_CIRCOM = 19
ERR_NULL_POINTER = 24
ERR_BAD_ARITY = 25
ERR_BAD_DEPTH = 26
ERR_BUFFER_TOO_SMALL = 27
ERR_BAD_FRONTIER = 28
ERR_BAD_ALIGNMENT = 29
ERR_NO_DEVICE = 64
ERR_CUDA = 65
ERR_OUT_OF_MEMORY = 66
ERR_NCCL = 67
MULTI_PEER_COPY = 1
FLAG_LITTLE_ENDIAN = 1

# every symbol include/infimum_b200.h declares
EXPORTS = [
    "inf_init", "inf_destroy", "inf_strerror", "inf_last_cuda_error", "inf_version",
    "inf_poseidon_hash_batch", "inf_poseidon_hash_batch_dev", "inf_poseidon_hash_bytes",
    "inf_poseidon_hash_batch_dense", "inf_registration_leaves", "inf_interaction_leaves",
    "inf_registration_leaves_dev", "inf_interaction_leaves_dev", "inf_merkle_zeroes", "inf_empty_ballot_roots",
    "inf_tree_merge", "inf_tree_merge_dev", "inf_tree_reduce_dev", "inf_tree_frontier", "inf_tree_build", "inf_tree_root", "inf_tree_paths",
    "inf_tree_destroy", "inf_merkle_roots_from_paths", "inf_multi_init", "inf_multi_destroy",
    "inf_multi_device_count", "inf_multi_tree_merge", "inf_multi_poseidon_hash_batch", "inf_merge_registrations",
    "inf_merge_interactions", "inf_debug_dense_params", "inf_debug_opt_table",
    "inf_measure_imad_peak", "inf_poseidon_hash_batch_params",
    "inf_tree_append", "inf_tree_append_dev", "inf_tree_merge_frontier",
    "inf_tree_node_paths", "inf_tree_level_nodes", "inf_replay_registrations", "inf_replay_interactions",
    "inf_host_alloc", "inf_host_free", "inf_host_register", "inf_host_unregister",
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library (no GPU needed just to load and inspect symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            "libinfimum_b200.so is not built (run `python -m infimum_b200.build`); "
            "infimum_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    u8p, u32p, u64p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_void_p
    ip = C.POINTER(C.c_int)
    lib.inf_init.argtypes = [C.c_int, C.POINTER(vp)]
    lib.inf_init.restype = C.c_int
    lib.inf_destroy.argtypes = [vp]
    lib.inf_destroy.restype = None
    lib.inf_strerror.argtypes = [C.c_int]
    lib.inf_strerror.restype = C.c_char_p
    lib.inf_last_cuda_error.argtypes = [vp]
    lib.inf_last_cuda_error.restype = C.c_char_p
    lib.inf_version.argtypes = []
    lib.inf_version.restype = C.c_char_p
    lib.inf_poseidon_hash_batch.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, vp]
    lib.inf_poseidon_hash_batch.restype = C.c_int
    lib.inf_poseidon_hash_batch_dense.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, vp]
    lib.inf_poseidon_hash_batch_dense.restype = C.c_int
    lib.inf_poseidon_hash_batch_params.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, C.c_uint32,
                                                   vp, vp, C.c_uint64, vp]
    lib.inf_poseidon_hash_batch_params.restype = C.c_int
    lib.inf_poseidon_hash_batch_dev.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, vp, vp]
    lib.inf_poseidon_hash_batch_dev.restype = C.c_int
    lib.inf_poseidon_hash_bytes.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t),
                                            C.c_uint32, vp]
    lib.inf_poseidon_hash_bytes.restype = C.c_int
    lib.inf_registration_leaves.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.inf_registration_leaves.restype = C.c_int
    lib.inf_interaction_leaves.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.inf_interaction_leaves.restype = C.c_int
    lib.inf_registration_leaves_dev.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    lib.inf_registration_leaves_dev.restype = C.c_int
    lib.inf_interaction_leaves_dev.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    lib.inf_interaction_leaves_dev.restype = C.c_int
    lib.inf_merkle_zeroes.argtypes = [vp, C.c_uint32, vp]
    lib.inf_merkle_zeroes.restype = C.c_int
    lib.inf_empty_ballot_roots.argtypes = [vp]
    lib.inf_empty_ballot_roots.restype = C.c_int
    lib.inf_tree_merge.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint64, vp, u32p, u32p, ip]
    lib.inf_tree_merge.restype = C.c_int
    lib.inf_tree_merge_dev.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint64, vp, u32p,
                                       u32p, ip, vp]
    lib.inf_tree_merge_dev.restype = C.c_int
    lib.inf_tree_reduce_dev.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp, C.c_uint64, vp,
                                        u64p, vp]
    lib.inf_tree_reduce_dev.restype = C.c_int
    lib.inf_tree_frontier.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_uint64, vp, vp, C.c_uint32, u32p,
                                      u32p, ip, vp]
    lib.inf_tree_frontier.restype = C.c_int
    lib.inf_tree_append.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp, vp,
                                    C.c_uint32, u32p, u32p, ip, vp]
    lib.inf_tree_append.restype = C.c_int
    lib.inf_tree_append_dev.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp,
                                        vp, C.c_uint32, u32p, u32p, ip, vp, vp]
    lib.inf_tree_append_dev.restype = C.c_int
    lib.inf_tree_merge_frontier.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_int, vp, ip, u32p]
    lib.inf_tree_merge_frontier.restype = C.c_int
    lib.inf_tree_build.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, vp, C.c_uint64, C.POINTER(vp)]
    lib.inf_tree_build.restype = C.c_int
    lib.inf_tree_root.argtypes = [vp, vp]
    lib.inf_tree_root.restype = C.c_int
    lib.inf_tree_paths.argtypes = [vp, vp, C.c_uint64, vp]
    lib.inf_tree_paths.restype = C.c_int
    lib.inf_tree_node_paths.argtypes = [vp, C.c_uint32, vp, C.c_uint64, vp]
    lib.inf_tree_node_paths.restype = C.c_int
    lib.inf_tree_level_nodes.argtypes = [vp, C.c_uint32, C.c_uint64, C.c_uint64, vp]
    lib.inf_tree_level_nodes.restype = C.c_int
    lib.inf_replay_registrations.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint64, vp, vp, u32p, vp, C.POINTER(vp)]
    lib.inf_replay_registrations.restype = C.c_int
    lib.inf_replay_interactions.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, ip,
                                            u32p, u32p, u32p, vp, C.POINTER(vp)]
    lib.inf_replay_interactions.restype = C.c_int
    lib.inf_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.inf_host_alloc.restype = C.c_int
    lib.inf_host_free.argtypes = [vp, vp]
    lib.inf_host_free.restype = C.c_int
    lib.inf_host_register.argtypes = [vp, vp, C.c_size_t]
    lib.inf_host_register.restype = C.c_int
    lib.inf_host_unregister.argtypes = [vp, vp]
    lib.inf_host_unregister.restype = C.c_int
    lib.inf_tree_destroy.argtypes = [vp]
    lib.inf_tree_destroy.restype = None
    lib.inf_merkle_roots_from_paths.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint64, vp]
    lib.inf_merkle_roots_from_paths.restype = C.c_int
    lib.inf_multi_init.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_uint32, C.POINTER(vp)]
    lib.inf_multi_init.restype = C.c_int
    lib.inf_multi_destroy.argtypes = [vp]
    lib.inf_multi_destroy.restype = None
    lib.inf_multi_device_count.argtypes = [vp]
    lib.inf_multi_device_count.restype = C.c_int
    lib.inf_multi_tree_merge.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, C.c_uint64, vp, u32p, u32p, ip]
    lib.inf_multi_tree_merge.restype = C.c_int
    lib.inf_multi_poseidon_hash_batch.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64, vp]
    lib.inf_multi_poseidon_hash_batch.restype = C.c_int
    lib.inf_merge_registrations.argtypes = [vp, C.c_uint32, vp, C.c_uint64, vp, vp, u32p]
    lib.inf_merge_registrations.restype = C.c_int
    lib.inf_merge_interactions.argtypes = [vp, C.c_uint32, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                           vp, ip, u32p, u32p]
    lib.inf_merge_interactions.restype = C.c_int
    lib.inf_debug_dense_params.argtypes = [C.c_uint32, u32p, C.c_size_t]
    lib.inf_debug_dense_params.restype = C.c_int
    lib.inf_debug_opt_table.argtypes = [C.c_uint32, u32p, C.c_size_t]
    lib.inf_debug_opt_table.restype = C.c_int
    lib.inf_measure_imad_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.inf_measure_imad_peak.restype = C.c_int
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().inf_strerror(code).decode()
