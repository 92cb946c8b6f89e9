"""Error values of the hot path, named as the reference names them.

PoseidonError  — pallet/src/hash/poseidon.rs:13-31
MerkleTreeError — pallet/src/poll/state.rs:94-118 (with its `u8` codes)
"""
from __future__ import annotations

from . import _lib


class PoseidonError(Exception):
    """`kind` is the reference's enum variant name."""

    def __init__(self, kind: str, **info):
        super().__init__(kind, info)
        self.kind = kind
        self.info = info

    def __eq__(self, other):
        return isinstance(other, PoseidonError) and self.kind == other.kind

    def __hash__(self):
        return hash(("PoseidonError", self.kind))


class MerkleTreeError(Exception):
    CODES = {"TreeAlreadyFull": 1, "TreeAlreadyMerged": 2, "HashFailed": 3, "MergeFailed": 4}

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind
        self.code = self.CODES[kind]          # `impl From<MerkleTreeError> for u8`


class DeviceError(RuntimeError):
    """CUDA-side failure (codes >= 64).  A Rust host would surface it as
    MerkleTreeError::HashFailed / PoseidonError; here it keeps the CUDA text."""


_POSEIDON = {
    _lib.ERR_INVALID_NUMBER_OF_INPUTS: "InvalidNumberOfInputs",
    _lib.ERR_EMPTY_INPUT: "EmptyInput",
    _lib.ERR_INVALID_INPUT_LENGTH: "InvalidInputLength",
    _lib.ERR_INVALID_WIDTH_CIRCOM: "InvalidWidthCircom",
}
_TREE = {1: "TreeAlreadyFull", 2: "TreeAlreadyMerged", 3: "HashFailed", 4: "MergeFailed"}


def raise_for(code: int, ctx=None):
    if code == _lib.OK:
        return
    if code in _POSEIDON:
        raise PoseidonError(_POSEIDON[code])
    if code in _TREE:
        raise MerkleTreeError(_TREE[code])
    if code >= 64:
        detail = ""
        if ctx is not None:
            detail = _lib.load().inf_last_cuda_error(ctx).decode()
        raise DeviceError("%s %s" % (_lib.strerror(code), detail))
    raise ValueError(_lib.strerror(code))
