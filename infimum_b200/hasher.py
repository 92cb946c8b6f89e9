"""Host-side mirror of the reference's Poseidon hasher interface, backed by the
CUDA kernels through the C ABI.

Same names, argument meaning and error behaviour as
pallet/src/hash/poseidon.rs:
    Poseidon::<Fr>::new_circom(nr_inputs)              :304-307
    Poseidon::<Fr>::with_domain_tag_circom(n, tag)     :309-326
    PoseidonHasher::hash(&[Fr])                        :162-208
    PoseidonBytesHasher::hash_bytes_be / hash_bytes_le :213-250
    PoseidonParameters::new, Poseidon::new(params)     :47-71, 105-108
plus the batch forms the GPU exists for (`hash_batch*`).  Field elements are
Python ints (canonical) at the `hash` level and 32-byte strings at the byte
level.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .context import Context, get_context
from .errors import PoseidonError

# p, quoted at pallet/src/hash/parameters.rs:14
MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
HASH_LEN = 32            # poseidon.rs:9
MAX_X5_LEN = 13          # poseidon.rs:10


def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if a.dtype != np.uint8:
        raise TypeError("expected bytes or a uint8 array")
    return np.ascontiguousarray(a).reshape(-1)


class PoseidonParameters:
    """poseidon.rs:33-71: ark indexed round*width + i, mds[i][j] row i column j,
    field elements as Python ints."""

    def __init__(self, ark: Sequence[int], mds: Sequence[Sequence[int]], full_rounds: int, partial_rounds: int,
                 width: int, alpha: int):
        self.ark = [int(x) % MODULUS for x in ark]
        self.mds = [[int(x) % MODULUS for x in row] for row in mds]
        self.full_rounds, self.partial_rounds, self.width, self.alpha = full_rounds, partial_rounds, width, alpha
        if len(self.ark) != (full_rounds + partial_rounds) * width:
            raise ValueError("ark must hold (full_rounds + partial_rounds) * width elements")
        if len(self.mds) != width or any(len(r) != width for r in self.mds):
            raise ValueError("mds must be width x width")

    def _packed(self):
        ark = b"".join(x.to_bytes(32, "big") for x in self.ark)
        mds = b"".join(x.to_bytes(32, "big") for row in self.mds for x in row)
        return ark, mds


class Poseidon:
    def __init__(self, nr_inputs: int, domain_tag: int = 0, ctx: Optional[Context] = None,
                 params: Optional[PoseidonParameters] = None):
        width = nr_inputs + 1
        if width > MAX_X5_LEN or width < 2:                       # poseidon.rs:315-320, parameters.rs:38-42
            raise PoseidonError("InvalidWidthCircom", width=width, max_limit=MAX_X5_LEN)
        self.width = width
        self.domain_tag = int(domain_tag) % MODULUS
        self.ctx = ctx or get_context()
        self.params = params
        self._packed = params._packed() if params is not None else None

    @classmethod
    def new(cls, params: PoseidonParameters, ctx: Optional[Context] = None) -> "Poseidon":
        """Poseidon::new(params) (poseidon.rs:105-108): caller-supplied parameters, domain tag zero."""
        return cls(params.width - 1, 0, ctx, params)

    @classmethod
    def with_domain_tag(cls, params: PoseidonParameters, domain_tag: int, ctx: Optional[Context] = None) -> "Poseidon":
        """poseidon.rs:110-118"""
        return cls(params.width - 1, domain_tag, ctx, params)

    @classmethod
    def new_circom(cls, nr_inputs: int, ctx: Optional[Context] = None) -> "Poseidon":
        return cls(nr_inputs, 0, ctx)

    @classmethod
    def with_domain_tag_circom(cls, nr_inputs: int, domain_tag: int, ctx: Optional[Context] = None) -> "Poseidon":
        return cls(nr_inputs, domain_tag, ctx)

    # ---- single hash, reference signatures -----------------------------------
    def hash(self, inputs: Sequence[int]) -> int:
        """PoseidonHasher::hash: field elements in, field element out."""
        if len(inputs) != self.width - 1:                         # poseidon.rs:164-171
            raise PoseidonError("InvalidNumberOfInputs", inputs=len(inputs),
                                max_limit=self.width - 1, width=self.width)
        buf = b"".join((int(x) % MODULUS).to_bytes(32, "big") for x in inputs)
        out = self.hash_batch(buf, 1)
        return int.from_bytes(out.tobytes(), "big")

    def _hash_bytes(self, inputs: Sequence[bytes], flags: int) -> bytes:
        n = len(inputs)
        if n != self.width - 1:
            # the byte front ends convert every input first (so length errors
            # win), then call hash(), which checks the count (poseidon.rs:215-226)
            for b in inputs:
                if len(b) == 0:
                    raise PoseidonError("EmptyInput")
                if len(b) != HASH_LEN:
                    raise PoseidonError("InvalidInputLength", len=len(b), modulus_bytes_len=HASH_LEN)
            raise PoseidonError("InvalidNumberOfInputs", inputs=n, max_limit=self.width - 1, width=self.width)
        if self.params is not None:
            for b in inputs:                                      # validate_bytes_length, bytes_to_prime_field_element
                if len(b) == 0:
                    raise PoseidonError("EmptyInput")
                if len(b) != HASH_LEN:
                    raise PoseidonError("InvalidInputLength", len=len(b), modulus_bytes_len=HASH_LEN)
            return self.hash_batch(b"".join(bytes(b) for b in inputs), 1,
                                   little_endian=bool(flags & _lib.FLAG_LITTLE_ENDIAN)).tobytes()
        ptrs = (C.c_char_p * n)(*[bytes(b) if len(b) else None for b in inputs])
        lens = (C.c_size_t * n)(*[len(b) for b in inputs])
        out = C.create_string_buffer(32)
        tag = self._tag_bytes(flags)
        rc = self.ctx.lib.inf_poseidon_hash_bytes(self.ctx.handle, flags, tag, ptrs, lens, n, out)
        self.ctx.check(rc)
        return out.raw

    def hash_bytes_be(self, inputs: Sequence[bytes]) -> bytes:
        return self._hash_bytes(inputs, 0)

    def hash_bytes_le(self, inputs: Sequence[bytes]) -> bytes:
        return self._hash_bytes(inputs, _lib.FLAG_LITTLE_ENDIAN)

    # ---- batch forms -----------------------------------------------------------
    def _tag_bytes(self, flags: int):
        if self.domain_tag == 0:
            return None
        return self.domain_tag.to_bytes(32, "little" if flags & _lib.FLAG_LITTLE_ENDIAN else "big")

    def hash_batch(self, inputs, n: Optional[int] = None, little_endian: bool = False,
                   dense: bool = False, out: Optional[np.ndarray] = None) -> np.ndarray:
        """n independent hashes.  `inputs`: bytes / uint8 array of n*(width-1)*32
        bytes (host memory; pinned memory makes the copies faster).  Returns an
        (n, 32) uint8 array (`out` if given, e.g. a pinned buffer)."""
        a = _as_u8(inputs)
        k = self.width - 1
        if n is None:
            if a.size % (k * 32):
                raise PoseidonError("InvalidInputLength", len=a.size, modulus_bytes_len=HASH_LEN)
            n = a.size // (k * 32)
        if a.size != n * k * 32:
            raise PoseidonError("InvalidInputLength", len=a.size, modulus_bytes_len=HASH_LEN)
        if out is None:
            out = np.empty((n, 32), dtype=np.uint8)
        elif out.dtype != np.uint8 or out.size != n * 32 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous uint8 array of n*32 bytes")
        flags = _lib.FLAG_LITTLE_ENDIAN if little_endian else 0
        if self.params is not None:
            pr = self.params
            rc = self.ctx.lib.inf_poseidon_hash_batch_params(self.ctx.handle, pr.width, pr.full_rounds, pr.partial_rounds,
                                                             pr.alpha, self._packed[0], self._packed[1], flags,
                                                             self._tag_bytes(flags), a.ctypes.data, n, out.ctypes.data)
            self.ctx.check(rc)
            return out.reshape(n, 32)
        fn = self.ctx.lib.inf_poseidon_hash_batch_dense if dense else self.ctx.lib.inf_poseidon_hash_batch
        rc = fn(self.ctx.handle, k, flags, self._tag_bytes(flags), a.ctypes.data, n, out.ctypes.data)
        self.ctx.check(rc)
        return out.reshape(n, 32)

    def hash_batch_device(self, d_in: int, n: int, d_out: int, stream: int = 0, little_endian: bool = False):
        """Same on device pointers (ints); enqueued on `stream` (a cudaStream_t
        value; 0 = the context's own stream, 1 = CUDA's legacy default stream);
        does not synchronise."""
        if self.params is not None:
            raise NotImplementedError("custom parameters run through the host-buffer entry point")
        flags = _lib.FLAG_LITTLE_ENDIAN if little_endian else 0
        rc = self.ctx.lib.inf_poseidon_hash_batch_dev(self.ctx.handle, self.width - 1, flags,
                                                      self._tag_bytes(flags), d_in, n, d_out, stream or None)
        self.ctx.check(rc)


def validate_bytes_length(b: bytes) -> None:
    """poseidon.rs:255-273"""
    if len(b) == 0:
        raise PoseidonError("EmptyInput")
    if len(b) > HASH_LEN:
        raise PoseidonError("InvalidInputLength", len=len(b), modulus_bytes_len=HASH_LEN)
