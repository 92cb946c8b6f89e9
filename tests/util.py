"""Shared helpers for the test-suite."""
import os
import numpy as np

P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
SEED = 0x494E46494D554D  # "INFIMUM" 


def be(x: int) -> bytes:
    return int(x).to_bytes(32, "big")


def random_fr_bytes(n: int, seed: int = SEED, canonical: bool = True) -> np.ndarray:
    """(n, 32) uint8: uniform field elements as canonical 32-byte big-endian
    (rejection sampled below p), or arbitrary 256-bit values if not canonical."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    if canonical:
        out[:, 0] &= 0x3F                      # 254 bits
        pb = np.frombuffer(be(P), dtype=np.uint8)
        while True:
            # lexicographic compare with p
            diff = out.astype(np.int16) - pb.astype(np.int16)
            nz = diff != 0
            first = np.where(nz.any(axis=1), nz.argmax(axis=1), 31)
            bad = diff[np.arange(n), first] >= 0
            bad &= nz.any(axis=1) | True
            if not bad.any():
                break
            k = int(bad.sum())
            fresh = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
            fresh[:, 0] &= 0x3F
            out[bad] = fresh
    return out


EDGE_VALUES = [0, 1, 2, P - 1, P, P + 1, 2 * P - 1, 2 * P, 2 * P + 1, 3 * P, 4 * P, 5 * P, 2 ** 256 - 1,
               2 ** 255, 2 ** 254, 2 ** 224 - 1, 2 ** 224, (2 ** 256 - 1) ^ (2 ** 128 - 1), 0xFFFFFFFF,
               0xFFFFFFFF00000000, 2 ** 256 - 2 ** 32]
