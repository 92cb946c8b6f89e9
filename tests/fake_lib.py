"""TEST INFRASTRUCTURE: a stand-in for libinfimum_b200.so backed by the oracle.

It lets the CPU suite (`-m "not gpu"`) drive the Python host mirror
(infimum_b200/hasher.py, tree.py, poll.py, leaves.py, paths.py) — argument
marshalling, the reference's error ordering, the tree / poll state machines —
where there is no GPU.  The entry points below follow include/infimum_b200.h
(same argument order, same return codes) and compute with oracle/, which only
tests may do.  Nothing under infimum_b200/ knows about this module; the GPU
suite runs the same tests against the real library.
"""
import ctypes as C

import numpy as np

from oracle import c_oracle
from oracle import poseidon_ref as O

OK, FULL, MERGED, HASH_FAILED, MERGE_FAILED = 0, 1, 2, 3, 4
E_NINPUTS, E_EMPTY, E_LEN, E_WIDTH = 16, 17, 18, 19
E_NULL, E_ARITY, E_DEPTH = 24, 25, 26
LE = 1


def _rd(p, n):
    if n == 0:
        return b""
    if p is None:
        raise ValueError("null pointer")
    if isinstance(p, (bytes, bytearray)):
        return bytes(p[:n])
    if isinstance(p, int):
        return C.string_at(p, n)
    if isinstance(p, C.c_void_p):
        return C.string_at(p.value, n)
    return C.string_at(C.addressof(p), n)


def _wr(p, data):
    if not data:
        return
    if isinstance(p, int):
        C.memmove(p, data, len(data))
    elif isinstance(p, C.c_void_p):
        C.memmove(p.value, data, len(data))
    else:
        C.memmove(C.addressof(p), data, len(data))


def _set(ref, v):
    if ref is not None:
        getattr(ref, "_obj", ref).value = v


def _fr(b, le=False):
    return int.from_bytes(b, "little" if le else "big") % O.P


class FakeLib:
    def __init__(self):
        self._trees = {}
        self._next = 1

    # ---- hashing ---------------------------------------------------------------------
    def _batch(self, n_inputs, flags, tag, src, n, dst, params=None):
        if n_inputs < 1 or n_inputs + 1 > 13:
            return E_WIDTH
        le = bool(flags & LE)
        data = _rd(src, n * n_inputs * 32)
        tagv = _fr(_rd(tag, 32), le) if tag else 0
        if params is None and tagv == 0 and not le:
            out = c_oracle.hash_batch(n_inputs, np.frombuffer(data, dtype=np.uint8), threads=1).tobytes() if n else b""
        else:
            out = bytearray()
            for i in range(n):
                ins = [_fr(data[(i * n_inputs + k) * 32:(i * n_inputs + k + 1) * 32], le) for k in range(n_inputs)]
                h = (O.poseidon_hash_with_params(*params, ins, tagv) if params is not None
                     else O.poseidon_permute_hash(ins, tagv))
                out += h.to_bytes(32, "little" if le else "big")
        _wr(dst, bytes(out))
        return OK

    def inf_poseidon_hash_batch(self, ctx, n_inputs, flags, tag, src, n, dst):
        return self._batch(n_inputs, flags, tag, src, n, dst)

    inf_poseidon_hash_batch_dense = inf_poseidon_hash_batch

    def inf_poseidon_hash_batch_params(self, ctx, width, full_rounds, partial_rounds, alpha, ark, mds, flags, tag,
                                       src, n, dst):
        if width < 2 or width > 13:
            return E_WIDTH
        rounds = full_rounds + partial_rounds
        a = _rd(ark, rounds * width * 32)
        m = _rd(mds, width * width * 32)
        arkv = [_fr(a[32 * i:32 * i + 32]) for i in range(rounds * width)]
        mdsv = [[_fr(m[32 * (i * width + j):32 * (i * width + j) + 32]) for j in range(width)] for i in range(width)]
        return self._batch(width - 1, flags, tag, src, n, dst, (arkv, mdsv, full_rounds, partial_rounds, width, alpha))

    def inf_poseidon_hash_bytes(self, ctx, flags, tag, ptrs, lens, n_inputs, out):
        if n_inputs < 1 or n_inputs + 1 > 13:
            return E_WIDTH
        buf = b""
        addrs = C.cast(ptrs, C.POINTER(C.c_void_p))          # raw addresses: c_char_p values stop at a NUL byte
        for i in range(n_inputs):
            if lens[i] == 0:
                return E_EMPTY
            if lens[i] != 32:
                return E_LEN
            buf += C.string_at(addrs[i], 32)
        return self._batch(n_inputs, flags, tag, buf, 1, out)

    # ---- tables -----------------------------------------------------------------------
    def inf_merkle_zeroes(self, ctx, arity, out):
        _wr(out, b"".join(O.merkle_zeroes(2 if arity == 2 else 5)))
        return OK

    # ---- trees ------------------------------------------------------------------------
    def inf_tree_merge(self, ctx, arity, full_depth, blank, to_depth, leaves, n, root, idepth, rdepth, has):
        if arity not in (2, 5):
            return E_ARITY
        if full_depth > 32:
            return E_DEPTH
        _set(idepth, 0), _set(rdepth, 0), _set(has, 0)
        total = n + (1 if blank else 0)
        if total > arity ** full_depth:
            return FULL
        if total == 0:
            return OK
        lv = np.frombuffer(_rd(leaves, n * 32), dtype=np.uint8).reshape(n, 32)
        rc, r, depth, _ = c_oracle.tree_insert_merge(arity, full_depth, bool(blank), bool(to_depth), lv)
        d = full_depth
        if not to_depth and total != arity ** full_depth:
            d = 0
            while arity ** d < total:
                d += 1
        _wr(root, r)
        _set(idepth, depth), _set(rdepth, d), _set(has, 1)
        return rc

    def inf_tree_frontier(self, ctx, arity, full_depth, blank, leaves, n, out_levels, out_hashes, cap, n_entries,
                          idepth, has, root):
        _set(n_entries, 0), _set(idepth, 0), _set(has, 0)
        total = n + (1 if blank else 0)
        if total > arity ** full_depth:
            return FULL
        data = _rd(leaves, n * 32)
        t = O.PollStateTree.new(arity, full_depth, (0, O.merkle_zeroes(arity)[0]) if blank else None)
        for i in range(n):
            t.insert(data[32 * i:32 * i + 32])
        _set(idepth, t.depth)
        if t.root is not None:
            _set(has, 1)
            _wr(root, t.root)
            return OK
        if len(t.hashes) > cap:
            return 27
        _wr(out_levels, bytes(l for l, _ in t.hashes))
        _wr(out_hashes, b"".join(h for _, h in t.hashes))
        _set(n_entries, len(t.hashes))
        return OK

    def inf_tree_append(self, ctx, arity, full_depth, in_levels, in_hashes, n_in, depth_in, leaves, n, out_levels,
                        out_hashes, cap, n_entries, depth_out, has, root):
        _set(n_entries, 0), _set(depth_out, depth_in), _set(has, 0)
        lv, hs = _rd(in_levels, n_in), _rd(in_hashes, n_in * 32)
        t = O.PollStateTree(arity=arity, full_depth=full_depth, depth=depth_in,
                            hashes=[(lv[i], hs[32 * i:32 * i + 32]) for i in range(n_in)])
        if sum(arity ** l for l in lv) + n > arity ** full_depth:
            return FULL
        data = _rd(leaves, n * 32)
        for i in range(n):
            t.insert(data[32 * i:32 * i + 32])
        _set(depth_out, t.depth)
        if t.root is not None:
            _set(has, 1)
            _wr(root, t.root)
            return OK
        if len(t.hashes) > cap:
            return 27
        _wr(out_levels, bytes(l for l, _ in t.hashes))
        _wr(out_hashes, b"".join(h for _, h in t.hashes))
        _set(n_entries, len(t.hashes))
        return OK

    def inf_tree_merge_frontier(self, ctx, arity, full_depth, levels, hashes, n, to_depth, root, has, rdepth):
        _set(has, 0), _set(rdepth, 0)
        lv, hs = _rd(levels, n), _rd(hashes, n * 32)
        t = O.PollStateTree(arity=arity, full_depth=full_depth,
                            hashes=[(lv[i], hs[32 * i:32 * i + 32]) for i in range(n)])
        top = max(lv) if n else 0
        t.merge(bool(to_depth))
        if t.root is not None:
            _set(has, 1)
            _wr(root, t.root)
            d = full_depth
            if not to_depth:
                d = top
                while sum(arity ** l for l in lv) > arity ** d:
                    d += 1
            _set(rdepth, d)
        return OK

    def inf_replay_registrations(self, ctx, depth, pk, ts, n, root, commitment, idepth, leaves_out, retained):
        if n + 1 > 2 ** depth:
            return FULL
        lv = c_oracle.registration_leaves(np.frombuffer(_rd(pk, n * 64), dtype=np.uint8),
                                          np.frombuffer(_rd(ts, n * 8), dtype=np.uint64)) if n else np.empty((0, 32), np.uint8)
        if leaves_out is not None:
            _wr(leaves_out, lv.tobytes())
        rc, r, d, _ = c_oracle.tree_insert_merge(2, depth, True, False, lv)
        _wr(root, r), _set(idepth, d)
        _wr(commitment, c_oracle.hash_one([r, O.EMPTY_BALLOT_ROOTS[1].to_bytes(32, "big"), bytes(32)]))
        return rc

    def inf_replay_interactions(self, ctx, depth, pk, data, n, regs, psd, tsd, root, has, idepth, ep, et, leaves_out, retained):
        _set(has, 0)
        if n > 5 ** depth:
            return FULL
        lv = c_oracle.interaction_leaves(np.frombuffer(_rd(pk, n * 64), dtype=np.uint8),
                                         np.frombuffer(_rd(data, n * 320), dtype=np.uint8)) if n else np.empty((0, 32), np.uint8)
        if leaves_out is not None:
            _wr(leaves_out, lv.tobytes())
        rc, r, d, _ = c_oracle.tree_insert_merge(5, depth, False, True, lv)
        if r is not None:
            _wr(root, r), _set(has, 1)
        _set(idepth, d), _set(ep, n // 5 ** psd + (1 if n % 5 ** psd else 0)), _set(et, 1 + regs // 2 ** tsd)
        return rc

    # ---- leaves -------------------------------------------------------------------------
    def inf_registration_leaves(self, ctx, pk, ts, n, out):
        keys = _rd(pk, 64 * n)
        stamps = np.frombuffer(_rd(ts, 8 * n), dtype=np.uint64)
        _wr(out, b"".join(O.registration_leaf(keys[64 * i:64 * i + 32], keys[64 * i + 32:64 * i + 64], int(stamps[i]))
                          for i in range(n)))
        return OK

    def inf_interaction_leaves(self, ctx, pk, data, n, out):
        keys, d = _rd(pk, 64 * n), _rd(data, 320 * n)
        _wr(out, b"".join(O.interaction_leaf(keys[64 * i:64 * i + 32], keys[64 * i + 32:64 * i + 64],
                                             [d[320 * i + 32 * k:320 * i + 32 * k + 32] for k in range(10)])
                          for i in range(n)))
        return OK

    # ---- paths ---------------------------------------------------------------------------
    def inf_merkle_roots_from_paths(self, ctx, arity, depth, indices, leaves, paths, n, roots):
        if arity not in (2, 5):
            return E_ARITY
        idx = np.frombuffer(_rd(indices, 8 * n), dtype=np.uint64)
        lv = _rd(leaves, 32 * n)
        per = depth * (arity - 1) * 32
        pb = _rd(paths, per * n)
        out = b""
        for i in range(n):
            p = pb[per * i:per * (i + 1)]
            path = [[p[(l * (arity - 1) + k) * 32:(l * (arity - 1) + k + 1) * 32] for k in range(arity - 1)]
                    for l in range(depth)]
            out += O.compute_merkle_root_from_path(depth, int(idx[i]), lv[32 * i:32 * i + 32], path, arity)
        _wr(roots, out)
        return OK

    def inf_tree_build(self, ctx, arity, depth, blank, leaves, n, out):
        if arity not in (2, 5):
            return E_ARITY
        total = n + (1 if blank else 0)
        if total > arity ** depth:
            return FULL
        if total == 0:
            return MERGE_FAILED
        data = _rd(leaves, n * 32)
        lv = ([O.merkle_zeroes(arity)[0]] if blank else []) + [data[32 * i:32 * i + 32] for i in range(n)]
        h = self._next
        self._next += 1
        self._trees[h] = (arity, depth, O.dense_tree_levels(lv, arity, depth))
        _set(out, h)
        return OK

    def inf_tree_root(self, tree, root):
        arity, depth, levels = self._trees[getattr(tree, "value", tree)]
        _wr(root, levels[depth][0])
        return OK

    def inf_tree_paths(self, tree, indices, n, out):
        arity, depth, levels = self._trees[getattr(tree, "value", tree)]
        idx = np.frombuffer(_rd(indices, 8 * n), dtype=np.uint64)
        if any(int(i) >= arity ** depth for i in idx):
            return E_DEPTH
        buf = b""
        for i in idx:
            for lvl in O.merkle_path(levels, arity, int(i)):
                buf += b"".join(lvl)
        _wr(out, buf)
        return OK

    def inf_tree_destroy(self, tree):
        self._trees.pop(getattr(tree, "value", tree), None)


class FakeContext:
    """Duck-types infimum_b200.context.Context."""

    def __init__(self):
        self.lib = FakeLib()
        self.handle = C.c_void_p(1)
        self.device = 0

    def check(self, rc):
        from infimum_b200.errors import raise_for
        raise_for(rc, None)

    def close(self):
        pass
