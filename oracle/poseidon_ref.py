"""Oracle A: pure-Python big-integer restatement of the reference hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Slow on purpose: every step is
the textbook operation on Python ints so that it is obviously right, and it is
pinned against the reference's own vectors by tests/test_oracle_golden.py.

What it restates (paths relative to the reference checkout):

* constants         pallet/src/hash/parameters.rs:16-19,35-43081 — NOT parsed or
                    copied: regenerated with the published Grain-LFSR procedure
                    the file's header names (hadeshash
                    generate_parameters_grain.sage 1 0 254 t 8 RP p);
                    tests/test_constants_vs_reference.py proves the two agree
                    element for element whenever /root/reference is present.
* hash              pallet/src/hash/poseidon.rs:123-208
* byte front ends   pallet/src/hash/poseidon.rs:213-300
* new_circom        pallet/src/hash/poseidon.rs:302-327
* tree              pallet/src/poll/state.rs:138-302
* zero tables       pallet/src/poll/zeroes.rs:1-85 (seeds only; chains recomputed)
* merge wrappers    pallet/src/poll/provider.rs:289-327
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

# BN254 scalar field modulus (quoted at parameters.rs:14).
P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
HASH_LEN = 32            # poseidon.rs:9
MAX_X5_LEN = 13          # poseidon.rs:10
FULL_ROUNDS = 8          # parameters.rs:16
PARTIAL_ROUNDS = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64]  # parameters.rs:17-18
ALPHA = 5                # parameters.rs:19


# ----------------------------------------------------------------------------
# Error values (poseidon.rs:13-31, state.rs:94-118)
# ----------------------------------------------------------------------------
class PoseidonError(Exception):
    def __init__(self, kind: str, **info):
        super().__init__(kind, info)
        self.kind = kind
        self.info = info

    def __eq__(self, other):
        return isinstance(other, PoseidonError) and self.kind == other.kind

    def __hash__(self):
        return hash(self.kind)


class MerkleTreeError(Exception):
    CODES = {"TreeAlreadyFull": 1, "TreeAlreadyMerged": 2, "HashFailed": 3, "MergeFailed": 4}

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind
        self.code = self.CODES[kind]


# ----------------------------------------------------------------------------
# Grain LFSR parameter generation (published algorithm; see module docstring)
# ----------------------------------------------------------------------------
class _Grain:
    """80-bit Grain LFSR in self-shrinking mode, as specified in the Poseidon
    paper (section "Concrete instantiations", and the hadeshash script)."""

    def __init__(self, field_bits: int, t: int, r_f: int, r_p: int):
        bits: List[int] = []

        def put(value: int, width: int):
            bits.extend((value >> (width - 1 - k)) & 1 for k in range(width))

        put(1, 2)            # prime field
        put(0, 4)            # x^alpha S-box
        put(field_bits, 12)
        put(t, 12)
        put(r_f, 10)
        put(r_p, 10)
        bits.extend([1] * 30)
        self.s = bits
        for _ in range(160):
            self._clock()

    def _clock(self) -> int:
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def next_bit(self) -> int:
        while True:
            keep = self._clock()
            out = self._clock()
            if keep:
                return out

    def next_int(self, nbits: int) -> int:
        v = 0
        for _ in range(nbits):
            v = (v << 1) | self.next_bit()
        return v


@lru_cache(maxsize=None)
def poseidon_parameters(t: int) -> Tuple[Tuple[int, ...], Tuple[Tuple[int, ...], ...], int, int]:
    """(ark, mds, full_rounds, partial_rounds) for state width t, 2 <= t <= 13.

    ark is indexed round*t+i (poseidon.rs:126), mds[i][j] row i col j
    (poseidon.rs:153).  Values are canonical integers in [0, P).
    """
    if not 2 <= t <= MAX_X5_LEN:
        raise PoseidonError("InvalidWidthCircom", width=t, max_limit=MAX_X5_LEN)
    r_p = PARTIAL_ROUNDS[t - 2]
    g = _Grain(254, t, FULL_ROUNDS, r_p)
    ark = []
    for _ in range((FULL_ROUNDS + r_p) * t):
        v = g.next_int(254)
        while v >= P:
            v = g.next_int(254)
        ark.append(v)
    while True:
        xy = [g.next_int(254) % P for _ in range(2 * t)]
        if len(set(xy)) == 2 * t:
            break
    xs, ys = xy[:t], xy[t:]
    mds = tuple(tuple(pow((xs[i] + ys[j]) % P, P - 2, P) for j in range(t)) for i in range(t))
    return tuple(ark), mds, FULL_ROUNDS, r_p


# ----------------------------------------------------------------------------
# The hash (poseidon.rs:94-208, 302-327)
# ----------------------------------------------------------------------------
def poseidon_permute_hash(inputs: Sequence[int], domain_tag: int = 0) -> int:
    """Poseidon hash of len(inputs) field elements with width len(inputs)+1.

    Follows PoseidonHasher::hash (poseidon.rs:162-208) step by step; inputs
    are already field elements (reduced) here.
    """
    t = len(inputs) + 1
    ark, mds, r_f, r_p = poseidon_parameters(t)
    state = [domain_tag % P] + [x % P for x in inputs]
    half = r_f // 2
    for rnd in range(r_f + r_p):
        state = [(s + ark[rnd * t + i]) % P for i, s in enumerate(state)]       # apply_ark :123-129
        if rnd < half or rnd >= half + r_p:
            state = [pow(s, ALPHA, P) for s in state]                          # apply_sbox_full :132-137
        else:
            state[0] = pow(state[0], ALPHA, P)                                 # apply_sbox_partial :140-143
        state = [sum(s * mds[i][j] for j, s in enumerate(state)) % P           # apply_mds :146-157
                 for i in range(t)]
    return state[0]


def poseidon_hash_with_params(ark: Sequence[int], mds: Sequence[Sequence[int]], full_rounds: int, partial_rounds: int,
                              width: int, alpha: int, inputs: Sequence[int], domain_tag: int = 0) -> int:
    """PoseidonHasher::hash for a hasher built with Poseidon::new(PoseidonParameters::new(ark, mds,
    full_rounds, partial_rounds, width, alpha)) (poseidon.rs:47-71, 105-108, 162-208): the same
    schedule with caller-supplied constants, round counts and S-box exponent."""
    if len(inputs) != width - 1:
        raise PoseidonError("InvalidNumberOfInputs", inputs=len(inputs), max_limit=width - 1, width=width)
    state = [domain_tag % P] + [x % P for x in inputs]
    half = full_rounds // 2                                                     # :183
    for rnd in range(full_rounds + partial_rounds):
        state = [(s + ark[rnd * width + i]) % P for i, s in enumerate(state)]
        if rnd < half or rnd >= half + partial_rounds:
            state = [pow(s, alpha, P) for s in state]
        else:
            state[0] = pow(state[0], alpha, P)
        state = [sum(s * mds[i][j] for j, s in enumerate(state)) % P for i in range(width)]
    return state[0]


class Poseidon:
    """Mirror of `Poseidon<Fr>` built with new_circom / with_domain_tag_circom."""

    def __init__(self, nr_inputs: int, domain_tag: int = 0):
        width = nr_inputs + 1
        if width > MAX_X5_LEN:                                                   # poseidon.rs:315-320
            raise PoseidonError("InvalidWidthCircom", width=width, max_limit=MAX_X5_LEN)
        if width < 2:
            # get_poseidon_parameters(1) falls off the end of the if-chain and
            # get_poseidon_parameters(0) is rejected explicitly (parameters.rs:38-42)
            raise PoseidonError("InvalidWidthCircom", width=width, max_limit=MAX_X5_LEN)
        self.width = width
        self.domain_tag = domain_tag % P

    @classmethod
    def new_circom(cls, nr_inputs: int) -> "Poseidon":
        return cls(nr_inputs, 0)

    @classmethod
    def with_domain_tag_circom(cls, nr_inputs: int, domain_tag: int) -> "Poseidon":
        return cls(nr_inputs, domain_tag)

    def hash(self, inputs: Sequence[int]) -> int:
        if len(inputs) != self.width - 1:                                        # poseidon.rs:164-171
            raise PoseidonError("InvalidNumberOfInputs", inputs=len(inputs),
                                max_limit=self.width - 1, width=self.width)
        return poseidon_permute_hash(inputs, self.domain_tag)

    # -- byte front ends (poseidon.rs:213-300) --------------------------------
    @staticmethod
    def _validate(b: bytes):
        if len(b) == 0:                                                          # :261-264
            raise PoseidonError("EmptyInput")
        if len(b) > HASH_LEN:                                                    # :265-271
            raise PoseidonError("InvalidInputLength", len=len(b), modulus_bytes_len=HASH_LEN)

    @staticmethod
    def _to_fr_le(b: bytes) -> int:
        if len(b) != HASH_LEN:                                                   # :282-288
            raise PoseidonError("InvalidInputLength", len=len(b), modulus_bytes_len=HASH_LEN)
        return int.from_bytes(b, "little") % P                                   # :291 (silently reduces)

    def hash_bytes_be(self, inputs: Sequence[bytes]) -> bytes:
        frs = []
        for b in inputs:
            self._validate(b)
            frs.append(self._to_fr_le(bytes(reversed(b))))
        return self.hash(frs).to_bytes(HASH_LEN, "big")

    def hash_bytes_le(self, inputs: Sequence[bytes]) -> bytes:
        frs = []
        for b in inputs:
            self._validate(b)
            frs.append(self._to_fr_le(bytes(b)))
        return self.hash(frs).to_bytes(HASH_LEN, "little")


def hash_be(inputs: Sequence[bytes]) -> bytes:
    """PollStateTree::hash (state.rs:284-302): from_be_bytes_mod_order on every
    32-byte input, circom hasher of matching width, canonical 32-byte BE out."""
    frs = [int.from_bytes(b, "big") % P for b in inputs]
    return Poseidon.new_circom(len(inputs)).hash(frs).to_bytes(HASH_LEN, "big")


# ----------------------------------------------------------------------------
# Zero tables (zeroes.rs) — seeds quoted, chains recomputed
# ----------------------------------------------------------------------------
# zeroes.rs:2  — MACI blank state leaf hash
BINARY_ZERO_LEAF = 6769006970205099520508948723718471724660867171122235270773600567925038008762
# zeroes.rs:38 — MACI "nothing up my sleeve" message-tree zero
QUINARY_ZERO_LEAF = 8370432830353022751713833565135785980866757267633941821328460903436894336785
# zeroes.rs:73-79
EMPTY_BALLOT_ROOTS = [
    16015576667038038422103932363190100635991292382181099511410843174865570503661,
    166510078825589460025300915201657086611944528317298994959376081297530246971,
    10057734083972610459557695472359628128485394923403014377687504571662791937025,
    4904828619307091008204672239231377290495002626534171783829482835985709082773,
    18694062287284245784028624966421731916526814537891066525886866373016385890569,
]
N_ZERO_LEVELS = 33


@lru_cache(maxsize=None)
def merkle_zeroes(arity: int) -> Tuple[bytes, ...]:
    """get_merkle_zeroes (zeroes.rs:81-85): any arity other than 2 gets the
    quinary table.  Z[l+1] = H(Z[l] x arity)."""
    if arity == 2:
        z, k = BINARY_ZERO_LEAF, 2
    else:
        z, k = QUINARY_ZERO_LEAF, 5
    out = [z.to_bytes(32, "big")]
    for _ in range(N_ZERO_LEVELS - 1):
        out.append(hash_be([out[-1]] * k))
    return tuple(out)


# ----------------------------------------------------------------------------
# The tree (state.rs:70-302)
# ----------------------------------------------------------------------------
@dataclass
class PollStateTree:
    arity: int
    full_depth: int
    depth: int = 0
    count: int = 0
    hashes: List[Tuple[int, bytes]] = field(default_factory=list)
    root: Optional[bytes] = None

    @classmethod
    def new(cls, arity: int, full_depth: int, zero_hash: Optional[Tuple[int, bytes]] = None):
        t = cls(arity=arity, full_depth=full_depth)                              # state.rs:142-170
        if zero_hash is not None:
            t.hashes.append(zero_hash)
        return t

    @staticmethod
    def hash(inputs: Sequence[bytes]) -> bytes:
        return hash_be(inputs)

    def insert(self, leaf: bytes) -> "PollStateTree":                            # state.rs:176-225
        if self.root is not None:
            raise MerkleTreeError("TreeAlreadyFull")
        self.count += 1
        self.hashes.append((0, leaf))
        k = self.arity
        while len(self.hashes) >= k:
            sub = self.hashes[-k:]
            d = sub[0][0]
            if not all(e[0] == d for e in sub):
                break
            h = self.hash([e[1] for e in sub])
            del self.hashes[-k:]
            self.hashes.append((d + 1, h))
            if self.depth < d + 1:
                self.depth = d + 1
        if len(self.hashes) == 1 and self.hashes[0][0] == self.full_depth:
            self.root = self.hashes[0][1]
            self.hashes = []
        return self

    def merge(self, to_depth: bool) -> "PollStateTree":                          # state.rs:230-281
        if self.root is not None:
            raise MerkleTreeError("TreeAlreadyMerged")
        zeroes = merkle_zeroes(self.arity)
        k = self.arity
        while self.hashes:
            d = self.hashes[-1][0]
            if len(self.hashes) == 1 and (not to_depth or d == self.full_depth):
                break
            run = []
            for e in reversed(self.hashes):
                if e[0] != d:
                    break
                run.append(e[1])
            run.reverse()
            size = len(run)
            if k >= size:
                run = run + [zeroes[d]] * (k - size)
            h = self.hash(run)
            del self.hashes[len(self.hashes) - size:]
            self.hashes.append((d + 1, h))
        if len(self.hashes) == 1:
            self.root = self.hashes[0][1]
            self.hashes = []
        return self


def new_registration_tree(registration_depth: int) -> PollStateTree:
    """PollState::new, registrations half (state.rs:48-52): arity 2, seeded with
    the blank state leaf at level 0."""
    return PollStateTree.new(2, registration_depth, (0, merkle_zeroes(2)[0]))


def new_interaction_tree(interaction_depth: int) -> PollStateTree:
    """PollState::new, interactions half (state.rs:53-57)."""
    return PollStateTree.new(5, interaction_depth, None)


def merge_registrations(tree: PollStateTree) -> Tuple[PollStateTree, bytes]:
    """provider.rs:289-311 — merge(false) then process commitment
    H3(root, EMPTY_BALLOT_ROOTS[1], 0)."""
    tree.merge(False)
    if tree.root is None:
        raise MerkleTreeError("MergeFailed")
    commitment = hash_be([tree.root, EMPTY_BALLOT_ROOTS[1].to_bytes(32, "big"), bytes(32)])
    return tree, commitment


def merge_interactions(tree: PollStateTree, registrations_count: int,
                       process_subtree_depth: int, tally_subtree_depth: int):
    """provider.rs:313-327 — merge(true) and the expected proof counts."""
    tree.merge(True)
    batch = tree.arity ** process_subtree_depth
    expected_process = tree.count // batch + (1 if tree.count % batch else 0)
    expected_tally = 1 + registrations_count // (2 ** tally_subtree_depth)
    return tree, expected_process, expected_tally


def prepare_public_inputs(registrations: PollStateTree, interactions: PollStateTree, process_commitment: Tuple[int, bytes],
                          tally_commitment: Tuple[int, bytes], process_subtree_depth: int, tally_subtree_depth: int,
                          voting_period_end: int, coordinator_pk: Tuple[bytes, bytes], new_commitment: bytes):
    """provider.rs:141-215 without the verify key: ("process" | "tally", public inputs as
    integers, (proof_index + 1, new_commitment)) or None.  u32 arithmetic as in the reference."""
    message_batch_size = interactions.arity ** process_subtree_depth                       # :150
    current_batch_index = interactions.count
    if current_batch_index > 0:                                                             # :152-157
        r = interactions.count % message_batch_size
        current_batch_index -= message_batch_size if r == 0 else r
    proof_index = process_commitment[0]
    index_offset = proof_index * message_batch_size
    if index_offset <= current_batch_index:                                                 # :162
        coord_hash = poseidon_permute_hash([int.from_bytes(coordinator_pk[0], "big") % P,
                                            int.from_bytes(coordinator_pk[1], "big") % P])  # :166-172
        if interactions.root is None:
            return None
        current_batch_index -= index_offset
        end_batch_index = min(current_batch_index + message_batch_size, interactions.count)
        inputs = [registrations.count + 1, voting_period_end, int.from_bytes(interactions.root, "big") % P,
                  registrations.depth, end_batch_index, current_batch_index, coord_hash,
                  int.from_bytes(process_commitment[1], "big") % P, int.from_bytes(new_commitment, "big") % P]
        return "process", inputs, (proof_index + 1, new_commitment)
    proof_index = tally_commitment[0]                                                       # :196-213
    batch_size = registrations.arity ** tally_subtree_depth
    current_batch_index = proof_index * batch_size
    if current_batch_index >= registrations.count + 1:
        return None
    inputs = [int.from_bytes(process_commitment[1], "big") % P, int.from_bytes(tally_commitment[1], "big") % P,
              int.from_bytes(new_commitment, "big") % P, current_batch_index, registrations.count + 1]
    return "tally", inputs, (proof_index + 1, new_commitment)


# -- leaf hashing (provider.rs:218-287), the "next" row ------------------------
def registration_leaf(pk_x: bytes, pk_y: bytes, timestamp: int) -> bytes:
    """hash4(pk.x, pk.y, 1, timestamp)  (provider.rs:224-233)."""
    return hash_be([pk_x, pk_y, (1).to_bytes(32, "big"), timestamp.to_bytes(32, "big")])


def interaction_leaf(pk_x: bytes, pk_y: bytes, data: Sequence[bytes]) -> bytes:
    """hash4(hash5(d[0..5]), hash5(d[5..10]), pk.x, pk.y)  (provider.rs:249-278)."""
    left = hash_be(list(data[0:5]))
    right = hash_be(list(data[5:10]))
    return hash_be([left, right, pk_x, pk_y])


# ----------------------------------------------------------------------------
# Batch-equivalent dense tree — what the GPU computes
# ----------------------------------------------------------------------------
def dense_tree_root(leaves: Sequence[bytes], arity: int, depth: int) -> bytes:
    """Root of the dense arity-ary tree of exactly `depth` levels over `leaves`,
    with missing children at level l replaced by zeroes[l]."""
    zeroes = merkle_zeroes(arity)
    level = list(leaves)
    if not level:
        return zeroes[depth]
    for l in range(depth):
        nxt = []
        for i in range(0, len(level), arity):
            grp = level[i:i + arity]
            grp = grp + [zeroes[l]] * (arity - len(grp))
            nxt.append(hash_be(grp))
        level = nxt
    assert len(level) == 1
    return level[0]


def batch_merge(arity: int, full_depth: int, leaves: Sequence[bytes], *,
                prepend_blank_leaf: bool, to_depth: bool):
    """One-shot equivalent of new + insert*N + merge(to_depth).

    Returns (root, insert_depth, root_depth, count) where insert_depth is the
    `depth` field after the inserts (what goes into the public signal,
    provider.rs:182) and root_depth the number of levels under `root`.
    """
    zeroes = merkle_zeroes(arity)
    lv = ([zeroes[0]] if prepend_blank_leaf else []) + list(leaves)
    n = len(lv)
    count = len(leaves)
    if n > arity ** full_depth:
        raise MerkleTreeError("TreeAlreadyFull")
    # depth reached by insert(): largest d with arity^d <= n
    insert_depth = 0
    while arity ** (insert_depth + 1) <= n:
        insert_depth += 1
    if n == 0:
        return None, 0, 0, 0
    if to_depth or n == arity ** full_depth:
        root_depth = full_depth
    else:
        root_depth = 0
        while arity ** root_depth < n:
            root_depth += 1
    return dense_tree_root(lv, arity, root_depth), insert_depth, root_depth, count


# ----------------------------------------------------------------------------
# Merkle paths and outcome verification (provider.rs:76-139, 396-436)
# ----------------------------------------------------------------------------
def compute_merkle_root_from_path(depth: int, index: int, leaf: bytes, path, arity: int = 5):
    """provider.rs:396-436 (the reference fixes arity 5, VOTE_TREE_ARITY): at
    each level the node takes position index % arity among the arity-1 siblings
    of path[level] (stored in order with that position skipped)."""
    idx, cur = index, leaf
    for i in range(depth):
        pos = idx % arity
        level = [cur if j == pos else path[i][j - 1 if j > pos else j] for j in range(arity)]
        cur = hash_be(level)
        idx //= arity
    return cur


def verify_outcome(vote_option_tree_depth: int, n_options: int, tally_commitment: bytes, outcome: dict):
    """verify_outcome (provider.rs:76-139) minus the `is_proven` guard: every
    option's tally result must hash, through its Merkle path and the two salts,
    to the tally commitment; so must the total spent.  Returns the index of the
    first option with the largest tally, or None."""
    best, best_val = 0, 0
    for i in range(n_options):
        res = outcome["tally_results"][i]
        root = compute_merkle_root_from_path(vote_option_tree_depth, i, res.to_bytes(32, "big"),
                                             outcome["tally_result_proofs"][i])
        h = hash_be([hash_be([root, outcome["tally_result_salt"]]), outcome["spent_votes_hash"]])
        if h != tally_commitment:
            return None
        if res > best_val:
            best, best_val = i, res
    h = hash_be([outcome["new_results_commitment"], hash_be([outcome["total_spent"], outcome["total_spent_salt"]])])
    if h != tally_commitment:
        return None
    return best


def dense_tree_levels(leaves: Sequence[bytes], arity: int, depth: int):
    """All levels (0 = leaves) of the dense zero-padded tree of `depth` levels."""
    zeroes = merkle_zeroes(arity)
    levels = [list(leaves)]
    for l in range(depth):
        cur, nxt = levels[-1], []
        for i in range(0, len(cur), arity):
            grp = cur[i:i + arity]
            nxt.append(hash_be(grp + [zeroes[l]] * (arity - len(grp))))
        levels.append(nxt)
    return levels


def merkle_path(levels, arity: int, index: int):
    """Sibling path of leaf `index` in the layout compute_merkle_root_from_path
    consumes; nodes to the right of a level's last node are its zero value."""
    zeroes = merkle_zeroes(arity)
    path, idx = [], index
    for l in range(len(levels) - 1):
        pos = idx % arity
        base = idx - pos
        sibs = []
        for j in range(arity):
            if j == pos:
                continue
            sibs.append(levels[l][base + j] if base + j < len(levels[l]) else zeroes[l])
        path.append(sibs)
        idx //= arity
    return path
