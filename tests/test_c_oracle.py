"""Pins the C oracle (oracle/poseidon_oracle.c) against the reference's golden
vectors and against the Python oracle."""
import random

import numpy as np
import pytest

from oracle import c_oracle
from oracle import poseidon_ref as O
from tests.util import EDGE_VALUES, be

H = bytes.fromhex


def test_kats(golden):
    assert c_oracle.hash_one([be(1), be(1)]) == H(golden["fr_one"]["expected_be"])
    g = golden["bytes_ones_twos"]
    assert c_oracle.hash_one([H(x) for x in g["inputs_be"]]) == H(g["expected_be"])
    g = golden["random_input"]
    assert c_oracle.hash_one([H(x) for x in g["inputs_be"]])[::-1] == H(g["expected_le"])
    assert c_oracle.hash_one([be(1), be(2)])[::-1] == H(golden["fr_one_two"]["expected_le"])
    for n in range(1, 13):
        for faithful in (False, True):
            assert c_oracle.hash_one([be(1)] * n, faithful=faithful) == H(golden["circomlibjs_ones"][n - 1])
    g = golden["with_domain_tag"]
    ins = [H(x) for x in g["inputs_be"]]
    assert c_oracle.hash_one(ins, be(0)) == H(g["expected_tag_zero_be"])
    assert c_oracle.hash_one(ins, be(1)) != H(g["expected_tag_zero_be"])


@pytest.mark.parametrize("arity,key", [(2, "binary_zeroes"), (5, "quinary_zeroes")])
def test_zero_chains(golden, arity, key):
    z = [H(x) for x in golden[key]]
    for l in range(32):
        assert c_oracle.hash_one([z[l]] * arity) == z[l + 1]
    # merging an empty quinary tree seeded with nothing but one zero leaf walks the table
    rc, root, depth, count = c_oracle.tree_insert_merge(arity, 32, False, True, z[0])
    assert rc == 0 and root == z[32]


def test_tree_pins(golden):
    blk = golden["merge_registration_state_success"]["registration_block"]
    leaves = b"".join(O.registration_leaf(H(p["x"]), H(p["y"]), blk) for p in golden["participants"])
    rc, root, depth, count = c_oracle.tree_insert_merge(2, golden["poll_config"]["registration_depth"], True,
                                                        False, leaves)
    assert rc == 0 and root == H(golden["merge_registration_state_success"]["registrations_root"])
    assert depth == golden["process_messages_public_signals"]["registrations_depth"] and count == 3
    assert c_oracle.hash_one([root, be(O.EMPTY_BALLOT_ROOTS[1]), bytes(32)]) == \
        H(golden["merge_registration_state_success"]["process_commitment"])
    p = golden["participant"]
    leaf = O.interaction_leaf(H(p["shared_pk"]["x"]), H(p["shared_pk"]["y"]), [H(x) for x in p["message"]])
    rc, root, depth, count = c_oracle.tree_insert_merge(5, golden["poll_config"]["interaction_depth"], False,
                                                        True, leaf)
    assert rc == 0 and root == H(golden["merge_interaction_state_success"]["interactions_root"])


def test_against_python_oracle_random_and_edges():
    rng = random.Random(9)
    for k in (1, 2, 3, 4, 5, 12):
        for _ in range(6):
            ins = [be(rng.choice(EDGE_VALUES) if rng.random() < 0.3 else rng.randrange(1 << 256)) for _ in range(k)]
            assert c_oracle.hash_one(ins) == O.hash_be(ins)


def test_batch_threads_agree():
    data = np.random.default_rng(4).integers(0, 256, size=512 * 64, dtype=np.uint8)
    a = c_oracle.hash_batch(2, data, threads=1)
    b = c_oracle.hash_batch(2, data, threads=3)
    assert (a == b).all()
    assert a[7].tobytes() == O.hash_be([data[7 * 64:7 * 64 + 32].tobytes(), data[7 * 64 + 32:8 * 64].tobytes()])


@pytest.mark.parametrize("arity,full_depth,blank,to_depth", [(2, 8, True, False), (5, 3, False, True),
                                                            (2, 5, False, True), (5, 3, True, False)])
def test_insert_merge_equals_dense_tree(arity, full_depth, blank, to_depth):
    rng = random.Random(arity * 100 + full_depth)
    cap = arity ** full_depth
    for n in [0, 1, 2, 3, 4, 5, 6, 24, 25, 26, 31, 32, 33, 100, 124, 125, 126, 200, cap - 2, cap - 1, cap, cap + 1]:
        if n < 0 or n > 300:
            continue
        leaves = [be(rng.randrange(O.P)) for _ in range(n)]
        total = n + (1 if blank else 0)
        rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, b"".join(leaves))
        if total > cap:
            assert rc == 1
            continue
        if total == cap:
            assert rc == 2 and root is not None          # completed by insert; merge refuses
            lv = ([O.merkle_zeroes(arity)[0]] if blank else []) + leaves
            assert root == c_oracle.dense_tree_root(arity, full_depth, b"".join(lv))
            continue
        r2, idp, rdp, cnt = O.batch_merge(arity, full_depth, leaves, prepend_blank_leaf=blank, to_depth=to_depth)
        assert rc == 0 and root == r2 and depth == idp and count == cnt


def test_leaf_helpers_match_python_oracle(golden):
    p = golden["participant"]
    pk = H(p["shared_pk"]["x"]) + H(p["shared_pk"]["y"])
    msg = b"".join(H(x) for x in p["message"])
    assert c_oracle.interaction_leaves(pk, msg)[0].tobytes() == \
        O.interaction_leaf(H(p["shared_pk"]["x"]), H(p["shared_pk"]["y"]), [H(x) for x in p["message"]])
    pks = b"".join(H(q["x"]) + H(q["y"]) for q in golden["participants"])
    got = c_oracle.registration_leaves(pks, [2, 2, 2 ** 40 + 5])
    for i, (q, ts) in enumerate(zip(golden["participants"], [2, 2, 2 ** 40 + 5])):
        assert got[i].tobytes() == O.registration_leaf(H(q["x"]), H(q["y"]), ts)
