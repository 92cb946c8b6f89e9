"""Parity tests proper: the CUDA path, called through the C ABI, against the
oracle and the reference's golden vectors.  Bit-exact everywhere (integer
arithmetic mod p).  Names follow the reference's tests
(pallet/src/tests/poseidon.rs, pallet/src/tests/extrinsics.rs)."""
import random

import numpy as np
import pytest

from oracle import c_oracle
from oracle import poseidon_ref as O
from tests.util import EDGE_VALUES, P, be, random_fr_bytes

pytestmark = pytest.mark.gpu
H = bytes.fromhex


@pytest.fixture(scope="module")
def ib():
    import infimum_b200
    infimum_b200.get_context(0)
    return infimum_b200


# ---- pallet/src/tests/poseidon.rs ------------------------------------------------
def test_fr_one(ib, golden):
    h = ib.Poseidon.new_circom(2)
    for raw in (b"\x01", b"\x00\x01", b"\x00\x00\x01"):
        x = int.from_bytes(raw, "big")
        assert be(h.hash([x, x])) == H(golden["fr_one"]["expected_be"])


def test_bytes_ones_twos(ib, golden):
    g = golden["bytes_ones_twos"]
    ins = [H(x) for x in g["inputs_be"]]
    h = ib.Poseidon.new_circom(2)
    assert be(h.hash([int.from_bytes(b, "big") % P for b in ins])) == H(g["expected_be"])
    assert h.hash_bytes_be(ins) == H(g["expected_be"])
    assert h.hash_bytes_le(ins) == H(g["expected_le"])


def test_with_domain_tag(ib, golden):
    g = golden["with_domain_tag"]
    ins = [int.from_bytes(H(x), "big") % P for x in g["inputs_be"]]
    assert be(ib.Poseidon.with_domain_tag_circom(2, 0).hash(ins)) == H(g["expected_tag_zero_be"])
    tagged = ib.Poseidon.with_domain_tag_circom(2, 1).hash(ins)
    assert be(tagged) != H(g["expected_tag_zero_be"])
    assert tagged == O.Poseidon.with_domain_tag_circom(2, 1).hash(ins)
    # byte front ends carry the tag too, in either wire order
    raw = [H(x) for x in g["inputs_be"]]
    assert ib.Poseidon.with_domain_tag_circom(2, 77).hash_bytes_be(raw) == \
        O.Poseidon.with_domain_tag_circom(2, 77).hash_bytes_be(raw)
    assert ib.Poseidon.with_domain_tag_circom(2, 77).hash_bytes_le(raw) == \
        O.Poseidon.with_domain_tag_circom(2, 77).hash_bytes_le(raw)


def test_fr_one_two(ib, golden):
    assert ib.Poseidon.new_circom(2).hash([1, 2]).to_bytes(32, "little") == H(golden["fr_one_two"]["expected_le"])


def test_random_input(ib, golden):
    # both inputs >= p: the kernels must reduce like from_be_bytes_mod_order
    g = golden["random_input"]
    raw = b"".join(H(x) for x in g["inputs_be"])
    out = ib.Poseidon.new_circom(2).hash_batch(raw, 1)
    assert out.tobytes()[::-1] == H(g["expected_le"])


def test_empty_input(ib):
    for n in range(1, 12):
        h = ib.Poseidon.new_circom(n)
        for ins in ([b""] * n, [b"\x01" * 32] * (n - 1) + [b""]):
            for fn in (h.hash_bytes_be, h.hash_bytes_le):
                with pytest.raises(ib.PoseidonError) as e:
                    fn(ins)
                assert e.value.kind == "EmptyInput"


def test_input_length_and_width_errors(ib):
    h = ib.Poseidon.new_circom(2)
    for bad in (b"\x01" * 33, b"\x01" * 31, b"\x01"):
        with pytest.raises(ib.PoseidonError) as e:
            h.hash_bytes_be([bad, b"\x01" * 32])
        assert e.value.kind == "InvalidInputLength"
    with pytest.raises(ib.PoseidonError) as e:
        h.hash([1])
    assert e.value.kind == "InvalidNumberOfInputs"
    with pytest.raises(ib.PoseidonError) as e:
        h.hash_bytes_be([b"\x01" * 32])
    assert e.value.kind == "InvalidNumberOfInputs"
    for n in (0, 13, 14):
        with pytest.raises(ib.PoseidonError) as e:
            ib.Poseidon.new_circom(n)
        assert e.value.kind == "InvalidWidthCircom"


def test_circomlibjs_compat_1_to_12_inputs(ib, golden):
    one, two = be(1), be(2)
    for n in range(1, 13):
        h = ib.Poseidon.new_circom(n)
        assert h.hash_bytes_be([one] * n) == H(golden["circomlibjs_ones"][n - 1])
        assert h.hash_bytes_be([two] * n) != H(golden["circomlibjs_ones"][n - 1])


# ---- pallet/src/poll/zeroes.rs ------------------------------------------------------
@pytest.mark.parametrize("arity,key", [(2, "binary_zeroes"), (5, "quinary_zeroes")])
def test_zero_tables(ib, golden, arity, key):
    table = [H(x) for x in golden[key]]
    assert ib.get_merkle_zeroes(arity) == table
    # one batch call re-derives all 32 links of the chain
    ins = b"".join(z * arity for z in table[:32])
    out = ib.Poseidon.new_circom(arity).hash_batch(ins)
    assert [out[i].tobytes() for i in range(32)] == table[1:]
    assert ib.get_merkle_zeroes(7) == ib.get_merkle_zeroes(5)
    assert ib.empty_ballot_roots() == [H(x) for x in golden["empty_ballot_roots"]]


# ---- pallet/src/tests/extrinsics.rs ---------------------------------------------------
def _registration_leaves(ib, golden):
    blk = golden["merge_registration_state_success"]["registration_block"]
    h4 = ib.Poseidon.new_circom(4)
    return [h4.hash_bytes_be([H(p["x"]), H(p["y"]), be(1), be(blk)]) for p in golden["participants"]]


def test_merge_registration_state_success(ib, golden):
    g = golden["merge_registration_state_success"]
    t = ib.new_registration_tree(golden["poll_config"]["registration_depth"])
    for leaf in _registration_leaves(ib, golden):
        t.insert(leaf)
    t, commitment = ib.merge_registrations(t)
    assert t.root == H(g["registrations_root"])
    assert commitment == H(g["process_commitment"])
    assert t.hashes == []


def test_merge_interaction_state_success(ib, golden):
    g = golden["merge_interaction_state_success"]
    cfg = golden["poll_config"]
    p = golden["participant"]
    h5, h4 = ib.Poseidon.new_circom(5), ib.Poseidon.new_circom(4)
    msg = [H(x) for x in p["message"]]
    leaf = h4.hash_bytes_be([h5.hash_bytes_be(msg[:5]), h5.hash_bytes_be(msg[5:]),
                             H(p["shared_pk"]["x"]), H(p["shared_pk"]["y"])])
    t = ib.new_interaction_tree(cfg["interaction_depth"]).insert(leaf)
    t, ep, et = ib.merge_interactions(t, 3, cfg["process_subtree_depth"], cfg["tally_subtree_depth"])
    assert t.root == H(g["interactions_root"])
    assert (ep, et) == (g["expected_process"], g["expected_tally"])


def test_process_messages_public_signals(ib, golden):
    g = golden["process_messages_public_signals"]
    t = ib.new_registration_tree(golden["poll_config"]["registration_depth"])
    for leaf in _registration_leaves(ib, golden):
        t.insert(leaf)
    t, commitment = ib.merge_registrations(t)
    assert t.count + 1 == g["registrations_count_plus_one"]
    assert t.depth == g["registrations_depth"]
    assert commitment == H(g["process_commitment"])
    pk = golden["coordinator_pk"]
    hsh = ib.Poseidon.new_circom(2).hash([int(pk["x"], 16), int(pk["y"], 16)])
    assert str(hsh) == g["coord_pub_key_hash_decimal"]
    # the whole sequence of the reference test through the Poll mirror, ending in the nine
    # public inputs listed at extrinsics.rs:621-633 (prepare_public_inputs, provider.rs:141-215)
    cfg = golden["poll_config"]
    poll = ib.Poll(ib.PollConfig(cfg["registration_depth"], cfg["interaction_depth"], cfg["process_subtree_depth"],
                                 cfg["tally_subtree_depth"], cfg["signup_period"], cfg["voting_period"]),
                   created_at=g["created_at_block"])
    for p in golden["participants"]:
        poll.register_participant((H(p["x"]), H(p["y"])), golden["merge_registration_state_success"]["registration_block"])
    poll.merge_registrations()
    msg = golden["participant"]
    poll.consume_interaction((H(msg["shared_pk"]["x"]), H(msg["shared_pk"]["y"])), [H(x) for x in msg["message"]])
    poll.merge_interactions()
    exp = [int(x) for x in g["expected_public_inputs_decimal"]]
    kind, inputs, nxt = poll.prepare_public_inputs((H(pk["x"]), H(pk["y"])), be(exp[8]))
    assert kind == "process" and inputs == exp and nxt.process == (1, be(exp[8]))
    poll.commitment = nxt
    kind, inputs, nxt = poll.prepare_public_inputs((H(pk["x"]), H(pk["y"])), bytes(32))
    assert kind == "tally" and inputs == [exp[8], 0, 0, 0, 4] and nxt.tally[0] == 1


def test_participant_limit_reached_quirk(ib):
    # extrinsics.rs:301-316: depth-2 tree, blank + 3 leaves is completed by insert
    t = ib.new_registration_tree(2)
    for i in range(3):
        t.insert(be(100 + i))
    assert t.root is not None and t.hashes == []
    o = O.new_registration_tree(2)
    for i in range(3):
        o.insert(be(100 + i))
    assert t.root == o.root and t.depth == o.depth
    with pytest.raises(ib.MerkleTreeError) as e:
        t.merge(False)
    assert e.value.code == 2
    with pytest.raises(ib.MerkleTreeError) as e:
        t.insert(be(7))
    assert e.value.code == 1


# ---- oracle parity, every width, edge values ---------------------------------------------
@pytest.mark.parametrize("n_inputs", range(1, 13))
def test_hash_batch_vs_oracle_all_widths(ib, n_inputs):
    rng = random.Random(300 + n_inputs)
    rows = [[v] * n_inputs for v in EDGE_VALUES]
    rows += [[rng.choice(EDGE_VALUES) for _ in range(n_inputs)] for _ in range(40)]
    rows += [[rng.randrange(1 << 256) for _ in range(n_inputs)] for _ in range(200)]
    raw = b"".join(be(x) for r in rows for x in r)
    got = ib.Poseidon.new_circom(n_inputs).hash_batch(raw)
    exp = c_oracle.hash_batch(n_inputs, raw)
    assert (got == exp).all()
    # little-endian wire order
    raw_le = b"".join(x.to_bytes(32, "little") for r in rows for x in r)
    got_le = ib.Poseidon.new_circom(n_inputs).hash_batch(raw_le, little_endian=True)
    assert (got_le[:, ::-1] == exp).all()
    # x and x + p hash alike whenever x + p still fits 256 bits
    shifted = b"".join(be(x + P if x + P < (1 << 256) else x) for r in rows for x in r)
    assert (ib.Poseidon.new_circom(n_inputs).hash_batch(shifted) == exp).all()


@pytest.mark.parametrize("n_inputs", [1, 2, 3, 4, 5, 7])
def test_optimised_kernel_equals_dense_kernel_on_device(ib, n_inputs):
    """Two device implementations that share no tables: sparse/lazy vs the
    reference schedule taken literally."""
    raw = random_fr_bytes(4096 * n_inputs, seed=n_inputs, canonical=False)
    h = ib.Poseidon.new_circom(n_inputs)
    assert (h.hash_batch(raw) == h.hash_batch(raw, dense=True)).all()


@pytest.mark.parametrize("n_inputs", range(1, 8))
def test_small_and_large_batches_take_different_kernels_same_results(ib, n_inputs):
    """Batches of at most 12 288 (widths <= 4) / 8 192 hashes run the warp-cooperative kernel (one
    hash spread over width + 1 warps), larger ones one hash per thread: both against the oracle, at the
    sizes where the cooperative kernel changes shape (one lane, a ragged warp, more than one block per
    SM = rotated role layout, its largest batch), with a domain tag and in either wire order."""
    n = 16384 + 4000
    raw = random_fr_bytes(n_inputs * n, seed=70 + n_inputs, canonical=False).reshape(-1)
    exp = c_oracle.hash_batch(n_inputs, raw)
    h = ib.Poseidon.new_circom(n_inputs)
    assert (h.hash_batch(raw) == exp).all()
    row = n_inputs * 32
    for m in (1, 31, 33, 4736 + 5, 8192, 8193, 12288, 12289, 16384):
        assert (h.hash_batch(raw[: m * row]) == exp[:m]).all(), m
    tag = 0x1234567890abcdef << 100
    ht = ib.Poseidon.with_domain_tag_circom(n_inputs, tag)
    rows = raw[: 40 * row].reshape(40, n_inputs, 32)
    exp_t = np.stack([np.frombuffer(c_oracle.hash_one([bytes(x) for x in r], be(tag)), dtype=np.uint8) for r in rows])
    assert (ht.hash_batch(raw[: 40 * row]) == exp_t).all()
    assert (ht.hash_batch(raw[: n * row])[:40] == exp_t).all()
    le = np.ascontiguousarray(rows[:, :, ::-1])
    assert (ht.hash_batch(le, little_endian=True)[:, ::-1] == exp_t).all()


@pytest.mark.parametrize("n_inputs", [2, 5])
def test_batches_at_wave_boundaries_of_the_wide_launch_shapes(ib, n_inputs):
    """Large launches run one 256-, 384- or 512-thread block per SM, chosen per launch; sizes just below, at
    and just above whole waves of every shape (the ragged last block must neither drop nor invent a hash)."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    sizes = sorted({sms * b * w + d for b in (256, 384, 512) for w in (1, 2) for d in (-1, 0, 1)})
    row = n_inputs * 32
    raw = random_fr_bytes(n_inputs * sizes[-1], seed=90 + n_inputs).reshape(-1)
    exp = c_oracle.hash_batch(n_inputs, raw)
    h = ib.Poseidon.new_circom(n_inputs)
    for n in sizes:
        out = np.full((n + 1, 32), 0xA5, dtype=np.uint8)
        h.hash_batch(raw[: n * row], n, out=out[:n])
        assert (out[:n] == exp[:n]).all(), n
        assert (out[n] == 0xA5).all(), n


@pytest.mark.parametrize("arity,depth", [(2, 18), (5, 8)])
def test_tree_levels_at_wave_boundaries(ib, arity, depth):
    """The same for tree levels: level 0 one parent more than a whole wave of 384-thread blocks (ragged last
    group included), the levels above at whatever shape their size picks."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for n in (arity * sms * 384 + 1, arity * sms * 512 - 1):
        leaves = random_fr_bytes(n, seed=n % 1000)
        t = ib.PollStateTree.new(arity, depth).extend(leaves).merge(True)
        rc, root, d, c = c_oracle.tree_insert_merge(arity, depth, False, True, leaves)
        assert rc == 0 and t.root == root and t.depth == d, (arity, n)


def test_hash2_2_17_pairs_bit_exact_vs_oracle(ib):
    """A quick 2^17-pair batch, every output compared (the full 2^24 pairs of
    BASELINE config 2 are in tests/test_gpu_fullsize.py)."""
    n = 1 << 17
    raw = random_fr_bytes(2 * n)
    got = ib.Poseidon.new_circom(2).hash_batch(raw)
    assert (got == c_oracle.hash_batch(2, raw)).all()


def test_hash5_bit_exact_vs_oracle(ib):
    n = 1 << 15
    raw = random_fr_bytes(5 * n, seed=5)
    got = ib.Poseidon.new_circom(5).hash_batch(raw)
    assert (got == c_oracle.hash_batch(5, raw)).all()


def test_pageable_pinned_and_registered_host_buffers_agree(ib):
    """Large batches from pageable memory go through the library's threaded pinned staging;
    the same batch from inf_host_alloc'ed and from inf_host_register'ed memory is copied by
    DMA directly.  Same outputs, ragged last chunk included, checked against the oracle."""
    import ctypes as C
    ctx = ib.get_context()
    n = (1 << 20) + 3 * (1 << 18) + 12345
    for k in (2, 5):
        raw = random_fr_bytes(k * n, seed=40 + k, canonical=False)
        h = ib.Poseidon.new_circom(k)
        got = h.hash_batch(raw)                                      # pageable numpy arrays in and out
        idx = np.r_[0:64, (1 << 18) - 32:(1 << 18) + 32, (1 << 20) - 32:(1 << 20) + 32, n - 64:n]
        exp = c_oracle.hash_batch(k, raw.reshape(n, k * 32)[idx])
        assert (got[idx] == exp).all()
        if not hasattr(ctx.lib, "_name"):
            continue
        p_in, p_out = C.c_void_p(), C.c_void_p()
        ctx.check(ctx.lib.inf_host_alloc(ctx.handle, raw.size, C.byref(p_in)))
        ctx.check(ctx.lib.inf_host_alloc(ctx.handle, n * 32, C.byref(p_out)))
        C.memmove(p_in.value, raw.ctypes.data, raw.size)
        ctx.check(ctx.lib.inf_poseidon_hash_batch(ctx.handle, k, 0, None, p_in.value, n, p_out.value))
        pinned = np.frombuffer(C.string_at(p_out.value, n * 32), dtype=np.uint8).reshape(n, 32)
        assert (pinned == got).all()
        ctx.check(ctx.lib.inf_host_free(ctx.handle, p_in))
        ctx.check(ctx.lib.inf_host_free(ctx.handle, p_out))
        out2 = np.empty((n, 32), dtype=np.uint8)
        ctx.check(ctx.lib.inf_host_register(ctx.handle, raw.ctypes.data, raw.size))
        ctx.check(ctx.lib.inf_host_register(ctx.handle, out2.ctypes.data, out2.size))
        h.hash_batch(raw, out=out2)
        ctx.check(ctx.lib.inf_host_unregister(ctx.handle, raw.ctypes.data))
        ctx.check(ctx.lib.inf_host_unregister(ctx.handle, out2.ctypes.data))
        assert (out2 == got).all()


# ---- trees -----------------------------------------------------------------------------------
SIZES = sorted(set(list(range(0, 41)) + [63, 64, 65, 100, 124, 125, 126, 127, 128, 129, 255, 256, 257,
                                          624, 625, 626, 1000, 1023, 1024, 1025, 3124, 3125, 3126, 5000]))


@pytest.mark.parametrize("arity,full_depth,blank,to_depth", [(2, 13, True, False), (5, 6, False, True),
                                                            (2, 13, False, True), (5, 6, True, False),
                                                            (2, 13, True, True), (5, 6, False, False)])
def test_tree_merge_equals_insert_merge(ib, arity, full_depth, blank, to_depth):
    """gpu_tree(leaves) == oracle new + insert*N + merge, for ragged sizes."""
    allv = random_fr_bytes(max(SIZES), seed=arity * 10 + full_depth)
    for n in SIZES:
        leaves = allv[:n]
        rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        t = ib.PollStateTree.new(arity, full_depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None)
        t.extend(leaves)
        t.merge(to_depth)
        assert rc == 0
        assert t.root == root, (arity, n)
        assert (t.depth, t.count) == (depth, count), (arity, n)


@pytest.mark.parametrize("arity,full_depth", [(2, 6), (5, 3)])
def test_tree_capacity_edges(ib, arity, full_depth):
    cap = arity ** full_depth
    allv = random_fr_bytes(cap + 1, seed=99)
    for blank in (False, True):
        for n in (cap - 2, cap - 1, cap, cap + 1):
            total = n + (1 if blank else 0)
            rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, True, allv[:n])
            t = ib.PollStateTree.new(arity, full_depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None)
            if total > cap:
                assert rc == 1
                with pytest.raises(ib.MerkleTreeError) as e:
                    t.extend(allv[:n])
                assert e.value.code == 1
                continue
            t.extend(allv[:n])
            if total == cap:
                assert rc == 2 and t.root == root and t.depth == depth
                with pytest.raises(ib.MerkleTreeError) as e:
                    t.merge(True)
                assert e.value.code == 2
            else:
                t.merge(True)
                assert rc == 0 and t.root == root and t.depth == depth


def test_state_tree_2_16_registrations(ib):
    """BASELINE config 3 shape at an oracle-affordable size: blank leaf + 2^16
    registrations, depth field 16, root depth 17."""
    leaves = random_fr_bytes(1 << 16, seed=3)
    t = ib.new_registration_tree(20).extend(leaves)
    t.merge(False)
    lv = np.concatenate([np.frombuffer(ib.get_merkle_zeroes(2)[0], dtype=np.uint8).reshape(1, 32), leaves])
    assert t.root == c_oracle.dense_tree_root(2, 17, lv)
    assert t.depth == 16 and t.count == 1 << 16


def test_message_tree_depth_12_padding(ib):
    """Quinary tree with full_depth 12 over 5^6+17 leaves: six levels of real
    work, then zero-sibling padding up to depth 12 (merge(true))."""
    n = 5 ** 6 + 17
    leaves = random_fr_bytes(n, seed=12)
    t = ib.new_interaction_tree(12).extend(leaves)
    t.merge(True)
    assert t.root == c_oracle.dense_tree_root(5, 12, leaves)
    rc, root, depth, count = c_oracle.tree_insert_merge(5, 12, False, True, leaves)
    assert rc == 0 and root == t.root and depth == t.depth == 6


def test_full_size_properties_2_24(ib):
    """At BASELINE's full size the oracle cannot follow; use size-independent
    properties: (1) a tree over [x]*2^24 must equal 24 chained self-hashes;
    (2) root(2^24 leaves) == H(root(left half), root(right half))."""
    n = 1 << 24
    x = random_fr_bytes(1, seed=24)
    leaves = np.repeat(x, n, axis=0)
    t = ib.PollStateTree.new(2, 24).extend(leaves)           # completed by insert: n == 2^24
    node = x[0].tobytes()
    for _ in range(24):
        node = c_oracle.hash_one([node, node])
    assert t.root == node
    leaves = random_fr_bytes(n, seed=2424)
    whole = ib.PollStateTree.new(2, 24).extend(leaves).root
    left = ib.PollStateTree.new(2, 23).extend(leaves[: n // 2]).root
    right = ib.PollStateTree.new(2, 23).extend(leaves[n // 2:]).root
    assert whole == c_oracle.hash_one([left, right])


# ---- Poseidon::new(params): caller-supplied PoseidonParameters (poseidon.rs:47-71, 105-108) ----
def test_custom_parameters_equal_new_circom_for_the_circom_tables(ib, golden):
    for n_inputs in (1, 2, 5, 12):
        ark, mds, rf, rp = O.poseidon_parameters(n_inputs + 1)
        h = ib.Poseidon.new(ib.PoseidonParameters(ark, mds, rf, rp, n_inputs + 1, 5))
        assert h.hash([1] * n_inputs) == int(golden["circomlibjs_ones"][n_inputs - 1], 16)
    ark, mds, rf, rp = O.poseidon_parameters(3)
    h = ib.Poseidon.new(ib.PoseidonParameters(ark, mds, rf, rp, 3, 5))
    a, b = H(golden["random_input"]["inputs_be"][0]), H(golden["random_input"]["inputs_be"][1])
    assert h.hash_bytes_le([a[::-1], b[::-1]]) == H(golden["random_input"]["expected_le"])
    with pytest.raises(ib.PoseidonError) as e:
        h.hash([1])
    assert e.value.kind == "InvalidNumberOfInputs"


@pytest.mark.parametrize("width,full_rounds,partial_rounds,alpha", [(4, 7, 3, 3), (2, 2, 0, 7), (13, 4, 2, 5),
                                                                    (3, 0, 5, 17), (5, 6, 9, 1), (3, 0, 0, 5)])
def test_custom_parameters_vs_oracle(ib, width, full_rounds, partial_rounds, alpha):
    """Arbitrary constants, odd full_rounds (the second half gets the extra round,
    poseidon.rs:183-203), other S-box exponents, with and without a domain tag."""
    rng = random.Random(width * 1000 + full_rounds * 10 + alpha)
    ark = [rng.randrange(P) for _ in range((full_rounds + partial_rounds) * width)]
    mds = [[rng.randrange(P) for _ in range(width)] for _ in range(width)]
    n = 37
    rows = [[rng.choice(EDGE_VALUES + [rng.randrange(P)]) for _ in range(width - 1)] for _ in range(n)]
    buf = b"".join(be(x % (1 << 256)) for r in rows for x in r)
    for tag in (0, 12345):
        h = ib.Poseidon.with_domain_tag(ib.PoseidonParameters(ark, mds, full_rounds, partial_rounds, width, alpha), tag)
        got = h.hash_batch(buf, n)
        for i, r in enumerate(rows):
            exp = O.poseidon_hash_with_params(ark, mds, full_rounds, partial_rounds, width, alpha, r, tag)
            assert got[i].tobytes() == be(exp), (i, tag)


# ---- frontier (`PollStateTree.hashes` before the merge, state.rs:85-86) -----------------------
@pytest.mark.parametrize("arity,full_depth,blank", [(2, 11, True), (5, 5, False), (2, 11, False), (5, 5, True)])
def test_frontier_equals_insert_cascade(ib, arity, full_depth, blank):
    allv = random_fr_bytes(700, seed=arity + 40)
    for n in [0, 1, 2, 3, 4, 5, 6, 7, 8, 24, 25, 26, 31, 32, 33, 124, 125, 126, 255, 256, 511, 624, 625, 626, 700]:
        leaves = allv[:n]
        o = O.PollStateTree.new(arity, full_depth, (0, O.merkle_zeroes(arity)[0]) if blank else None)
        for i in range(n):
            o.insert(leaves[i].tobytes())
        t = ib.PollStateTree.new(arity, full_depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None)
        t.extend(leaves)
        assert t.hashes == o.hashes, (arity, n)
        assert (t.depth, t.count) == (o.depth, o.count)


def _resume(ib, o):
    """A mirror tree resumed from the oracle tree's persisted fields."""
    return ib.PollStateTree.from_state(o.arity, o.full_depth, o.depth, o.count, list(o.hashes), o.root)


@pytest.mark.parametrize("arity,full_depth,blank", [(2, 12, True), (5, 5, False), (2, 12, False), (5, 5, True)])
def test_insert_and_merge_on_a_stored_frontier(ib, arity, full_depth, blank):
    """state.rs:176-281 on persisted state (lib.rs:706-714): resume from the frontier
    after k inserts, append the rest in one batch, and get the oracle's incremental
    state — frontier, depth, count — then merge from that frontier, both ways."""
    import copy
    allv = random_fr_bytes(640, seed=arity + 70)
    rng = random.Random(arity * 100 + full_depth)
    for n in [0, 1, 2, 5, 24, 25, 26, 124, 125, 126, 255, 256, 257, 626, 640]:
        splits = sorted({0, n, rng.randrange(n + 1), rng.randrange(n + 1)})
        o = O.PollStateTree.new(arity, full_depth, (0, O.merkle_zeroes(arity)[0]) if blank else None)
        states = {}
        for i in range(n + 1):
            if i in splits:
                states[i] = copy.deepcopy(o)
            if i < n:
                o.insert(allv[i].tobytes())
        for k in splits:
            t = _resume(ib, states[k]).extend(allv[k:n])
            assert t.hashes == o.hashes, (arity, n, k)
            assert (t.depth, t.count, t.root) == (o.depth, o.count, o.root), (arity, n, k)
        for to_depth in (False, True):
            m = copy.deepcopy(o).merge(to_depth)
            t = _resume(ib, o).merge(to_depth)
            assert (t.root, t.hashes, t.depth, t.count) == (m.root, m.hashes, m.depth, m.count), (arity, n, to_depth)


@pytest.mark.parametrize("arity,full_depth", [(2, 6), (5, 3)])
def test_stored_frontier_capacity_and_errors(ib, arity, full_depth):
    import copy
    import ctypes as C
    cap = arity ** full_depth
    allv = random_fr_bytes(cap + 1, seed=98)
    for blank in (False, True):
        n = cap - (1 if blank else 0)
        o = O.PollStateTree.new(arity, full_depth, (0, O.merkle_zeroes(arity)[0]) if blank else None)
        k = n // 3
        for i in range(k):
            o.insert(allv[i].tobytes())
        mid = copy.deepcopy(o)
        with pytest.raises(ib.MerkleTreeError) as e:            # one leaf too many: nothing is inserted
            _resume(ib, mid).extend(allv[k:n + 1])
        assert e.value.code == 1
        for i in range(k, n):
            o.insert(allv[i].tobytes())
        t = _resume(ib, mid).extend(allv[k:n])                   # completed by insert (state.rs:218-222)
        assert o.root is not None and (t.root, t.hashes, t.depth, t.count) == (o.root, [], o.depth, o.count)
        with pytest.raises(ib.MerkleTreeError) as e:
            t.merge(True)
        assert e.value.code == 2
    # malformed frontiers and short output arrays, at the C ABI
    ctx = ib.get_context()
    if not hasattr(ctx.lib, "_name"):
        return                                                   # (the CPU stand-in does not model these)
    h = allv[:8].tobytes()
    out_l, out_h = C.create_string_buffer(8), C.create_string_buffer(8 * 32)
    nn, dd, has = C.c_uint32(), C.c_uint32(), C.c_int()
    root = C.create_string_buffer(32)

    def append(levels, n_in, leaves, n, capn=8):
        return ctx.lib.inf_tree_append(ctx.handle, arity, full_depth, bytes(levels), h, n_in, 0, leaves, n, out_l, out_h,
                                       capn, C.byref(nn), C.byref(dd), C.byref(has), root)
    lv = allv[4:10].tobytes()
    assert append([0, 1], 2, lv, 1) == 28                       # levels must not increase towards the tail
    assert append([1] * arity, arity, lv, 1) == 28              # a full group would have been hashed
    assert append([full_depth, 0], 2, lv, 1) == 28
    assert append([1, 0], 2, None, 0, capn=1) == 27
    assert append([1, 0], 2, lv, 1) == 0 and nn.value >= 1
    assert ctx.lib.inf_tree_merge_frontier(ctx.handle, arity, full_depth, bytes([0, 1]), h, 2, 0, root, C.byref(has),
                                           C.byref(dd)) == 28


def test_hypothesis_tree_and_reduction_properties(ib):
    """gpu_tree(leaves) == oracle insert*N + merge, and
    hash(x) == hash(x + p) for x < 2^256 - p, on hypothesis-drawn cases."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    h2 = ib.Poseidon.new_circom(2)

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(0, (1 << 256) - 1 - P), st.integers(0, (1 << 256) - 1))
    def reduction(x, y):
        a = h2.hash_batch(be(x) + be(y))
        b = h2.hash_batch(be(x + P) + be(y))
        assert (a == b).all()
        assert a.tobytes() == c_oracle.hash_one([be(x), be(y)])

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.sampled_from([2, 5]), st.integers(0, 400), st.booleans(), st.booleans(), st.integers(0, 2 ** 32))
    def tree(arity, n, blank, to_depth, seed):
        full_depth = 9 if arity == 2 else 4
        if n + blank >= arity ** full_depth:
            n = arity ** full_depth - 2
        leaves = random_fr_bytes(max(n, 1), seed=seed, canonical=False)[:n]
        rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        t = ib.PollStateTree.new(arity, full_depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None)
        t.extend(leaves).merge(to_depth)
        assert rc == 0 and t.root == root and (t.depth, t.count) == (depth, count)

    reduction()
    tree()


# ---- leaf hashing (provider.rs:218-287) ------------------------------------------------------------
def test_leaf_hashing_pins_and_oracle(ib, golden):
    """The reference pins leaf hashing only through the tree roots: rebuild both
    pinned trees from raw keys / messages with the fused leaf kernels."""
    cfg = golden["poll_config"]
    poll = ib.Poll(ib.PollConfig(cfg["registration_depth"], cfg["interaction_depth"],
                                 cfg["process_subtree_depth"], cfg["tally_subtree_depth"]))
    blk = golden["merge_registration_state_success"]["registration_block"]
    for i, q in enumerate(golden["participants"]):
        assert poll.register_participant((H(q["x"]), H(q["y"])), blk) == i + 1
    poll.merge_registrations()
    assert poll.registrations.root == H(golden["merge_registration_state_success"]["registrations_root"])
    assert poll.commitment.process == (0, H(golden["merge_registration_state_success"]["process_commitment"]))
    assert poll.registrations.depth == golden["process_messages_public_signals"]["registrations_depth"]
    p = golden["participant"]
    assert poll.consume_interaction((H(p["shared_pk"]["x"]), H(p["shared_pk"]["y"])), [H(x) for x in p["message"]]) == 1
    poll.merge_interactions()
    assert poll.interactions.root == H(golden["merge_interaction_state_success"]["interactions_root"])
    assert (poll.commitment.expected_process, poll.commitment.expected_tally) == \
        (golden["merge_interaction_state_success"]["expected_process"], golden["merge_interaction_state_success"]["expected_tally"])


def test_leaf_hashing_batches_vs_oracle(ib):
    n = 20000
    pk = random_fr_bytes(2 * n, seed=61, canonical=False).reshape(n, 64)
    data = random_fr_bytes(10 * n, seed=62, canonical=False).reshape(n, 320)
    ts = np.random.default_rng(63).integers(0, 2 ** 63, size=n, dtype=np.uint64)
    ts[:4] = [0, 1, 2 ** 32 - 1, 2 ** 64 - 1]
    assert (ib.interaction_leaves(pk, data) == c_oracle.interaction_leaves(pk, data)).all()
    assert (ib.registration_leaves(pk, ts) == c_oracle.registration_leaves(pk, ts)).all()
    # ragged tail of the chunked pipeline
    m = (1 << 17) + 77
    pk = random_fr_bytes(2 * m, seed=64).reshape(m, 64)
    ts = np.arange(m, dtype=np.uint64)
    got = ib.registration_leaves(pk, ts)
    idx = np.r_[0:50, (1 << 17) - 25:(1 << 17) + 77]
    assert (got[idx] == c_oracle.registration_leaves(pk[idx], ts[idx])).all()


# ---- replay: raw rows -> leaves -> merged tree on the device (provider.rs:218-327) ---------------------
@pytest.mark.parametrize("n", [0, 1, 4, 26, 700, 20000, (1 << 17) + 77])
def test_replay_interactions_equals_leaf_hash_then_insert_merge(ib, n):
    depth, psd = 8, 2
    pk = random_fr_bytes(2 * max(n, 1), seed=71 + n, canonical=False).reshape(-1, 64)[:n]
    data = random_fr_bytes(10 * max(n, 1), seed=72 + n, canonical=False).reshape(-1, 320)[:n]
    t, ep, et, leaves, kept = ib.replay_interactions(depth, pk, data, 1000, psd, 1, want_leaves=True, retain=True)
    exp_leaves = c_oracle.interaction_leaves(pk, data) if n else np.empty((0, 32), dtype=np.uint8)
    assert (leaves == exp_leaves).all()
    rc, root, d, count = c_oracle.tree_insert_merge(5, depth, False, True, exp_leaves)
    assert rc == 0 and (t.root, t.depth, t.count, t.hashes) == (root, d, n, [])
    assert (ep, et) == (-(-n // 25), 1 + 1000 // 2)
    if n == 0:
        assert kept is None and t.root is None
        return
    # the retained levels: every batch subroot and its path to the root, all batches in one call
    assert kept.root == root
    n_batches = -(-n // 5 ** psd)
    sub = kept.level_nodes(psd)
    assert sub.shape[0] == n_batches
    zero = np.frombuffer(ib.get_merkle_zeroes(5)[0], dtype=np.uint8)
    for b in {0, n_batches - 1, n_batches // 2}:
        chunk = exp_leaves[b * 25:(b + 1) * 25]
        assert sub[b].tobytes() == c_oracle.dense_tree_root(5, psd, chunk)
    idx = np.unique(np.r_[0, n_batches - 1, np.random.default_rng(n).integers(0, n_batches, size=50)]).astype(np.uint64)
    paths = kept.node_paths(psd, idx)
    roots = ib.merkle_roots_from_paths(5, depth - psd, idx, sub[idx.astype(np.int64)], paths)
    assert all(roots[k].tobytes() == root for k in range(len(idx)))
    # a batch of all-zero leaves to the right of the stored nodes reads as zeroes[psd]
    if n_batches < 5 ** (depth - psd):
        assert kept.level_nodes(psd, n_batches, 1)[0].tobytes() == ib.get_merkle_zeroes(5)[psd]
    kept.close()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 700, (1 << 19) + 5])
def test_replay_registrations_equals_leaf_hash_then_insert_merge(ib, n):
    depth = 21
    pk = random_fr_bytes(2 * max(n, 1), seed=81 + n, canonical=False).reshape(-1, 64)[:n]
    ts = np.random.default_rng(82 + n).integers(0, 2 ** 63, size=n, dtype=np.uint64)
    t, commitment, leaves, kept = ib.replay_registrations(depth, pk, ts, want_leaves=True, retain=True)
    exp_leaves = c_oracle.registration_leaves(pk, ts) if n else np.empty((0, 32), dtype=np.uint8)
    assert (leaves == exp_leaves).all()
    rc, root, d, count = c_oracle.tree_insert_merge(2, depth, True, False, exp_leaves)
    assert rc == 0 and (t.root, t.depth, t.count, t.hashes) == (root, d, n, [])
    assert commitment == c_oracle.hash_one([root, ib.empty_ballot_roots()[1], bytes(32)])
    assert kept.root == root
    if n:
        idx = np.unique(np.r_[0, 1, n, np.random.default_rng(n).integers(0, n + 1, size=30)]).astype(np.uint64)
        logical = np.concatenate([np.frombuffer(ib.get_merkle_zeroes(2)[0], dtype=np.uint8).reshape(1, 32), exp_leaves])
        roots = ib.merkle_roots_from_paths(2, kept.depth, idx, logical[idx.astype(np.int64)], kept.paths(idx))
        assert all(roots[k].tobytes() == root for k in range(len(idx)))
    kept.close()


def test_replay_capacity_errors(ib):
    pk = random_fr_bytes(2 * 26, seed=91).reshape(-1, 64)
    data = random_fr_bytes(10 * 26, seed=92).reshape(-1, 320)
    with pytest.raises(ib.MerkleTreeError) as e:
        ib.replay_interactions(2, pk, data, 0, 1, 1)             # 26 messages, capacity 25
    assert e.value.code == 1
    t, _, _, _, _ = ib.replay_interactions(2, pk[:25], data[:25], 0, 1, 1)     # exactly full: completed by insert
    rc, root, d, _ = c_oracle.tree_insert_merge(5, 2, False, True, c_oracle.interaction_leaves(pk[:25], data[:25]))
    assert rc == 2 and t.root == root and t.depth == d
    with pytest.raises(ib.MerkleTreeError) as e:
        ib.replay_registrations(2, pk[:4], np.arange(4, dtype=np.uint64))      # blank + 4 > 4
    assert e.value.code == 1


# ---- Merkle paths and verify_outcome (provider.rs:76-139, 396-436) ---------------------------------
@pytest.mark.parametrize("sid", [1, 2])
def test_verify_outcome_scenarios(ib, golden, sid):
    o = golden["scenario_%d_outcome" % sid]
    outcome = {k: H(o[k]) for k in ("total_spent", "total_spent_salt", "tally_result_salt",
                                    "new_results_commitment", "spent_votes_hash")}
    outcome["tally_results"] = o["tally_results"]
    outcome["tally_result_proofs"] = [[[H(x) for x in lvl] for lvl in opt] for opt in o["tally_result_proofs"]]
    depth = golden["poll_config"]["vote_option_tree_depth"]
    commitment = H(o["final_tally_commitment"])
    assert ib.verify_outcome(depth, 25, commitment, outcome) == o["expected_outcome_index"]
    assert ib.verify_outcome(depth, 25, commitment, dict(outcome, spent_votes_hash=bytes(32))) is None
    for i in (0, 5, 23, 24):
        got = ib.compute_merkle_root_from_path(depth, i, be(o["tally_results"][i]), outcome["tally_result_proofs"][i])
        assert got == O.compute_merkle_root_from_path(depth, i, be(o["tally_results"][i]), outcome["tally_result_proofs"][i])


@pytest.mark.parametrize("arity,depth,blank,n", [(5, 4, False, 611), (2, 10, True, 1000), (2, 9, False, 512), (5, 3, False, 1)])
def test_retained_tree_paths(ib, arity, depth, blank, n):
    leaves = random_fr_bytes(n, seed=arity * 7 + depth)
    tree = ib.RetainedTree(arity, depth, leaves, prepend_blank_leaf=blank)
    logical = ([ib.get_merkle_zeroes(arity)[0]] if blank else []) + [leaves[i].tobytes() for i in range(n)]
    levels = O.dense_tree_levels(logical, arity, depth) if n <= 1000 else None
    assert tree.root == levels[-1][0]
    rc, root, _, _ = c_oracle.tree_insert_merge(arity, depth, blank, True, leaves)
    assert tree.root == root
    rng = np.random.default_rng(5)
    idx = np.unique(np.r_[0, len(logical) - 1, rng.integers(0, len(logical), size=40)]).astype(np.uint64)
    paths = tree.paths(idx)
    for k, i in enumerate(idx):
        exp = O.merkle_path(levels, arity, int(i))
        got = [[paths[k, l, s].tobytes() for s in range(arity - 1)] for l in range(depth)]
        assert got == exp, (arity, int(i))
    # and the batched verifier closes the loop on the device
    roots = ib.merkle_roots_from_paths(arity, depth, idx, np.stack([np.frombuffer(logical[int(i)], dtype=np.uint8) for i in idx]), paths)
    assert all(roots[k].tobytes() == tree.root for k in range(len(idx)))
    tree.close()


# ---- extremes --------------------------------------------------------------------------------------
def test_maximum_depths_and_empty_inputs(ib):
    """Deepest trees the pallet admits (lib.rs:390-399: 2^depth and 5^depth must
    fit u32 -> registration depth <= 31, interaction depth <= 13) plus the 33-level
    zero table's limit, with a handful of leaves: all padding levels are walked."""
    lv = random_fr_bytes(3, seed=404)
    for arity, full_depth, blank in ((2, 31, True), (2, 32, False), (5, 13, False), (5, 32, False)):
        rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, True, lv)
        t = ib.PollStateTree.new(arity, full_depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None)
        t.extend(lv).merge(True)
        assert rc == 0 and t.root == root and t.depth == depth
    with pytest.raises(ValueError):
        ib.PollStateTree.new(2, 33).extend(lv).merge(True)          # beyond the zero table
    # empty batch, empty tree
    h = ib.Poseidon.new_circom(2)
    assert h.hash_batch(b"").shape == (0, 32)
    t = ib.new_interaction_tree(3)
    t.merge(True)
    assert t.root is None and t.count == 0 and t.hashes == []       # merge on an empty frontier (state.rs:240-248)
    o = O.new_interaction_tree(3)
    o.merge(True)
    assert o.root is None
    # merge_registrations on an empty registration tree: the blank leaf alone is the root
    t, c = ib.merge_registrations(ib.new_registration_tree(10))
    o, oc = O.merge_registrations(O.new_registration_tree(10))
    assert t.root == o.root and c == oc
