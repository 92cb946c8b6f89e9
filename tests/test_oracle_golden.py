"""Pins the Python oracle (oracle/poseidon_ref.py) against every vector the
reference's own tests hold for the hot path.  Test names follow
the reference's (pallet/src/tests/poseidon.rs, pallet/src/tests/extrinsics.rs)."""
import pytest

from oracle import poseidon_ref as O

H = bytes.fromhex


def be(i):
    return i.to_bytes(32, "big")


def test_fr_one(golden):
    # poseidon.rs:17-40 — 1, 0x0001 and 0x000001 are the same field element
    g = golden["fr_one"]
    for raw in (b"\x01", b"\x00\x01", b"\x00\x00\x01"):
        x = int.from_bytes(raw, "big") % O.P
        assert be(O.Poseidon.new_circom(2).hash([x, x])) == H(g["expected_be"])


def test_bytes_ones_twos(golden):
    g = golden["bytes_ones_twos"]
    ins = [H(x) for x in g["inputs_be"]]
    h = O.Poseidon.new_circom(2)
    assert be(h.hash([int.from_bytes(b, "big") % O.P for b in ins])) == H(g["expected_be"])
    assert h.hash_bytes_be(ins) == H(g["expected_be"])
    assert h.hash_bytes_le(ins) == H(g["expected_le"])


def test_with_domain_tag(golden):
    g = golden["with_domain_tag"]
    ins = [int.from_bytes(H(x), "big") % O.P for x in g["inputs_be"]]
    assert be(O.Poseidon.with_domain_tag_circom(2, 0).hash(ins)) == H(g["expected_tag_zero_be"])
    assert be(O.Poseidon.with_domain_tag_circom(2, 1).hash(ins)) != H(g["expected_tag_zero_be"])


def test_fr_one_two(golden):
    g = golden["fr_one_two"]
    assert O.Poseidon.new_circom(2).hash([1, 2]).to_bytes(32, "little") == H(g["expected_le"])


def test_random_input(golden):
    # inputs are >= p: exercises from_be_bytes_mod_order
    g = golden["random_input"]
    ins = [int.from_bytes(H(x), "big") for x in g["inputs_be"]]
    assert all(x >= O.P for x in ins)
    assert O.Poseidon.new_circom(2).hash([x % O.P for x in ins]).to_bytes(32, "little") == H(g["expected_le"])


def test_empty_input():
    # poseidon.rs tests :134-175
    for n in range(1, 12):
        h = O.Poseidon.new_circom(n)
        for ins in ([b""] * n, [b"\x01" * 32] * (n - 1) + [b""]):
            for fn in (h.hash_bytes_be, h.hash_bytes_le):
                with pytest.raises(O.PoseidonError) as e:
                    fn(ins)
                assert e.value.kind == "EmptyInput"


def test_input_length_and_width_errors():
    h = O.Poseidon.new_circom(2)
    with pytest.raises(O.PoseidonError) as e:
        h.hash_bytes_be([b"\x01" * 33, b"\x01" * 32])
    assert e.value.kind == "InvalidInputLength"
    with pytest.raises(O.PoseidonError) as e:
        h.hash_bytes_be([b"\x01" * 31, b"\x01" * 32])
    assert e.value.kind == "InvalidInputLength"
    with pytest.raises(O.PoseidonError) as e:
        h.hash([1])
    assert e.value.kind == "InvalidNumberOfInputs"
    with pytest.raises(O.PoseidonError) as e:
        O.Poseidon.new_circom(13)
    assert e.value.kind == "InvalidWidthCircom"


def test_circomlibjs_compat_1_to_12_inputs(golden):
    one, two = be(1), be(2)
    for n in range(1, 13):
        h = O.Poseidon.new_circom(n)
        assert h.hash_bytes_be([one] * n) == H(golden["circomlibjs_ones"][n - 1])
        assert h.hash_bytes_be([two] * n) != H(golden["circomlibjs_ones"][n - 1])


@pytest.mark.parametrize("arity,key", [(2, "binary_zeroes"), (5, "quinary_zeroes")])
def test_zero_chains(golden, arity, key):
    # zeroes.rs:1-71 — 32 self-checking links per table
    table = [H(x) for x in golden[key]]
    assert list(O.merkle_zeroes(arity)) == table
    for l in range(32):
        assert O.hash_be([table[l]] * arity) == table[l + 1]
    assert O.merkle_zeroes(7) == O.merkle_zeroes(5)      # zeroes.rs:83-84


def test_empty_ballot_roots(golden):
    assert [be(x) for x in O.EMPTY_BALLOT_ROOTS] == [H(x) for x in golden["empty_ballot_roots"]]


def _registered_tree(golden):
    cfg = golden["poll_config"]
    t = O.new_registration_tree(cfg["registration_depth"])
    blk = golden["merge_registration_state_success"]["registration_block"]
    for pk in golden["participants"]:
        t.insert(O.registration_leaf(H(pk["x"]), H(pk["y"]), blk))
    return t


def test_merge_registration_state_success(golden):
    g = golden["merge_registration_state_success"]
    t, commitment = O.merge_registrations(_registered_tree(golden))
    assert t.root == H(g["registrations_root"])
    assert commitment == H(g["process_commitment"])
    assert t.hashes == []


def test_merge_interaction_state_success(golden):
    g = golden["merge_interaction_state_success"]
    cfg = golden["poll_config"]
    reg = _registered_tree(golden)
    p = golden["participant"]
    t = O.new_interaction_tree(cfg["interaction_depth"])
    t.insert(O.interaction_leaf(H(p["shared_pk"]["x"]), H(p["shared_pk"]["y"]), [H(x) for x in p["message"]]))
    t, ep, et = O.merge_interactions(t, reg.count, cfg["process_subtree_depth"], cfg["tally_subtree_depth"])
    assert t.root == H(g["interactions_root"])
    assert (ep, et) == (g["expected_process"], g["expected_tally"])


def test_process_messages_public_signals(golden):
    g = golden["process_messages_public_signals"]
    reg, commitment = O.merge_registrations(_registered_tree(golden))
    assert reg.count + 1 == g["registrations_count_plus_one"]
    assert reg.depth == g["registrations_depth"]
    assert commitment == H(g["process_commitment"])
    assert int(g["interactions_root_decimal"]) == int(g["interactions_root"], 16)
    pk = golden["coordinator_pk"]
    hsh = O.Poseidon.new_circom(2).hash([int(pk["x"], 16) % O.P, int(pk["y"], 16) % O.P])
    assert str(hsh) == g["coord_pub_key_hash_decimal"]
    # the nine public inputs listed at extrinsics.rs:621-633 (prepare_public_inputs, provider.rs:141-215)
    msg = golden["participant"]
    it = O.new_interaction_tree(golden["poll_config"]["interaction_depth"])
    it.insert(O.interaction_leaf(H(msg["shared_pk"]["x"]), H(msg["shared_pk"]["y"]), [H(x) for x in msg["message"]]))
    cfg = golden["poll_config"]
    it, _, _ = O.merge_interactions(it, reg.count, cfg["process_subtree_depth"], cfg["tally_subtree_depth"])
    exp = [int(x) for x in g["expected_public_inputs_decimal"]]
    end = g["created_at_block"] + cfg["signup_period"] + cfg["voting_period"]                 # provider.rs:355-358
    kind, inputs, nxt = O.prepare_public_inputs(reg, it, (0, commitment), (0, bytes(32)), cfg["process_subtree_depth"],
                                                cfg["tally_subtree_depth"], end, (H(pk["x"]), H(pk["y"])),
                                                exp[8].to_bytes(32, "big"))
    assert kind == "process" and inputs == exp and nxt[0] == 1
    # after the only process proof the tally branch is taken (index_offset 5 > batch index 0)
    kind, inputs, nxt = O.prepare_public_inputs(reg, it, nxt, (0, bytes(32)), cfg["process_subtree_depth"],
                                                cfg["tally_subtree_depth"], end, (H(pk["x"]), H(pk["y"])), bytes(32))
    assert kind == "tally" and inputs == [exp[8], 0, 0, 0, 4]


def test_participant_limit_reached_quirk():
    # extrinsics.rs:301-316: depth 2 tree, blank + 3 leaves completes in insert,
    # so merge() then reports TreeAlreadyMerged and a 5th insert TreeAlreadyFull.
    t = O.new_registration_tree(2)
    for i in range(3):
        t.insert(be(100 + i))
    assert t.root is not None and t.hashes == []
    with pytest.raises(O.MerkleTreeError) as e:
        t.merge(False)
    assert e.value.code == 2
    with pytest.raises(O.MerkleTreeError) as e:
        t.insert(be(7))
    assert e.value.code == 1


@pytest.mark.parametrize("arity,full_depth,to_depth,blank", [(2, 7, False, True), (5, 3, True, False),
                                                            (2, 6, True, False), (5, 3, False, False)])
def test_batch_equivalence(arity, full_depth, to_depth, blank):
    """new + insert*N + merge == dense zero-padded tree."""
    cap = arity ** full_depth
    ns = sorted(set(list(range(0, 40)) + [63, 64, 65, 100, 124, 125, 126, 127, 128]))
    for n in ns:
        total = n + (1 if blank else 0)
        if total > cap or (total == cap):
            continue
        leaves = [be((i * 7919 + 13) % O.P) for i in range(n)]
        t = O.PollStateTree.new(arity, full_depth, (0, O.merkle_zeroes(arity)[0]) if blank else None)
        for lf in leaves:
            t.insert(lf)
        depth_after_insert = t.depth
        t.merge(to_depth)
        root, insert_depth, root_depth, count = O.batch_merge(
            arity, full_depth, leaves, prepend_blank_leaf=blank, to_depth=to_depth)
        assert t.root == root, (arity, n)
        assert depth_after_insert == insert_depth, (arity, n)
        assert t.count == count


def _outcome(golden, sid):
    o = golden["scenario_%d_outcome" % sid]
    out = {k: H(o[k]) for k in ("total_spent", "total_spent_salt", "tally_result_salt", "new_results_commitment",
                               "spent_votes_hash")}
    out["tally_results"] = o["tally_results"]
    out["tally_result_proofs"] = [[[H(x) for x in lvl] for lvl in opt] for opt in o["tally_result_proofs"]]
    return out, H(o["final_tally_commitment"]), o["expected_outcome_index"]


@pytest.mark.parametrize("sid", [1, 2])
def test_verify_outcome_scenarios(golden, sid):
    """data.rs:241-275: the circuits' own tally proofs and salts must hash to the
    final tally commitment through compute_merkle_root_from_path (hash5 paths)
    and the hash2 chain of verify_outcome; expected winners 5 and 23."""
    outcome, commitment, expected = _outcome(golden, sid)
    depth = golden["poll_config"]["vote_option_tree_depth"]
    assert O.verify_outcome(depth, 25, commitment, outcome) == expected
    bad = dict(outcome, tally_result_salt=bytes(32))
    assert O.verify_outcome(depth, 25, commitment, bad) is None


def test_merkle_path_roundtrip():
    leaves = [be(1000 + i) for i in range(38)]
    for arity, depth in ((5, 3), (2, 6)):
        levels = O.dense_tree_levels(leaves, arity, depth)
        root = levels[-1][0]
        assert root == O.dense_tree_root(leaves, arity, depth)
        for idx in (0, 1, 4, 5, 24, 37):
            path = O.merkle_path(levels, arity, idx)
            assert O.compute_merkle_root_from_path(depth, idx, leaves[idx], path, arity) == root


def test_custom_parameters_restatement_agrees_with_the_circom_hasher():
    """Poseidon::new(params) (poseidon.rs:105-108) with the parameters.rs tables must be
    new_circom (poseidon.rs:304-326): pins the generic-parameter oracle to the same KATs."""
    for t in (2, 3, 6, 13):
        ark, mds, rf, rp = O.poseidon_parameters(t)
        ins = [(7 * i + 1) % O.P for i in range(t - 1)]
        assert O.poseidon_hash_with_params(ark, mds, rf, rp, t, 5, ins) == O.poseidon_permute_hash(ins)
        assert O.poseidon_hash_with_params(ark, mds, rf, rp, t, 5, ins, 9) == O.poseidon_permute_hash(ins, 9)
