#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the Poseidon / poll-tree path.

Run in the build container (where /root/reference is mounted, read-only):

    python tests/golden/extract_reference_vectors.py

It reads ONLY test expectations and fixtures (numbers inside the reference's
test files and its zero tables) and writes them to
tests/golden/reference_vectors.json.  The JSON travels to the GPU box; the
reference checkout does not.  No reference code is copied.

Sources (relative to /root/reference):
  pallet/src/tests/poseidon.rs:17-251     hasher KATs
  pallet/src/poll/zeroes.rs:1-79          zero tables + empty ballot roots
  pallet/src/tests/extrinsics.rs:481-647  tree roots / commitments / signals
  pallet/src/tests/data.rs:15-275         keys, poll config, messages
"""
import json
import os
import re
import sys

REF = os.environ.get("INFIMUM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")

ARR32 = re.compile(r"\[\s*((?:\d+\s*,\s*){31}\d+)\s*,?\s*\]")


def arrays32(text):
    """All 32-element u8 array literals in `text`, as hex strings, in order."""
    out = []
    for m in ARR32.finditer(text):
        vals = [int(x) for x in m.group(1).split(",")]
        assert all(0 <= v < 256 for v in vals)
        out.append(bytes(vals).hex())
    return out


def fn_body(src, name):
    """Source text of `fn name(...) {...}` (brace matched)."""
    m = re.search(r"fn\s+" + re.escape(name) + r"\s*\(", src)
    assert m, name
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return src[i:j + 1]
        j += 1


def const_body(src, name):
    m = re.search(r"const\s+" + re.escape(name) + r"\b[^=]*=", src)
    assert m, name
    j = src.index(";", m.end())
    return src[m.end():j]


def main():
    rd = lambda p: open(os.path.join(REF, p)).read()
    tp = rd("pallet/src/tests/poseidon.rs")
    zr = rd("pallet/src/poll/zeroes.rs")
    ex = rd("pallet/src/tests/extrinsics.rs")
    da = rd("pallet/src/tests/data.rs")

    v = {"_source": "rhysbalevicius/infimum reference tests; see extract_reference_vectors.py"}

    # ---- hasher KATs ---------------------------------------------------------
    v["fr_one"] = {"inputs_int": [1, 1], "expected_be": arrays32(fn_body(tp, "fr_one"))[0]}
    a = arrays32(fn_body(tp, "bytes_ones_twos"))
    assert a[0] == a[1]
    v["bytes_ones_twos"] = {"inputs_be": ["01" * 32, "02" * 32], "expected_be": a[0], "expected_le": a[2]}
    v["with_domain_tag"] = {"inputs_be": ["01" * 32, "02" * 32],
                            "expected_tag_zero_be": arrays32(fn_body(tp, "with_domain_tag"))[0]}
    v["fr_one_two"] = {"inputs_int": [1, 2], "expected_le": arrays32(fn_body(tp, "fr_one_two"))[0]}
    a = arrays32(fn_body(tp, "random_input"))
    v["random_input"] = {"inputs_be": a[0:2], "expected_le": a[2]}
    a = arrays32(const_body(tp, "CIRCOMLIBJS_TEST_CASES"))
    assert len(a) == 12
    v["circomlibjs_ones"] = a          # entry n-1 = poseidon([1]*n), 32-byte BE

    # ---- zero tables ---------------------------------------------------------
    b = arrays32(const_body(zr, "BINARY_ZEROES"))
    q = arrays32(const_body(zr, "QUINARY_ZEROES"))
    e = arrays32(const_body(zr, "EMPTY_BALLOT_ROOTS"))
    assert (len(b), len(q), len(e)) == (33, 33, 5)
    v["binary_zeroes"], v["quinary_zeroes"], v["empty_ballot_roots"] = b, q, e

    # ---- fixtures (data.rs) --------------------------------------------------
    a = arrays32(fn_body(da, "get_coordinator_data"))
    v["coordinator_pk"] = {"x": a[0], "y": a[1]}
    cfg = fn_body(da, "get_poll_config")
    g = lambda k: int(re.search(k + r"\s*=\s*(\d+)", cfg).group(1))
    v["poll_config"] = {k: g(k) for k in ("signup_period", "voting_period", "registration_depth",
                                         "interaction_depth", "process_subtree_depth",
                                         "tally_subtree_depth", "vote_option_tree_depth")}
    a = arrays32(fn_body(da, "get_participant"))
    assert len(a) == 14
    v["participant"] = {"pk": {"x": a[0], "y": a[1]}, "shared_pk": {"x": a[2], "y": a[3]},
                        "message": a[4:14]}
    a = arrays32(fn_body(da, "get_participants"))
    assert len(a) == 6
    v["participants"] = [{"x": a[2 * i], "y": a[2 * i + 1]} for i in range(3)]
    # scenario interactions (pk + 10 x 32-byte message words each); used as
    # extra realistic leaves — their roots are pinned only through Groth16
    # proofs in the reference, which is out of scope here.
    for sid in (1, 2):
        body = fn_body(da, "poll_scenario_%d" % sid)
        inter = body[body.index("interactions:"):body.index("proof_batches:")]
        a = arrays32(inter)
        assert len(a) % 12 == 0, len(a)
        v["scenario_%d_interactions" % sid] = [
            {"pk": {"x": a[12 * i], "y": a[12 * i + 1]}, "message": a[12 * i + 2:12 * i + 12]}
            for i in range(len(a) // 12)]

    # ---- outcome fixtures (data.rs:223-275): pin compute_merkle_root_from_path and
    # the tally-commitment hash chain of verify_outcome (provider.rs:76-139, 396-436).
    # The last commitment of proof_batches is what commit_outcome leaves in
    # commitment.tally.1 before verify_outcome runs (lib.rs:567-640).
    for sid in (1, 2):
        body = fn_body(da, "poll_scenario_%d" % sid)
        batches = arrays32(body[body.index("proof_batches:"):body.index("expected:")])
        out = body[body.index("outcome:"):]
        fld = lambda k: arrays32(re.search(k + r":\s*\[[^\]]*\]", out).group(0))[0]
        proofs = arrays32(out[out.index("tally_result_proofs:"):])
        results = [int(x) for x in re.search(r"tally_results:\s*vec::Vec::from\(\[([^\]]*)\]\)", out).group(1).split(",") if x.strip()]
        depth = v["poll_config"]["vote_option_tree_depth"]
        assert len(proofs) == len(results) * depth * 4
        v["scenario_%d_outcome" % sid] = {
            "expected_outcome_index": int(re.search(r"expected:\s*Some\((\d+)\)", body).group(1)),
            "final_tally_commitment": batches[-1], "batch_commitments": batches,
            "tally_results": results,
            "tally_result_proofs": [[proofs[(o * depth + l) * 4:(o * depth + l) * 4 + 4] for l in range(depth)]
                                    for o in range(len(results))],
            "total_spent": fld("total_spent"), "total_spent_salt": fld("total_spent_salt"),
            "tally_result_salt": fld("tally_result_salt"), "new_results_commitment": fld("new_results_commitment"),
            "spent_votes_hash": fld("spent_votes_hash")}

    # ---- tree pins (extrinsics.rs) -------------------------------------------
    a = arrays32(fn_body(ex, "merge_registration_state_success"))
    v["merge_registration_state_success"] = {
        "registration_block": 2,            # run_to_block(2) before registering
        "registrations_root": a[0], "process_commitment": a[1]}
    body = fn_body(ex, "merge_interaction_state_success")
    a = arrays32(body)
    v["merge_interaction_state_success"] = {
        "interactions_root": a[0],
        "expected_process": int(re.search(r"expected_process,\s*(\d+)", body).group(1)),
        "expected_tally": int(re.search(r"expected_tally,\s*(\d+)", body).group(1))}
    body = fn_body(ex, "process_messages_public_signals")
    a = arrays32(body)
    dec = re.findall(r'"(\d{20,})"', body)
    v["process_messages_public_signals"] = {
        "registrations_count_plus_one": 4,
        "registrations_depth": int(re.search(r"registrations\.depth,\s*(\d+)", body).group(1)),
        "interactions_root": a[0], "process_commitment": a[1],
        "interactions_root_decimal": dec[0],
        "coord_pub_key_hash_decimal": re.search(r'coord_pub_key_hash,\s*"(\d+)"', body).group(1),
        # the nine public inputs of the first process-messages proof, as listed in the test's comment
        # (extrinsics.rs:621-633): what prepare_public_inputs (provider.rs:141-215) must produce
        "expected_public_inputs_decimal": re.findall(r'//\s+"(\d+)"', body),
        "created_at_block": 1}

    with open(OUT, "w") as f:
        json.dump(v, f, indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
