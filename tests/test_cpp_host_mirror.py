"""Builds and runs the C++ parity tests (tests/cpp/parity_tests.cpp) that drive
the C++ host mirror include/infimum_b200.hpp over the C ABI.  The compile-only
part runs on the CPU box; the run needs the GPU."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BIN = os.path.join(CPP, "_build", "parity_tests")


def _arr(name, hexstr):
    b = bytes.fromhex(hexstr)
    return "static const uint8_t %s[32] = {%s};" % (name, ", ".join(str(x) for x in b))


def _arr2(name, rows):
    return "static const uint8_t %s[%d][32] = {%s};" % (
        name, len(rows), ", ".join("{%s}" % ", ".join(str(x) for x in bytes.fromhex(r)) for r in rows))


def _write_golden_header(path):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    L = ["// generated from tests/golden/reference_vectors.json by tests/test_cpp_host_mirror.py", "#pragma once",
         "#include <stdint.h>"]
    L.append(_arr("G_FR_ONE_EXPECTED_BE", g["fr_one"]["expected_be"]))
    L.append(_arr("G_BYTES_ONES_TWOS_BE", g["bytes_ones_twos"]["expected_be"]))
    L.append(_arr("G_BYTES_ONES_TWOS_LE", g["bytes_ones_twos"]["expected_le"]))
    L.append(_arr("G_WITH_DOMAIN_TAG_ZERO_BE", g["with_domain_tag"]["expected_tag_zero_be"]))
    L.append(_arr("G_FR_ONE_TWO_LE", g["fr_one_two"]["expected_le"]))
    L.append(_arr("G_RANDOM_INPUT_1", g["random_input"]["inputs_be"][0]))
    L.append(_arr("G_RANDOM_INPUT_2", g["random_input"]["inputs_be"][1]))
    L.append(_arr("G_RANDOM_INPUT_LE", g["random_input"]["expected_le"]))
    L.append(_arr2("G_CIRCOMLIBJS", g["circomlibjs_ones"]))
    L.append(_arr2("G_BINARY_ZEROES", g["binary_zeroes"]))
    L.append(_arr2("G_QUINARY_ZEROES", g["quinary_zeroes"]))
    L.append(_arr2("G_EMPTY_BALLOT_ROOTS", g["empty_ballot_roots"]))
    L.append("static const uint8_t G_PARTICIPANTS[3][2][32] = {%s};" % ", ".join(
        "{{%s}, {%s}}" % (", ".join(str(x) for x in bytes.fromhex(p["x"])), ", ".join(str(x) for x in bytes.fromhex(p["y"])))
        for p in g["participants"]))
    L.append(_arr2("G_COORDINATOR_PK", [g["coordinator_pk"]["x"], g["coordinator_pk"]["y"]]))
    L.append(_arr2("G_SHARED_PK", [g["participant"]["shared_pk"]["x"], g["participant"]["shared_pk"]["y"]]))
    L.append(_arr2("G_MESSAGE", g["participant"]["message"]))
    cfg = g["poll_config"]
    L.append("static const int G_REGISTRATION_DEPTH = %d, G_INTERACTION_DEPTH = %d, G_PROCESS_SUBTREE_DEPTH = %d, "
             "G_TALLY_SUBTREE_DEPTH = %d;" % (cfg["registration_depth"], cfg["interaction_depth"],
                                              cfg["process_subtree_depth"], cfg["tally_subtree_depth"]))
    L.append("static const uint64_t G_REGISTRATION_BLOCK = %d;" % g["merge_registration_state_success"]["registration_block"])
    L.append(_arr("G_REGISTRATIONS_ROOT", g["merge_registration_state_success"]["registrations_root"]))
    L.append(_arr("G_PROCESS_COMMITMENT", g["merge_registration_state_success"]["process_commitment"]))
    L.append(_arr("G_INTERACTIONS_ROOT", g["merge_interaction_state_success"]["interactions_root"]))
    L.append("static const uint32_t G_EXPECTED_PROCESS = %d, G_EXPECTED_TALLY = %d, G_REGISTRATIONS_DEPTH = %d;" % (
        g["merge_interaction_state_success"]["expected_process"], g["merge_interaction_state_success"]["expected_tally"],
        g["process_messages_public_signals"]["registrations_depth"]))
    L.append('static const char* G_COORD_PUB_KEY_HASH_DECIMAL = "%s";' % g["process_messages_public_signals"]["coord_pub_key_hash_decimal"])
    pm = g["process_messages_public_signals"]
    L.append(_arr2("G_EXPECTED_PUBLIC_INPUTS", ["%064x" % int(x) for x in pm["expected_public_inputs_decimal"]]))
    L.append("static const uint64_t G_CREATED_AT = %d, G_SIGNUP_PERIOD = %d, G_VOTING_PERIOD = %d;" % (
        pm["created_at_block"], cfg["signup_period"], cfg["voting_period"]))
    sc = [g["scenario_1_outcome"], g["scenario_2_outcome"]]
    L.append("static const uint32_t G_OUTCOME_TALLY_RESULTS[2][25] = {%s};" % ", ".join(
        "{%s}" % ", ".join(str(x) for x in o["tally_results"]) for o in sc))
    L.append("static const uint32_t G_OUTCOME_EXPECTED[2] = {%d, %d};" % tuple(o["expected_outcome_index"] for o in sc))
    bb = lambda h: "{%s}" % ", ".join(str(x) for x in bytes.fromhex(h))
    L.append("static const uint8_t G_OUTCOME_FIELDS[2][6][32] = {%s};" % ", ".join(
        "{%s}" % ", ".join(bb(o[k]) for k in ("total_spent", "total_spent_salt", "tally_result_salt",
                                              "new_results_commitment", "spent_votes_hash", "final_tally_commitment"))
        for o in sc))
    L.append("static const uint8_t G_OUTCOME_PROOFS[2][25][2][4][32] = {%s};" % ", ".join(
        "{%s}" % ", ".join("{%s}" % ", ".join("{%s}" % ", ".join(bb(x) for x in lvl) for lvl in opt)
                           for opt in o["tally_result_proofs"]) for o in sc))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write("\n".join(L) + "\n")


def build():
    import __graft_entry__ as entry
    entry.build()
    out = os.path.join(CPP, "_build")
    _write_golden_header(os.path.join(out, "golden_vectors.h"))
    cmd = ["g++", "-O1", "-std=c++17", "-I", out, "-I", os.path.join(ROOT, "include"),
           os.path.join(CPP, "parity_tests.cpp"), "-o", BIN,
           "-L", os.path.join(ROOT, "infimum_b200"), "-linfimum_b200", "-L", os.path.join(ROOT, "oracle"), "-loracle",
           "-Wl,-rpath," + os.path.join(ROOT, "infimum_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return BIN


def build_cpu():
    """The same test source linked against tests/cpp/fake_infimum.c (the oracle behind the C ABI's
    signatures; test infrastructure) ahead of the real library: runs without a GPU."""
    build()
    out = os.path.join(CPP, "_build")
    exe = os.path.join(out, "parity_tests_cpu")
    inc = ["-I", out, "-I", os.path.join(ROOT, "include")]
    r = subprocess.run(["gcc", "-O2", "-std=c11", "-Wno-unused-variable", "-Wno-unused-const-variable"] + inc +
                       ["-c", os.path.join(CPP, "fake_infimum.c"), "-o", os.path.join(out, "fake_infimum.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cmd = ["g++", "-O1", "-std=c++17"] + inc + [os.path.join(CPP, "parity_tests.cpp"), os.path.join(out, "fake_infimum.o"),
           "-o", exe, "-L", os.path.join(ROOT, "infimum_b200"), "-linfimum_b200", "-L", os.path.join(ROOT, "oracle"),
           "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "infimum_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_parity_tests_pass_on_cpu_over_the_fake_library():
    exe = build_cpu()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout


def test_cpp_host_mirror_compiles_and_links():
    assert os.path.exists(build())


@pytest.mark.gpu
def test_cpp_parity_tests_pass_on_gpu():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout
