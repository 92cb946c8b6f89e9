"""The Python host mirror on the CPU box: the reference-named tests of
tests/test_gpu_parity.py, run with tests/fake_lib.py (the oracle behind the C
ABI's signatures) injected as the context.  This covers the host logic —
argument marshalling, the reference's error ordering, the tree and poll state
machines, public-input assembly, outcome verification — without a GPU; the GPU
suite runs the very same test functions against the real library."""
import pytest

import tests.test_gpu_parity as G
from tests.fake_lib import FakeContext


@pytest.fixture(scope="module")
def ib():
    import infimum_b200
    from infimum_b200 import context
    saved = dict(context._contexts)
    context._contexts.clear()
    context._contexts[0] = FakeContext()
    yield infimum_b200
    context._contexts.clear()
    context._contexts.update(saved)


# pallet/src/tests/poseidon.rs
test_fr_one = G.test_fr_one
test_bytes_ones_twos = G.test_bytes_ones_twos
test_with_domain_tag = G.test_with_domain_tag
test_fr_one_two = G.test_fr_one_two
test_random_input = G.test_random_input
test_empty_input = G.test_empty_input
test_input_length_and_width_errors = G.test_input_length_and_width_errors
test_circomlibjs_compat_1_to_12_inputs = G.test_circomlibjs_compat_1_to_12_inputs
# pallet/src/poll/zeroes.rs, pallet/src/tests/extrinsics.rs
test_zero_tables = G.test_zero_tables
test_merge_registration_state_success = G.test_merge_registration_state_success
test_merge_interaction_state_success = G.test_merge_interaction_state_success
test_process_messages_public_signals = G.test_process_messages_public_signals
test_participant_limit_reached_quirk = G.test_participant_limit_reached_quirk
# trees, frontier, leaves, paths, custom parameters
test_tree_capacity_edges = G.test_tree_capacity_edges
test_frontier_equals_insert_cascade = G.test_frontier_equals_insert_cascade
test_insert_and_merge_on_a_stored_frontier = G.test_insert_and_merge_on_a_stored_frontier
test_stored_frontier_capacity_and_errors = G.test_stored_frontier_capacity_and_errors
test_leaf_hashing_pins_and_oracle = G.test_leaf_hashing_pins_and_oracle
test_verify_outcome_scenarios = G.test_verify_outcome_scenarios
test_custom_parameters_equal_new_circom_for_the_circom_tables = G.test_custom_parameters_equal_new_circom_for_the_circom_tables
test_maximum_depths_and_empty_inputs = G.test_maximum_depths_and_empty_inputs
test_retained_tree_paths = G.test_retained_tree_paths
test_custom_parameters_vs_oracle = G.test_custom_parameters_vs_oracle
