"""The optimised schedule (sparse partial rounds, compressed constants) is an
exact rewriting of the reference algorithm: model == dense oracle."""
import random

import pytest

from oracle import poseidon_ref as O
from tests import opt_model


@pytest.mark.parametrize("t", list(range(2, 14)))
def test_optimised_schedule_equals_dense(t):
    rng = random.Random(1000 + t)
    tables = opt_model.derive(t)
    cases = [[1] * (t - 1), [0] * (t - 1), [O.P - 1] * (t - 1)]
    cases += [[rng.randrange(O.P) for _ in range(t - 1)] for _ in range(3)]
    for ins in cases:
        for tag in (0, 5):
            assert opt_model.hash_opt(ins, tag, tables) == O.poseidon_permute_hash(ins, tag)


@pytest.mark.parametrize("t", list(range(2, 14)))
def test_paired_partial_rounds_equal_dense(t):
    rng = random.Random(2000 + t)
    tables = opt_model.derive(t)
    for ins in ([1] * (t - 1), [O.P - 1] * (t - 1), [rng.randrange(O.P) for _ in range(t - 1)]):
        for tag in (0, 9):
            assert opt_model.hash_opt_paired(ins, tag, tables) == O.poseidon_permute_hash(ins, tag)


@pytest.mark.parametrize("t", list(range(2, 14)))
def test_unit_leading_coefficient_schedule_equals_dense(t):
    """Round 2's schedule: s0 carried as u = s0 / lambda_j so that a partial round
    is u' = u^5 + v'.s[1:] + k' — paired and unpaired forms against the dense oracle."""
    rng = random.Random(3000 + t)
    tables = opt_model.derive(t)
    for ins in ([1] * (t - 1), [0] * (t - 1), [O.P - 1] * (t - 1), [rng.randrange(O.P) for _ in range(t - 1)]):
        for tag in (0, 11):
            exp = O.poseidon_permute_hash(ins, tag)
            assert opt_model.hash_opt_scaled(ins, tag, tables, True) == exp
            assert opt_model.hash_opt_scaled(ins, tag, tables, False) == exp
