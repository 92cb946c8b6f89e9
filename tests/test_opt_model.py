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


def test_functional_basis_schedule_equals_dense():
    """Width 3 with the passive state carried as the next pair's two functionals (Layout<3>::FB)."""
    T = opt_model.derive(3)
    F = opt_model.derive_fb(3, T)
    assert len(F["pairs"]) == 28 and F["last"] is not None
    rng = random.Random(33)
    for i in range(12):
        ins = [rng.randrange(O.P) for _ in range(2)]
        tag = 0 if i % 3 else rng.randrange(O.P)
        assert opt_model.hash_opt_fb(ins, tag, T, F) == O.poseidon_permute_hash(ins, tag)


def test_rows_over_existing_values_schedule_equals_dense():
    """Width 3 as the kernels run it: b eliminated, rows over Q = (a, u', z_a, z_b), odd round first."""
    T = opt_model.derive(3)
    F = opt_model.derive_fb2(3, T)
    assert len(F["pairs"]) == 28
    rng = random.Random(34)
    for i in range(12):
        ins = [rng.randrange(O.P) for _ in range(2)]
        tag = 0 if i % 3 else rng.randrange(O.P)
        assert opt_model.hash_opt_fb2(ins, tag, T, F) == O.poseidon_permute_hash(ins, tag)


@pytest.mark.parametrize("t", list(range(2, 9)))
def test_history_recurrence_schedule_equals_dense(t):
    """The partial rounds as a recurrence over the last t-1 S-box inputs and outputs (Layout<T>::HR)."""
    T = opt_model.derive(t)
    H = opt_model.derive_hr(t, T)
    assert len(H["steady"]) == T["rp"] - (t - 1) and len(H["exit"]) == t - 1
    rng = random.Random(35 + t)
    for i in range(6):
        ins = [rng.randrange(O.P) for _ in range(t - 1)]
        tag = 0 if i % 3 else rng.randrange(O.P)
        assert opt_model.hash_opt_hr(ins, tag, T, H) == O.poseidon_permute_hash(ins, tag)
