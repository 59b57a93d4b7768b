"""CPU CI of the book surface through the host emulation harness of the product logic (a test tool, never a fallback)."""
import pytest

import book_cases
from helpers import build_emu


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def test_kat1(emu):
    book_cases.kat1(emu)


@pytest.mark.parametrize("fixture,n_ops", [("env_IBM_2003-01-14_s789.npz", 10000), ("ddqn_IBM_2003-01-14_s4242.npz", 12000)])
def test_recorded_operation_tape(emu, golden_dir, fixture, n_ops):
    fills, modifies = book_cases.recorded_tape(golden_dir, fixture, n_ops, emu)
    assert fills > 500 and modifies > 1000


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_adversarial_tape(emu, seed):
    mods, execs = book_cases.random_tape_vs_oracle(emu, seed=seed)
    assert mods > 200 and execs > 500
