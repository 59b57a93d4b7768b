"""GPU parity of the book surface through the C ABI (abx_book_create / abx_book_replay): the reference's known-answer vector and
operation tapes recorded at the reference's exchange boundary replayed through the GPU books, bit-exact in fills, notifications
and book snapshots."""
import pytest

import book_cases

pytestmark = pytest.mark.gpu


def test_kat1():
    book_cases.kat1()


@pytest.mark.parametrize("fixture,n_ops", [("env_IBM_2003-01-14_s789.npz", 10000), ("ddqn_IBM_2003-01-14_s4242.npz", 15000)])
def test_recorded_operation_tape(golden_dir, fixture, n_ops):
    fills, modifies = book_cases.recorded_tape(golden_dir, fixture, n_ops, n_envs=64)
    assert fills > 500 and modifies > 1000


@pytest.mark.parametrize("seed", [0, 5])
def test_adversarial_tape(seed):
    mods, execs = book_cases.random_tape_vs_oracle(seed=seed, n_ops=20000)
    assert mods > 500 and execs > 1500


@pytest.mark.parametrize("seed,reprice", [(0, 0.3), (2, 1.0), (7, 0.6)])
def test_adversarial_tape_with_repricing_modifies(seed, reprice):
    """MODIFY_ORDER with a new price: slot 0 of the level takes the new order where it stands, the level shows the new price, the level lists go
    unsorted and several levels may show one price (util/OrderBook.py:350-352,381,393) -- every notification and book snapshot equal to the oracle's."""
    mods, execs = book_cases.random_tape_vs_oracle(seed=seed, n_ops=12000, reprice=reprice)
    assert mods > 300 and execs > 800


def test_repricing_modifies_match_the_live_reference(golden_dir):
    assert book_cases.reprice_golden(golden_dir) > 300
