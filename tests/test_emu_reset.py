"""Per-environment reset, day rotation and auto-reset (ABIDESEnv.reset is per environment object in the reference, ABIDESEnv.py:51-57; the date sweep
of config/execution/marketreplay/..._parallel.py maps onto "environment e replays day (e + resets) % n_days") -- CPU emulation of the product logic;
the GPU suite repeats the checks on the CUDA kernels."""
import os

import numpy as np
import pytest

from helpers import build_emu
from marl_optimal_execution_b200 import _lib
from reset_cases import abidesenv_masked_reset_and_rotation, ddqn_auto_reset, order_level_one


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def test_abidesenv_masked_reset_and_day_rotation(emu, golden_dir):
    abidesenv_masked_reset_and_rotation(golden_dir, emu)


def test_ddqn_shape_auto_reset_next_day(emu, golden_dir):
    ddqn_auto_reset(golden_dir, emu)


def test_order_level_one_actions(emu, golden_dir):
    order_level_one(golden_dir, emu)


def test_queue_cap_below_the_on_chip_queue_is_rejected(emu, golden_dir):
    from marl_optimal_execution_b200.env import ABIDESEnv, env_config
    g = np.load(os.path.join(golden_dir, "env_IBM_2003-01-14_s789.npz"))
    with pytest.raises(_lib.AbxError):
        ABIDESEnv(g["stream"], n_envs=2, cfg=env_config(_lib.load(emu), queue_cap=32), lib_path=emu)
