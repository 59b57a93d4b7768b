"""marl_optimal_execution_b200 -- B200-native batched ABIDES market simulator (hot path only).

The package holds the CUDA kernels + C ABI (csrc/, include/abides_b200.h at the repo root) and the thin
Python host mirror of the reference's config / Kernel / ABIDESEnv surfaces.  No CPU fallback exists.
"""
from ._lib import AbxError, SimConfig, EnvStats  # noqa: F401
from .sim import BatchedSim, sparse_zi_config  # noqa: F401
from .env import ABIDESEnv, DDQNExecutionEnv, dq_config, env_config  # noqa: F401
from .qnet import QNetwork  # noqa: F401
