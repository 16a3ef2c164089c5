"""The workload builders live in the package (unbiasedmcmc_b200/workloads.py); tests import them under the old name."""
from unbiasedmcmc_b200.workloads import *  # noqa: F401,F403
from unbiasedmcmc_b200.workloads import SEED  # noqa: F401
