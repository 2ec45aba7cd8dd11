"""syntropy_b200 -- B200-native drop-in for the ``syntropy.knn`` KSG hot path.

Only what the path needs lives here: ``knn`` (the reference's public functions), ``mixed`` (the
knn branch of ``syntropy.mixed``: the per-class caller of ``knn.differential_entropy``),
``csrc/`` (sm_100a kernels + the C ABI of ``include/ksg_b200.h``), ``dist`` (query-row
sharding over NCCL) and ``workloads`` (the synthetic BASELINE configurations).
"""
from . import knn  # noqa: F401
from . import mixed  # noqa: F401
from . import dist  # noqa: F401

__version__ = "0.1.0"
