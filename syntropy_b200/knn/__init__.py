"""``syntropy_b200.knn`` -- the ten public names of ``syntropy.knn``
(reference: syntropy/knn/__init__.py:20-31), same signatures and return conventions."""
from .shannon import (
    differential_entropy,
    mutual_information,
    conditional_mutual_information,
    kullback_leibler_divergence,
)
from .multivariate_mi import (
    total_correlation,
    dual_total_correlation,
    s_information,
    o_information,
)
from .temporal import sample_entropy, cross_sample_entropy

__all__ = [
    "differential_entropy",
    "mutual_information",
    "conditional_mutual_information",
    "kullback_leibler_divergence",
    "total_correlation",
    "dual_total_correlation",
    "s_information",
    "o_information",
    "sample_entropy",
    "cross_sample_entropy",
]
