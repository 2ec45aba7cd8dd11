"""``syntropy_b200.mixed`` -- the knn branch of ``syntropy.mixed`` (reference:
syntropy/mixed/__init__.py): discrete x continuous entropy, conditional entropy and mutual
information with ``continuous_estimator="knn"``, i.e. the callers of
``knn.differential_entropy`` per discrete class."""
from .shannon import shannon_entropy, conditional_entropy, mutual_information

__all__ = ["shannon_entropy", "conditional_entropy", "mutual_information"]
