"""``pointops`` drop-in for the query the PT-v3 / SpUNet evaluation path calls (pointcept/engines/hooks/evaluator.py:569-575,
libs/pointops/functions/query.py:7-26): ``knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None) -> (idx, dist)``."""
from .ops import knn_query  # noqa: F401
