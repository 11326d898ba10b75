"""``pointrope`` drop-in (libs/pointrope/pointrope.cpp): ``pointrope(tokens [B,N,H,D], positions [B,N,3] int64, base, F0)`` rotates
the tokens in place -- the entry point litept_v1.py:27-46 wraps in its PointROPE_func (forward F0, backward -F0)."""
from .ops import pointrope_ as pointrope  # noqa: F401
