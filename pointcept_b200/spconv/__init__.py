"""Drop-in for the ``spconv`` package surface Pointcept imports (``import spconv.pytorch as spconv``)."""
from . import pytorch  # noqa: F401
