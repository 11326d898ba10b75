"""pointcept_b200 -- B200 (sm_100a) operators behind Pointcept's PT-v3 / SpUNet hot path.

``install()`` registers the drop-in modules under the import names the reference uses
(``spconv``, ``spconv.pytorch`` and optionally ``flash_attn``) so unmodified Pointcept model files
and configs resolve to these operators.
"""
import sys

__version__ = "0.1.0"


def install(flash_attn=True):
    from . import spconv as _spconv
    sys.modules["spconv"] = _spconv
    sys.modules["spconv.pytorch"] = _spconv.pytorch
    sys.modules["spconv.pytorch.modules"] = _spconv.pytorch.modules
    sys.modules["spconv.pytorch.conv"] = _spconv.pytorch.conv
    sys.modules["spconv.pytorch.core"] = _spconv.pytorch.core
    if flash_attn:
        from . import flash_attn_interface as _fa
        import types
        m = types.ModuleType("flash_attn")
        m.flash_attn_varlen_qkvpacked_func = _fa.flash_attn_varlen_qkvpacked_func
        m.__version__ = "2.8.3+b2pc"
        m.flash_attn_interface = _fa
        sys.modules["flash_attn"] = m
        sys.modules["flash_attn.flash_attn_interface"] = _fa
