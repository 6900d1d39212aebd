"""pyorc_b200 - B200-native LSPIV cross-correlation engine behind pyorc's ``Frames.get_piv`` / ``get_ffpiv``.

Public surface (mirrors the reference's for this one path):

* :func:`pyorc_b200.frames.get_piv`          <- ``pyorc.api.frames.Frames.get_piv``
* :func:`pyorc_b200.velocimetry.get_b2piv`   <- ``pyorc.velocimetry.get_ffpiv``
* :mod:`pyorc_b200.window`                   <- ``ffpiv.window``
* :class:`pyorc_b200.engine.Engine`          <- ``ffpiv.cross_corr`` + ``ffpiv.u_v_displacement`` (fused, CUDA)

The CUDA library (``libb2piv.so``) is required; nothing here computes PIV on the CPU.
"""

from . import window  # noqa: F401

__version__ = "0.1.0"
__all__ = ["window", "get_piv", "get_b2piv", "Engine", "get_engine"]


def __getattr__(name):  # lazy: importing the package must not need the built library
    if name in ("Engine", "get_engine"):
        from . import engine

        return getattr(engine, name)
    if name == "get_b2piv":
        from .velocimetry import get_b2piv

        return get_b2piv
    if name == "get_piv":
        from .frames import get_piv

        return get_piv
    raise AttributeError(name)
