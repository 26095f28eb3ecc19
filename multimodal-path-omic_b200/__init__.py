"""B200-native slide hot path of mattiagualtieri/multimodal-path-omic (MCAT / NaCAGaT / GE-NaCAGaT).

The directory name carries a hyphen (it mirrors the reference repository's name), so import it through the
`mpo_b200` alias module at the repository root:  `import mpo_b200 as mpo`.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
