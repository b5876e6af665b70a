"""mbm-b200: B200-native time-stepping hot path of MovingBoundaryMinerals.jl behind a C ABI.

The directory name contains a dot (it mirrors the reference's repository name), so it cannot be
imported with a plain `import`; use the repo-root shim `mbm_b200` which aliases this package.
"""
from . import abi, host, drivers, ensembles  # noqa: F401
from .abi import load_cuda_library  # noqa: F401
from .host import CoupleEnsemble, SimpleEnsemble  # noqa: F401

__all__ = ["abi", "host", "drivers", "load_cuda_library", "CoupleEnsemble", "SimpleEnsemble"]
