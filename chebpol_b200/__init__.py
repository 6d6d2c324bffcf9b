"""chebpol_b200 -- B200 (sm_100a) batched evaluation of chebpol interpolants.

The package is the host side of ``libchebpol_cuda`` (C ABI:
``include/chebpol_cuda.h``): a ctypes binding (``_lib``) and a mirror of the R
interface for the accelerated path (``interp``: ``ipol`` / ``chebappx`` /
``mlappx`` / ``fhappx`` / ``ucappx`` returning a callable interpolant).

There is no CPU evaluation path: the first call that needs the library raises
ImportError if ``libchebpol_cuda.so`` has not been built, and every compute
entry point fails with a ChebpolCudaError when no B200 is visible.
"""
from . import _lib
from ._lib import ChebpolCudaError, Plan, device_count, fp64_peak_tflops, launch_count
from .interp import (Interpolant, chebappx, chebappxf, chebappxg, chebappxgf, hyman_spline, chebcoef, chebeval, chebknots, evalongrid, evalongridV,
                     fhappx, fhweights, hstalkerappx, ipol, mlappx, oldstalkerappx, polyh, slappx, stalkerappx, ucappx, ucappxf)

__version__ = "0.1.0"
__all__ = [
    "ipol", "chebappx", "chebappxf", "chebappxg", "chebappxgf", "hyman_spline", "mlappx", "fhappx", "polyh", "slappx", "oldstalkerappx", "stalkerappx", "hstalkerappx", "ucappx", "ucappxf", "chebeval", "chebknots",
    "chebcoef", "evalongrid", "evalongridV", "fhweights", "Interpolant", "Plan", "ChebpolCudaError", "device_count",
    "fp64_peak_tflops", "launch_count",
]
