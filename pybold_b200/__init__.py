"""pybold_b200 -- B200-native batched solver for pyBOLD's deconvolution hot path.

Mirror of the reference's module layout for the path only:
``bold_signal`` (deconv, bd, hrf_estim, hrf_fit_err), ``linear`` (DiscretInteg, ConvAndLinear),
``convolution`` (short-kernel convolve / retro-convolve), ``hrf_model`` (spm_hrf),
``utils`` (spectral_radius_est).  Importing the package loads ``libpybold_b200.so`` and fails
loudly if it is missing: there is no CPU implementation.
"""
from . import _lib  # noqa: F401  (loads the shared library, raises ImportError if absent)
from .bold_signal import bd, deconv, hrf_estim, hrf_fit_err  # noqa: F401
from .hrf_model import MAX_DELTA, MIN_DELTA, spm_hrf  # noqa: F401
from .linear import ConvAndLinear, DiscretInteg  # noqa: F401

__version__ = "0.1.0"
