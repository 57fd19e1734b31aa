"""futuresdr_b200 -- B200-native (sm_100a) backend for FutureSDR's FIR / decimator / resampler /
FFT / Apply / PfbArbResampler hot path.

Python host layer above the C ABI (include/b200sdr.h).  Class and method names mirror the
reference's Rust API for this path (futuredsp::{FirFilter, DecimatingFirFilter,
PolyphaseResamplingFir}, futuresdr::blocks::{Fir, FirBuilder, Fft, Apply, PfbArbResampler},
runtime::mocker::Mocker) so the parity tests read like the reference's own tests.
Importing this package loads libb200sdr.so and raises if it is missing: no CPU fallback.
"""
from ._lib import (  # noqa: F401
    B200SdrError,
    INSUFFICIENT_INPUT, INSUFFICIENT_OUTPUT, BOTH_SUFFICIENT,
    ALGO_AUTO, ALGO_DIRECT, ALGO_TENSOR, ALGO_FFT,
)
from .context import Context, default_context  # noqa: F401
from .filters import (  # noqa: F401
    ComputationStatus, FirFilter, DecimatingFirFilter, PolyphaseResamplingFir,
)
from . import firdes  # noqa: F401
# host edges (VectorSource/Sink, FileSource/Sink, H2D/D2H ring, run_chain): futuresdr_b200.edges
