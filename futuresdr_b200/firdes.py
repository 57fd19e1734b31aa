"""Tap design (host, f64): futuredsp::firdes::kaiser (crates/futuredsp/src/firdes/basic.rs:310-459)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib


class kaiser:
    @staticmethod
    def lowpass(cutoff: float, transition_bw: float, max_ripple: float) -> np.ndarray:
        """firdes::kaiser::lowpass::<f32> (basic.rs:310-321)."""
        assert cutoff > 0.0, "cutoff must be greater than 0"
        assert transition_bw > 0.0, "transition_bw must be greater than 0"
        assert cutoff + transition_bw < 0.5, "cutoff+transition_bw must be less than 1/2"
        n = lib.b2s_firdes_kaiser_lowpass(cutoff, transition_bw, max_ripple, None, 0)
        t = np.zeros(n, np.float32)
        lib.b2s_firdes_kaiser_lowpass(cutoff, transition_bw, max_ripple,
                                      t.ctypes.data_as(C.POINTER(C.c_float)), n)
        return t

    @staticmethod
    def multirate(interp: int, decim: int, half_polyphase_len: int, max_ripple: float) -> np.ndarray:
        """firdes::kaiser::multirate::<f32> (basic.rs:412-442)."""
        assert interp > 0 and decim > 0 and half_polyphase_len > 0
        n = lib.b2s_firdes_kaiser_multirate(interp, decim, half_polyphase_len, max_ripple, None, 0)
        t = np.zeros(n, np.float32)
        lib.b2s_firdes_kaiser_multirate(interp, decim, half_polyphase_len, max_ripple,
                                        t.ctypes.data_as(C.POINTER(C.c_float)), n)
        return t
