"""ctypes binding of libb200sdr.so (include/b200sdr.h).

The product path is the CUDA library and nothing else: if ``libb200sdr.so`` is missing or does
not export a symbol the header declares, importing this module raises -- there is no CPU
fallback anywhere under ``futuresdr_b200/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2S_LIB lets a developer A/B a differently-built library (kernel tuning); default is the in-tree build
SO_PATH = os.environ.get("B2S_LIB") or os.path.join(_HERE, "libb200sdr.so")

OK, EINVAL, ECUDA, ENOMEM, EAGAIN, EUNSUPPORTED, ESTATE, ETIMEOUT = 0, -1, -2, -3, -4, -5, -6, -7
INSUFFICIENT_INPUT, INSUFFICIENT_OUTPUT, BOTH_SUFFICIENT = 0, 1, 2
F32_F32, C32_F32, C32_C32, F64_F64 = 0, 1, 2, 3
ALGO_AUTO, ALGO_DIRECT, ALGO_TENSOR, ALGO_FFT = 0, 1, 2, 3
(OP_SCALE_F32, OP_SCALE_C32, OP_QUAD_DEMOD, OP_NORM_SQR, OP_QUAD_DEMOD_C32, OP_EXP_F32,
 OP_MAG_C32, OP_LOG10_F32) = range(8)

_vp, _sz, _i32, _f32 = C.c_void_p, C.c_size_t, C.c_int32, C.c_float
_szp, _i32p, _vpp, _f32p = C.POINTER(C.c_size_t), C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_float)

# name -> (restype, argtypes): one entry per function declared in include/b200sdr.h
SIGNATURES = {
    "b2s_version": (_i32, []),
    "b2s_ctx_create": (_i32, [C.c_int, _vpp]),
    "b2s_ctx_create_on_stream": (_i32, [C.c_int, _vp, _vpp]),
    "b2s_ctx_destroy": (None, [_vp]),
    "b2s_last_error": (C.c_char_p, [_vp]),
    "b2s_ctx_sync": (_i32, [_vp]),
    "b2s_ctx_stream": (_vp, [_vp]),
    "b2s_ctx_sm_count": (_i32, [_vp]),
    "b2s_ctx_launch_count": (C.c_uint64, [_vp]),
    "b2s_malloc": (_i32, [_vp, _sz, _vpp]),
    "b2s_free": (_i32, [_vp, _vp]),
    "b2s_host_alloc": (_i32, [_vp, _sz, _vpp]),
    "b2s_host_free": (_i32, [_vp, _vp]),
    "b2s_memcpy_h2d": (_i32, [_vp, _vp, _vp, _sz]),
    "b2s_memcpy_d2h": (_i32, [_vp, _vp, _vp, _sz]),
    "b2s_fir_plan": (_i32, [_vp, C.c_int, _f32p, _sz, _sz, _vpp]),
    "b2s_fir_plan_f32_f32": (_i32, [_vp, _f32p, _sz, _sz, _vpp]),
    "b2s_fir_plan_c32_f32": (_i32, [_vp, _f32p, _sz, _sz, _vpp]),
    "b2s_fir_plan_c32_c32": (_i32, [_vp, _f32p, _sz, _sz, _vpp]),
    "b2s_fir_plan_f64_f64": (_i32, [_vp, C.POINTER(C.c_double), _sz, _sz, _vpp]),
    "b2s_fir_destroy": (None, [_vp]),
    "b2s_fir_length": (_sz, [_vp]),
    "b2s_fir_set_algo": (_i32, [_vp, C.c_int]),
    "b2s_fir_get_algo": (_i32, [_vp]),
    "b2s_fir_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp, _i32p]),
    "b2s_fir_filter_host": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp, _i32p]),
    "b2s_fir_exec_hist": (_i32, [_vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _szp, _szp, _i32p]),
    "b2s_resamp_plan": (_i32, [_vp, C.c_int, _f32p, _sz, _sz, _sz, _vpp]),
    "b2s_resamp_destroy": (None, [_vp]),
    "b2s_resamp_length": (_sz, [_vp]),
    "b2s_resamp_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp, _i32p]),
    "b2s_pfbarb_plan_c32": (_i32, [_vp, _f32p, _sz, _sz, _f32, _vpp]),
    "b2s_pfbarb_destroy": (None, [_vp]),
    "b2s_pfbarb_reset": (_i32, [_vp]),
    "b2s_pfbarb_period": (_i32, [_f32, _sz, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b2s_pfbarb_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp, _i32p]),
    "b2s_fft_plan_c32": (_i32, [_vp, _sz, _i32, _i32, _i32, _f32, _vpp]),
    "b2s_fft_destroy": (None, [_vp]),
    "b2s_fft_length": (_sz, [_vp]),
    "b2s_fft_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp]),
    "b2s_apply_create": (_i32, [_vp, C.c_int, _f32, _vpp]),
    "b2s_apply_destroy": (None, [_vp]),
    "b2s_apply_reset": (_i32, [_vp]),
    "b2s_apply_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp]),
    "b2s_rotator_create": (_i32, [_vp, _f32, _vpp]),
    "b2s_rotator_destroy": (None, [_vp]),
    "b2s_rotator_reset": (_i32, [_vp]),
    "b2s_rotator_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _i32p]),
    "b2s_xlating_taps": (_i32, [_f32p, _sz, _f32, _f32, _sz, _f32p, _f32p]),
    "b2s_chan_plan_c32": (_i32, [_vp, _sz, _f32p, _sz, _f32, _vpp]),
    "b2s_chan_destroy": (None, [_vp]),
    "b2s_chan_decimation": (_sz, [_vp]),
    "b2s_chan_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _sz, _szp, _szp, _i32p]),
    "b2s_synth_plan_c32": (_i32, [_vp, _sz, _f32p, _sz, _vpp]),
    "b2s_synth_destroy": (None, [_vp]),
    "b2s_synth_exec": (_i32, [_vp, _vp, _sz, _sz, _vp, _sz, _szp, _szp]),
    "b2s_mavg_create": (_i32, [_vp, _sz, _f32, _sz, _vpp]),
    "b2s_mavg_destroy": (None, [_vp]),
    "b2s_mavg_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp]),
    "b2s_spectrum_plan": (_i32, [_vp, _sz, _i32, _f32, _sz, _f32, _vpp]),
    "b2s_spectrum_destroy": (None, [_vp]),
    "b2s_spectrum_reset": (_i32, [_vp]),
    "b2s_spectrum_exec": (_i32, [_vp, _vp, _sz, _vp, _sz, _szp, _szp]),
    "b2s_ring_create": (_i32, [_vp, _sz, _sz, _sz, _i32, _i32, _vpp]),
    "b2s_ring_destroy": (None, [_vp]),
    "b2s_ring_acquire_empty": (_i32, [_vp, _vpp]),
    "b2s_ring_submit_full": (_i32, [_vp, _vp, _sz, _i32]),
    "b2s_ring_acquire_full": (_i32, [_vp, _vpp, _szp]),
    "b2s_ring_release": (_i32, [_vp, _vp]),
    "b2s_ring_carry_halo": (_i32, [_vp, _vp, _sz, _sz, _vp]),
    "b2s_slot_device_ptr": (_vp, [_vp]),
    "b2s_slot_host_ptr": (_vp, [_vp]),
    "b2s_slot_halo_valid": (_sz, [_vp]),
    "b2s_slot_fetch_to_host": (_i32, [_vp, _sz]),
    "b2s_slot_wait": (_i32, [_vp]),
    "b2s_ring_free_slots": (_sz, [_vp]),
    "b2s_ring_full_slots": (_sz, [_vp]),
    "b2s_ring_base": (_vp, [_vp]),
    "b2s_ring_bytes": (_sz, [_vp]),
    "b2s_ring_slot_offset": (_sz, [_vp, _i32]),
    "b2s_ring_flags_offset": (_sz, [_vp]),
    "b2s_slot_index": (_i32, [_vp]),
    "b2s_ipc_export": (_i32, [_vp, _vp, _vp]),
    "b2s_ipc_open": (_i32, [_vp, _vp, _vpp]),
    "b2s_ipc_close": (_i32, [_vp, _vp]),
    "b2s_peer_enable": (_i32, [_vp, _i32]),
    "b2s_flag_set": (_i32, [_vp, _vp, C.c_uint32]),
    "b2s_flag_wait": (_i32, [_vp, _vp, C.c_uint32]),
    "b2s_flag_read": (_i32, [_vp, _vp, C.POINTER(C.c_uint32)]),
    "b2s_memcpy_d2d": (_i32, [_vp, _vp, _vp, _sz]),
    "b2s_memset": (_i32, [_vp, _vp, _i32, _sz]),
    "b2s_firdes_kaiser_lowpass": (_sz, [C.c_double, C.c_double, C.c_double, _f32p, _sz]),
    "b2s_firdes_kaiser_multirate": (_sz, [_sz, _sz, _sz, C.c_double, _f32p, _sz]),
}


class Handshake(C.Structure):
    """b2s_handshake (include/b200sdr.h): the cross-GPU flags of one b2s_fir_exec_hist call."""
    _fields_ = [("publish_flag", C.c_void_p), ("publish_value", C.c_uint32),
                ("wait_flag", C.c_void_p), ("wait_value", C.c_uint32),
                ("done_flag", C.c_void_p), ("done_value", C.c_uint32)]


class B200SdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200sdr error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C futuresdr_b200/csrc`). futuresdr_b200 has no CPU fallback."
        )
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{SO_PATH} does not export {name} (stale build?)") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, ctx=None):
    if rc < 0:
        msg = lib.b2s_last_error(ctx)
        raise B200SdrError(rc, msg.decode() if msg else "")
    return rc
