"""ctypes binding of libofdm_b200.so (include/ofdm_b200.h).  No fallback: if the shared library
is missing or a call fails, the error is raised -- there is no CPU path in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))           # lte-gnu-radio-code_b200/
LIB_PATH = os.environ.get("OFDM_B200_LIB", os.path.join(_PKG_ROOT, "libofdm_b200.so"))


class OfdmParams(C.Structure):
    """struct ofdm_params"""
    _fields_ = [
        ("nfft", C.c_int32), ("cp_len", C.c_int32), ("num_synch_bins", C.c_int32),
        ("num_data_bins", C.c_int32), ("m0", C.c_int32), ("d", C.c_int32), ("bps", C.c_int32),
        ("zc_root", C.c_int32), ("zc_parity_of", C.c_int32), ("equalizer", C.c_int32),
        ("sync_advance", C.c_int32), ("gate", C.c_float), ("inv_reg_chanest", C.c_float),
        ("inv_reg_eq", C.c_float), ("inv_reg_eq_synch", C.c_float), ("llr_scale", C.c_float),
        ("pilot_spacing", C.c_int32), ("pilot_offset", C.c_int32), ("pilot_root", C.c_int32),
        ("cp_first", C.c_int32), ("slot_syms", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class RxBlockResult(C.Structure):
    """struct ofdm_rx_block_result"""
    _fields_ = [
        ("n_new_locks", C.c_int32), ("n_demod", C.c_int32), ("obs_after_search", C.c_int32), ("n_state_rows", C.c_int32),
        ("time_synch_ref", C.POINTER(C.c_int64)), ("new_rows", C.POINTER(C.c_int32)), ("chan_time", C.POINTER(C.c_float)),
        ("synch_freq", C.POINTER(C.c_float)), ("demod_rows", C.POINTER(C.c_int32)), ("data_freq", C.POINTER(C.c_float)),
    ]


_P = C.c_void_p
_I64 = C.c_int64
SIGNATURES = {
    # name: (restype, argtypes)
    "ofdm_abi_version": (C.c_int, []),
    "ofdm_last_error": (C.c_char_p, []),
    "ofdm_launch_count": (C.c_uint64, []),
    "ofdm_plan_create": (C.c_int, [C.POINTER(OfdmParams), C.POINTER(_P)]),
    "ofdm_plan_destroy": (C.c_int, [_P]),
    "ofdm_plan_hard_stride": (C.c_int, [_P]),
    "ofdm_plan_frame_len": (C.c_int, [_P]),
    "ofdm_plan_payload_bins": (C.c_int, [_P]),
    "ofdm_tx_modulate": (C.c_int, [_P, _P, _P, _I64, _P]),
    "ofdm_rx_sync_search": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "ofdm_rx_chanest": (C.c_int, [_P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "ofdm_rx_demod": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "ofdm_rx_frames": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ofdm_plan_set_cfo_bank": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "ofdm_rx_sync_search_cfo": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P]),
    "ofdm_rx_chanest_cfo": (C.c_int, [_P, _P, _I64, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "ofdm_rx_demod_cfo": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "ofdm_rx_frames_cfo": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ofdm_dsss_despread": (C.c_int, [_P, _P, _I64, C.c_int32, _P, C.c_int32, _P, _P]),
    "ofdm_bit_recovery": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _P]),
    "ofdm_counters": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P, _P]),
    "ofdm_channel_awgn": (C.c_int, [_P, _P, _P, _I64, _P, C.c_int32, C.c_int32, C.c_float, C.c_uint64, _P]),
    "ofdm_rx_frames_host": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "ofdm_tx_modulate_host": (C.c_int, [_P, _P, _P, _I64]),
    "ofdm_rx_block_create": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "ofdm_rx_block_destroy": (C.c_int, [_P]),
    "ofdm_rx_block_work": (C.c_int, [_P, _P, _I64, C.POINTER(RxBlockResult)]),
    "ofdm_rx_block_work_device": (C.c_int, [_P, _P, _I64, _P, _P, _I64, _I64, _P, C.POINTER(RxBlockResult),
                                           C.POINTER(_I64)]),
    "ofdm_rx_block_max_rows": (C.c_int, [_P]),
    "ofdm_rx_block_chan_freq": (C.c_int, [_P, _P, C.c_int32]),
}

_lib = None


class OfdmError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OfdmError("libofdm_b200.so not found at %s -- build it with `python -c 'import "
                            "__graft_entry__ as g; g.build()'` or `make -C lte-gnu-radio-code_b200/csrc`"
                            % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.ofdm_abi_version() != 5:
            raise OfdmError("libofdm_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise OfdmError("libofdm_b200: error %d: %s" % (rc, lib().ofdm_last_error().decode()))
