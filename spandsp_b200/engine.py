"""ctypes binding of libspandsp_b200.so (the C ABI in include/spandsp_b200.h).

This is plumbing for tests and the benchmark: it passes raw device pointers (e.g. from
torch tensors) and plain ints through the C ABI.  There is no Python fallback of any
kernel: if the shared library is missing this module raises at import of the symbols,
and if no sm_100 GPU is present ``Context()`` raises with the library's error text.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libspandsp_b200.so")

DET_DTMF, DET_BELL_MF, DET_R2_MF, DET_SUPER_TONE = 0, 1, 2, 3
EV_DIGIT, EV_TONE, EV_SEGMENT = 1, 2, 5

EVENT_DTYPE = np.dtype([("channel", "<i4"), ("block", "<i4"), ("kind", "<i4"),
                        ("a", "<i4"), ("b", "<i4"), ("c", "<i4")])
WIRE_DTYPE = np.dtype([("channel", "<u4"), ("c", "<i4"), ("block_kind", "<u2"), ("a", "i1"), ("b", "i1")])
assert WIRE_DTYPE.itemsize == 12


class SuperToneDesc(C.Structure):
    _fields_ = [("tones", C.c_int32), ("tone_segs", C.POINTER(C.c_int32)), ("elements", C.POINTER(C.c_int32))]


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing - build it with `python -m spandsp_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    sig = {
        "span_b200_abi_version": (i32, []),
        "span_b200_last_error": (C.c_char_p, []),
        "span_b200_ctx_create": (vp, [i32]),
        "span_b200_ctx_destroy": (None, [vp]),
        "span_b200_ctx_device": (i32, [vp]),
        "span_b200_ctx_sm_count": (i32, [vp]),
        "span_b200_dtmf_bank_create": (vp, [vp, i32]),
        "span_b200_bell_mf_bank_create": (vp, [vp, i32]),
        "span_b200_r2_mf_bank_create": (vp, [vp, i32, i32]),
        "span_b200_super_tone_bank_create": (vp, [vp, i32, C.POINTER(SuperToneDesc), i32]),
        "span_b200_bank_destroy": (None, [vp]),
        "span_b200_bank_channels": (i32, [vp]),
        "span_b200_bank_detector": (i32, [vp]),
        "span_b200_bank_block_len": (i32, [vp]),
        "span_b200_bank_bins": (i32, [vp]),
        "span_b200_bank_coefficients": (i32, [vp, vp, i32]),
        "span_b200_bank_reset": (i32, [vp, i32, i32]),
        "span_b200_dtmf_bank_parms": (i32, [vp, i32, i32, i32, f32, f32, f32]),
        "span_b200_dtmf_bank_realtime": (i32, [vp, i32, i32, i32]),
        "span_b200_dtmf_bank_fillin": (i32, [vp, i32, i32]),
        "span_b200_bank_status": (i32, [vp, i32, i32, vp]),
        "span_b200_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_bank_rx_device_g711": (i32, [vp, vp, i64, i32, i32, vp]),
        "span_b200_bank_rx_host_g711": (i32, [vp, vp, i64, i32, i32, vp]),
        "span_b200_bank_event_count": (i64, [vp, C.POINTER(i32)]),
        "span_b200_bank_events": (i64, [vp, vp, i64]),
        "span_b200_bank_events_device": (vp, [vp]),
        "span_b200_bank_set_event_capacity": (i32, [vp, i64]),
        "span_b200_bank_events_to_device": (i64, [vp, vp, i64, vp]),
        "span_b200_bank_kernel_ms": (C.c_double, [vp, C.POINTER(i32)]),
        "span_b200_bank_set_wire": (i32, [vp, i32, C.c_uint32]),
        "span_b200_bank_events_wire": (i64, [vp, vp, i64]),
        "span_b200_wire_expand": (None, [vp, vp, i64, C.c_uint32]),
        "span_b200_comm_unique_id": (i32, [vp]),
        "span_b200_comm_create": (vp, [vp, vp, i32, i32, i32]),
        "span_b200_comm_destroy": (None, [vp]),
        "span_b200_comm_rank": (i32, [vp]),
        "span_b200_comm_set_transport": (i32, [vp, i32]),
        "span_b200_comm_transport": (i32, [vp]),
        "span_b200_comm_nranks": (i32, [vp]),
        "span_b200_comm_sync": (i32, [vp]),
        "span_b200_bank_attach_comm": (i32, [vp, vp, i32]),
        "span_b200_bank_gather_begin": (i32, [vp]),
        "span_b200_bank_gather_end": (i64, [vp, vp]),
        "span_b200_bank_gathered": (i64, [vp, vp]),
        "span_b200_bank_gathered_host": (i64, [vp, vp, i64]),
        "span_b200_ctx_numa_node": (i32, [vp]),
        "span_b200_host_alloc": (vp, [vp, C.c_size_t]),
        "span_b200_host_free": (None, [vp, vp]),
        "span_b200_bank_last_blocks": (i32, [vp]),
        "span_b200_bank_block_codes": (i32, [vp, vp, i64]),
        "span_b200_bank_tune": (i32, [vp, i32, i32]),
        "span_b200_bank_last_path": (C.c_char_p, [vp]),
        "span_b200_bank_last_launches": (i32, [vp]),
        "span_b200_goertzel_blocks_device": (i32, [vp, vp, i32, i32, vp, i64, i32, i32, vp, i64, vp]),
        "span_b200_goertzel_blocks_energy_device": (i32, [vp, vp, i32, i32, vp, i64, i32, i32, vp, i64, vp, vp]),
        "span_b200_goertzel_tone_set": (i32, [i32, vp, i32, vp]),
        "span_b200_goertzel_coefficient": (f32, [f32]),
        "span_b200_rfc4733_event_code": (i32, [i32]),
        "span_b200_rfc4733_pack": (None, [vp, i32, i32, i32, i32]),
        "span_b200_rfc4733_dtmf": (i32, [vp, i32, i32, i32, vp]),
        "span_b200_v29_bank_create": (vp, [vp, i32, i32, i32]),
        "span_b200_v29_bank_destroy": (None, [vp]),
        "span_b200_v29_bank_channels": (i32, [vp]),
        "span_b200_v29_bank_restart": (i32, [vp, i32, i32, i32]),
        "span_b200_v29_bank_set_signal_cutoff": (i32, [vp, i32, i32, f32]),
        "span_b200_v29_bank_fillin": (i32, [vp, i32, i32, i32]),
        "span_b200_v29_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v29_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v29_bank_counts": (i32, [vp, vp, vp]),
        "span_b200_v29_bank_bits": (i64, [vp, i32, vp, i64]),
        "span_b200_v29_bank_symbols": (i64, [vp, i32, vp, i64]),
        "span_b200_v29_bank_output_packed": (i64, [vp, vp, i64, vp, vp, i64, vp]),
        "span_b200_v29_bank_output_layout": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "span_b200_v29_bank_bits_all": (i64, [vp, vp, i64, vp]),
        "span_b200_v29_bank_channel_state": (i32, [vp, i32, vp, vp]),
        "span_b200_v29_tables": (i32, [vp, vp, vp, vp, vp, vp]),
        "span_b200_v29_bank_restart_ex": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_v17_bank_create": (vp, [vp, i32, i32, i32]),
        "span_b200_v17_bank_destroy": (None, [vp]),
        "span_b200_v17_bank_channels": (i32, [vp]),
        "span_b200_v17_bank_restart": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_v17_bank_set_signal_cutoff": (i32, [vp, i32, i32, f32]),
        "span_b200_v17_bank_fillin": (i32, [vp, i32, i32, i32]),
        "span_b200_v17_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v17_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v17_bank_counts": (i32, [vp, vp, vp]),
        "span_b200_v17_bank_bits": (i64, [vp, i32, vp, i64]),
        "span_b200_v17_bank_symbols": (i64, [vp, i32, vp, i64]),
        "span_b200_v17_bank_bits_all": (i64, [vp, vp, i64, vp]),
        "span_b200_v17_bank_output_packed": (i64, [vp, vp, i64, vp, vp, i64, vp]),
        "span_b200_v17_bank_output_layout": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "span_b200_v17_bank_channel_state": (i32, [vp, i32, vp, vp]),
        "span_b200_v17_tables": (i32, [vp, vp, vp, vp, vp, vp, vp]),
        "span_b200_v27ter_bank_create": (vp, [vp, i32, i32, i32]),
        "span_b200_v27ter_bank_destroy": (None, [vp]),
        "span_b200_v27ter_bank_channels": (i32, [vp]),
        "span_b200_v27ter_bank_restart": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_v27ter_bank_set_signal_cutoff": (i32, [vp, i32, i32, f32]),
        "span_b200_v27ter_bank_fillin": (i32, [vp, i32, i32, i32]),
        "span_b200_v27ter_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v27ter_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_v27ter_bank_counts": (i32, [vp, vp, vp]),
        "span_b200_v27ter_bank_bits": (i64, [vp, i32, vp, i64]),
        "span_b200_v27ter_bank_symbols": (i64, [vp, i32, vp, i64]),
        "span_b200_v27ter_bank_bits_all": (i64, [vp, vp, i64, vp]),
        "span_b200_v27ter_bank_output_packed": (i64, [vp, vp, i64, vp, vp, i64, vp]),
        "span_b200_v27ter_bank_output_layout": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "span_b200_v27ter_bank_channel_state": (i32, [vp, i32, vp, vp]),
        "span_b200_v27ter_tables": (i32, [vp, vp, vp, vp, vp]),
        "span_b200_fsk_preset": (vp, [i32]),
        "span_b200_fsk_bank_create": (vp, [vp, i32, vp, i32]),
        "span_b200_fsk_bank_destroy": (None, [vp]),
        "span_b200_fsk_bank_channels": (i32, [vp]),
        "span_b200_fsk_bank_restart": (i32, [vp, i32, i32, vp, i32]),
        "span_b200_fsk_bank_set_signal_cutoff": (i32, [vp, i32, i32, f32]),
        "span_b200_fsk_bank_set_frame_parameters": (i32, [vp, i32, i32, i32, i32, i32]),
        "span_b200_fsk_bank_fillin": (i32, [vp, i32, i32, i32]),
        "span_b200_fsk_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_fsk_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_fsk_bank_counts": (i32, [vp, vp]),
        "span_b200_fsk_bank_output": (i64, [vp, i32, vp, i64]),
        "span_b200_fsk_bank_errors": (i32, [vp, i32, vp, vp, i32]),
        "span_b200_fsk_bank_signal_power": (f32, [vp, i32]),
        "span_b200_fsk_bank_channel_state": (i32, [vp, i32, vp, vp]),
        "span_b200_dds_int_table": (i32, [vp]),
        "span_b200_dtmf_tx_bank_create": (vp, [vp, i32]),
        "span_b200_dtmf_tx_bank_destroy": (None, [vp]),
        "span_b200_dtmf_tx_bank_channels": (i32, [vp]),
        "span_b200_dtmf_tx_bank_init": (i32, [vp, i32, i32]),
        "span_b200_dtmf_tx_bank_set_level": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_dtmf_tx_bank_set_timing": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_dtmf_tx_bank_put": (i32, [vp, i32, i32, C.c_char_p, i32]),
        "span_b200_dtmf_tx_bank_put_each": (i32, [vp, i32, i32, vp, i64, vp]),
        "span_b200_dtmf_tx_bank_tx_device": (i32, [vp, vp, i64, i32, i32, vp]),
        "span_b200_dtmf_tx_bank_tx_host": (i32, [vp, vp, i64, i32, i32]),
        "span_b200_dtmf_tx_bank_lens": (i32, [vp, vp]),
        "span_b200_awgn_bank_create": (vp, [vp, i32, vp, i32, f32]),
        "span_b200_awgn_bank_destroy": (None, [vp]),
        "span_b200_awgn_bank_channels": (i32, [vp]),
        "span_b200_awgn_bank_init_dbm0": (i32, [vp, i32, i32, vp, i32, f32]),
        "span_b200_awgn_bank_init_dbov": (i32, [vp, i32, i32, vp, i32, f32]),
        "span_b200_awgn_bank_add_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_awgn_bank_fill_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_dds_float_table": (i32, [vp]),
        "span_b200_dtmf_tx_bank_sync": (i32, [vp]),
        "span_b200_tone_gen_bank_create": (vp, [vp, i32]),
        "span_b200_tone_gen_bank_destroy": (None, [vp]),
        "span_b200_tone_gen_bank_channels": (i32, [vp]),
        "span_b200_tone_gen_bank_init": (i32, [vp, i32, i32, vp]),
        "span_b200_tone_gen_bank_init_each": (i32, [vp, i32, i32, vp]),
        "span_b200_tone_gen_bank_tx_device": (i32, [vp, vp, i64, i32, i32, vp]),
        "span_b200_tone_gen_bank_lens": (i32, [vp, vp]),
        "span_b200_tone_gen_bank_sync": (i32, [vp]),
        "span_b200_v29_tx_bank_create": (vp, [vp, i32, i32, i32]),
        "span_b200_v29_tx_bank_destroy": (None, [vp]),
        "span_b200_v29_tx_bank_channels": (i32, [vp]),
        "span_b200_v29_tx_bank_restart": (i32, [vp, i32, i32, i32, i32]),
        "span_b200_v29_tx_bank_power": (i32, [vp, i32, i32, f32]),
        "span_b200_v29_tx_bank_set_prbs": (i32, [vp, i32, i32, vp, C.c_uint32]),
        "span_b200_v29_tx_bank_set_bits": (i32, [vp, i32, i32, vp, i64, vp]),
        "span_b200_v29_tx_bank_tx_device": (i32, [vp, vp, i64, i32, i32, vp]),
        "span_b200_v29_tx_bank_lens": (i32, [vp, vp]),
        "span_b200_v29_tx_bank_status": (i32, [vp, vp]),
        "span_b200_v29_tx_bank_sync": (i32, [vp]),
        "span_b200_v29_tx_tables": (i32, [vp]),
        "span_b200_awgn_bank_sync": (i32, [vp]),
        "span_b200_sig_bank_create": (vp, [vp, i32, i32]),
        "span_b200_sig_bank_destroy": (None, [vp]),
        "span_b200_sig_bank_channels": (i32, [vp]),
        "span_b200_sig_bank_init": (i32, [vp, i32, i32, i32]),
        "span_b200_sig_bank_set_mode": (i32, [vp, i32, i32, i32]),
        "span_b200_sig_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_sig_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_sig_bank_events": (i64, [vp, vp, i64]),
        "span_b200_sig_bank_channel_state": (i32, [vp, i32, vp]),
        "span_b200_mct_bank_create": (vp, [vp, i32, i32]),
        "span_b200_mct_bank_destroy": (None, [vp]),
        "span_b200_mct_bank_channels": (i32, [vp]),
        "span_b200_mct_bank_init": (i32, [vp, i32, i32, i32]),
        "span_b200_mct_bank_rx_device": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_mct_bank_rx_host": (i32, [vp, vp, i64, i32, vp]),
        "span_b200_mct_bank_events": (i64, [vp, vp, i64]),
        "span_b200_mct_bank_get": (i32, [vp, i32, i32, vp]),
        "span_b200_mct_bank_channel_state": (i32, [vp, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


def _err():
    return lib().span_b200_last_error().decode(errors="replace")


class Context:
    """One CUDA device (span_b200_ctx_create)."""

    def __init__(self, device=-1):
        self.h = lib().span_b200_ctx_create(device)
        if not self.h:
            raise EngineError(_err())

    @property
    def device(self):
        return lib().span_b200_ctx_device(self.h)

    @property
    def sm_count(self):
        return lib().span_b200_ctx_sm_count(self.h)

    def close(self):
        if self.h:
            lib().span_b200_ctx_destroy(self.h)
            self.h = None

    @property
    def numa_node(self):
        return lib().span_b200_ctx_numa_node(self.h)

    def host_alloc(self, shape, dtype):
        """Pinned host memory on the GPU's NUMA node (span_b200_host_alloc) as a numpy array; free with host_free()."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = lib().span_b200_host_alloc(self.h, n)
        if not p:
            raise EngineError(_err())
        buf = (C.c_char * n).from_address(p)
        arr = np.frombuffer(buf, dtype=dt).reshape(shape)
        self._host = getattr(self, "_host", {})
        self._host[arr.ctypes.data] = p
        return arr

    def host_free(self, arr):
        p = getattr(self, "_host", {}).pop(arr.ctypes.data, None)
        if p:
            lib().span_b200_host_free(self.h, p)

    # raw Goertzel bank -------------------------------------------------------------------
    def goertzel_blocks(self, fac, block_len, d_amp_ptr, stride, channels, samples, d_out_ptr, out_capacity, stream=None, d_energy_ptr=None):
        fac = np.ascontiguousarray(fac, dtype=np.float32)
        if d_energy_ptr is not None:
            rc = lib().span_b200_goertzel_blocks_energy_device(self.h, fac.ctypes.data, len(fac), block_len, d_amp_ptr, stride,
                                                               channels, samples, d_out_ptr, out_capacity, d_energy_ptr, stream)
        else:
            rc = lib().span_b200_goertzel_blocks_device(self.h, fac.ctypes.data, len(fac), block_len, d_amp_ptr, stride,
                                                        channels, samples, d_out_ptr, out_capacity, stream)
        if rc < 0:
            raise EngineError(_err())
        return rc


class Bank:
    """A group of channels of one detector type (span_b200_*_bank_create)."""

    def __init__(self, ctx, handle):
        if not handle:
            raise EngineError(_err())
        self.ctx = ctx
        self.h = handle

    # constructors ------------------------------------------------------------------------
    @classmethod
    def dtmf(cls, ctx, channels):
        return cls(ctx, lib().span_b200_dtmf_bank_create(ctx.h, channels))

    @classmethod
    def bell_mf(cls, ctx, channels):
        return cls(ctx, lib().span_b200_bell_mf_bank_create(ctx.h, channels))

    @classmethod
    def r2_mf(cls, ctx, channels, fwd=True):
        return cls(ctx, lib().span_b200_r2_mf_bank_create(ctx.h, channels, int(bool(fwd))))

    @classmethod
    def super_tone(cls, ctx, channels, tones, want_segments=False):
        """tones: list of tones, each a list of (f1_hz, f2_hz, min_ms, max_ms)."""
        segs = np.asarray([len(t) for t in tones], dtype=np.int32)
        flat = np.asarray([x for t in tones for e in t for x in e], dtype=np.int32)
        if flat.size == 0:
            flat = np.zeros(4, dtype=np.int32)
        d = SuperToneDesc()
        d.tones = len(tones)
        d.tone_segs = segs.ctypes.data_as(C.POINTER(C.c_int32))
        d.elements = flat.ctypes.data_as(C.POINTER(C.c_int32))
        return cls(ctx, lib().span_b200_super_tone_bank_create(ctx.h, channels, C.byref(d), int(bool(want_segments))))

    # properties --------------------------------------------------------------------------
    @property
    def channels(self):
        return lib().span_b200_bank_channels(self.h)

    @property
    def block_len(self):
        return lib().span_b200_bank_block_len(self.h)

    @property
    def bins(self):
        return lib().span_b200_bank_bins(self.h)

    @property
    def last_path(self):
        return lib().span_b200_bank_last_path(self.h).decode()

    @property
    def last_launches(self):
        return lib().span_b200_bank_last_launches(self.h)

    @property
    def last_blocks(self):
        return lib().span_b200_bank_last_blocks(self.h)

    def coefficients(self):
        out = np.zeros(64, dtype=np.float32)
        n = lib().span_b200_bank_coefficients(self.h, out.ctypes.data, 64)
        return out[:n]

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    # control plane -----------------------------------------------------------------------
    def reset(self, first=0, count=None):
        self._ck(lib().span_b200_bank_reset(self.h, first, self.channels - first if count is None else count))

    def dtmf_parms(self, filter_dialtone=-1, twist=-1.0, reverse_twist=-1.0, threshold=-99.0, first=0, count=None):
        self._ck(lib().span_b200_dtmf_bank_parms(self.h, first, self.channels - first if count is None else count,
                                                 filter_dialtone, twist, reverse_twist, threshold))

    def dtmf_realtime(self, on=True, first=0, count=None):
        self._ck(lib().span_b200_dtmf_bank_realtime(self.h, first, self.channels - first if count is None else count, int(on)))

    def dtmf_fillin(self, first=0, count=None):
        self._ck(lib().span_b200_dtmf_bank_fillin(self.h, first, self.channels - first if count is None else count))

    def status(self, first=0, count=None):
        count = self.channels - first if count is None else count
        out = np.zeros(count, dtype=np.int32)
        self._ck(lib().span_b200_bank_status(self.h, first, count, out.ctypes.data))
        return out

    def tune(self, what, value):
        self._ck(lib().span_b200_bank_tune(self.h, what, value))

    def set_event_capacity(self, n):
        self._ck(lib().span_b200_bank_set_event_capacity(self.h, n))

    # processing --------------------------------------------------------------------------
    def rx_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_bank_rx_device(self.h, d_ptr, stride, samples, stream))

    def rx_host(self, amp, stream=None, samples=None):
        """amp: int16 numpy array [channels, n] (rows contiguous) or a (ptr, stride) pair with samples."""
        if isinstance(amp, tuple):
            ptr, stride = amp
            self._ck(lib().span_b200_bank_rx_host(self.h, ptr, stride, samples, stream))
            return
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        self._ck(lib().span_b200_bank_rx_host(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], stream))

    def rx_device_g711(self, d_ptr, stride, samples, alaw=False, stream=None):
        self._ck(lib().span_b200_bank_rx_device_g711(self.h, d_ptr, stride, samples, int(alaw), stream))

    def rx_host_g711(self, data, alaw=False, stream=None, samples=None):
        """data: uint8 numpy array [channels, n] or a (ptr, stride) pair with samples."""
        if isinstance(data, tuple):
            ptr, stride = data
            self._ck(lib().span_b200_bank_rx_host_g711(self.h, ptr, stride, samples, int(alaw), stream))
            return
        assert data.dtype == np.uint8 and data.ndim == 2 and data.shape[0] == self.channels and data.strides[1] == 1
        self._ck(lib().span_b200_bank_rx_host_g711(self.h, data.ctypes.data, data.strides[0], data.shape[1], int(alaw), stream))

    def event_count(self):
        ov = C.c_int(0)
        n = lib().span_b200_bank_event_count(self.h, C.byref(ov))
        if n < 0:
            raise EngineError(_err())
        return n, bool(ov.value)

    def events(self, out=None):
        n, overflow = self.event_count()
        if overflow:
            raise EngineError("event buffer overflow")
        if out is None:
            out = np.zeros(n, dtype=EVENT_DTYPE)
        got = lib().span_b200_bank_events(self.h, out.ctypes.data, len(out))
        if got < 0:
            raise EngineError(_err())
        return out[:got]

    def events_to_device(self, d_ptr, max_events, stream=None):
        n = lib().span_b200_bank_events_to_device(self.h, d_ptr, max_events, stream)
        if n < 0:
            raise EngineError(_err())
        return n

    # wire records and the multi-GPU gather ------------------------------------------------
    def set_wire(self, on=True, channel_base=0):
        self._ck(lib().span_b200_bank_set_wire(self.h, int(on), channel_base))

    def events_wire(self, out=None):
        n, overflow = self.event_count()
        if overflow:
            raise EngineError("event buffer overflow")
        if out is None:
            out = np.zeros(n, dtype=WIRE_DTYPE)
        got = lib().span_b200_bank_events_wire(self.h, out.ctypes.data, len(out))
        if got < 0:
            raise EngineError(_err())
        return out[:got]

    def attach_comm(self, comm, root=0):
        self._ck(lib().span_b200_bank_attach_comm(self.h, comm.h, root))
        self._comm = comm

    def gather_begin(self):
        self._ck(lib().span_b200_bank_gather_begin(self.h))

    def gather_end(self):
        """Returns (total records over all ranks, per-rank counts)."""
        counts = np.zeros(self._comm.nranks, dtype=np.int64)
        total = lib().span_b200_bank_gather_end(self.h, counts.ctypes.data)
        if total < 0:
            raise EngineError(_err())
        return total, counts

    def gathered(self):
        """Root: (total, device pointer) of the last completed gather (waits for it)."""
        p = C.c_void_p(0)
        total = lib().span_b200_bank_gathered(self.h, C.byref(p))
        if total < 0:
            raise EngineError(_err())
        return total, p.value

    def gathered_host(self, out=None):
        total, _ = self.gathered()
        if out is None:
            out = np.zeros(total, dtype=WIRE_DTYPE)
        got = lib().span_b200_bank_gathered_host(self.h, out.ctypes.data, len(out))
        if got < 0:
            raise EngineError(_err())
        return out[:got]

    def kernel_ms(self):
        """(summed filter-bank kernel time in ms, launches) since the last query; needs tune(4, 1)."""
        k = C.c_int(0)
        ms = lib().span_b200_bank_kernel_ms(self.h, C.byref(k))
        if ms < 0:
            raise EngineError(_err())
        return ms, k.value

    def events_device_ptr(self):
        return lib().span_b200_bank_events_device(self.h)

    def block_codes(self):
        nb = self.last_blocks
        out = np.zeros(nb * self.channels, dtype=np.uint16)
        n = lib().span_b200_bank_block_codes(self.h, out.ctypes.data, out.size)
        if n < 0:
            raise EngineError(_err())
        return out[:n].reshape(nb, self.channels)

    def close(self):
        if self.h:
            lib().span_b200_bank_destroy(self.h)
            self.h = None


class Comm:
    """One rank of a multi-GPU job (span_b200_comm_create): an NCCL communicator the library resolves at run time."""

    @staticmethod
    def unique_id():
        buf = np.zeros(128, dtype=np.uint8)
        if lib().span_b200_comm_unique_id(buf.ctypes.data) != 0:
            raise EngineError(_err())
        return buf

    def __init__(self, ctx, unique_id, nranks, rank, max_ctas=0):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        self.h = lib().span_b200_comm_create(ctx.h, uid.ctypes.data, nranks, rank, max_ctas)
        if not self.h:
            raise EngineError(_err())
        self.nranks = nranks
        self.rank = rank

    def sync(self):
        if lib().span_b200_comm_sync(self.h) != 0:
            raise EngineError(_err())

    def set_transport(self, which):
        """'peer_copy' (the root's copy engines pull the records over NVLink) or 'nccl' (exact-count send / recv)."""
        if lib().span_b200_comm_set_transport(self.h, {"nccl": 0, "peer_copy": 1}[which]) != 0:
            raise EngineError(_err())

    @property
    def transport(self):
        return ("nccl", "peer_copy")[lib().span_b200_comm_transport(self.h)]

    def close(self):
        if self.h:
            lib().span_b200_comm_destroy(self.h)
            self.h = None


def wire_unpack(w):
    """Wire records -> (channel, block, kind, a, b, c) int64 columns."""
    bk = w["block_kind"].astype(np.int64)
    kind = bk >> 14
    kind = np.where(kind == 3, 5, kind)
    return (w["channel"].astype(np.int64), bk & 0x3FFF, kind, w["a"].astype(np.int64), w["b"].astype(np.int64), w["c"].astype(np.int64))


V29_SYMBOL_DTYPE = np.dtype([("re", "<f4"), ("im", "<f4"), ("tre", "<f4"), ("tim", "<f4"), ("state", "<i4"), ("bit_pos", "<i4")])


class V29Bank:
    """N V.29 receivers (span_b200_v29_bank_create)."""
    PREFIX = "span_b200_v29_bank_"
    INFO = 10

    def __init__(self, ctx, channels, bit_rate=9600, want_symbols=False):
        self.ctx = ctx
        self.h = self._fn("create")(ctx.h, channels, bit_rate, int(want_symbols))
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _fn(self, name):
        return getattr(lib(), self.PREFIX + name)

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def restart(self, bit_rate, first=0, count=None, mode=0):
        """mode: V.29 and V.27ter old_train (0/1); V.17 short_train (0/1/2)."""
        n = self.channels - first if count is None else count
        if self.PREFIX == "span_b200_v29_bank_":
            self._ck(lib().span_b200_v29_bank_restart_ex(self.h, first, n, bit_rate, mode))
        else:
            self._ck(self._fn("restart")(self.h, first, n, bit_rate, mode))

    def fillin(self, samples, first=0, count=None):
        self._ck(self._fn("fillin")(self.h, first, self.channels - first if count is None else count, samples))

    def set_signal_cutoff(self, cutoff, first=0, count=None):
        self._ck(self._fn("set_signal_cutoff")(self.h, first, self.channels - first if count is None else count, cutoff))

    def rx_device(self, d_ptr, stride, samples, stream=None):
        self._ck(self._fn("rx_device")(self.h, d_ptr, stride, samples, stream))

    def rx_host(self, amp, stream=None, samples=None):
        """amp: int16 numpy array [channels, n] (rows contiguous) or a (ptr, stride) pair with samples."""
        if isinstance(amp, tuple):
            ptr, stride = amp
            self._ck(self._fn("rx_host")(self.h, ptr, stride, samples, stream))
            return
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        self._ck(self._fn("rx_host")(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], stream))

    def counts(self):
        nb = np.zeros(self.channels, dtype=np.int32)
        ns = np.zeros(self.channels, dtype=np.int32)
        self._ck(self._fn("counts")(self.h, nb.ctypes.data, ns.ctypes.data))
        return nb, ns

    def output_packed(self, words=None, status=None):
        """The bulk read-back: (words uint32 [C][W], nbits [C], status int32 [C][S][2], nstatus [C]); pass pre-allocated
        (pinned) arrays to avoid the allocation."""
        nb = np.zeros(self.channels, dtype=np.int32)
        ns = np.zeros(self.channels, dtype=np.int32)
        if words is None:
            cnt, _ = self.counts()
            words = np.zeros((self.channels, (int(cnt.max()) + 31) // 32 + 1), dtype=np.uint32)
        if status is None:
            status = np.zeros((self.channels, 64, 2), dtype=np.int32)
        mx = self._fn("output_packed")(self.h, words.ctypes.data, words.shape[1], nb.ctypes.data, status.ctypes.data, status.shape[1], ns.ctypes.data)
        if mx < 0:
            raise EngineError(_err())
        return words, nb, status, ns

    def bits(self, channel, cap=1 << 22):
        out = np.zeros(cap, dtype=np.int8)
        n = self._fn("bits")(self.h, channel, out.ctypes.data, cap)
        if n < 0:
            raise EngineError(_err())
        return out[:n]

    def symbols(self, channel, cap=1 << 20):
        out = np.zeros(cap, dtype=V29_SYMBOL_DTYPE)
        n = self._fn("symbols")(self.h, channel, out.ctypes.data, cap)
        if n < 0:
            raise EngineError(_err())
        return out[:n]

    def channel_state(self, channel):
        eq = np.zeros(66, dtype=np.float32)
        info = np.zeros(self.INFO, dtype=np.int32)
        self._ck(self._fn("channel_state")(self.h, channel, eq.ctypes.data, info.ctypes.data))
        return eq, info

    def close(self):
        if self.h:
            self._fn("destroy")(self.h)
            self.h = None


class V17Bank(V29Bank):
    """N V.17 receivers (span_b200_v17_bank_create)."""
    PREFIX = "span_b200_v17_bank_"
    INFO = 12

    def __init__(self, ctx, channels, bit_rate=14400, want_symbols=False):
        V29Bank.__init__(self, ctx, channels, bit_rate, want_symbols)


class V27terBank(V29Bank):
    """N V.27ter receivers (span_b200_v27ter_bank_create)."""
    PREFIX = "span_b200_v27ter_bank_"
    INFO = 12

    def __init__(self, ctx, channels, bit_rate=4800, want_symbols=False):
        V29Bank.__init__(self, ctx, channels, bit_rate, want_symbols)


class FskSpec(C.Structure):
    """span_b200_fsk_spec_t (= the reference's fsk_spec_t)."""
    _fields_ = [("name", C.c_char_p), ("freq_zero", C.c_int), ("freq_one", C.c_int), ("tx_level", C.c_int),
                ("min_level", C.c_int), ("baud_rate", C.c_int)]


def fsk_preset(which):
    p = lib().span_b200_fsk_preset(which)
    if not p:
        raise EngineError("no such FSK preset")
    return C.cast(p, C.POINTER(FskSpec)).contents


class FskBank:
    """N FSK receivers (span_b200_fsk_bank_create).  spec: a preset index or an FskSpec."""

    def __init__(self, ctx, channels, spec=1, framing_mode=1):
        self.ctx = ctx
        self._spec = fsk_preset(spec) if isinstance(spec, int) else spec
        self.h = lib().span_b200_fsk_bank_create(ctx.h, channels, C.addressof(self._spec), framing_mode)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def _range(self, first, count):
        return first, (self.channels - first if count is None else count)

    def restart(self, spec, framing_mode, first=0, count=None):
        sp = fsk_preset(spec) if isinstance(spec, int) else spec
        f, n = self._range(first, count)
        self._ck(lib().span_b200_fsk_bank_restart(self.h, f, n, C.addressof(sp), framing_mode))

    def set_signal_cutoff(self, cutoff, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_fsk_bank_set_signal_cutoff(self.h, f, n, cutoff))

    def set_frame_parameters(self, data_bits, parity, stop_bits, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_fsk_bank_set_frame_parameters(self.h, f, n, data_bits, parity, stop_bits))

    def fillin(self, samples, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_fsk_bank_fillin(self.h, f, n, samples))

    def rx_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_fsk_bank_rx_device(self.h, d_ptr, stride, samples, stream))

    def rx_host(self, amp, stream=None):
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        self._ck(lib().span_b200_fsk_bank_rx_host(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], stream))

    def counts(self):
        n = np.zeros(self.channels, dtype=np.int32)
        self._ck(lib().span_b200_fsk_bank_counts(self.h, n.ctypes.data))
        return n

    def output(self, channel, cap=1 << 22):
        out = np.zeros(cap, dtype=np.int16)
        n = lib().span_b200_fsk_bank_output(self.h, channel, out.ctypes.data, cap)
        if n < 0:
            raise EngineError(_err())
        return out[:n]

    def errors(self, channel, reset=False):
        p = C.c_int32(0)
        f = C.c_int32(0)
        self._ck(lib().span_b200_fsk_bank_errors(self.h, channel, C.addressof(p), C.addressof(f), int(reset)))
        return p.value, f.value

    def signal_power(self, channel):
        return lib().span_b200_fsk_bank_signal_power(self.h, channel)

    def channel_state(self, channel):
        info = np.zeros(28, dtype=np.int32)
        win = np.zeros((2, 128, 2), dtype=np.int32)
        self._ck(lib().span_b200_fsk_bank_channel_state(self.h, channel, info.ctypes.data, win.ctypes.data))
        return info, win

    def close(self):
        if self.h:
            lib().span_b200_fsk_bank_destroy(self.h)
            self.h = None


MCT_EVENT_DTYPE = np.dtype([("channel", "<i4"), ("tone", "<i4"), ("level", "<i4")])


class MctBank:
    """N modem connect tone detectors (span_b200_mct_bank_create); tone_type as the reference's MODEM_CONNECT_TONES_*."""

    def __init__(self, ctx, channels, tone_type):
        self.ctx = ctx
        self.h = lib().span_b200_mct_bank_create(ctx.h, channels, tone_type)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def init(self, tone_type, first=0, count=None):
        self._ck(lib().span_b200_mct_bank_init(self.h, first, self.channels - first if count is None else count, tone_type))

    def rx_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_mct_bank_rx_device(self.h, d_ptr, stride, samples, stream))

    def rx_host(self, amp, stream=None):
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        self._ck(lib().span_b200_mct_bank_rx_host(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], stream))

    def events(self):
        n = self._ck(lib().span_b200_mct_bank_events(self.h, None, 0))
        ev = np.zeros(n, dtype=MCT_EVENT_DTYPE)
        if n:
            self._ck(lib().span_b200_mct_bank_events(self.h, ev.ctypes.data, n))
        return ev

    def get(self, first=0, count=None):
        n = self.channels - first if count is None else count
        hits = np.zeros(n, dtype=np.int32)
        self._ck(lib().span_b200_mct_bank_get(self.h, first, n, hits.ctypes.data))
        return hits

    def channel_state(self, channel):
        info = np.zeros(17, dtype=np.int32)
        fsk = np.zeros(28, dtype=np.int32)
        self._ck(lib().span_b200_mct_bank_channel_state(self.h, channel, info.ctypes.data, fsk.ctypes.data))
        return info, fsk

    def close(self):
        if self.h:
            lib().span_b200_mct_bank_destroy(self.h)
            self.h = None


SIG_EVENT_DTYPE = np.dtype([("channel", "<i4"), ("signalling_state", "<i4"), ("duration", "<i4")])


class SigBank:
    """N in-band signalling tone receivers (span_b200_sig_bank_create); tone_type 1 = 2280 Hz, 2 = 2600 Hz, 3 = 2400 + 2600 Hz.
    rx_* rewrite the audio in place, as sig_tone_rx() does."""

    def __init__(self, ctx, channels, tone_type):
        self.ctx = ctx
        self.h = lib().span_b200_sig_bank_create(ctx.h, channels, tone_type)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def init(self, tone_type, first=0, count=None):
        self._ck(lib().span_b200_sig_bank_init(self.h, first, self.channels - first if count is None else count, tone_type))

    def set_mode(self, mode, first=0, count=None):
        self._ck(lib().span_b200_sig_bank_set_mode(self.h, first, self.channels - first if count is None else count, mode))

    def rx_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_sig_bank_rx_device(self.h, d_ptr, stride, samples, stream))

    def rx_host(self, amp, stream=None):
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        assert amp.flags["WRITEABLE"]
        self._ck(lib().span_b200_sig_bank_rx_host(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], stream))

    def events(self):
        n = self._ck(lib().span_b200_sig_bank_events(self.h, None, 0))
        ev = np.zeros(n, dtype=SIG_EVENT_DTYPE)
        if n:
            self._ck(lib().span_b200_sig_bank_events(self.h, ev.ctypes.data, n))
        return ev

    def channel_state(self, channel):
        info = np.zeros(31, dtype=np.int32)
        self._ck(lib().span_b200_sig_bank_channel_state(self.h, channel, info.ctypes.data))
        return info

    def close(self):
        if self.h:
            lib().span_b200_sig_bank_destroy(self.h)
            self.h = None


class DtmfTxBank:
    """N DTMF transmitters (span_b200_dtmf_tx_bank_create): dtmf_tx_init/put/set_level/set_timing/dtmf_tx per channel."""

    def __init__(self, ctx, channels):
        self.ctx = ctx
        self.h = lib().span_b200_dtmf_tx_bank_create(ctx.h, channels)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def _range(self, first, count):
        return first, (self.channels - first if count is None else count)

    def init(self, first=0, count=None):
        self._ck(lib().span_b200_dtmf_tx_bank_init(self.h, *self._range(first, count)))

    def set_level(self, level, twist, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_dtmf_tx_bank_set_level(self.h, f, n, level, twist))

    def set_timing(self, on_time, off_time, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_dtmf_tx_bank_set_timing(self.h, f, n, on_time, off_time))

    def put(self, digits, first=0, count=None):
        f, n = self._range(first, count)
        return self._ck(lib().span_b200_dtmf_tx_bank_put(self.h, f, n, digits.encode(), -1))

    def put_each(self, strings, first=0):
        """One digit string per channel, channels [first, first + len(strings))."""
        n = len(strings)
        stride = max(1, max(len(x) for x in strings))
        buf = np.zeros((n, stride), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.int32)
        for i, x in enumerate(strings):
            b = np.frombuffer(x.encode(), dtype=np.uint8)
            buf[i, :len(b)] = b
            lens[i] = len(b)
        return self._ck(lib().span_b200_dtmf_tx_bank_put_each(self.h, first, n, buf.ctypes.data, stride, lens.ctypes.data))

    def tx_device(self, d_ptr, stride, max_samples, zero_fill=False, stream=None):
        self._ck(lib().span_b200_dtmf_tx_bank_tx_device(self.h, d_ptr, stride, max_samples, int(zero_fill), stream))

    def tx_host(self, amp, zero_fill=False):
        assert amp.dtype == np.int16 and amp.ndim == 2 and amp.shape[0] == self.channels and amp.strides[1] == 2
        self._ck(lib().span_b200_dtmf_tx_bank_tx_host(self.h, amp.ctypes.data, amp.strides[0] // 2, amp.shape[1], int(zero_fill)))

    def lens(self):
        n = np.zeros(self.channels, dtype=np.int32)
        self._ck(lib().span_b200_dtmf_tx_bank_lens(self.h, n.ctypes.data))
        return n

    def sync(self):
        self._ck(lib().span_b200_dtmf_tx_bank_sync(self.h))

    def close(self):
        if self.h:
            lib().span_b200_dtmf_tx_bank_destroy(self.h)
            self.h = None


class AwgnBank:
    """N noise sources (span_b200_awgn_bank_create): awgn_init_dbm0(seed) / awgn() per channel."""

    def __init__(self, ctx, channels, level_dbm0, seeds=None, seed0=0):
        self.ctx = ctx
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.int32)
            assert len(self._seeds) == channels
            sp = self._seeds.ctypes.data
        self.h = lib().span_b200_awgn_bank_create(ctx.h, channels, sp, seed0, level_dbm0)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def init(self, level, seeds=None, seed0=0, first=0, count=None, dbov=False):
        n = self.channels - first if count is None else count
        sp = None
        if seeds is not None:
            s = np.ascontiguousarray(seeds, dtype=np.int32)
            assert len(s) == n
            sp = s.ctypes.data
        fn = lib().span_b200_awgn_bank_init_dbov if dbov else lib().span_b200_awgn_bank_init_dbm0
        self._ck(fn(self.h, first, n, sp, seed0, level))

    def add_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_awgn_bank_add_device(self.h, d_ptr, stride, samples, stream))

    def fill_device(self, d_ptr, stride, samples, stream=None):
        self._ck(lib().span_b200_awgn_bank_fill_device(self.h, d_ptr, stride, samples, stream))

    def sync(self):
        self._ck(lib().span_b200_awgn_bank_sync(self.h))

    def close(self):
        if self.h:
            lib().span_b200_awgn_bank_destroy(self.h)
            self.h = None


class ToneGenBank:
    """N cadenced tone generators (span_b200_tone_gen_bank_create): tone_gen_descriptor_init / tone_gen_init / tone_gen."""

    def __init__(self, ctx, channels):
        self.ctx = ctx
        self.h = lib().span_b200_tone_gen_bank_create(ctx.h, channels)
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def init(self, desc, first=0, count=None):
        """desc = (f1, l1, f2, l2, d1, d2, d3, d4, repeat) for every channel of the range."""
        d = np.asarray(desc, dtype=np.int32).reshape(9)
        n = self.channels - first if count is None else count
        self._ck(lib().span_b200_tone_gen_bank_init(self.h, first, n, d.ctypes.data))

    def init_each(self, descs, first=0):
        d = np.ascontiguousarray(descs, dtype=np.int32).reshape(-1, 9)
        self._ck(lib().span_b200_tone_gen_bank_init_each(self.h, first, len(d), d.ctypes.data))

    def tx_device(self, d_ptr, stride, max_samples, zero_fill=False, stream=None):
        self._ck(lib().span_b200_tone_gen_bank_tx_device(self.h, d_ptr, stride, max_samples, int(zero_fill), stream))

    def lens(self):
        n = np.zeros(self.channels, dtype=np.int32)
        self._ck(lib().span_b200_tone_gen_bank_lens(self.h, n.ctypes.data))
        return n

    def sync(self):
        self._ck(lib().span_b200_tone_gen_bank_sync(self.h))

    def close(self):
        if self.h:
            lib().span_b200_tone_gen_bank_destroy(self.h)
            self.h = None


class V29TxBank:
    """N V.29 transmitters (span_b200_v29_tx_bank_create): v29_tx_init / restart / power / v29_tx per channel."""

    def __init__(self, ctx, channels, bit_rate=9600, tep=False):
        self.ctx = ctx
        self.h = lib().span_b200_v29_tx_bank_create(ctx.h, channels, bit_rate, int(tep))
        if not self.h:
            raise EngineError(_err())
        self.channels = channels

    def _ck(self, rc):
        if rc < 0:
            raise EngineError(_err())
        return rc

    def _range(self, first, count):
        return first, (self.channels - first if count is None else count)

    def restart(self, bit_rate, tep=False, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_v29_tx_bank_restart(self.h, f, n, bit_rate, int(tep)))

    def power(self, dbm0, first=0, count=None):
        f, n = self._range(first, count)
        self._ck(lib().span_b200_v29_tx_bank_power(self.h, f, n, dbm0))

    def set_prbs(self, seeds=None, seed0=1, first=0, count=None):
        f, n = self._range(first, count)
        sp = None
        if seeds is not None:
            s = np.ascontiguousarray(seeds, dtype=np.uint32)
            assert len(s) == n
            sp = s.ctypes.data
        self._ck(lib().span_b200_v29_tx_bank_set_prbs(self.h, f, n, sp, seed0))

    def set_bits(self, bits, nbits, first=0):
        """bits: uint8 [count][bytes], LSB first; nbits: bits per channel."""
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        nb = np.ascontiguousarray(nbits, dtype=np.int32)
        assert b.ndim == 2 and len(nb) == b.shape[0]
        self._ck(lib().span_b200_v29_tx_bank_set_bits(self.h, first, b.shape[0], b.ctypes.data, b.shape[1], nb.ctypes.data))

    def tx_device(self, d_ptr, stride, max_samples, zero_fill=False, stream=None):
        self._ck(lib().span_b200_v29_tx_bank_tx_device(self.h, d_ptr, stride, max_samples, int(zero_fill), stream))

    def lens(self):
        n = np.zeros(self.channels, dtype=np.int32)
        self._ck(lib().span_b200_v29_tx_bank_lens(self.h, n.ctypes.data))
        return n

    def status(self):
        n = np.zeros(self.channels, dtype=np.int32)
        self._ck(lib().span_b200_v29_tx_bank_status(self.h, n.ctypes.data))
        return n

    def sync(self):
        self._ck(lib().span_b200_v29_tx_bank_sync(self.h))

    def close(self):
        if self.h:
            lib().span_b200_v29_tx_bank_destroy(self.h)
            self.h = None


def v29_tx_tables():
    t = np.zeros((10, 9), dtype=np.float32)
    lib().span_b200_v29_tx_tables(t.ctypes.data)
    return t


def events_by_channel(ev, channels):
    """Split an event array into per-channel lists of (kind, a, b, c), keeping each channel's order."""
    out = [[] for _ in range(channels)]
    for e in ev:
        out[int(e["channel"])].append((int(e["kind"]), int(e["a"]), int(e["b"]), int(e["c"])))
    return out
