"""ctypes binding of libamps_b200.so (the C ABI declared in include/amps_b200.h).

This is plumbing for tests/ and bench.py: the product is the shared library.  There is no
fallback of any kind -- if the library is missing or no sm_100 device is usable, loading or
handle creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMPS_B200_LIB") or os.path.join(PKG_DIR, "libamps_b200.so")      # (the override: A/B runs of two builds)

TRIGGER_SYMS = 74
CAPTURE_SYMS = 3374

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


class AmpsError(RuntimeError):
    def __init__(self, status: int, detail: str):
        super().__init__(f"libamps_b200 status {status}: {detail}")
        self.status = status


class ReccWords(C.Structure):
    _fields_ = [
        ("dcc", C.c_uint8 * 7), ("dcc_errs", C.c_uint8),
        ("words", (C.c_uint8 * 240) * 7),
        ("errs", C.c_uint16 * 7),
        ("valid", C.c_uint8 * 7), ("valid_repeat", C.c_uint8 * 7),
        ("F", C.c_uint8), ("NAWC", C.c_uint8), ("T", C.c_uint8), ("S", C.c_uint8), ("E", C.c_uint8),
        ("ER", C.c_uint8), ("SCM", C.c_uint8), ("pad0", C.c_uint8),
        ("MIN1", C.c_uint32),
        ("B_F", C.c_uint8), ("B_NAWC", C.c_uint8), ("MSG_TYPE", C.c_uint8), ("ORDQ", C.c_uint8),
        ("ORDER", C.c_uint8), ("LT", C.c_uint8), ("EP", C.c_uint8), ("SCM4", C.c_uint8), ("MPCI", C.c_uint8),
        ("SDCC1", C.c_uint8), ("SDCC2", C.c_uint8), ("pad1", C.c_uint8),
        ("MIN2", C.c_uint16), ("pad2", C.c_uint16),
        ("word_c_serial", C.c_uint32),
        ("kind", C.c_int32),
        ("esn", C.c_uint32),
        ("min", C.c_char * 12),
        ("dialed", C.c_char * 36),
    ]


class Burst(C.Structure):
    _fields_ = [
        ("sample_index", C.c_uint64), ("demod_index", C.c_uint64),
        ("corr", C.c_float), ("run_length", C.c_uint32),
        ("symbols", C.c_uint8 * CAPTURE_SYMS), ("pad", C.c_uint8 * 2),
        ("decoded", ReccWords),
    ]

    def symbols_np(self) -> np.ndarray:
        return np.frombuffer(bytes(self.symbols), np.uint8).copy()


class ReccIqParams(C.Structure):
    _fields_ = [
        ("samp_rate", C.c_double), ("center_freq", C.c_double), ("device", C.c_int),
        ("max_samples", C.c_uint32), ("max_bursts", C.c_uint32), ("flags", C.c_uint32),
        ("lpf_taps", f32p), ("n_lpf_taps", C.c_uint32), ("sc16_scale", C.c_float),
    ]


class FwdParams(C.Structure):
    _fields_ = [
        ("samp_rate", C.c_double), ("symrate", C.c_double), ("max_deviation", C.c_double),
        ("device", C.c_int), ("ncarriers", C.c_int),
        ("carrier_freq", C.c_double * 3), ("lpf_transition", C.c_double * 3),
        ("out_scale", C.c_float), ("max_samples", C.c_uint32),
    ]


class FwdVoiceParams(C.Structure):
    _fields_ = [("carrier_gated", C.c_int), ("carrier_open", C.c_int), ("audio_rate", C.c_double), ("max_dev", C.c_double),
                ("tau", C.c_double), ("sat_freq", C.c_double), ("sat_amp", C.c_double)]


BURST_CB = C.CFUNCTYPE(None, C.POINTER(Burst), C.c_void_p)
BLOB_CB = C.CFUNCTYPE(None, u8p, C.c_void_p)
BATCH_BURST_CB = C.CFUNCTYPE(None, C.c_int, C.POINTER(Burst), C.c_void_p)

RX_DUMP_BASEBAND = 1
RX_TIME_KERNELS = 2
RX_TIMING_MM = 4
RX_INPUT_SC16 = 8
RX_FUSED_SEARCH = 16

# every symbol include/amps_b200.h declares
EXPORTS = [
    "amps_b200_version", "amps_b200_strerror", "amps_b200_last_error", "amps_b200_device_count", "amps_b200_abi_sizes",
    "amps_recc_iq_create", "amps_recc_iq_destroy", "amps_recc_iq_reset", "amps_recc_iq_work",
    "amps_recc_iq_submit_dev", "amps_recc_iq_work_sc16", "amps_recc_iq_submit_sc16_dev", "amps_recc_iq_collect", "amps_recc_iq_peek", "amps_recc_iq_consume", "amps_recc_iq_poll", "amps_recc_iq_granularity", "amps_recc_iq_read_demod",
    "amps_recc_iq_read_baseband", "amps_recc_iq_stats", "amps_recc_iq_front_times", "amps_recc_iq_get_taps", "amps_recc_iq_debug_prof", "amps_b200_debug_deal",
    "amps_recc_iq_batch_create", "amps_recc_iq_batch_destroy", "amps_recc_iq_batch_size", "amps_recc_iq_batch_submit_dev",
    "amps_recc_iq_batch_work_shared", "amps_recc_iq_batch_front_times", "amps_recc_iq_batch_stats",
    "amps_recc_decode_create", "amps_recc_decode_destroy", "amps_recc_decode_burst", "amps_recc_decode_bursts",
    "amps_recc_create", "amps_recc_destroy", "amps_recc_work", "amps_recc_work_chunks",
    "amps_focc_create", "amps_focc_destroy", "amps_focc_work", "amps_focc_generate", "amps_focc_generate_dev",
    "amps_focc_push_words", "amps_focc_set_busy_idle", "amps_focc_generate_bits", "amps_focc_generate_bits_dev",
    "amps_fvc_create", "amps_fvc_destroy", "amps_fvc_push_words", "amps_fvc_work", "amps_fvc_work_bits",
    "amps_fwd_create", "amps_fwd_destroy", "amps_fwd_reset", "amps_fwd_work", "amps_fwd_submit_dev",
    "amps_fwd_interp", "amps_fwd_get_taps", "amps_fwd_work_bits", "amps_fwd_submit_bits_dev",
    "amps_fwd_enable_voice", "amps_fwd_work_voice", "amps_fwd_submit_voice_dev",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  gr_amps_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.amps_b200_strerror.restype = C.c_char_p
    L.amps_b200_last_error.restype = C.c_char_p
    L.amps_recc_iq_create.argtypes = [C.POINTER(ReccIqParams), C.POINTER(C.c_void_p)]
    L.amps_recc_iq_destroy.argtypes = [C.c_void_p]
    L.amps_recc_iq_reset.argtypes = [C.c_void_p]
    L.amps_recc_iq_work.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, BURST_CB, C.c_void_p]
    L.amps_recc_iq_submit_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.amps_recc_iq_work_sc16.argtypes = L.amps_recc_iq_work.argtypes
    L.amps_recc_iq_submit_sc16_dev.argtypes = L.amps_recc_iq_submit_dev.argtypes
    L.amps_recc_iq_collect.argtypes = [C.c_void_p, C.POINTER(Burst), C.c_int, C.POINTER(C.c_int)]
    L.amps_recc_iq_granularity.argtypes = [C.c_void_p]
    L.amps_recc_iq_peek.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Burst)), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.amps_recc_iq_consume.argtypes = [C.c_void_p, C.c_uint64]
    L.amps_recc_iq_poll.argtypes = L.amps_recc_iq_peek.argtypes
    L.amps_b200_abi_sizes.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    bs, ws = C.c_size_t(0), C.c_size_t(0)
    L.amps_b200_abi_sizes(C.byref(bs), C.byref(ws))
    if bs.value != C.sizeof(Burst) or ws.value != C.sizeof(ReccWords):
        raise ImportError("gr_amps_b200.capi struct layout (%d, %d) does not match libamps_b200.so (%d, %d)"
                          % (C.sizeof(Burst), C.sizeof(ReccWords), bs.value, ws.value))
    L.amps_recc_iq_read_demod.argtypes = [C.c_void_p, C.c_uint64, f32p, C.c_size_t]
    L.amps_recc_iq_read_baseband.argtypes = [C.c_void_p, C.c_uint64, f32p, C.c_size_t]
    L.amps_recc_iq_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
    L.amps_recc_iq_get_taps.argtypes = [C.c_void_p, f32p, C.c_int]
    L.amps_recc_iq_front_times.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
    L.amps_recc_iq_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
    L.amps_recc_iq_batch_destroy.argtypes = [C.c_void_p]
    L.amps_recc_iq_batch_size.argtypes = [C.c_void_p]
    L.amps_recc_iq_batch_submit_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p]
    L.amps_recc_iq_batch_work_shared.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, BATCH_BURST_CB, C.c_void_p]
    L.amps_recc_iq_batch_front_times.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
    L.amps_recc_iq_batch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.amps_recc_decode_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.amps_recc_decode_destroy.argtypes = [C.c_void_p]
    L.amps_recc_decode_burst.argtypes = [C.c_void_p, u8p, C.POINTER(ReccWords)]
    L.amps_recc_decode_bursts.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(ReccWords)]
    if hasattr(L, "amps_recc_create"):
        L.amps_recc_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.amps_recc_destroy.argtypes = [C.c_void_p]
        L.amps_recc_work.argtypes = [C.c_void_p, u8p, C.c_int, BLOB_CB, C.c_void_p]
        L.amps_recc_work_chunks.argtypes = [C.c_void_p, u8p, C.POINTER(C.c_int), C.c_int, BLOB_CB, C.c_void_p]
    if hasattr(L, "amps_focc_create"):
        L.amps_focc_create.argtypes = [C.c_ulong, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.amps_focc_destroy.argtypes = [C.c_void_p]
        L.amps_focc_work.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_int)]
        L.amps_focc_generate.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.amps_focc_generate_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.amps_focc_push_words.argtypes = [C.c_void_p, C.c_long, u8p, C.c_long]
        L.amps_focc_generate_bits.argtypes = [C.c_void_p, u8p, C.c_size_t]
        L.amps_focc_generate_bits_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.amps_focc_set_busy_idle.argtypes = [C.c_void_p, C.c_int]
    if hasattr(L, "amps_fvc_create"):
        L.amps_fvc_create.argtypes = [C.c_ulong, C.c_int, C.POINTER(C.c_void_p)]
        L.amps_fvc_destroy.argtypes = [C.c_void_p]
        L.amps_fvc_push_words.argtypes = [C.c_void_p, u8p, C.c_long, C.c_int, C.c_uint64]
        L.amps_fvc_work.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.amps_fvc_work_bits.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    if hasattr(L, "amps_fwd_create"):
        L.amps_fwd_create.argtypes = [C.POINTER(FwdParams), C.POINTER(C.c_void_p)]
        L.amps_fwd_destroy.argtypes = [C.c_void_p]
        L.amps_fwd_reset.argtypes = [C.c_void_p]
        L.amps_fwd_work.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, f32p]
        L.amps_fwd_submit_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p, C.c_void_p]
        L.amps_fwd_interp.argtypes = [C.c_void_p]
        L.amps_fwd_work_bits.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, f32p]
        L.amps_fwd_submit_bits_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p, C.c_void_p]
        L.amps_fwd_get_taps.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int]
    _lib = L
    return L


def check(status: int) -> None:
    if status != 0:
        raise AmpsError(status, lib().amps_b200_last_error().decode(errors="replace") or lib().amps_b200_strerror(status).decode())


class ReccIq:
    """Fused RECC receive path on IQ (amps_recc_iq_*)."""

    def __init__(self, max_samples: int, center_freq=-160e3, samp_rate=10e6, device=0, max_bursts=256,
                 dump_baseband=False, lpf_taps: np.ndarray | None = None, time_kernels=False, timing_mm=False,
                 sc16=False, sc16_scale=0.0, fused_search=False):
        self._taps = None if lpf_taps is None else np.ascontiguousarray(lpf_taps, dtype=np.float32)
        p = ReccIqParams(samp_rate, center_freq, device, max_samples, max_bursts,
                         (RX_DUMP_BASEBAND if dump_baseband else 0) | (RX_TIME_KERNELS if time_kernels else 0) | (RX_TIMING_MM if timing_mm else 0)
                         | (RX_INPUT_SC16 if sc16 else 0) | (RX_FUSED_SEARCH if fused_search else 0),
                         None if self._taps is None else self._taps.ctypes.data_as(f32p),
                         0 if self._taps is None else len(self._taps), sc16_scale)
        self.sc16 = sc16
        self.h = C.c_void_p()
        check(lib().amps_recc_iq_create(C.byref(p), C.byref(self.h)))
        self.granularity = lib().amps_recc_iq_granularity(self.h)
        self.max_bursts = max_bursts

    def close(self):
        if getattr(self, "h", None):
            lib().amps_recc_iq_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        check(lib().amps_recc_iq_reset(self.h))

    def work(self, iq: np.ndarray) -> list[Burst]:
        """Host-buffer call; iq complex64 (or its float32 view), or interleaved int16 I,Q for an sc16 handle.
        Returns the bursts published."""
        iq = np.ascontiguousarray(iq)
        if self.sc16 != (iq.dtype == np.int16):
            raise TypeError("sc16 handles take int16 I,Q; fc32 handles take complex64 / float32")
        n = iq.size if iq.dtype == np.complex64 else iq.size // 2
        got: list[Burst] = []

        def on(bp, _user):
            b = Burst()
            C.memmove(C.byref(b), bp, C.sizeof(Burst))
            got.append(b)

        cb = BURST_CB(on)
        fn = lib().amps_recc_iq_work_sc16 if self.sc16 else lib().amps_recc_iq_work
        check(fn(self.h, iq.ctypes.data_as(C.c_void_p), n, cb, None))
        return got

    def work_ptr(self, host_ptr: int, nsamples: int, cb=None):
        fn = lib().amps_recc_iq_work_sc16 if self.sc16 else lib().amps_recc_iq_work
        check(fn(self.h, C.c_void_p(host_ptr), nsamples, cb or C.cast(None, BURST_CB), None))

    def submit_dev(self, dev_ptr: int, nsamples: int, stream: int = 0):
        fn = lib().amps_recc_iq_submit_sc16_dev if self.sc16 else lib().amps_recc_iq_submit_dev
        check(fn(self.h, C.c_void_p(dev_ptr), nsamples, C.c_void_p(stream)))

    def collect(self, max_bursts: int | None = None) -> list[Burst]:
        m = max_bursts or self.max_bursts
        arr = (Burst * m)()
        n = C.c_int(0)
        check(lib().amps_recc_iq_collect(self.h, arr, m, C.byref(n)))
        return [arr[i] for i in range(n.value)]

    def peek(self):
        """Zero-copy view of the published bursts: (ring pointer, ring_len, first, count); call consume(count) after."""
        ring = C.POINTER(Burst)()
        rl, first, count = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().amps_recc_iq_peek(self.h, C.byref(ring), C.byref(rl), C.byref(first), C.byref(count)))
        return ring, rl.value, first.value, count.value

    def poll(self):
        """Like peek() but without synchronising the stream: what has been published so far."""
        ring = C.POINTER(Burst)()
        rl, first, count = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        check(lib().amps_recc_iq_poll(self.h, C.byref(ring), C.byref(rl), C.byref(first), C.byref(count)))
        return ring, rl.value, first.value, count.value

    def consume(self, count: int):
        check(lib().amps_recc_iq_consume(self.h, count))

    def read_demod(self, first: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.float32)
        check(lib().amps_recc_iq_read_demod(self.h, first, out.ctypes.data_as(f32p), n))
        return out

    def read_baseband(self, first: int, n: int) -> np.ndarray:
        out = np.zeros(2 * n, np.float32)
        check(lib().amps_recc_iq_read_baseband(self.h, first, out.ctypes.data_as(f32p), n))
        return out.view(np.complex64)

    def stats(self) -> dict:
        v = [C.c_uint64(0) for _ in range(4)]
        check(lib().amps_recc_iq_stats(self.h, *[C.byref(x) for x in v]))
        return dict(samples_in=v[0].value, demod_out=v[1].value, bursts=v[2].value, kernel_launches=v[3].value)


    def front_times_ms(self, cap: int = 256) -> np.ndarray:
        out = np.zeros(cap, np.float32)
        n = C.c_int(0)
        check(lib().amps_recc_iq_front_times(self.h, out.ctypes.data_as(f32p), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_prof(self, ctas: int) -> np.ndarray:
        out = np.zeros((ctas, 16), np.uint64)
        lib().amps_recc_iq_debug_prof.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        check(lib().amps_recc_iq_debug_prof(self.h, out.ctypes.data_as(C.c_void_p), ctas))
        return out

    def taps(self) -> np.ndarray:
        n = lib().amps_recc_iq_get_taps(self.h, None, 0)
        t = np.zeros(n, np.float32)
        lib().amps_recc_iq_get_taps(self.h, t.ctypes.data_as(f32p), n)
        return t


class ReccIqBatch:
    """K ReccIq handles of one GPU served by one front launch + one capture launch per call (amps_recc_iq_batch_*)."""

    def __init__(self, handles: list, time_kernels=False):
        self.handles = list(handles)
        arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
        self.b = C.c_void_p()
        check(lib().amps_recc_iq_batch_create(arr, len(handles), RX_TIME_KERNELS if time_kernels else 0, C.byref(self.b)))

    def close(self):
        if getattr(self, "b", None):
            lib().amps_recc_iq_batch_destroy(self.b)
            self.b = None

    __del__ = close

    def submit_dev(self, dev_ptrs, nsamples, stream: int = 0):
        k = len(self.handles)
        ns = [nsamples] * k if isinstance(nsamples, int) else list(nsamples)
        ptrs = (C.c_void_p * k)(*dev_ptrs)
        cnt = (C.c_size_t * k)(*ns)
        check(lib().amps_recc_iq_batch_submit_dev(self.b, ptrs, cnt, C.c_void_p(stream)))

    def prepare(self, dev_ptrs, nsamples):
        """The two argument arrays of submit_dev built once (a caller that reuses its buffers saves the marshalling)."""
        k = len(self.handles)
        ns = [nsamples] * k if isinstance(nsamples, int) else list(nsamples)
        return (C.c_void_p * k)(*dev_ptrs), (C.c_size_t * k)(*ns)

    def submit_prepared(self, prepared, stream: int = 0):
        check(lib().amps_recc_iq_batch_submit_dev(self.b, prepared[0], prepared[1], C.c_void_p(stream)))

    def work_shared(self, iq: np.ndarray) -> list:
        """One host buffer for every channel; returns [(channel, Burst), ...]."""
        iq = np.ascontiguousarray(iq)
        n = iq.size if iq.dtype == np.complex64 else iq.size // 2
        got = []

        def on(ch, bp, _user):
            b = Burst()
            C.memmove(C.byref(b), bp, C.sizeof(Burst))
            got.append((ch, b))

        cb = BATCH_BURST_CB(on)
        check(lib().amps_recc_iq_batch_work_shared(self.b, iq.ctypes.data_as(C.c_void_p), n, cb, None))
        return got

    def work_shared_ptr(self, host_ptr: int, nsamples: int, cb=None):
        check(lib().amps_recc_iq_batch_work_shared(self.b, C.c_void_p(host_ptr), nsamples, cb or C.cast(None, BATCH_BURST_CB), None))

    def front_times_ms(self, cap=256) -> np.ndarray:
        out = np.zeros(cap, np.float32)
        n = C.c_int(0)
        check(lib().amps_recc_iq_batch_front_times(self.b, out.ctypes.data_as(f32p), cap, C.byref(n)))
        return out[:n.value]

    def stats(self) -> dict:
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(lib().amps_recc_iq_batch_stats(self.b, C.byref(a), C.byref(b)))
        return dict(calls=a.value, kernel_launches=b.value)

class ReccDecode:
    """Message-only burst decoder (amps_recc_decode_*)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().amps_recc_decode_create(device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().amps_recc_decode_destroy(self.h)
            self.h = None

    __del__ = close

    def decode(self, blobs: np.ndarray) -> list[ReccWords]:
        b = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(-1, CAPTURE_SYMS)
        out = (ReccWords * len(b))()
        check(lib().amps_recc_decode_bursts(self.h, b.ctypes.data_as(u8p), len(b), out))
        return list(out)


class Recc:
    """amps.recc byte-stream sink, compat mode (amps_recc_*)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().amps_recc_create(device, C.byref(self.h)))
        self.bursts: list[np.ndarray] = []
        self._cb = BLOB_CB(lambda p, u: self.bursts.append(np.ctypeslib.as_array(p, shape=(CAPTURE_SYMS,)).copy()))

    def close(self):
        if getattr(self, "h", None):
            lib().amps_recc_destroy(self.h)
            self.h = None

    __del__ = close

    def work(self, syms) -> int:
        s = np.ascontiguousarray(syms, dtype=np.uint8)
        check(lib().amps_recc_work(self.h, s.ctypes.data_as(u8p), len(s), self._cb, None))
        return 0    # recc_impl::work always returns 0 (lib/recc_impl.cc:144)

    def work_chunks(self, syms, sizes) -> None:
        s = np.ascontiguousarray(syms, dtype=np.uint8)
        z = (C.c_int * len(sizes))(*[int(v) for v in sizes])
        check(lib().amps_recc_work_chunks(self.h, s.ctypes.data_as(u8p), z, len(sizes), self._cb, None))


class Focc:
    """amps.focc half-symbol source (amps_focc_*)."""

    def __init__(self, symrate=100000, aggressive=False, device=0):
        self.h = C.c_void_p()
        check(lib().amps_focc_create(symrate, int(aggressive), device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().amps_focc_destroy(self.h)
            self.h = None

    __del__ = close

    def work(self, n: int):
        buf = np.zeros(max(n, 1), np.uint8)
        produced = C.c_int(0)
        check(lib().amps_focc_work(self.h, buf.ctypes.data_as(u8p), n, C.byref(produced)))
        return produced.value, buf[:max(produced.value, 0)].copy()

    def generate(self, n: int) -> np.ndarray:
        buf = np.zeros(n, np.uint8)
        check(lib().amps_focc_generate(self.h, buf.ctypes.data_as(u8p), n))
        return buf

    def generate_dev(self, dev_ptr: int, n: int, stream: int = 0):
        check(lib().amps_focc_generate_dev(self.h, C.c_void_p(dev_ptr), n, C.c_void_p(stream)))

    def generate_bits(self, nbits: int) -> np.ndarray:
        buf = np.zeros(nbits, np.uint8)
        check(lib().amps_focc_generate_bits(self.h, buf.ctypes.data_as(u8p), nbits))
        return buf

    def generate_bits_dev(self, dev_ptr: int, nbits: int, stream: int = 0):
        check(lib().amps_focc_generate_bits_dev(self.h, C.c_void_p(dev_ptr), nbits, C.c_void_p(stream)))

    def push_words(self, stream: int, words):
        w = np.ascontiguousarray(words, dtype=np.uint8).reshape(-1)
        check(lib().amps_focc_push_words(self.h, stream, w.ctypes.data_as(u8p), len(w) // 28))

    def set_busy_idle(self, idle: bool):
        check(lib().amps_focc_set_busy_idle(self.h, int(idle)))


class Fvc:
    """amps.fvc half-symbol source (amps_fvc_*)."""

    def __init__(self, symrate=100000, device=0):
        self.h = C.c_void_p()
        check(lib().amps_fvc_create(symrate, device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().amps_fvc_destroy(self.h)
            self.h = None

    __del__ = close

    def push_words(self, words, timer=None):
        w = np.ascontiguousarray(words, dtype=np.uint8).reshape(-1)
        check(lib().amps_fvc_push_words(self.h, w.ctypes.data_as(u8p), len(w) // 28, int(timer is not None), int(timer or 0)))

    def work(self, n: int, fill=0x55):
        buf = np.full(max(n, 1), fill, np.uint8)
        produced, off = C.c_int(0), C.c_int(0)
        check(lib().amps_fvc_work(self.h, buf.ctypes.data_as(u8p), n, C.byref(produced), C.byref(off)))
        return produced.value, buf[:max(produced.value, 0)].copy(), bool(off.value)


class Fwd:
    """Fused forward path (amps_fwd_*): half-symbol bytes of up to 3 carriers -> complex baseband @10 MS/s."""

    def __init__(self, max_samples: int, carrier_freq=(0.0, 60e3, 90e3), lpf_transition=(5e3, 3e3, 3e3),
                 out_scale=0.5, device=0, max_deviation=8000.0):
        n = len(carrier_freq)
        p = FwdParams(10e6, 100e3, max_deviation, device, n, (C.c_double * 3)(*(list(carrier_freq) + [0.0] * (3 - n))),
                      (C.c_double * 3)(*(list(lpf_transition)[:n] + [5e3] * (3 - n))), out_scale, max_samples)
        self.h = C.c_void_p()
        self.ncar = n
        check(lib().amps_fwd_create(C.byref(p), C.byref(self.h)))
        self.interp = lib().amps_fwd_interp(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib().amps_fwd_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        check(lib().amps_fwd_reset(self.h))

    def taps(self, carrier: int) -> np.ndarray:
        n = lib().amps_fwd_get_taps(self.h, carrier, None, 0)
        t = np.zeros(n, np.float32)
        lib().amps_fwd_get_taps(self.h, carrier, t.ctypes.data_as(f32p), n)
        return t

    def work(self, syms) -> np.ndarray:
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
        nsym = len(arrs[0])
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs] + [None] * (3 - len(arrs)))
        out = np.zeros(2 * nsym * self.interp, np.float32)
        check(lib().amps_fwd_work(self.h, ptrs, nsym, out.ctypes.data_as(f32p)))
        return out.view(np.complex64)

    def submit_dev(self, dev_ptrs, nsym: int, out_ptr: int, stream: int = 0):
        ptrs = (C.c_void_p * 3)(*list(dev_ptrs) + [None] * (3 - len(dev_ptrs)))
        check(lib().amps_fwd_submit_dev(self.h, ptrs, nsym, C.c_void_p(out_ptr), C.c_void_p(stream)))

    def enable_voice(self, carrier_gated=1, carrier_open=2, sat_amp=0.05):
        """Voice legs of the reference graph: nbfm_tx(16 kS/s) -> x25 arb resampler in front of the carriers' mixers."""
        vp = FwdVoiceParams(carrier_gated, carrier_open, 16000.0, 8e3, 75e-6, 6000.0, sat_amp)
        lib().amps_fwd_enable_voice.argtypes = [C.c_void_p, C.POINTER(FwdVoiceParams)]
        check(lib().amps_fwd_enable_voice(self.h, C.byref(vp)))

    def work_voice(self, syms, audio, audio_mute=False) -> np.ndarray:
        arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
        nsym = len(arrs[0])
        a = np.ascontiguousarray(audio, dtype=np.float32)
        assert len(a) == nsym * 4 // 25
        ptrs = (C.c_void_p * 3)(*[x.ctypes.data for x in arrs] + [None] * (3 - len(arrs)))
        out = np.zeros(2 * nsym * self.interp, np.float32)
        lib().amps_fwd_work_voice.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), f32p, C.c_size_t, C.c_int, f32p]
        check(lib().amps_fwd_work_voice(self.h, ptrs, a.ctypes.data_as(f32p), nsym, int(bool(audio_mute)), out.ctypes.data_as(f32p)))
        return out.view(np.complex64)

    def submit_voice_dev(self, dev_ptrs, audio_ptr: int, nsym: int, out_ptr: int, audio_mute=False, stream: int = 0):
        ptrs = (C.c_void_p * 3)(*list(dev_ptrs) + [None] * (3 - len(dev_ptrs)))
        lib().amps_fwd_submit_voice_dev.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
        check(lib().amps_fwd_submit_voice_dev(self.h, ptrs, C.c_void_p(audio_ptr), nsym, int(bool(audio_mute)), C.c_void_p(out_ptr), C.c_void_p(stream)))

    def work_bits(self, bits) -> np.ndarray:
        """Manchester-bit fast path: bits = list of uint8 arrays (0, 1, 0xFF = muted); 1000 output samples per bit."""
        arrs = [np.ascontiguousarray(b, dtype=np.uint8) for b in bits]
        nbits = len(arrs[0])
        ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in arrs] + [None] * (3 - len(arrs)))
        out = np.zeros(2 * nbits * 1000, np.float32)
        check(lib().amps_fwd_work_bits(self.h, ptrs, nbits, out.ctypes.data_as(f32p)))
        return out.view(np.complex64)

    def submit_bits_dev(self, dev_ptrs, nbits: int, out_ptr: int, stream: int = 0):
        ptrs = (C.c_void_p * 3)(*list(dev_ptrs) + [None] * (3 - len(dev_ptrs)))
        check(lib().amps_fwd_submit_bits_dev(self.h, ptrs, nbits, C.c_void_p(out_ptr), C.c_void_p(stream)))
