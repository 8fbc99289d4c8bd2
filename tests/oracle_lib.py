"""ctypes binding of oracle/liboracle.so -- the CPU checker.  Imported by tests only
(and by bench.py's cpu_baseline leg / __graft_entry__.smoke through the same file)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class ReccResult(C.Structure):
    _fields_ = [
        ("dcc", C.c_uint8 * 7), ("dcc_errs", C.c_uint8),
        ("words", (C.c_uint8 * 240) * 7),
        ("errs", C.c_uint16 * 7),
        ("valid", C.c_uint8 * 7), ("valid_repeat", C.c_uint8 * 7),
        ("F", C.c_uint8), ("NAWC", C.c_uint8), ("T", C.c_uint8), ("S", C.c_uint8), ("E", C.c_uint8),
        ("ER", C.c_uint8), ("SCM", C.c_uint8),
        ("MIN1", C.c_uint32),
        ("B_F", C.c_uint8), ("B_NAWC", C.c_uint8), ("MSG_TYPE", C.c_uint8), ("ORDQ", C.c_uint8),
        ("ORDER", C.c_uint8), ("LT", C.c_uint8), ("EP", C.c_uint8), ("SCM4", C.c_uint8), ("MPCI", C.c_uint8),
        ("SDCC1", C.c_uint8), ("SDCC2", C.c_uint8),
        ("MIN2", C.c_uint16),
        ("word_c_serial", C.c_uint32),
        ("kind", C.c_int32),
        ("esn", C.c_uint32),
        ("min", C.c_char * 11),
        ("dialed", C.c_char * 33),
    ]


class ReccActions(C.Structure):
    _fields_ = [("n_focc", C.c_int32), ("focc_stream", C.c_int64), ("focc_words", (C.c_uint8 * 28) * 2),
                ("has_fvc", C.c_int32), ("fvc_word", C.c_uint8 * 28), ("fvc_timer", C.c_uint64),
                ("fvc_mute", C.c_int32), ("audio_mute", C.c_int32), ("command", C.c_char * 48)]


class CmdActions(C.Structure):
    _fields_ = [("n_focc", C.c_int32), ("focc_stream", C.c_int64), ("focc_words", (C.c_uint8 * 28) * 2),
                ("has_fvc", C.c_int32), ("fvc_word", C.c_uint8 * 28), ("fvc_mute", C.c_int32), ("audio_mute", C.c_int32),
                ("n_debug", C.c_int32), ("debug", (C.c_char * 48) * 2)]


class Burst(C.Structure):
    _fields_ = [("d_index", C.c_uint64), ("corr", C.c_float), ("symbols", C.c_uint8 * 3374)]


BURST_CB = C.CFUNCTYPE(None, u8p, C.c_void_p)

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B" if force else "-s", "liboracle.so"])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    L.orc_bch_encode_40_28.argtypes = [u8p, u8p]
    L.orc_bch_encode_48_36.argtypes = [u8p, u8p]
    L.orc_bch_decode_48.argtypes = [u8p, u8p]
    L.orc_bch_decode_48.restype = C.c_int
    L.orc_bch_syndromes63.argtypes = [u8p]
    L.orc_bch_syndromes63.restype = C.c_uint32
    L.orc_focc_new.restype = C.c_void_p
    L.orc_focc_new.argtypes = [C.c_ulong, C.c_int]
    L.orc_focc_free.argtypes = [C.c_void_p]
    L.orc_focc_work.argtypes = [C.c_void_p, u8p, C.c_int]
    L.orc_focc_push_words.argtypes = [C.c_void_p, C.c_long, u8p, C.c_long]
    L.orc_focc_superframe_frames.argtypes = [C.c_void_p]
    L.orc_fvc_new.restype = C.c_void_p
    L.orc_fvc_new.argtypes = [C.c_ulong]
    L.orc_fvc_free.argtypes = [C.c_void_p]
    L.orc_fvc_push_words.argtypes = [C.c_void_p, u8p, C.c_long, C.c_int, C.c_uint64]
    L.orc_fvc_work.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_int)]
    L.orc_recc_new.restype = C.c_void_p
    L.orc_recc_free.argtypes = [C.c_void_p]
    L.orc_recc_trigger.argtypes = [u8p]
    L.orc_recc_work.argtypes = [C.c_void_p, u8p, C.c_int, BURST_CB, C.c_void_p]
    L.orc_recc_buflen.argtypes = [C.c_void_p]
    L.orc_recc_buflen.restype = C.c_size_t
    L.orc_manchester_decode.argtypes = [u8p, u8p, C.c_size_t]
    L.orc_manchester_decode.restype = C.c_size_t
    L.orc_recc_decode.argtypes = [u8p, C.POINTER(ReccResult)]
    L.orc_recc_actions_for.argtypes = [C.POINTER(ReccResult), C.POINTER(ReccActions)]
    L.orc_firdes_low_pass.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, f32p, C.c_int]
    L.orc_nco_fcw.argtypes = [C.c_double, C.c_double]
    L.orc_nco_fcw.restype = C.c_uint32
    L.orc_rx_chain_f64.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, f64p, f64p]
    L.orc_rx_chain_f32.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, f32p, f32p]
    L.orc_rx_chain_f32_at.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, C.c_uint32, f32p, f32p]
    L.orc_rx_chain400_f64.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, f64p, f64p]
    L.orc_rx_chain400_f32.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, f32p, f32p]
    L.orc_rx_detect.argtypes = [f32p, C.c_size_t, C.POINTER(Burst), C.c_int]
    L.orc_cpu_baseline_run.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.orc_cpu_baseline_run.restype = C.c_double
    L.orc_fwd_chain_f64.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_uint32, C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_double, f64p]
    ui = C.c_uint
    L.orc_overhead_word_1.argtypes = [u8p, ui, ui, C.c_int, C.c_int, C.c_int, ui]
    L.orc_overhead_word_2.argtypes = [u8p, ui, C.c_int, C.c_int, C.c_int, C.c_int, ui, ui, C.c_int, C.c_int, ui, C.c_int]
    L.orc_control_filler_word.argtypes = [u8p]
    L.orc_access_type_global_action.argtypes = [u8p, ui, C.c_int]
    L.orc_reg_increment_global_action.argtypes = [u8p, ui, ui, C.c_int]
    L.orc_registration_id.argtypes = [u8p, ui, C.c_ulong, C.c_int]
    L.orc_focc_word1.argtypes = [u8p, C.c_int, ui, C.c_uint64]
    L.orc_focc_word2_general.argtypes = [u8p, C.c_uint64, ui, ui, ui]
    L.orc_fvc_word1_general.argtypes = [u8p, ui, ui, ui, ui]
    L.orc_focc_word2_voice_channel.argtypes = [u8p, ui, C.c_uint64, ui, ui]
    L.orc_compute_min_3.argtypes = [C.c_char, C.c_char, C.c_char]
    L.orc_compute_min_3.restype = C.c_uint64
    L.orc_extract_min_3.argtypes = [C.c_uint64, C.c_char_p]
    L.orc_calc_min.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]
    L.orc_parse_min.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def word(builder: str, *args) -> np.ndarray:
    """Call one of the 28-bit word builders, e.g. word('orc_overhead_word_1', 0, 16, 1, 0, 0, 3)."""
    w = np.zeros(28, np.uint8)
    getattr(lib(), builder)(w.ctypes.data_as(u8p), *args)
    return w


def as_u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def bch_encode_40_28(bits28) -> np.ndarray:
    i = as_u8(bits28)
    o = np.zeros(40, np.uint8)
    lib().orc_bch_encode_40_28(ptr(i, u8p), ptr(o, u8p))
    return o


def bch_encode_48_36(bits36) -> np.ndarray:
    i = as_u8(bits36)
    o = np.zeros(48, np.uint8)
    lib().orc_bch_encode_48_36(ptr(i, u8p), ptr(o, u8p))
    return o


def bch_decode_48(bits48):
    i = as_u8(bits48)
    o = np.zeros(48, np.uint8)
    ok = lib().orc_bch_decode_48(ptr(i, u8p), ptr(o, u8p))
    return bool(ok), o


def lpf_taps() -> np.ndarray:
    """lpf_taps of grc/ampsbs.grc:138-184: firdes.low_pass(3, 400e3, 10e3, 4.5e3, BLACKMAN)."""
    n = lib().orc_firdes_low_pass(3.0, 400e3, 10e3, 4500.0, 2, None, 0)
    t = np.zeros(n, np.float32)
    lib().orc_firdes_low_pass(3.0, 400e3, 10e3, 4500.0, 2, ptr(t, f32p), n)
    return t


def firdes_low_pass(gain, fs, fc, tw, window=0) -> np.ndarray:
    n = lib().orc_firdes_low_pass(gain, fs, fc, tw, window, None, 0)
    t = np.zeros(n, np.float32)
    lib().orc_firdes_low_pass(gain, fs, fc, tw, window, ptr(t, f32p), n)
    return t


def iq_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x.astype(np.complex64, copy=False))
    return x.view(np.float32)


def rx_chain_f32(x: np.ndarray, center=-160e3, fs=10e6, taps=None, blk0=0):
    """blk0: absolute index of the stream's 25-sample block x[0] starts (the NCO phase follows the absolute sample index)."""
    iq = iq_f32(x)
    n = len(x) - len(x) % 50
    taps = lpf_taps() if taps is None else taps
    fcw = lib().orc_nco_fcw(center, fs)
    y = np.zeros(2 * (n // 50), np.float32)
    d = np.zeros(n // 50, np.float32)
    lib().orc_rx_chain_f32_at(ptr(iq, f32p), n, fcw, ptr(taps, f32p), len(taps), blk0 & 0xFFFFFFFF, ptr(y, f32p), ptr(d, f32p))
    return y.view(np.complex64), d


def rx_chain_f64(x: np.ndarray, center=-160e3, fs=10e6, taps=None):
    iq = iq_f32(x)
    n = len(x) - len(x) % 50
    taps = lpf_taps() if taps is None else taps
    fcw = lib().orc_nco_fcw(center, fs)
    y = np.zeros(2 * (n // 50), np.float64)
    d = np.zeros(n // 50, np.float64)
    lib().orc_rx_chain_f64(ptr(iq, f32p), n, fcw, ptr(taps, f32p), len(taps), ptr(y, f64p), ptr(d, f64p))
    return y.view(np.complex128), d


def rx_chain400_f32(x: np.ndarray, center=-160e3, fs=400e3, taps=None):
    iq = iq_f32(x)
    n = len(x) - len(x) % 2
    taps = lpf_taps() if taps is None else taps
    fcw = lib().orc_nco_fcw(center, fs)
    y = np.zeros(2 * (n // 2), np.float32)
    d = np.zeros(n // 2, np.float32)
    lib().orc_rx_chain400_f32(ptr(iq, f32p), n, fcw, ptr(taps, f32p), len(taps), ptr(y, f32p), ptr(d, f32p))
    return y.view(np.complex64), d


def rx_chain400_f64(x: np.ndarray, center=-160e3, fs=400e3, taps=None):
    iq = iq_f32(x)
    n = len(x) - len(x) % 2
    taps = lpf_taps() if taps is None else taps
    fcw = lib().orc_nco_fcw(center, fs)
    y = np.zeros(2 * (n // 2), np.float64)
    d = np.zeros(n // 2, np.float64)
    lib().orc_rx_chain400_f64(ptr(iq, f32p), n, fcw, ptr(taps, f32p), len(taps), ptr(y, f64p), ptr(d, f64p))
    return y.view(np.complex128), d


def rx_detect(d: np.ndarray, max_bursts=64):
    d = np.ascontiguousarray(d, dtype=np.float32)
    arr = (Burst * max_bursts)()
    n = lib().orc_rx_detect(ptr(d, f32p), len(d), arr, max_bursts)
    return [(int(arr[i].d_index), float(arr[i].corr), np.frombuffer(bytes(arr[i].symbols), np.uint8).copy()) for i in range(n)]


class MmState(C.Structure):
    _fields_ = [("mu", C.c_float), ("omega", C.c_float), ("last", C.c_float), ("pad", C.c_uint32), ("pos", C.c_uint64)]


def mmse_table() -> np.ndarray:
    t = np.zeros(129 * 8, np.float32)
    lib().orc_mmse_table.argtypes = [f32p]
    lib().orc_mmse_table(ptr(t, f32p))
    return t.reshape(129, 8)


class MmTiming:
    """Serial clock_recovery_mm_ff + binary_slicer_fb on a growing demod stream (oracle/mm_timing.c)."""

    def __init__(self):
        L = lib()
        L.orc_mm_init.argtypes = [C.POINTER(MmState)]
        L.orc_mm_process.argtypes = [C.POINTER(MmState), f32p, C.c_uint64, f32p, u8p, C.c_size_t]
        L.orc_mm_process.restype = C.c_size_t
        self.st = MmState()
        L.orc_mm_init(C.byref(self.st))
        self.table = np.ascontiguousarray(mmse_table().reshape(-1))

    def process(self, d: np.ndarray, total: int) -> np.ndarray:
        """d = the whole stream so far (float32, from its first sample); returns the new half-symbols."""
        d = np.ascontiguousarray(d, dtype=np.float32)
        out = np.zeros(total // 8 + 16, np.uint8)
        n = lib().orc_mm_process(C.byref(self.st), ptr(d, f32p), total, ptr(self.table, f32p), ptr(out, u8p), len(out))
        return out[:n]


_native = None


def native_lib():
    """The oracle sources rebuilt -O3 -march=native ON THIS MACHINE, for the CPU-baseline legs of
    bench.py only; falls back to the portable parity build if the compiler is not there.  Returns (CDLL, flags string)."""
    global _native
    if _native is not None:
        return _native
    path = os.path.join(ORACLE_DIR, "liboracle_native.so")
    try:
        subprocess.check_call(["make", "-s", "-B", "-C", ORACLE_DIR, "liboracle_native.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        L = C.CDLL(path)
        flags = "-O3 -march=native -ffp-contract=off"
    except Exception:
        L = lib()
        flags = "-O2 -march=x86-64-v3 -ffp-contract=off (portable parity build: native rebuild failed)"
    L.orc_cpu_baseline_run.argtypes = [f32p, C.c_size_t, C.c_uint32, f32p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.orc_cpu_baseline_run.restype = C.c_double
    _native = (L, flags)
    return _native


def cpu_baseline_run(x: np.ndarray, threads: int, reps: int, center=-160e3, fs=10e6, native=False):
    """Time the fp32 oracle chain (+detect+decode) on `threads` host threads; returns (seconds, bursts)."""
    iq = iq_f32(x)
    n = len(x) - len(x) % 50
    taps = lpf_taps()
    fcw = lib().orc_nco_fcw(center, fs)
    nb = C.c_int(0)
    L = native_lib()[0] if native else lib()
    sec = L.orc_cpu_baseline_run(ptr(iq, f32p), n, fcw, ptr(taps, f32p), len(taps), threads, reps, C.byref(nb))
    return sec, nb.value


def command_actions(cmd: str) -> CmdActions:
    a = CmdActions()
    lib().orc_command_actions.argtypes = [C.c_char_p, C.POINTER(CmdActions)]
    lib().orc_command_actions.restype = None
    lib().orc_command_actions(cmd.encode(), C.byref(a))
    return a


def recc_actions(result: ReccResult) -> ReccActions:
    a = ReccActions()
    lib().orc_recc_actions_for(C.byref(result), C.byref(a))
    return a


def fwd_chain_f64(syms, carrier_freq=(0.0, 60e3, 90e3), lpf_transition=(5e3, 3e3, 3e3), scale=0.5,
                  max_deviation=8000.0, symrate=100e3, fs=10e6) -> np.ndarray:
    """float64 forward chain of BASELINE config 3 (see oracle/dsp_chain.c); syms = list of +-1/0 byte arrays."""
    n = len(syms)
    arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
    nsym = len(arrs[0])
    taps = [firdes_low_pass(1.0, 400e3, 10e3, lpf_transition[c], 0) for c in range(n)]
    symp = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    tapp = (C.c_void_p * n)(*[t.ctypes.data for t in taps])
    nt = (C.c_int * n)(*[len(t) for t in taps])
    fcw = (C.c_uint32 * n)(*[lib().orc_nco_fcw(-carrier_freq[c], fs) for c in range(n)])
    fcw_fm = int(round(max_deviation / symrate * 4294967296.0)) & 0xffffffff
    out = np.zeros(2 * nsym * 100, np.float64)
    lib().orc_fwd_chain_f64(symp, n, nsym, fcw_fm, tapp, nt, fcw, scale, ptr(out, f64p))
    return out.view(np.complex128)


def voice_lpf_taps() -> np.ndarray:
    """variable_low_pass_filter_taps voice_lpf_taps: firdes.low_pass(3, 400e3, 15e3, 6e3, BLACKMAN)."""
    return firdes_low_pass(3.0, 400e3, 15e3, 6e3, 2)


def voice_tx_f64(audio, mute=None, sat_amp=0.05) -> np.ndarray:
    """float64 voice leg (oracle/voice_tx.c): audio @16 kS/s -> complex baseband @400 kS/s (25 per audio sample)."""
    a = np.ascontiguousarray(audio, dtype=np.float32)
    taps = voice_lpf_taps()
    out = np.zeros(2 * 25 * len(a), np.float64)
    m = None if mute is None else np.ascontiguousarray(mute, dtype=np.uint8)
    L = lib()
    L.orc_voice_tx_f64.argtypes = [f32p, C.c_size_t, C.c_double, C.c_void_p, f32p, C.c_int, f64p]
    L.orc_voice_tx_f64.restype = None
    L.orc_voice_tx_f64(ptr(a, f32p), len(a), sat_amp, None if m is None else m.ctypes.data, ptr(taps, f32p), len(taps), ptr(out, f64p))
    return out.view(np.complex128)


def fwd_chain_voice_f64(syms, extra400, carrier_freq=(0.0, 60e3, 90e3), lpf_transition=(5e3, 3e3, 3e3), scale=0.5,
                        max_deviation=8000.0, symrate=100e3, fs=10e6) -> np.ndarray:
    """fwd_chain_f64 with extra400[c] (complex128 @400 kS/s or None) added to carrier c in front of its mixer."""
    n = len(syms)
    arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in syms]
    nsym = len(arrs[0])
    taps = [firdes_low_pass(1.0, 400e3, 10e3, lpf_transition[c], 0) for c in range(n)]
    symp = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    tapp = (C.c_void_p * n)(*[t.ctypes.data for t in taps])
    nt = (C.c_int * n)(*[len(t) for t in taps])
    fcw = (C.c_uint32 * n)(*[lib().orc_nco_fcw(-carrier_freq[c], fs) for c in range(n)])
    fcw_fm = int(round(max_deviation / symrate * 4294967296.0)) & 0xffffffff
    ex = [None if e is None else np.ascontiguousarray(e, dtype=np.complex128) for e in extra400]
    exp = (C.c_void_p * n)(*[None if e is None else e.ctypes.data for e in ex])
    out = np.zeros(2 * nsym * 100, np.float64)
    L = lib()
    L.orc_fwd_chain_voice_f64.argtypes = L.orc_fwd_chain_f64.argtypes[:8] + [C.POINTER(C.c_void_p), f64p]
    L.orc_fwd_chain_voice_f64.restype = None
    L.orc_fwd_chain_voice_f64(symp, n, nsym, fcw_fm, tapp, nt, fcw, scale, exp, ptr(out, f64p))
    return out.view(np.complex128)


def recc_decode(blob) -> ReccResult:
    b = as_u8(blob)
    assert len(b) == 3374
    r = ReccResult()
    lib().orc_recc_decode(ptr(b, u8p), C.byref(r))
    return r


class Focc:
    def __init__(self, symrate=100000, aggressive=False):
        self.h = lib().orc_focc_new(symrate, int(aggressive))

    def work(self, n):
        buf = np.zeros(max(n, 1), np.uint8)
        r = lib().orc_focc_work(self.h, ptr(buf, u8p), n)
        return r, buf[:max(r, 0)].copy()

    def generate(self, total, chunk=4096):
        out = bytearray()
        while len(out) < total:
            r, b = self.work(min(chunk, total - len(out)))
            out += b.tobytes()
        return np.frombuffer(bytes(out), np.uint8)

    def push_words(self, stream, words):
        w = as_u8(words).reshape(-1)
        return lib().orc_focc_push_words(self.h, stream, ptr(w, u8p), len(w) // 28)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_focc_free(self.h)
            self.h = None


class Fvc:
    def __init__(self, symrate=100000):
        self.h = lib().orc_fvc_new(symrate)

    def push_words(self, words, timer=None):
        w = as_u8(words).reshape(-1)
        return lib().orc_fvc_push_words(self.h, ptr(w, u8p), len(w) // 28, int(timer is not None), int(timer or 0))

    def work(self, n, fill=0x55):
        buf = np.full(max(n, 1), fill, np.uint8)
        off = C.c_int(0)
        r = lib().orc_fvc_work(self.h, ptr(buf, u8p), n, C.byref(off))
        return r, buf[:max(r, 0)].copy(), bool(off.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fvc_free(self.h)
            self.h = None


class Recc:
    def __init__(self):
        self.h = lib().orc_recc_new()
        self.bursts = []
        self._cb = BURST_CB(self._on)

    def _on(self, p, user):
        self.bursts.append(np.ctypeslib.as_array(p, shape=(3374,)).copy())

    def work(self, syms):
        s = as_u8(syms)
        return lib().orc_recc_work(self.h, ptr(s, u8p), len(s), self._cb, None)

    def buflen(self):
        return lib().orc_recc_buflen(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_recc_free(self.h)
            self.h = None
