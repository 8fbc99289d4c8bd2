"""ctypes binding of oracle/_ref/libamps_ref.so: the reference's OWN block sources (compiled from /root/reference/lib
by `make -C oracle _ref` against stand-in GNU Radio / Boost / IT++ headers) behind the C harness oracle/ref_harness.cc.
Test infrastructure: used to pin oracle/*.c and to generate tests/golden/ref_vectors.json.  The classes mirror
tests/oracle_lib.py so that a test can run the same schedule through both."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

from . import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libamps_ref.so")
REF_TREE = os.environ.get("AMPS_REFERENCE", "/root/reference")

u8p = O.u8p
_lib = None


def available() -> bool:
    """True when the compiled reference exists (prebuilt .so travels to the GPU box) or can be built here."""
    if os.path.exists(REF_SO):
        return True
    if os.path.isdir(os.path.join(REF_TREE, "lib")):
        try:
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_ref", "REF=" + REF_TREE])
        except (subprocess.CalledProcessError, OSError):
            return False
        return os.path.exists(REF_SO)
    return False


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libamps_ref.so is not built and there is no reference tree to build it from")
    L = C.CDLL(REF_SO)
    L.ref_focc_new.restype = C.c_void_p
    L.ref_focc_new.argtypes = [C.c_ulong, C.c_int]
    L.ref_focc_free.argtypes = [C.c_void_p]
    L.ref_focc_work.argtypes = [C.c_void_p, u8p, C.c_int]
    L.ref_focc_push_words.argtypes = [C.c_void_p, C.c_long, u8p, C.c_long]
    L.ref_set_busy_idle.argtypes = [C.c_int]
    L.ref_fvc_new.restype = C.c_void_p
    L.ref_fvc_new.argtypes = [C.c_ulong]
    L.ref_fvc_free.argtypes = [C.c_void_p]
    L.ref_fvc_push_words.argtypes = [C.c_void_p, u8p, C.c_long, C.c_int, C.c_uint64]
    L.ref_fvc_work.argtypes = [C.c_void_p, u8p, C.c_int, C.POINTER(C.c_int)]
    L.ref_recc_new.restype = C.c_void_p
    L.ref_recc_free.argtypes = [C.c_void_p]
    L.ref_recc_work.argtypes = [C.c_void_p, u8p, C.c_int, O.BURST_CB, C.c_void_p]
    L.ref_recc_trigger.argtypes = [u8p]
    L.ref_recc_bursts_message.argtypes = [u8p, C.c_size_t, C.POINTER(O.ReccActions), C.c_char_p, C.c_size_t]
    L.ref_recc_bursts_message.restype = None
    L.ref_recc_fields.argtypes = [u8p, C.POINTER(O.ReccResult)]
    L.ref_recc_fields.restype = None
    L.ref_called_digits.argtypes = [u8p, C.c_char_p]
    L.ref_command.argtypes = [C.c_char_p, C.POINTER(O.CmdActions)]
    L.ref_command.restype = None
    ui = C.c_uint
    L.ref_focc_word1.argtypes = [u8p, C.c_int, ui, C.c_uint64]
    L.ref_focc_word2_general.argtypes = [u8p, C.c_uint64, ui, ui, ui]
    L.ref_fvc_word1_general.argtypes = [u8p, ui, ui, ui, ui]
    L.ref_focc_word2_voice_channel.argtypes = [u8p, ui, C.c_uint64, ui, ui]
    L.ref_expandbits.argtypes = [u8p, C.c_size_t, C.c_uint64]
    L.ref_manchester_decode.argtypes = [u8p, u8p, C.c_size_t]
    L.ref_manchester_decode.restype = C.c_size_t
    L.ref_extract_min_3.argtypes = [C.c_uint64, C.c_char_p]
    L.ref_compute_min_3.argtypes = [C.c_char, C.c_char, C.c_char]
    L.ref_compute_min_3.restype = C.c_uint64
    L.ref_parse_min.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ref_calc_min.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]
    L.ref_bch_encode_40_28.argtypes = [u8p, u8p]
    L.ref_bch_decode_48.argtypes = [u8p, u8p]
    L.ref_bch_decode_48.restype = C.c_int
    _lib = L
    return L


def word(builder: str, *args) -> np.ndarray:
    w = np.zeros(28, np.uint8)
    getattr(lib(), builder)(O.ptr(w, u8p), *args)
    return w


def bch_encode_40_28(bits28) -> np.ndarray:
    i = O.as_u8(bits28)
    o = np.zeros(40, np.uint8)
    lib().ref_bch_encode_40_28(O.ptr(i, u8p), O.ptr(o, u8p))
    return o


def bch_decode_48(bits48) -> bool:
    i = O.as_u8(bits48)
    return bool(lib().ref_bch_decode_48(O.ptr(i, u8p), None))


def recc_fields(blob) -> O.ReccResult:
    b = O.as_u8(blob)
    assert len(b) == 3374
    r = O.ReccResult()
    lib().ref_recc_fields(O.ptr(b, u8p), C.byref(r))
    return r


_LOG_PATTERNS = [  # (kind as in oracle/amps_oracle.h:orc_recc_result.kind, regex on the reference's log text)
    (0, re.compile(r"got a burst with an invalid Word A")),
    (1, re.compile(r"got a RECC message with E=0")),
    (6, re.compile(r"invalid NAWC value in RECC origination")),
    (4, re.compile(r"origination: MIN=(?P<min>\d+) ESN=(?P<esn>[0-9a-f]+) dialed (?P<dialed>[0-9*#]*)")),
    (3, re.compile(r"got registration from MIN=(?P<min>\d+)")),
    (2, re.compile(r"got a response from MIN=(?P<min>\d+)")),
    (5, re.compile(r"got unknown RECC message")),
]


def recc_bursts_message(blob):
    """Run the reference's recc_decode_impl::bursts_message on one blob.  Returns (actions, info) where info has the
    dispatch `kind` and whatever of min/esn/dialed the reference logged (they are not published on any port)."""
    b = O.as_u8(blob)
    a = O.ReccActions()
    log = C.create_string_buffer(1 << 16)
    lib().ref_recc_bursts_message(O.ptr(b, u8p), len(b), C.byref(a), log, len(log))
    text = log.value.decode("ascii", "replace")
    info = {"kind": None, "log": text}
    for kind, rx in _LOG_PATTERNS:
        m = rx.search(text)
        if m:
            info["kind"] = kind
            info.update(m.groupdict())
            break
    m = re.search(r"registration included S; ESN=([0-9a-f]+)", text)
    if m:
        info["esn"] = m.group(1)
    return a, info


def command_actions(cmd: str) -> O.CmdActions:
    a = O.CmdActions()
    lib().ref_command(cmd.encode(), C.byref(a))
    return a


class Focc:
    def __init__(self, symrate=100000, aggressive=False):
        self.h = lib().ref_focc_new(symrate, int(aggressive))

    def work(self, n):
        buf = np.zeros(max(n, 1), np.uint8)
        r = lib().ref_focc_work(self.h, O.ptr(buf, u8p), n)
        return r, buf[:max(r, 0)].copy()

    def generate(self, total, chunk=4096):
        out = bytearray()
        while len(out) < total:
            r, b = self.work(min(chunk, total - len(out)))
            out += b.tobytes()
        return np.frombuffer(bytes(out), np.uint8)

    def push_words(self, stream, words):
        w = O.as_u8(words).reshape(-1)
        lib().ref_focc_push_words(self.h, stream, O.ptr(w, u8p), len(w) // 28)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_focc_free(self.h)
            self.h = None


class Fvc:
    def __init__(self, symrate=100000):
        self.h = lib().ref_fvc_new(symrate)

    def push_words(self, words, timer=None):
        w = O.as_u8(words).reshape(-1)
        lib().ref_fvc_push_words(self.h, O.ptr(w, u8p), len(w) // 28, int(timer is not None), int(timer or 0))

    def work(self, n, fill=0x55):
        buf = np.full(max(n, 1), fill, np.uint8)
        off = C.c_int(0)
        r = lib().ref_fvc_work(self.h, O.ptr(buf, u8p), n, C.byref(off))
        return r, buf[:max(r, 0)].copy(), bool(off.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_fvc_free(self.h)
            self.h = None


class Recc:
    def __init__(self):
        self.h = lib().ref_recc_new()
        self.bursts = []
        self._cb = O.BURST_CB(self._on)

    def _on(self, p, user):
        self.bursts.append(np.ctypeslib.as_array(p, shape=(3374,)).copy())

    def work(self, syms):
        s = O.as_u8(syms)
        return lib().ref_recc_work(self.h, O.ptr(s, u8p), len(s), self._cb, None)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_recc_free(self.h)
            self.h = None
