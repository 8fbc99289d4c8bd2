"""Seeded scenarios that drive the byte/bit-level gr-amps blocks.  Each `run_*` takes an implementation module --
tests/oracle_lib.py (our C restatement) or tests/ref_lib.py (the reference's own sources, oracle/_ref) -- and returns
a transcript (bytes) of everything observable: return values, output bytes, published blobs/messages.  The same
scenario through both must give the same transcript; tests/golden/ref_vectors.json holds the SHA-256 of the
reference's transcripts (written by tests/golden/make_ref_golden.py in the container that has /root/reference)."""
from __future__ import annotations

import ctypes as C
import hashlib
import struct

import numpy as np

from gr_amps_b200 import synth

from . import oracle_lib as O


def _pack_i(v: int) -> bytes:
    return struct.pack("<i", int(v))


# ------------------------------------------------------------------------------------------------ FOCC
def run_focc(impl, symrate: int, aggressive: bool, seed: int, total: int, inject: bool) -> bytes:
    """work() with a random schedule of request sizes (incl. 0 -> WORK_DONE); optionally random focc_words injections
    (lib/focc_impl.cc:521-563) between calls."""
    rng = np.random.default_rng(seed)
    f = impl.Focc(symrate, aggressive)
    out = bytearray()
    produced = 0
    while produced < total:
        if inject and rng.random() < 0.02:
            nwords = int(rng.integers(1, 4))
            stream = int(rng.integers(1, 4))
            f.push_words(stream, rng.integers(0, 2, 28 * nwords).astype(np.uint8))
            out += b"P" + _pack_i(stream) + _pack_i(nwords)
        n = int(rng.choice([0, 1, 2, 7, 64, 463, 4096, 8192])) if rng.random() < 0.5 else int(rng.integers(1, 20000))
        r, b = f.work(n)
        out += _pack_i(n) + _pack_i(r) + b.tobytes()
        produced += max(r, 0)
    return bytes(out)


# ------------------------------------------------------------------------------------------------ FVC
def run_fvc(impl, symrate: int, seed: int, calls: int) -> bytes:
    """idle calls (buffer untouched), word pushes with/without the timer hack, growth of d_curdata by later pushes
    (lib/fvc_impl.cc:109-193)."""
    rng = np.random.default_rng(seed)
    f = impl.Fvc(symrate)
    out = bytearray()
    per_train = 2064 * (symrate // 20000)
    for c in range(calls):
        if c in (3, calls // 2) or (c > 3 and rng.random() < 0.01):
            nwords = int(rng.integers(1, 3))
            timer = int(rng.integers(1, 4)) if rng.random() < 0.5 else None
            f.push_words(rng.integers(0, 2, 28 * nwords).astype(np.uint8), timer)
            out += b"P" + _pack_i(nwords) + _pack_i(-1 if timer is None else timer)
        n = int(rng.choice([1, 2, per_train - 1, per_train, per_train + 1, 4096])) if rng.random() < 0.5 else int(rng.integers(1, 3 * per_train))
        r, b, off = f.work(n)
        out += _pack_i(n) + _pack_i(r) + _pack_i(int(off)) + b.tobytes()
    return bytes(out)


# ------------------------------------------------------------------------------------------------ RECC capture
def recc_symbol_stream(seed: int, n_bursts: int, noisy: bool) -> np.ndarray:
    """Random half-symbols with embedded trigger + payload sequences; some bursts truncated, some triggers back to
    back, some payloads containing another trigger (lib/recc_impl.cc:115-134 quirks)."""
    rng = np.random.default_rng(seed)
    trig = synth.trigger_symbols()
    parts = [rng.integers(0, 2, int(rng.integers(0, 9000))).astype(np.uint8)]
    for b in range(n_bursts):
        kind = int(rng.integers(0, 6))
        payload = rng.integers(0, 2, 3374 + int(rng.integers(0, 200))).astype(np.uint8)
        if kind == 0:       # truncated: another trigger arrives before 3374 symbols have passed
            payload = payload[: int(rng.integers(10, 3000))]
        elif kind == 1:     # a trigger inside the payload
            p = int(rng.integers(0, 3200))
            payload[p:p + 74] = trig
        parts.append(trig)
        parts.append(payload)
        gap = int(rng.integers(0, 30000)) if rng.random() < 0.7 else 0
        g = rng.integers(0, 2, gap).astype(np.uint8)
        if noisy and gap > 200:
            q = int(rng.integers(0, gap - 100))
            g[q:q + 73] = trig[:73]     # near miss
        parts.append(g)
    return np.concatenate(parts)


def run_recc(impl, seed: int, n_bursts: int, max_chunk: int, noisy: bool = True) -> bytes:
    rng = np.random.default_rng(seed ^ 0x5EED)
    s = recc_symbol_stream(seed, n_bursts, noisy)
    r = impl.Recc()
    out = bytearray()
    pos = 0
    while pos < len(s):
        n = int(rng.integers(1, max_chunk + 1)) if rng.random() < 0.9 else int(rng.choice([1, 73, 74, 75, 3374, 3448, 3449, 4096]))
        n = min(n, len(s) - pos, 61439)
        before = len(r.bursts)
        ret = r.work(s[pos:pos + n])
        pos += n
        out += _pack_i(n) + _pack_i(ret) + _pack_i(len(r.bursts) - before)
        for b in r.bursts[before:]:
            out += b.tobytes()
    return bytes(out)


# ------------------------------------------------------------------------------------------------ RECC decode
def _x_rem(i: int) -> int:
    """x^i mod g(x), g = octal 12471, by GF(2) long division."""
    v = 1 << i
    for b in range(i, 11, -1):
        if (v >> b) & 1:
            v ^= synth.G_BCH << (b - 12)
    return v


_X_REM = [_x_rem(i) for i in range(63)]
_CORRECTABLE = {0} | set(_X_REM) | {a ^ b for a in _X_REM for b in _X_REM}


def recc_blobs(seed: int, count: int):
    """3374-symbol blobs: well-formed originations / page responses / registrations with random MIN/ESN/digits, with
    0..3 bit errors sprinkled into each repeat, broken Manchester pairs, odd field values, and plain noise."""
    rng = np.random.default_rng(seed)
    blobs = []
    for i in range(count):
        kind = i % 8
        min10 = "".join(str(int(d)) for d in rng.integers(0, 10, 10))
        esn = int(rng.integers(0, 1 << 32))
        if kind in (0, 1, 2):
            nd = int(rng.integers(1, 33))
            dialed = "".join("0123456789*#"[int(d)] for d in rng.integers(0, 12, nd))
            words = synth.origination_words(min10=min10, esn=esn, dialed=dialed, scm=int(rng.integers(0, 16)))
        elif kind == 3:
            words = synth.page_response_words(min10=min10)
        elif kind == 4:
            words = synth.registration_words(min10=min10, esn=esn)
        elif kind == 5:     # arbitrary field values in words A and B
            words = synth.pad_words([list(rng.integers(0, 2, 36)) for _ in range(int(rng.integers(2, 8)))])
            words[0][0] = 1
            words[0][6] = int(rng.random() < 0.8)      # E
        elif kind == 6:     # origination whose NAWC/S combination walks the odd branches (lib/recc_decode_impl.cc:139-158)
            words = synth.origination_words(min10=min10, esn=esn, dialed="0" + min10)
            words[0][1:4] = synth.bits_msb(int(rng.integers(0, 8)), 3)
            words[0][5] = int(rng.integers(0, 2))      # S
        elif i % 16 == 7:   # word A damaged beyond repair in every repeat -> "invalid Word A" (lib/recc_decode_impl.cc:108-111)
            words = synth.origination_words(min10=min10, esn=esn, dialed="411")
        else:
            words = None
        if words is None:
            blob = rng.integers(0, 2, 3374).astype(np.uint8)
        else:
            bits = synth.recc_message_bits(words, dcc7=tuple(int(b) for b in rng.integers(0, 2, 7)))
            blob = synth.manchester(bits)[82:82 + 3374].copy()   # 8 leading dotting symbols + the 74-symbol trigger
            assert len(blob) == 3374
            if kind != 0:
                # bit errors: flip BOTH half-symbols of a pair so that the pair stays valid Manchester
                for w in range(1 if kind == 7 else 0, 7):
                    for rpt in range(5):
                        for _ in range(int(rng.choice([0, 0, 1, 2, 3]))):
                            b = 14 + 480 * w + 96 * rpt + 2 * int(rng.integers(0, 48))
                            blob[b] ^= 1
                            blob[b + 1] ^= 1
            if kind == 7:
                for rpt in range(5):
                    while True:     # an error pattern whose syndrome no pattern of weight <= 2 shares
                        pos = rng.choice(48, size=int(rng.integers(3, 7)), replace=False)
                        rem = 0
                        for p in pos:
                            rem ^= _X_REM[47 - int(p)]
                        if rem not in _CORRECTABLE:
                            break
                    for p in pos:
                        b = 14 + 96 * rpt + 2 * int(p)
                        blob[b] ^= 1
                        blob[b + 1] ^= 1
            if kind in (2, 5):
                # invalid Manchester pairs: (1,1) and (0,0)  (lib/utils.cc:36-44)
                for _ in range(int(rng.integers(1, 12))):
                    b = 2 * int(rng.integers(0, 1687))
                    blob[b] = blob[b + 1] = int(rng.integers(0, 2))
        blobs.append(blob)
    return blobs


def actions_bytes(a) -> bytes:
    return bytes(a)   # ctypes structure memory (both sides memset their struct before filling it)


def result_bytes(r: O.ReccResult, with_dispatch: bool) -> bytes:
    """Fields of an orc_recc_result that the reference computes too (ref_recc_fields leaves kind/esn/dialed zero)."""
    out = bytearray()
    out += bytes(r.dcc) + bytes([r.dcc_errs])
    for w in range(7):
        out += bytes(r.words[w])
    out += struct.pack("<7H", *r.errs) + bytes(r.valid) + bytes(r.valid_repeat)
    for f in ("F", "NAWC", "T", "S", "E", "ER", "SCM", "MIN1", "B_F", "B_NAWC", "MSG_TYPE", "ORDQ", "ORDER", "LT", "EP",
              "SCM4", "MPCI", "SDCC1", "SDCC2", "MIN2", "word_c_serial"):
        out += struct.pack("<I", int(getattr(r, f)))
    out += r.min.split(b"\0")[0] + b"|"
    if with_dispatch:
        out += _pack_i(r.kind) + struct.pack("<I", r.esn) + r.dialed.split(b"\0")[0]
    return bytes(out)


def dispatch_tuple_oracle(r: O.ReccResult):
    """(kind, min, esn, dialed) with the parts the reference does not log for that kind blanked out."""
    k = int(r.kind)
    mn = r.min.split(b"\0")[0].decode()
    dialed = r.dialed.split(b"\0")[0].decode()
    if k == 4:
        return (4, mn, int(r.esn), dialed)
    if k in (2, 3):
        return (k, mn, None, None)
    return (k, None, None, None)


def dispatch_tuple_ref(info: dict):
    k = info["kind"]
    if k == 4:
        return (4, info["min"], int(info["esn"], 16), info["dialed"])
    if k in (2, 3):
        return (k, info["min"], None, None)
    return (k, None, None, None)


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


COMMANDS = ["fvc off", "fvc on", "fvc alert", "fvc offline", "page 2125551234", "PAGE 8005550000  ", "page ", "page 12ab",
            "page 12345678901", "page 0000000000", "bogus", "", "fvc", "page  4155550100\n"]

FOCC_CASES = [  # (symrate, aggressive, seed, total bytes, inject)
    (20000, False, 11, 200000, False), (20000, True, 12, 200000, False), (200000, False, 13, 1500000, False),
    (100000, False, 14, 1200000, True), (20000, False, 15, 300000, True), (100000, True, 16, 1000000, True),
]
FVC_CASES = [(20000, 21, 120), (100000, 22, 120), (200000, 23, 60)]   # (symrate, seed, calls)
RECC_CASES = [(31, 24, 4096), (32, 24, 61439), (33, 40, 300), (34, 12, 20000), (35, 30, 9000)]   # (seed, bursts, max chunk)
DECODE_CASES = [(41, 160), (42, 160)]   # (seed, blobs)


# ------------------------------------------------------------------------------------------------ BCH / words / MIN
def bch_decode_inputs(seed: int, count: int):
    """48-bit words: codewords with 0..4 flipped bits (anywhere in the 48) and random words."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        if i % 5 == 4:
            w = rng.integers(0, 2, 48).astype(np.uint8)
        else:
            w = np.asarray(synth.bch_encode(rng.integers(0, 2, 36)), np.uint8)
            for p in rng.choice(48, size=i % 5, replace=False):
                w[p] ^= 1
        out.append(w)
    return out


def word_builder_transcript(impl, prefix: str, seed: int) -> bytes:
    """lib/amps_packet.cc:26-95 with random field values (incl. values wider than the field)."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    for _ in range(200):
        out += impl.word(prefix + "focc_word1", int(rng.integers(0, 2)), int(rng.integers(0, 4)), int(rng.integers(0, 1 << 24))).tobytes()
        out += impl.word(prefix + "focc_word2_general", int(rng.integers(0, 1 << 10)), int(rng.integers(0, 32)), int(rng.integers(0, 8)), int(rng.integers(0, 32))).tobytes()
        out += impl.word(prefix + "fvc_word1_general", int(rng.integers(0, 4)), int(rng.integers(0, 32)), int(rng.integers(0, 8)), int(rng.integers(0, 32))).tobytes()
        out += impl.word(prefix + "focc_word2_voice_channel", int(rng.integers(0, 4)), int(rng.integers(0, 1 << 10)), int(rng.integers(0, 8)), int(rng.integers(0, 2048))).tobytes()
    return bytes(out)


def min_transcript(L, prefix: str, seed: int) -> bytes:
    """lib/amps_packet.h:277-366: parse_min / calc_min / extract_min_3 / compute_min_3 over valid, short and bad strings
    and over every 10-bit group value."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    parse = getattr(L, prefix + "parse_min")
    calc = getattr(L, prefix + "calc_min")
    ext = getattr(L, prefix + "extract_min_3")
    for v in range(1024):
        b = C.create_string_buffer(4)
        ext(v, b)
        out += b.raw[:3]
    cases = ["", "12ab567890", "12345678901", "0000000000", "9999999999", "x", "123456789012345678901234567890"]
    cases += ["".join(str(int(d)) for d in rng.integers(0, 10, 10)) for _ in range(300)]
    for s in cases:
        m1, m2 = C.c_uint64(0), C.c_uint64(0)
        ok = parse(s.encode(), C.byref(m1), C.byref(m2))
        out += _pack_i(ok)
        if ok and len(s) == 10:
            b = C.create_string_buffer(16)
            calc(m1.value, m2.value, b)
            out += struct.pack("<QQ", m1.value, m2.value) + b.value
    for _ in range(300):
        m1, m2 = int(rng.integers(0, 1 << 24)), int(rng.integers(0, 1 << 10))
        b = C.create_string_buffer(16)
        calc(m1, m2, b)
        out += b.value + b"|"
    return bytes(out)
