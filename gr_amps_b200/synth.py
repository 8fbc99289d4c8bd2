"""Synthetic RECC test-signal generator (numpy).  Not part of the reference: gr-amps has no
mobile-side transmitter and ships no captures (grc/recctest.grc:591 points at a file that is not
in the repository), so BASELINE config 2 is driven by this generator.

Message layout follows what the reference's receiver expects:
  * seizure precursor: 30-bit dotting 1010..10, 11-bit word sync 11100010010, 7-bit coded DCC
    (lib/recc_impl.cc:76 keeps the last 26 dotting bits + word sync as its trigger);
  * up to 7 words, each 36 info + 12 BCH parity bits, each repeated 5 times
    (lib/recc_decode_impl.cc:89-107 reads dcc at symbols 0..13 and word w at 14+480w);
  * Manchester: bit 0 -> (high, low), bit 1 -> (low, high) (lib/recc_impl.cc:51-65,
    lib/utils.cc:36-50); "high" = positive frequency deviation;
  * FSK/FM with +-8 kHz peak deviation (grc/ampsbs.grc:209), 20 k half-symbols/s.

This module is independent of oracle/ and of the CUDA path; both are fed by it.
"""
from __future__ import annotations

import numpy as np

HALF_SYMBOL_RATE = 20000.0
G_BCH = 0b1010100111001  # x^12+x^10+x^8+x^5+x^4+x^3+1 (octal 12471)
WORD_SYNC = [1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0]
TRIGGER_SYMS = 74
CAPTURE_SYMS = 3374


def bits_msb(val: int, n: int) -> list[int]:
    return [(val >> (n - 1 - i)) & 1 for i in range(n)]


def bch_parity12(info_bits) -> list[int]:
    """(m(x) * x^12) mod g(x) by plain GF(2) long division -- independent of the oracle's LFSR."""
    m = 0
    for b in info_bits:
        m = (m << 1) | (int(b) & 1)
    m <<= 12
    n = len(info_bits) + 12
    for i in range(n - 1, 11, -1):
        if (m >> i) & 1:
            m ^= G_BCH << (i - 12)
    return bits_msb(m & 0xFFF, 12)


def bch_encode(info_bits) -> list[int]:
    return [int(b) & 1 for b in info_bits] + bch_parity12(info_bits)


def min_to_fields(min10: str) -> tuple[int, int]:
    """10-digit MIN -> (MIN1 24 bits, MIN2 10 bits), TIA/EIA-553 2.3.1 (inverse of
    lib/amps_packet.h:277-302)."""
    def d(c):
        v = ord(c) - 48
        return 10 if v == 0 else v

    def three(a, b, c):
        return 100 * d(a) + 10 * d(b) + d(c) - 111

    assert len(min10) == 10 and min10.isdigit()
    min2 = three(*min10[0:3])
    thous = d(min10[6])
    min1 = (three(*min10[3:6]) << 14) | (thous << 10) | three(*min10[7:10])
    return min1, min2


def digit_code(c: str) -> int:
    if c == "0":
        return 10
    if c == "*":
        return 11
    if c == "#":
        return 12
    return ord(c) - 48


def origination_words(min10="2125551234", esn=0x82ABCDEF, dialed="18005551212", scm=0b0010) -> list[list[int]]:
    """7-word origination: A, B, C(serial), 4 x called-address (36 info bits each)."""
    min1, min2 = min_to_fields(min10)
    # word A: F=1 NAWC=6 T=1 S=1 E=1 ER=0 SCM MIN1   (lib/amps_packet.h:154-161)
    wa = [1] + bits_msb(6, 3) + [1, 1, 1, 0] + bits_msb(scm, 4) + bits_msb(min1, 24)
    # word B: F=0 NAWC=5 LOCAL=0 ORDQ=0 ORDER=0 LT=0 EP=0 SCM4=0 MPCI=0 SDCC1=0 SDCC2=0 MIN2 (:177-188)
    wb = [0] + bits_msb(5, 3) + bits_msb(0, 5) + bits_msb(0, 3) + bits_msb(0, 5) + [0, 0, 0] + bits_msb(0, 2) + bits_msb(0, 2) + bits_msb(0, 2) + bits_msb(min2, 10)
    wc = [0] + bits_msb(4, 3) + bits_msb(esn, 32)
    digs = [digit_code(c) for c in dialed] + [0] * 32
    words = [wa, wb, wc]
    for w in range(4):
        v = 0
        for k in range(8):
            v = (v << 4) | digs[8 * w + k]
        words.append([0] + bits_msb(3 - w, 3) + bits_msb(v, 32))
    assert all(len(w) == 36 for w in words)
    return words


def pad_words(words, n=7) -> list[list[int]]:
    """A RECC capture is always 7 words long (lib/recc_impl.cc:70); shorter messages are followed by idle (all-zero) words."""
    return list(words) + [[0] * 36 for _ in range(n - len(words))]


def page_response_words(min10="2125551234", scm=0b0010) -> list[list[int]]:
    """T=0 response with all-zero order fields -> recc_decode treats it as a page response (lib/recc_decode_impl.cc:121)."""
    min1, min2 = min_to_fields(min10)
    wa = [1] + bits_msb(1, 3) + [0, 0, 1, 0] + bits_msb(scm, 4) + bits_msb(min1, 24)
    wb = [0] + bits_msb(0, 3) + bits_msb(0, 5) + bits_msb(0, 3) + bits_msb(0, 5) + [0, 0, 0] + bits_msb(0, 6) + bits_msb(min2, 10)
    return pad_words([wa, wb])


def registration_words(min10="2125551234", esn=0x82ABCDEF, scm=0b0010) -> list[list[int]]:
    """T=1, ORDER=01101: word-C-included registration order (lib/recc_decode_impl.cc:123-138)."""
    min1, min2 = min_to_fields(min10)
    wa = [1] + bits_msb(2, 3) + [1, 1, 1, 0] + bits_msb(scm, 4) + bits_msb(min1, 24)
    wb = [0] + bits_msb(1, 3) + bits_msb(0, 5) + bits_msb(0, 3) + bits_msb(0xd, 5) + [0, 0, 0] + bits_msb(0, 6) + bits_msb(min2, 10)
    wc = [0] + bits_msb(0, 3) + bits_msb(esn, 32)
    return pad_words([wa, wb, wc])


def recc_message_bits(words36, dcc7=(0, 0, 0, 0, 0, 0, 0)) -> np.ndarray:
    bits = [1, 0] * 15 + WORD_SYNC + list(dcc7)
    for w in words36:
        enc = bch_encode(w)
        bits += enc * 5
    return np.asarray(bits, dtype=np.uint8)


def manchester(bits: np.ndarray) -> np.ndarray:
    """bit 0 -> (1, 0), bit 1 -> (0, 1) hard half-symbols."""
    out = np.empty(2 * len(bits), dtype=np.uint8)
    out[0::2] = 1 - bits
    out[1::2] = bits
    return out


def trigger_symbols() -> np.ndarray:
    return manchester(np.asarray([1, 0] * 13 + WORD_SYNC, dtype=np.uint8))


def fm_burst(halfsyms: np.ndarray, n_total: int, lead: int, samp_rate=10e6, center=-160e3, dev=8e3, amp=0.5,
             snr_db=None, seed=0xA3B5, chan_bw=30e3, carrier_phase=0.0) -> np.ndarray:
    """Complex64 baseband of one FM burst placed `lead` samples into an n_total-sample buffer.
    Outside the burst there is no carrier (noise only)."""
    sps = int(round(samp_rate / HALF_SYMBOL_RATE))
    nrz = np.repeat(halfsyms.astype(np.float64) * 2.0 - 1.0, sps)
    nb = len(nrz)
    assert lead + nb <= n_total, "burst does not fit"
    phi = np.cumsum(nrz) * (2.0 * np.pi * dev / samp_rate)
    n = np.arange(lead, lead + nb, dtype=np.float64)
    ang = 2.0 * np.pi * ((center / samp_rate * n) % 1.0) + phi + carrier_phase
    x = np.zeros(n_total, dtype=np.complex64)
    x[lead:lead + nb] = (amp * np.exp(1j * ang)).astype(np.complex64)
    if snr_db is not None:
        rng = np.random.default_rng(seed)
        sigma2 = (amp * amp) / (10.0 ** (snr_db / 10.0)) * (samp_rate / chan_bw)
        s = np.float32(np.sqrt(sigma2 / 2.0))
        x.real += s * rng.standard_normal(n_total, dtype=np.float32)
        x.imag += s * rng.standard_normal(n_total, dtype=np.float32)
    return x


def burst_period(words36, n_total=55 * 38400, lead=20000, snr_db=None, seed=0xA3B5, center=-160e3):
    """One period carrying an arbitrary 7-word RECC message."""
    hs = manchester(recc_message_bits(words36))
    return fm_burst(hs, n_total, lead, center=center, snr_db=snr_db, seed=seed), hs


def config2_period(n_total=1 << 21, lead=20000, snr_db=None, seed=0xA3B5, center=-160e3, **kw):
    """One BASELINE config-2 period: a 7-word origination burst in n_total samples at 10 MS/s.
    (SURVEY 8d says 'one burst per 2^19 samples'; a full 7-word burst is 3456 half-symbols =
    1.728 M samples at 10 MS/s, so the period is 2^21.)"""
    words = origination_words(**kw)
    bits = recc_message_bits(words)
    hs = manchester(bits)
    x = fm_burst(hs, n_total, lead, center=center, snr_db=snr_db, seed=seed)
    return x, hs, words
