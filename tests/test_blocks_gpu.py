"""GPU parity of the byte-level blocks (focc, fvc, recc, recc_decode) against the oracle: bit-exact,
same call schedules.  Everything goes through the C ABI."""
import os

import numpy as np
import pytest

from gr_amps_b200 import synth
from tests.helpers import words_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def capi():
    from gr_amps_b200 import capi as c
    c.lib()
    return c


# ------------------------------------------------------------------ focc
def test_focc_config1_million_symbols(capi, oracle):
    """BASELINE config 1 on the GPU source: 1e6 half-symbols at symrate 20000, byte-for-byte."""
    ref = oracle.Focc(20000, False).generate(1_000_000, chunk=4096)
    g = capi.Focc(20000, False)
    assert np.array_equal(g.generate(1_000_000), ref)
    g2 = capi.Focc(20000, False)
    out = bytearray()
    while len(out) < 200_000:
        r, b = g2.work(4096)
        assert 0 <= r <= 46
        out += b.tobytes()
    assert np.array_equal(np.frombuffer(bytes(out), np.uint8), ref[:len(out)])


@pytest.mark.parametrize("symrate,aggr", [(100000, False), (200000, True), (10_000_000, False)])
def test_focc_work_schedule_parity(capi, oracle, symrate, aggr):
    rng = np.random.default_rng(symrate % 97 + aggr)
    o, g = oracle.Focc(symrate, aggr), capi.Focc(symrate, aggr)
    sps = symrate // 20000
    for step in range(400):
        n = int(rng.choice([1, 7, sps, 2 * sps, 22 * 2 * sps, 23 * 2 * sps, 46 * sps + 3, 100000]))
        if step == 50:       # inject words mid-stream on every stream type
            w1 = oracle.word("orc_focc_word1", 1, 0, 0xABCDE)
            w2 = oracle.word("orc_focc_word2_voice_channel", 1, 0x155, 0, 355)
            for st in (1, 2, 3):
                o.push_words(st, np.concatenate([w1, w2]))
                g.push_words(st, np.concatenate([w1, w2]))
        ro, bo = o.work(n)
        rg, bg = g.work(n)
        assert ro == rg and np.array_equal(bo, bg), (step, n, ro, rg)
    assert g.work(0)[0] == -1 and o.work(0)[0] == -1


def test_focc_generate_then_work_state(capi, oracle):
    o, g = oracle.Focc(100000, False), capi.Focc(100000, False)
    for n in (1, 229, 230, 231, 4630 * 19 + 17, 5):
        a = o.generate(n, chunk=1 << 20)
        b = g.generate(n)
        assert np.array_equal(a, b)
        ro, bo = o.work(1000)
        rg, bg = g.work(1000)
        assert ro == rg and np.array_equal(bo, bg)


def test_focc_busy_idle_and_device_output(capi, oracle):
    torch = pytest.importorskip("torch")
    g = capi.Focc(100000, False)
    g.set_busy_idle(False)
    out = g.generate(4630)
    ref = oracle.Focc(100000, False).generate(4630)
    hs_ref = ref.reshape(-1, 5)[:, 0]
    hs = out.reshape(-1, 5)[:, 0]
    bi_slots = [0, 11] + [23 + 11 * i for i in range(40)]
    diff = np.nonzero(hs != hs_ref)[0]
    assert sorted(set(diff // 2)) == bi_slots                    # only the busy/idle bits flipped
    g2 = capi.Focc(10_000_000, False)
    t = torch.empty(5_000_000, dtype=torch.uint8, device="cuda")
    g2.generate_dev(t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy(), oracle.Focc(10_000_000, False).generate(5_000_000, chunk=1 << 22))


def test_focc_bits_match_bytes(capi, oracle):
    """generate_bits is the byte stream seen as data bits, and leaves the state where the bytes would."""
    g = capi.Focc(100000, False)
    w = oracle.word("orc_focc_word1", 1, 0, 0x13579)
    g.push_words(3, w)
    o = oracle.Focc(100000, False)
    o.push_words(3, w)
    ref = o.generate(463 * 25 * 10, chunk=1 << 20)               # 25 frames
    hs = ref.reshape(-1, 10)
    assert np.array_equal(g.generate_bits(463 * 20), (hs[:463 * 20, 5] == 1).astype(np.uint8))
    assert np.array_equal(g.generate(4630 * 5), ref[4630 * 20:])   # the byte stream continues where the bits stopped
    assert g.work(3)[0] == 0                                     # the frame just ended: this call only steps over FOCC_END
    assert g.work(3)[0] == 3
    with pytest.raises(capi.AmpsError):
        g.generate_bits(10)                                      # not on a bit boundary


# ------------------------------------------------------------------ fvc
def test_fvc_parity(capi, oracle):
    rng = np.random.default_rng(8)
    o, g = oracle.Fvc(100000), capi.Fvc(100000)
    r, b, off = g.work(512)
    assert r == 512 and not off and np.all(b == 0x55)            # idle: claims n, buffer untouched (lib/fvc_impl.cc:159-161)
    alert = oracle.word("orc_fvc_word1_general", 1, 0, 0, 1)
    for step in range(300):
        if step == 0:
            o.push_words(alert, timer=3); g.push_words(alert, timer=3)
        if step == 120:
            w = np.concatenate([oracle.word("orc_fvc_word1_general", 1, 0, 0, 3), alert])
            o.push_words(w); g.push_words(w)
        n = int(rng.integers(1, 9000))
        ro, bo, fo = o.work(n)
        rg, bg, fg = g.work(n)
        assert (ro, fo) == (rg, fg) and np.array_equal(bo, bg), step


# ------------------------------------------------------------------ recc (compat)
def _stream(rng, n_bursts, gaps):
    hs = np.load(os.path.join(GOLD, "recc_origination_halfsymbols.npy"))
    parts = []
    for i in range(n_bursts):
        parts += [rng.integers(0, 2, gaps[i % len(gaps)]).astype(np.uint8), hs]
    parts.append(rng.integers(0, 2, 9000).astype(np.uint8))
    return np.concatenate(parts)


@pytest.mark.parametrize("seed,maxchunk", [(1, 512), (2, 4096), (3, 20000), (4, 61439)])
def test_recc_chunk_schedule_parity(capi, oracle, seed, maxchunk):
    """Same chunk schedule -> same blobs, including the wrap / stale-buffer quirks (lib/recc_impl.cc:104-134)."""
    rng = np.random.default_rng(seed)
    s = _stream(rng, 12, [3000, 40000, 500, 61000, 12000])
    o, g = oracle.Recc(), capi.Recc()
    pos, sizes = 0, []
    while pos < len(s):
        n = int(rng.integers(1, maxchunk + 1))
        n = min(n, len(s) - pos)
        o.work(s[pos:pos + n])
        g.work(s[pos:pos + n])
        sizes.append(n)
        pos += n
        assert len(o.bursts) == len(g.bursts), (pos, n)
    assert len(o.bursts) >= 1
    for a, b in zip(o.bursts, g.bursts):
        assert np.array_equal(a, b)
    # the whole schedule in ONE launch gives the same result
    g2 = capi.Recc()
    g2.work_chunks(s, sizes)
    assert len(g2.bursts) == len(o.bursts) and all(np.array_equal(a, b) for a, b in zip(o.bursts, g2.bursts))


def test_recc_rejects_oversized_chunk(capi):
    g = capi.Recc()
    with pytest.raises(capi.AmpsError):
        g.work(np.zeros(61440, np.uint8))


# ------------------------------------------------------------------ recc_decode
def test_decode_batch_with_errors(capi, oracle):
    rng = np.random.default_rng(10)
    hs = np.load(os.path.join(GOLD, "recc_origination_halfsymbols.npy"))
    base = hs[82:82 + 3374]
    blobs = []
    for i in range(96):
        b = base.copy()
        nflip = int(rng.integers(0, 60)) if i else 0
        idx = rng.choice(3374, nflip, replace=False)
        b[idx] ^= 1
        if i % 7 == 3:                                   # T / E / ORDER variations -> other message classes
            for bit in rng.choice(36, 3, replace=False):
                for rep in range(5):
                    j = 14 + 2 * (48 * rep + int(bit))
                    b[j], b[j + 1] = b[j + 1], b[j]
        blobs.append(b)
    blobs = np.stack(blobs)
    dec = capi.ReccDecode()
    out = dec.decode(blobs)
    kinds = set()
    for b, g in zip(blobs, out):
        r = oracle.recc_decode(b)
        assert words_equal(g, r) == []
        kinds.add(r.kind)
    assert 4 in kinds and len(kinds) >= 2


def test_bch_validity_matches_oracle_on_random_error_patterns(capi, oracle):
    """Every 48-bit repeat of every word is an independent BCH(63,51) validity check: fill blobs with
    codewords hit by 0..4 errors (includes the S1 == 0 / S3-cube case IT++ accepts) and compare."""
    rng = np.random.default_rng(12)
    n_blobs = 64
    blobs = np.zeros((n_blobs, 3374), np.uint8)
    expect = []
    for bi in range(n_blobs):
        bits = np.zeros(7 + 7 * 240, np.uint8)
        for w in range(7):
            for rep in range(5):
                cw = oracle.bch_encode_48_36(rng.integers(0, 2, 36).astype(np.uint8))
                nerr = int(rng.integers(0, 5))
                cw[rng.choice(48, nerr, replace=False)] ^= 1
                bits[7 + 240 * w + 48 * rep: 7 + 240 * w + 48 * rep + 48] = cw
        blobs[bi] = synth.manchester(bits)
    dec = capi.ReccDecode()
    out = dec.decode(blobs)
    nvalid = 0
    for b, g in zip(blobs, out):
        r = oracle.recc_decode(b)
        assert list(g.valid) == list(r.valid) and list(g.valid_repeat) == list(r.valid_repeat)
        nvalid += sum(r.valid)
    assert 0 < nvalid
    # exhaustive check of the single/double/triple-error rule on one codeword, all repeats independent
    cw = oracle.bch_encode_48_36(rng.integers(0, 2, 36).astype(np.uint8))
    pats = []
    for _ in range(35 * 40):
        e = cw.copy()
        e[rng.choice(48, 3, replace=False)] ^= 1
        pats.append(e)
    pats = np.array(pats).reshape(40, 35, 48)
    blobs = np.zeros((40, 3374), np.uint8)
    for bi in range(40):
        bits = np.zeros(7 + 1680, np.uint8)
        bits[7:] = pats[bi].reshape(-1)
        blobs[bi] = synth.manchester(bits)
    out = dec.decode(blobs)
    for bi in range(40):
        r = oracle.recc_decode(blobs[bi])
        assert list(out[bi].valid) == list(r.valid) and list(out[bi].valid_repeat) == list(r.valid_repeat)


def test_focc_queue_capacity_keeps_the_frame_in_flight(capi, oracle):
    """The pool of ephemeral frames holds 4096; one slot stays reserved for the frame that has been popped and is being
    transmitted, so 4095 words can be queued, the 4096th is refused (AMPS_E_OVERFLOW) -- and a device-side generate that is
    still reading a slot is waited for before the slot is rewritten (the stream stays the oracle's)."""
    torch = pytest.importorskip("torch")
    f, o = capi.Focc(20000, False), oracle.Focc(20000, False)
    w = lambda i: np.array([int(c) for c in format(0x5A00000 | i, "028b")], np.uint8)
    many = np.concatenate([w(i) for i in range(4095)])
    f.push_words(3, many)
    o.push_words(3, many.reshape(-1, 28))
    with pytest.raises(capi.AmpsError) as e:
        f.push_words(3, w(4095))
    assert e.value.status == -6
    # drain part of the queue through the device-resident entry point, push again right behind it, keep generating
    n1 = 300 * 926
    t = torch.empty(n1, dtype=torch.uint8, device="cuda")
    f.generate_dev(t.data_ptr(), n1, torch.cuda.current_stream().cuda_stream)
    more = np.concatenate([w(5000 + i) for i in range(200)])
    f.push_words(3, more)                                  # rewrites pool slots the generate above may still be reading
    o_first = o.generate(n1, chunk=4096)
    o.push_words(3, more.reshape(-1, 28))
    assert np.array_equal(t.cpu().numpy(), o_first)
    n2 = 6000 * 926
    assert np.array_equal(f.generate(n2), o.generate(n2, chunk=1 << 16))
