"""The C++ host layer (gr::amps::{focc,fvc,recc_iq,recc_decode} over the C ABI, gr_amps_b200/host/) driven by
its C++ QA program, compared with the oracle.  The `loop` scenario is the reference's closed loop
(grc/ampsbs.grc:4404-4470): IQ -> recc_iq -> bursts -> recc_decode -> focc_words / fvc_words -> focc / fvc."""
import json
import os
import subprocess

import numpy as np
import pytest

from gr_amps_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QA = os.path.join(ROOT, "gr_amps_b200", "host", "qa_blocks")


def run(args, cwd):
    return subprocess.run([QA] + [str(a) for a in args], cwd=cwd, check=True, capture_output=True, text=True, timeout=600).stdout


@pytest.mark.parametrize("symrate,aggr,seed", [(100000, 0, 1), (200000, 1, 7)])
def test_focc_block_work_schedule(oracle, tmp_path, symrate, aggr, seed):
    total = 500_000
    out = tmp_path / "focc.bin"
    run(["focc", symrate, aggr, total, seed, out], tmp_path)
    got = np.fromfile(out, np.uint8)
    # same LCG request schedule against the oracle block
    o = oracle.Focc(symrate, bool(aggr))
    lcg, ref = seed, bytearray()
    while len(ref) < total:
        lcg = (lcg * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        n = min(1 + (lcg >> 33) % 9000, total - len(ref))
        r, b = o.work(n)
        ref += b.tobytes()
    assert np.array_equal(got, np.frombuffer(bytes(ref), np.uint8))


@pytest.mark.parametrize("sc16", [False, True], ids=["fc32", "sc16"])
def test_closed_loop_flowgraph(oracle, tmp_path, sc16):
    """sc16: the recc_iq block takes interleaved int16 I,Q (item size 4) and converts on the GPU; the oracle is fed the
    floats those integers stand for."""
    period = 55 * 38400
    msgs = [synth.origination_words(min10="2125550101", dialed="4155551212"),
            synth.page_response_words(min10="2125550102"),
            synth.registration_words(min10="2125550103"),
            synth.origination_words(min10="2125550104", dialed="0")]
    parts = [synth.burst_period(w, n_total=period, snr_db=20.0, seed=40 + i)[0] for i, w in enumerate(msgs)]
    x = np.concatenate(parts + [np.zeros(38400, np.complex64)])
    iq = tmp_path / "iq.bin"
    x.tofile(iq)
    focc_bytes = 3 * 19 * 4630
    out = run(["loop", iq, len(x), 262144, focc_bytes, tmp_path / "loop"] + (["sc16"] if sc16 else []), tmp_path)
    lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    if sc16:
        q = np.clip(np.rint(x.view(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int16)
        x = (q.astype(np.float32) * np.float32(1.0 / 32768.0)).view(np.complex64)

    # oracle side of the same loop
    _, d = oracle.rx_chain_f32(x)
    bursts = oracle.rx_detect(d)
    assert len(bursts) == 4
    expect, ofocc, ofvc = [], oracle.Focc(100000, False), oracle.Fvc(100000)
    bits = lambda w: "".join(str(int(v)) for v in w)
    for _, _, blob in bursts:
        r = oracle.recc_decode(blob)
        a = oracle.recc_actions(r)
        if a.n_focc:
            words = [bits(a.focc_words[i]) for i in range(a.n_focc)]
            expect.append({"port": "focc_words", "stream": a.focc_stream, "n": a.n_focc, "words": words})
            ofocc.push_words(a.focc_stream, np.array([[int(c) for c in w] for w in words], np.uint8))
        if a.has_fvc:
            expect.append({"port": "fvc_words", "words": [bits(a.fvc_word)], "timer": a.fvc_timer})
            ofvc.push_words(np.array(list(a.fvc_word), np.uint8), timer=a.fvc_timer)
        if a.fvc_mute >= 0:
            expect.append({"port": "fvc_mute", "value": bool(a.fvc_mute)})
        if a.audio_mute >= 0:
            expect.append({"port": "audio_mute", "value": bool(a.audio_mute)})
        if a.command:
            expect.append({"port": "command_out", "text": a.command.decode()})
        expect.append({"port": "bursts", "len": 3374})
    assert lines == expect
    kinds = [oracle.recc_decode(b[2]).kind for b in bursts]
    assert kinds == [4, 2, 3, 4]
    assert [l["text"] for l in lines if l["port"] == "command_out"] == ["page 4155551212", "page 0"]

    got = np.fromfile(str(tmp_path / "loop.focc.bin"), np.uint8)
    assert np.array_equal(got, ofocc.generate(focc_bytes, chunk=65536))
    # the four injected word pairs occupy the first filler slots of the superframe, two frames each
    plain = oracle.Focc(100000, False).generate(focc_bytes, chunk=65536)
    fb = 4630
    changed = sorted(set(np.nonzero(got != plain)[0] // fb))
    assert changed == [4, 5, 6, 7, 8, 9, 10, 11]
    gv = np.fromfile(str(tmp_path / "loop.fvc.bin"), np.uint8)
    ref = bytearray()
    while len(ref) < 30000:
        r, b, _ = ofvc.work(4096)
        ref += b.tobytes()
    assert np.array_equal(gv, np.frombuffer(bytes(ref), np.uint8)[:len(gv)])


def test_forward_iq_block(oracle, tmp_path):
    """The composite forward_iq block (FOCC + FVC sources -> fused modulator, data-bit fast path) against the oracle:
    oracle focc/fvc byte streams -> float64 forward chain."""
    nbits = 4000
    out = run(["txblock", nbits, tmp_path / "tx.bin"], tmp_path)
    y = np.fromfile(str(tmp_path / "tx.bin"), np.complex64)
    assert len(y) == nbits * 1000
    w1 = [0,1,0,0, 0,0,0,1,0,0,1,0,0,0,1,1,0,1,0,0,0,1,0,1,0,1,1,0]
    w2 = [1,0,1,1, 0,1,0,1,0,1,0,1,0,1, 0, 0,0,0,0,0, 0,0,0, 0,0,0,0,0]
    alert = [1,0,1,1,0,1,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1]
    fo = oracle.Focc(100000, False)
    fo.push_words(3, np.array([w1, w2], np.uint8))
    s0 = fo.generate(nbits * 10, chunk=1 << 20)
    fv = oracle.Fvc(100000)
    fv.push_words(np.array(alert, np.uint8), timer=2)
    s1, offs = bytearray(), 0
    while len(s1) < nbits * 10:
        r, b, off = fv.work(min(8192, nbits * 10 - len(s1)))
        s1 += b.tobytes()
        offs += off
    s1 = np.frombuffer(bytes(s1), np.uint8).copy()
    s1[:nbits // 2 * 10] = 0                                   # the FVC leg is muted for the first half
    s2 = np.zeros(nbits * 10, np.uint8)
    ref = oracle.fwd_chain_f64([s0, s1, s2])
    err = float(np.sqrt(np.mean(np.abs(y.astype(np.complex128) - ref) ** 2)))
    assert err <= 1e-6, err
    lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    assert lines == [{"port": "command_out", "text": "fvc off"}] * offs and offs == 1


# ------------------------------------------------------------------ concurrency and malformed messages
def focc_frames(stream_bytes, sps=5):
    """FOCC byte stream (+1 = 0x01, -1 = 0xFF, sps bytes per half-symbol) -> list of (word A info bits, word B info bits, ok):
    frame = [BI][dotting 1010101010][BI][sync 11100010010] + 5 x (A, B) words, each 4 x ([BI] + 10 bits) (lib/focc_impl.cc:178-218);
    Manchester: bit 0 -> (+1, -1), bit 1 -> (-1, +1) (lib/amps_packet.h:52-70)."""
    hs = np.asarray(stream_bytes)[::sps]
    bits = (hs.reshape(-1, 2)[:, 0] == 0xFF).astype(np.uint8)
    frames = []
    for f in range(len(bits) // 463):
        b = bits[463 * f:463 * (f + 1)]
        ok = list(b[1:11]) == [1, 0] * 5 and list(b[12:23]) == [1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0]
        words = []
        for w in range(10):
            seg = b[23 + 44 * w:23 + 44 * (w + 1)].reshape(4, 11)[:, 1:].reshape(-1)       # drop the busy/idle bit in front of each 10
            words.append(seg)
        a, bb = words[0::2], words[1::2]
        ok = ok and all(np.array_equal(a[0], x) for x in a) and all(np.array_equal(bb[0], x) for x in bb)
        frames.append((a[0], bb[0], ok))
    return frames


def test_focc_words_from_another_thread_while_work_runs(oracle, tmp_path):
    """Messages posted from a second thread while the scheduler thread is inside focc::work(): the stream stays a valid
    sequence of frames, every word is BCH-consistent, and the 2 x nmsgs injected words appear exactly once, in order."""
    nmsgs, total = 40, 130 * 4630
    out = tmp_path / "thr.bin"
    run(["threads", total, nmsgs, out], tmp_path)
    frames = focc_frames(np.fromfile(out, np.uint8))
    assert len(frames) == 130 and all(ok for _, _, ok in frames)
    injected = []
    for a, b, _ in frames:
        for w in (a, b):
            assert np.array_equal(oracle.bch_encode_40_28(w[:28]), w)               # parity of every transmitted word
        v = int("".join(map(str, a[:28])), 2)
        if (v >> 20) == 0x5A:
            assert np.array_equal(a, b)                                              # stream BOTH: the word on A and on B
            injected.append(v & 0xFFFFF)
    assert injected == list(range(2 * nmsgs))


def test_malformed_word_messages_are_dropped(tmp_path):
    out = subprocess.run([QA, "badmsg"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert json.loads(out.stdout.strip().splitlines()[-1]) == {"fvc_work": 64, "fvc_buffer_untouched": True}
    assert out.stderr.count("dropped") == 7
