"""CPU tests of the PRODUCT's host-side helpers (gr_amps_b200/csrc/proto.cc, design.cc) through host/qa_proto:
word builders + BCH(40,28) against the golden KATs, frame/train layouts and filter designs against the oracle."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def proto():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gr_amps_b200"), "host/qa_proto"])
    out = subprocess.check_output([os.path.join(ROOT, "gr_amps_b200", "host", "qa_proto")], text=True)
    return json.loads(out)


def test_words_and_bch_match_kats(proto, oracle):
    kat = json.load(open(os.path.join(GOLD, "kat_bch.json")))
    for name, (info, parity) in kat.items():
        assert proto["words"][name] == [info, parity], name
    # the message words recc_decode's responses are built from, against the oracle builders
    bits = lambda w: "".join(map(str, w))
    for name, w in [("focc_word1", oracle.word("orc_focc_word1", 1, 0, 0xABCDE)),
                    ("focc_word2_general", oracle.word("orc_focc_word2_general", 0x155, 0, 0, 7)),
                    ("focc_word2_voice_channel", oracle.word("orc_focc_word2_voice_channel", 1, 0x2AA, 0, 355))]:
        assert proto["words"][name][0] == bits(w)
        assert proto["words"][name][1] == bits(oracle.bch_encode_40_28(w)[28:])


def test_frame_and_train_layouts(proto, oracle):
    # frame 0 of the default superframe, B/I slots marked '2': compare with the oracle's byte stream at sps = 1
    ref = oracle.Focc(20000, False).generate(926)
    bits_ref = (ref.reshape(-1, 2)[:, 1] == 1).astype(int)
    slots = proto["frame_slots"]
    assert len(slots) == 463 and slots.count("2") == 42
    assert all(s == "2" or int(s) == b for s, b in zip(slots, bits_ref))
    assert all(bits_ref[i] == 1 for i, s in enumerate(slots) if s == "2")        # busy/idle = idle = 1
    v = oracle.Fvc(20000)
    v.push_words(oracle.word("orc_fvc_word1_general", 1, 0, 0, 1))
    r, b, _ = v.work(4096)
    assert r == 2064
    assert proto["fvc_train"] == "".join(str(int(x)) for x in (b.reshape(-1, 2)[:, 1] == 1))


def test_firdes_reproduces_gnuradios_own_qa_vector(proto, oracle):
    """firdes.low_pass is GNU Radio code that is not in the reference tree; GNU Radio's QA suite publishes the taps of
    low_pass(1, 1, 0.4, 0.2) (qa_firdes.py, test_low_pass).  Both the product's design code and the oracle's restatement
    reproduce them bit for bit (float32)."""
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_gnuradio_firdes.json")))
    want = np.asarray(kat["taps"], np.float64).astype(np.float32)
    assert np.array_equal(want.astype(np.float64), np.asarray(kat["taps"]))       # the published values ARE float32 numbers
    assert np.array_equal(np.float32(proto["taps"]["gr_qa_firdes_low_pass"]), want)
    assert np.array_equal(oracle.firdes_low_pass(1.0, 1.0, 0.4, 0.2, 0), want)


def test_filter_designs_match_oracle(proto, oracle):
    L = oracle.lib()
    assert proto["fcw"] == [L.orc_nco_fcw(-160e3, 10e6), L.orc_nco_fcw(-160e3, 400e3)]
    t = proto["taps"]
    assert np.array_equal(np.float32(t["lpf"]), oracle.lpf_taps()) and len(t["lpf"]) == 299
    assert np.array_equal(np.float32(t["focc_interp"]), oracle.firdes_low_pass(1.0, 400e3, 10e3, 5e3, 0)) and len(t["focc_interp"]) == 193
    assert np.array_equal(np.float32(t["fvc_interp"]), oracle.firdes_low_pass(1.0, 400e3, 10e3, 3e3, 0)) and len(t["fvc_interp"]) == 321
    assert np.array_equal(np.float32(t["mmse"]).reshape(129, 8), oracle.mmse_table())
    # voice leg designs: pre-emphasis IIR, its impulse response, and the x25 form of the arb resampler
    L.orc_fm_preemph_taps.argtypes = [C.c_double] * 3 + [C.POINTER(C.c_double)] * 2
    b, a = (C.c_double * 2)(), (C.c_double * 2)()
    L.orc_fm_preemph_taps(16000.0, 75e-6, -1.0, b, a)
    assert np.allclose(t["preemph"], [b[0], b[1], a[0], a[1]], rtol=1e-14, atol=0)
    imp = np.zeros(192); y = 0.0
    for k in range(192):
        y = (b[0] if k == 0 else b[1] if k == 1 else 0.0) - a[1] * y
        imp[k] = y
    assert np.allclose(t["preemph_impulse"], imp, rtol=1e-13, atol=1e-30) and abs(imp[-1]) < 1e-18
    vt = oracle.voice_lpf_taps()
    assert len(vt) == 225 and np.array_equal(np.float32(t["voice_lpf"]), vt)
    E = np.zeros(25 * 29)
    L.orc_arb25_taps.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    assert L.orc_arb25_taps(vt.ctypes.data, 225, E.ctypes.data) == 29
    assert np.array_equal(np.float32(t["voice_arb25"]), E.astype(np.float32))
    c = np.float32(t["cic25"])
    assert len(c) == 73 and abs(float(c.astype(np.float64).sum()) - 1.0) < 1e-6 and np.array_equal(c, c[::-1])
    assert c[0] == np.float32(1 / 15625) and c[36] == np.float32(469 / 15625)


def test_command_processor_matches_oracle(oracle):
    """gr::amps::command_processor (host only) against the oracle's restatement of lib/command_processor_impl.cc:52-117."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "gr_amps_b200"), "host"])
    cmds = ["fvc off", "fvc on now", "fvc alert", "PAGE  2125551234 ", "page 0000000000", "page 9075550199\n", "page 12", "page ",
            "page", "Fvc off", "page 212555123x", "page 21255512345", "", "fvc", "pAgE 3105550100"]
    out = subprocess.check_output([os.path.join(ROOT, "gr_amps_b200", "host", "qa_blocks"), "cmd"] + cmds, text=True)
    lines = out.strip().split("\n")
    assert len(lines) == len(cmds)
    bits = lambda w: "".join(str(int(b)) for b in w)
    n_pages = 0
    for cmd, line in zip(cmds, lines):
        got = json.loads(line)
        a = oracle.command_actions(cmd)
        want = []                                            # publication order of the reference
        if a.fvc_mute >= 0:
            want += [{"port": "fvc_mute", "value": bool(a.fvc_mute)}, {"port": "audio_mute", "value": bool(a.audio_mute)}]
        if a.has_fvc:
            want.append({"port": "fvc_words", "words": [bits(a.fvc_word)]})
        dbg = [{"port": "debug_output", "text": a.debug[i].value.decode()} for i in range(a.n_debug)]
        if a.n_focc:
            want += dbg + [{"port": "focc_words", "stream": a.focc_stream, "n": a.n_focc, "words": [bits(a.focc_words[i]) for i in range(a.n_focc)]}]
            n_pages += 1
        else:
            want += dbg
        assert got == want, cmd
        # and, where the compiled reference is at hand, the reference's own command_processor_impl says the same (short
        # MINs excepted: the reference reads past the end of the string there, tests/test_ref_pin_cpu.py)
        from tests import ref_lib as R
        digits = cmd.strip()[5:].strip() if cmd.lower().startswith("page ") else ""
        if R.available() and not (digits.isdigit() and len(digits) < 10):
            assert bytes(R.command_actions(cmd)) == bytes(a), cmd
    assert n_pages == 4
    # the page words carry the MIN: decode them back with the oracle's MIN arithmetic
    a = oracle.command_actions("page 2125551234")
    w1, w2 = np.array(a.focc_words[0]), np.array(a.focc_words[1])
    min1 = int("".join(map(str, w1[4:28])), 2)
    min2 = int("".join(map(str, w2[4:14])), 2)
    buf = (C.c_char * 11)()
    oracle.lib().orc_calc_min(min1, min2, buf)
    assert buf.value == b"2125551234"


# ------------------------------------------------------------------ the block surface against the reference's own headers
REF = os.environ.get("AMPS_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "amps")), reason="the reference tree is not here")
def test_host_layer_compiles_against_the_references_public_headers():
    """gr::amps::{focc,fvc,recc,recc_decode,command_processor} are declared by the REFERENCE's include/amps/*.h in this build
    (boost::shared_ptr sptrs, AMPS_API through gnuradio/attributes.h, virtual public gr::sync_block / gr::block); the host
    layer and the QA driver compile and link against those declarations, and the host-only scenario gives the same answers
    as the regular build.  (GNU Radio 3.7 itself is not installed: its base classes and Boost come from host/gr_shim.)"""
    d = os.path.join(ROOT, "gr_amps_b200")
    subprocess.check_call(["make", "-s", "-C", d, "libamps_b200.so", "host", "host/qa_blocks_refhdr", "REF=" + REF])
    src = open(os.path.join(REF, "include", "amps", "focc.h")).read()
    assert "boost::shared_ptr<focc> sptr" in src
    cmds = ["page 2125551234", "fvc on", "fvc off", "fvc alert", "page 12345", "bogus"]
    a = subprocess.check_output([os.path.join(d, "host", "qa_blocks"), "cmd"] + cmds, text=True)
    b = subprocess.check_output([os.path.join(d, "host", "qa_blocks_refhdr"), "cmd"] + cmds, text=True)
    assert a == b and a.count("\n") == len(cmds) and "paging!" in a
    # our own public headers spell the pointer type the same way
    for h in ("focc", "fvc", "recc", "recc_decode", "command_processor", "recc_iq", "forward_iq"):
        assert "boost::shared_ptr<%s> sptr" % h in open(os.path.join(d, "host", "include", "amps", h + ".h")).read()
