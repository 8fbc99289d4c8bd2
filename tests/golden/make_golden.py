#!/usr/bin/env python
"""Generates the fixtures under tests/golden/ from the oracle (oracle/liboracle.so) and the numpy
generator (gr_amps_b200/synth.py).  The reference ships no golden vectors (lib/qa_amps.cc:9-15 is an
empty suite) and cannot be built here, so these fixtures pin OUR restatement; the KAT values in
kat_bch.json / kat_focc.json were derived independently in SURVEY.md Appendix A and the oracle is
checked against them.

    python tests/golden/make_golden.py        (re-run after changing oracle/ or synth.py)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gr_amps_b200 import synth  # noqa: E402
from tests import oracle_lib as O  # noqa: E402


def bits(s):
    return [int(c) for c in s]


def main():
    # --- SURVEY App. A.2 KATs (independent derivation) -------------------------------------
    kat = {
        "OW1 nawc=3": ["1100000000000010001000011110", "010110101011"],
        "OW1 nawc=4": ["1100000000000010001000100110", "111110001101"],
        "OW2": ["1100111100101111100101110111", "111100101000"],
        "control filler": ["1100010111000001100111111001", "001000000011"],
        "access-type GA END=0": ["1100100100000000000000000100", "010101000100"],
        "REGINCR=100 END=0": ["1100001000000110010000000100", "100110010010"],
        "REGID=0 END=1": ["1100000000000000000000001000", "110100011010"],
        "REGID=500 END=1": ["1100000000000001111101001000", "001100010100"],
        "FVC alert order scc=1": ["1011010000000000000000000001", "110111110111"],
    }
    json.dump(kat, open(os.path.join(HERE, "kat_bch.json"), "w"), indent=1)
    json.dump({
        "first48_symrate20000": "ff01ff0101ffff0101ffff0101ffff0101ffff0101ffff01ff01ff01ff0101ff01ff01ffff0101ff01ffff0101ffff01",
        "superframe_bytes_sps1": 17594,
        "superframe_sha256_sps1": "15840e34af8a8ece0615bbb6815e2ea984f46cdf748039e256f44513824705cf",
        "recc_trigger": "01100110011001100110011001100110011001100110011001100101011010100110100110",
    }, open(os.path.join(HERE, "kat_focc.json"), "w"), indent=1)

    # --- oracle-generated fixtures ------------------------------------------------------------
    f = O.Focc(20000, False)
    sf = f.generate(3 * 17594)
    open(os.path.join(HERE, "focc_3superframes_sps1.bin"), "wb").write(np.packbits((sf == 1).astype(np.uint8)).tobytes())
    fa = O.Focc(20000, True)
    sfa = fa.generate(38 * 926)
    fix = {"focc_3superframes_sps1_sha256": hashlib.sha256(sf.tobytes()).hexdigest(),
           "focc_aggressive_superframe_sps1_sha256": hashlib.sha256(sfa.tobytes()).hexdigest()}

    v = O.Fvc(100000)
    w = np.zeros(28, np.uint8)
    O.lib().orc_fvc_word1_general(O.ptr(w, O.u8p), 1, 0, 0, 1)
    v.push_words(w)
    out = bytearray()
    while len(out) < 2064 * 5:
        r, b, _ = v.work(4096)
        out += b.tobytes()
    fix["fvc_alert_train_sps5_sha256"] = hashlib.sha256(bytes(out[:10320])).hexdigest()
    fix["fvc_alert_word"] = "".join(str(int(x)) for x in w)

    # synthetic 7-word origination burst: transmitted half-symbols + expected decode
    words = synth.origination_words()
    msg = synth.recc_message_bits(words)
    hs = synth.manchester(msg)
    blob = hs[82:82 + 3374]
    r = O.recc_decode(blob)
    np.save(os.path.join(HERE, "recc_origination_halfsymbols.npy"), hs)
    fix["recc_origination"] = {
        "blob_sha256": hashlib.sha256(blob.tobytes()).hexdigest(),
        "valid": list(r.valid), "errs": list(r.errs), "kind": r.kind,
        "min": r.min.decode(), "dialed": r.dialed.decode(), "esn": r.esn,
        "NAWC": r.NAWC, "T": r.T, "S": r.S, "E": r.E, "SCM": r.SCM, "MIN1": r.MIN1, "MIN2": r.MIN2,
    }
    # RX chain: one config-2 period, SNR 15 dB -> detection position, correlation, demod checksum
    x, hs2, _ = synth.config2_period(n_total=55 * 38400, snr_db=15.0, seed=0xA3B5)
    y, d = O.rx_chain_f32(x)
    b = O.rx_detect(d)
    fix["rx_config2_snr15"] = {
        "n": len(x), "d_sha256": hashlib.sha256(d.tobytes()).hexdigest(),
        "bursts": [[int(p), float(np.float32(c)), hashlib.sha256(s.tobytes()).hexdigest()] for p, c, s in b],
    }
    taps = O.lpf_taps()
    fix["lpf_taps_sha256"] = hashlib.sha256(taps.tobytes()).hexdigest()
    fix["lpf_taps_first8"] = [float(t) for t in taps[:8]]
    json.dump(fix, open(os.path.join(HERE, "oracle_fixtures.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
