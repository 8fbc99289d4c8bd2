"""Writes tests/golden/ref_vectors.json from the REFERENCE ITSELF run here: oracle/_ref/libamps_ref.so is gr-amps's own
lib/*.cc compiled from /root/reference (make -C oracle _ref) behind oracle/ref_harness.cc.  Run in the container that
has /root/reference:   python -m tests.golden.make_ref_golden
The file pins oracle/*.c (tests/test_ref_pin_cpu.py) wherever the compiled reference is not at hand."""
import json
import os

import numpy as np

from tests import ref_cases as K
from tests import ref_lib as R

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")


def main():
    R.lib()
    g = {"generator": "tests/golden/make_ref_golden.py", "source": "gr-amps lib/*.cc via oracle/_ref/libamps_ref.so"}
    g["focc"] = {repr(c): K.sha(K.run_focc(R, *c)) for c in K.FOCC_CASES}
    # the reference's own first bytes, in the clear: config 1 (symrate 20000) and the three superframes at sps=1
    f = R.Focc(20000, False)
    first = f.generate(3 * 19 * 926)
    g["focc_first_96_bytes_hex"] = first[:96].tobytes().hex()
    g["focc_3_superframes_sha256"] = K.sha(first.tobytes())
    g["focc_1e6_sha256"] = K.sha(R.Focc(20000, False).generate(1000000).tobytes())
    g["fvc"] = {repr(c): K.sha(K.run_fvc(R, *c)) for c in K.FVC_CASES}
    alert = R.word("ref_fvc_word1_general", 1, 0, 0, 1)
    g["fvc_alert_word"] = "".join(str(int(b)) for b in alert)      # fvc_word1_general(scc=1, 0, 0, 1), lib/recc_decode_impl.cc:214
    v = R.Fvc(20000)
    v.push_words(alert)
    g["fvc_alert_train_hex"] = v.work(2064)[1].tobytes().hex()
    g["recc"] = {repr(c): K.sha(K.run_recc(R, *c)) for c in K.RECC_CASES}
    trig = np.zeros(74, np.uint8)
    R.lib().ref_recc_trigger(trig.ctypes.data_as(R.u8p))
    g["recc_trigger"] = "".join(str(int(b)) for b in trig)
    dec = {}
    for seed, count in K.DECODE_CASES:
        rows = []
        for blob in K.recc_blobs(seed, count):
            fields = R.recc_fields(blob)
            a, info = R.recc_bursts_message(blob)
            rows.append({"fields": K.sha(K.result_bytes(fields, False)), "actions": K.sha(K.actions_bytes(a)),
                         "dispatch": list(K.dispatch_tuple_ref(info))})
        dec[str(seed)] = rows
    g["decode"] = dec
    # one decode in the clear: the config-2 origination burst
    _, hs, _ = __import__("gr_amps_b200.synth", fromlist=["x"]).config2_period(n_total=1 << 21)
    a, info = R.recc_bursts_message(hs[82:82 + 3374])
    g["config2_burst"] = {"dispatch": list(K.dispatch_tuple_ref(info)), "focc_stream": int(a.focc_stream), "n_focc": int(a.n_focc),
                          "focc_words": [bytes(a.focc_words[i]).hex() for i in range(2)], "command": a.command.decode(),
                          "fvc_mute": int(a.fvc_mute), "audio_mute": int(a.audio_mute)}
    g["commands"] = {c: K.sha(bytes(R.command_actions(c))) for c in K.COMMANDS}
    rng = np.random.default_rng(51)
    words = rng.integers(0, 2, (64, 28)).astype(np.uint8)
    g["bch_40_28_sha256"] = K.sha(b"".join(R.bch_encode_40_28(w).tobytes() for w in words))
    g["bch_valid_bits"] = "".join(str(int(R.bch_decode_48(w))) for w in K.bch_decode_inputs(52, 4000))
    g["word_builders_sha256"] = K.sha(K.word_builder_transcript(R, "ref_", 53))
    g["min_sha256"] = K.sha(K.min_transcript(R.lib(), "ref_", 54))
    with open(OUT, "w") as fh:
        json.dump(g, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
