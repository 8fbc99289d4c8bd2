"""The CUDA byte/bit-level blocks against THE REFERENCE'S OWN CODE, with no oracle in between.

The scenarios of tests/ref_cases.py run through the C ABI (gr_amps_b200.capi -> libamps_b200.so -> kernels) and must
reproduce, byte for byte, the transcripts that gr-amps's own lib/*.cc produced (tests/golden/ref_vectors.json, written
by tests/golden/make_ref_golden.py from oracle/_ref); where the compiled reference travelled to this box
(oracle/_ref/libamps_ref.so) further seeds are compared live."""
import json
import os
import types

import numpy as np
import pytest

from tests import ref_cases as K
from tests import ref_lib as R

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")))
live = pytest.mark.skipif(not os.path.exists(R.REF_SO), reason="oracle/_ref/libamps_ref.so did not travel to this box")


@pytest.fixture(scope="module")
def G():
    """The product behind the same three class names the scenarios use."""
    from gr_amps_b200 import capi as c
    c.lib()
    return types.SimpleNamespace(Focc=c.Focc, Fvc=c.Fvc, Recc=c.Recc, capi=c)


@pytest.mark.parametrize("case", K.FOCC_CASES, ids=repr)
def test_focc_schedules_reproduce_the_reference(G, case):
    assert K.sha(K.run_focc(G, *case)) == GOLDEN["focc"][repr(case)]


def test_focc_config1_is_the_reference_stream(G):
    s = G.Focc(20000, False).generate(1000000)
    assert K.sha(s.tobytes()) == GOLDEN["focc_1e6_sha256"]
    assert s[:96].tobytes().hex() == GOLDEN["focc_first_96_bytes_hex"]


@pytest.mark.parametrize("case", K.FVC_CASES, ids=repr)
def test_fvc_schedules_reproduce_the_reference(G, case):
    """Includes the idle calls: n claimed, buffer untouched (lib/fvc_impl.cc:159-161)."""
    assert K.sha(K.run_fvc(G, *case)) == GOLDEN["fvc"][repr(case)]


def test_fvc_alert_train_is_the_reference_train(G):
    """The alert-order word train of BASELINE config 3 (lib/recc_decode_impl.cc:214), word and train from the reference."""
    v = G.Fvc(20000)
    v.push_words(np.asarray([int(c) for c in GOLDEN["fvc_alert_word"]], np.uint8))
    assert v.work(2064)[1].tobytes().hex() == GOLDEN["fvc_alert_train_hex"]


@pytest.mark.parametrize("case", K.RECC_CASES, ids=repr)
def test_recc_capture_reproduces_the_reference(G, case):
    assert K.sha(K.run_recc(G, *case)) == GOLDEN["recc"][repr(case)]


@pytest.mark.parametrize("seed,count", K.DECODE_CASES)
def test_recc_decode_reproduces_the_reference(G, seed, count):
    blobs = K.recc_blobs(seed, count)
    out = G.capi.ReccDecode().decode(np.stack(blobs))
    for g, row in zip(out, GOLDEN["decode"][str(seed)]):
        assert K.sha(K.result_bytes(g, False)) == row["fields"]
        assert list(K.dispatch_tuple_oracle(g)) == row["dispatch"]


@live
@pytest.mark.parametrize("symrate,aggressive,seed", [(20000, True, 601), (100000, False, 602), (10000000, False, 603)])
def test_live_focc_vs_compiled_reference(G, symrate, aggressive, seed):
    total = 40 * 926 * (symrate // 20000) if symrate < 10000000 else 2 * 926 * 500
    assert K.run_focc(G, symrate, aggressive, seed, total, True) == K.run_focc(R, symrate, aggressive, seed, total, True)


@live
@pytest.mark.parametrize("symrate,seed", [(20000, 611), (100000, 612)])
def test_live_fvc_vs_compiled_reference(G, symrate, seed):
    assert K.run_fvc(G, symrate, seed, 60) == K.run_fvc(R, symrate, seed, 60)


@live
@pytest.mark.parametrize("seed,bursts,chunk", [(621, 20, 2048), (622, 30, 61439)])
def test_live_recc_vs_compiled_reference(G, seed, bursts, chunk):
    assert K.run_recc(G, seed, bursts, chunk) == K.run_recc(R, seed, bursts, chunk)


@live
def test_live_decode_vs_compiled_reference(G):
    blobs = K.recc_blobs(631, 128)
    out = G.capi.ReccDecode().decode(np.stack(blobs))
    for blob, g in zip(blobs, out):
        assert K.result_bytes(g, False) == K.result_bytes(R.recc_fields(blob), False)
        _, info = R.recc_bursts_message(blob)
        assert K.dispatch_tuple_oracle(g) == K.dispatch_tuple_ref(info)


def test_testalloc_properties_hold_for_the_cuda_focc(G):
    """The reference's only executable checks (apps/testalloc.cc:64-92: focc at symrate 200000, 10240-byte requests,
    every sample of a half-symbol equal and non-zero, every pair (+1,-1) or (-1,+1)), on the CUDA source."""
    sps = 10
    f = G.Focc(200000, False)
    bits = 0
    while bits < 20000:
        r, b = f.work(10240)
        assert r % sps == 0 and r % 2 == 0
        s = b.view(np.int8).reshape(-1, sps)
        assert np.all(s == s[:, :1]) and np.all(s[:, 0] != 0)
        sym = s[:, 0]
        assert np.all(sym[0::2] == -sym[1::2])
        bits += r // sps // 2
