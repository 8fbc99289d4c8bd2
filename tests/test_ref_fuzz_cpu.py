"""Property-based comparison of the oracle with the compiled reference (oracle/_ref): hypothesis generates the
scenarios, both implementations must agree on every observable.  Skipped where the compiled reference is not at hand."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from gr_amps_b200 import synth
from tests import oracle_lib as O
from tests import ref_cases as K
from tests import ref_lib as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libamps_ref.so not built (no reference tree here)")
FUZZ = settings(max_examples=int(__import__("os").environ.get("AMPS_FUZZ_EXAMPLES", "60")), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
TRIG = synth.trigger_symbols()


@FUZZ
@given(st.lists(st.tuples(st.sampled_from(["noise", "trigger", "near", "payload"]), st.integers(0, 5000)), min_size=1, max_size=14),
       st.lists(st.integers(1, 61439), min_size=1, max_size=40), st.integers(0, 2 ** 32 - 1))
def test_recc_capture_any_stream_any_chunking(parts, chunks, seed):
    """recc_impl::work (lib/recc_impl.cc:93-145): any mix of noise, triggers, near-miss triggers and payloads, cut into
    any chunk sizes below the reference's 61440 limit."""
    rng = np.random.default_rng(seed)
    segs = []
    for kind, n in parts:
        if kind == "noise":
            segs.append(rng.integers(0, 2, n).astype(np.uint8))
        elif kind == "trigger":
            segs.append(TRIG)
        elif kind == "near":
            t = TRIG.copy()
            t[n % 74] ^= 1
            segs.append(t)
        else:
            segs.append(rng.integers(0, 2, 3374 + n % 300).astype(np.uint8))
    s = np.concatenate(segs)
    a, b = R.Recc(), O.Recc()
    pos, i = 0, 0
    while pos < len(s):
        n = min(chunks[i % len(chunks)], len(s) - pos)
        i += 1
        assert a.work(s[pos:pos + n]) == b.work(s[pos:pos + n]) == 0
        pos += n
        assert len(a.bursts) == len(b.bursts)
    assert all(np.array_equal(x, y) for x, y in zip(a.bursts, b.bursts))


@FUZZ
@given(st.sampled_from([20000, 40000, 100000, 200000]), st.booleans(),
       st.lists(st.one_of(st.integers(0, 12000), st.tuples(st.integers(1, 3), st.integers(1, 3), st.integers(0, 2 ** 31))), min_size=1, max_size=60))
def test_focc_any_schedule(symrate, aggressive, script):
    """focc_impl::work + focc_words_message: an integer is a work(n) request (0 -> WORK_DONE), a tuple injects
    (stream, nwords, seed) through the focc_words port."""
    a, b = R.Focc(symrate, aggressive), O.Focc(symrate, aggressive)
    for step in script:
        if isinstance(step, tuple):
            stream, nwords, seed = step
            w = np.random.default_rng(seed).integers(0, 2, 28 * nwords).astype(np.uint8)
            a.push_words(stream, w)
            b.push_words(stream, w)
        else:
            ra, ba = a.work(step)
            rb, bb = b.work(step)
            assert ra == rb and np.array_equal(ba, bb)


@FUZZ
@given(st.integers(0, 2 ** 32 - 1), st.lists(st.tuples(st.integers(0, 3373), st.integers(0, 1)), max_size=40),
       st.sampled_from(["orig", "page", "reg", "raw"]))
def test_recc_decode_any_damage(seed, hits, kind):
    """bursts_message on a well-formed message with arbitrary half-symbols forced to 0/1 (bit errors and invalid
    Manchester pairs alike), or on raw noise: fields, validity, dispatch and everything published."""
    rng = np.random.default_rng(seed)
    min10 = "".join(str(int(d)) for d in rng.integers(0, 10, 10))
    if kind == "raw":
        blob = rng.integers(0, 2, 3374).astype(np.uint8)
    else:
        words = {"orig": lambda: synth.origination_words(min10=min10, esn=int(rng.integers(0, 2 ** 32)),
                                                         dialed="".join("0123456789*#"[int(d)] for d in rng.integers(0, 12, int(rng.integers(1, 33))))),
                 "page": lambda: synth.page_response_words(min10=min10),
                 "reg": lambda: synth.registration_words(min10=min10, esn=int(rng.integers(0, 2 ** 32)))}[kind]()
        blob = synth.manchester(synth.recc_message_bits(words))[82:82 + 3374].copy()
    for pos, val in hits:
        blob[pos] = val
    fr = R.recc_fields(blob)
    acts, info = R.recc_bursts_message(blob)
    r = O.recc_decode(blob)
    assert K.result_bytes(fr, False) == K.result_bytes(r, False)
    assert K.actions_bytes(acts) == K.actions_bytes(O.recc_actions(r))
    assert K.dispatch_tuple_ref(info) == K.dispatch_tuple_oracle(r)


@FUZZ
@given(st.lists(st.integers(0, 1), min_size=48, max_size=48))
def test_bch_validity_any_word(bits):
    w = np.asarray(bits, np.uint8)
    assert R.bch_decode_48(w) == O.bch_decode_48(w)[0]


@FUZZ
@given(st.sampled_from([20000, 100000]), st.lists(st.one_of(st.integers(1, 30000), st.tuples(st.integers(1, 2), st.integers(0, 3), st.integers(0, 2 ** 31))),
                                                  min_size=1, max_size=40))
def test_fvc_any_schedule(symrate, script):
    """fvc_impl::work + fvc_words_message: integers are work(n) requests, tuples push (nwords, timer or 0 = none, seed)."""
    a, b = R.Fvc(symrate), O.Fvc(symrate)
    for step in script:
        if isinstance(step, tuple):
            nwords, timer, seed = step
            w = np.random.default_rng(seed).integers(0, 2, 28 * nwords).astype(np.uint8)
            a.push_words(w, timer or None)
            b.push_words(w, timer or None)
        else:
            assert [x if not isinstance(x, np.ndarray) else x.tobytes() for x in a.work(step)] == \
                   [x if not isinstance(x, np.ndarray) else x.tobytes() for x in b.work(step)]
