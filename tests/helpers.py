"""Shared comparison helpers for the parity tests."""
import numpy as np

WORD_FIELDS = ["dcc_errs", "F", "NAWC", "T", "S", "E", "ER", "SCM", "MIN1", "B_F", "B_NAWC", "MSG_TYPE", "ORDQ",
               "ORDER", "LT", "EP", "SCM4", "MPCI", "SDCC1", "SDCC2", "MIN2", "word_c_serial", "kind", "esn"]


def words_equal(gpu, orc) -> list:
    """Compare a capi.ReccWords with an oracle_lib.ReccResult; returns the list of differing fields."""
    bad = []
    if bytes(gpu.dcc) != bytes(orc.dcc):
        bad.append("dcc")
    for w in range(7):
        if bytes(gpu.words[w]) != bytes(orc.words[w]):
            bad.append(f"words[{w}]")
    for name in ("errs", "valid", "valid_repeat"):
        if list(getattr(gpu, name)) != list(getattr(orc, name)):
            bad.append(name)
    for f in WORD_FIELDS:
        if int(getattr(gpu, f)) != int(getattr(orc, f)):
            bad.append(f)
    if gpu.min.split(b"\0")[0] != orc.min.split(b"\0")[0]:
        bad.append("min")
    if gpu.dialed.split(b"\0")[0] != orc.dialed.split(b"\0")[0]:
        bad.append("dialed")
    return bad


def bits_equal_f32(a: np.ndarray, b: np.ndarray) -> bool:
    """Numerical equality of two float32 arrays, -0 == +0, NaN never equal."""
    return a.shape == b.shape and bool(np.all(a == b))
