"""Throughput of the M&M timing mode (AMPS_RX_TIMING_MM) on device-resident IQ: one thread walks the recurrence,
so this reports Msamples/s and Msymbols/s of the serial tail, not a roofline number.
usage: python tools/mm_rate.py [periods] [reps]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from gr_amps_b200 import capi, synth  # noqa: E402

periods = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
x, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=30.0)
xd = torch.from_numpy(np.tile(x, periods).view(np.float32).copy()).cuda()
n = len(x) * periods
out = {}
for mode in ("mm", "feed_forward"):
    rx = capi.ReccIq(max_samples=n, timing_mm=(mode == "mm"), max_bursts=4096)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    rx.submit_dev(xd.data_ptr(), n, st.cuda_stream)
    nb = len(rx.collect(4096))
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        rx.submit_dev(xd.data_ptr(), n, st.cuda_stream)
    got = rx.collect(4096)                       # waits for both streams
    dt = (time.perf_counter() - t0) / reps
    out[mode] = {"bursts_first_call": nb, "bursts_timed": len(got), "ms_per_call": 1e3 * dt,
                 "Msamples_per_s": n / dt / 1e6, "Msymbols_per_s": n / 500 / dt / 1e6}
    rx.close()
print(json.dumps({"samples_per_call": n, "reps": reps, **out}))
