#!/usr/bin/env python
"""BASELINE config 5: HBM-roofline sweep of the fused RECC front end over buffer sizes.

For each size (whole passes of 38 400 samples, from one pass up to 2^24+ samples) the device-resident
path is timed over >= 20 launches after 3 warm-ups, rotating through enough distinct buffers that the
working set is > 3 x L2 (126 MB), so every launch streams from HBM.  Prints one JSON object per size:
front-kernel time (CUDA events around the kernel, AMPS_RX_TIME_KERNELS), whole-call time, algorithmic GB/s and
the fraction of the measured HBM peak.  Small buffers are launch-latency / occupancy bound (one CTA per pass).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PASS = 38400
L2_BYTES = 126e6


def main():
    import torch
    from gr_amps_b200 import capi, synth
    from bench import load_peaks
    peak, src = load_peaks()
    period, _, _ = synth.config2_period(n_total=55 * PASS, snr_db=20.0)
    base = torch.from_numpy(period.view(np.float32).copy()).cuda()
    sizes = [1, 2, 4, 8, 16, 32, 64, 128, 256, 437, 874, 1748]          # passes: 38 400 ... 67 M samples
    stream = torch.cuda.current_stream()
    for npass in sizes:
        n = npass * PASS
        nbuf = max(2, int(np.ceil(3 * L2_BYTES / (8 * n))))
        nbuf = min(nbuf, 4096)
        reps = int(np.ceil(n / len(period)))
        one = base.repeat(reps)[:2 * n].contiguous()
        bufs = [one.clone() for _ in range(nbuf)]
        rx = capi.ReccIq(max_samples=n, time_kernels=True, max_bursts=4096)
        launches = max(20, nbuf)
        for i in range(3):
            rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
        rx.peek()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(launches):
            rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
        _, _, first, count = rx.peek()
        e1.record(stream)
        torch.cuda.synchronize()
        rx.consume(count)
        call_ms = e0.elapsed_time(e1) / launches
        front = rx.front_times_ms(256)[-min(launches, 256):]
        fm = float(np.median(front))
        alg = 8.0036 * n
        print(json.dumps({"samples": n, "log2": round(float(np.log2(n)), 2), "buffers": nbuf, "launches": launches,
                          "front_kernel_us": 1e3 * fm, "call_us": 1e3 * call_ms,
                          "front_GBps": alg / (fm * 1e-3) / 1e9, "front_frac_of_peak": alg / (fm * 1e-3) / 1e9 / peak,
                          "call_Msamples_s": n / (call_ms * 1e-3) / 1e6, "peak_GBps": peak, "peak_source": src}), flush=True)
        rx.close()
        del bufs, one


if __name__ == "__main__":
    main()
