#!/usr/bin/env python
"""Front kernel as a train of launches (AMPS_RX_FRONT_ONLY) at arbitrary buffer sizes for several grid caps (AMPS_RX_GRID):
where between one and two CTAs per SM the break-even lies.  usage: python tools/grid_probe.py n1,n2,... g1,g2,..."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gr_amps_b200 import capi, synth
    sizes = [int(v) for v in sys.argv[1].split(",")]
    grids = [int(v) for v in sys.argv[2].split(",")]
    period, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0)
    base = torch.from_numpy(period.view(np.float32).copy()).cuda()
    stream = torch.cuda.current_stream()
    os.environ["AMPS_RX_FRONT_ONLY"] = "1"
    for n in sizes:
        n -= n % 2
        nbuf = min(max(2, int(np.ceil(3 * 126e6 / (8 * n)))), 256)
        reps = int(np.ceil(n / len(period))) + 1
        bufs = [base.repeat(reps)[:2 * n].contiguous().clone() for _ in range(nbuf)]
        out = {"samples": n, "tiles": (n // 1600 + 2) // 3}
        for g in grids:
            os.environ["AMPS_RX_GRID"] = str(g)
            rx = capi.ReccIq(max_samples=n, max_bursts=64)
            launches = max(40, nbuf)
            for i in range(4):
                rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for i in range(launches):
                rx.submit_dev(bufs[(4 + i) % nbuf].data_ptr(), n, stream.cuda_stream)
            t1.record(stream)
            torch.cuda.synchronize()
            out["grid_%d_us" % g] = round(t0.elapsed_time(t1) * 1e3 / launches, 2)
            rx.close()
        print(json.dumps(out), flush=True)
        del bufs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
