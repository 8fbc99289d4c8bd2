#!/usr/bin/env python
"""Phase breakdown of one rx_front_kernel launch from per-CTA %globaltimer stamps (AMPS_RX_PROF=1):
  0 entry | 1 first tile landed | 2 warm-up done (first own tile landed) | 3 segment done | 4 predecessors' flags seen |
  5 search done | 6 completion counted | 8 channel finished (last CTA only) | 7 exit
usage: python tools/front_phases.py [log2 sizes ...]"""
import json
import os
import sys

import numpy as np

os.environ["AMPS_RX_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from gr_amps_b200 import capi, synth
    period, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0)
    base = torch.from_numpy(period.view(np.float32).copy()).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for lg in [int(a) for a in sys.argv[1:]] or [15, 16, 21, 23, 24]:
        n = 1 << lg
        reps = int(np.ceil(n / len(period)))
        bufs = [base.repeat(reps)[:2 * n].contiguous().clone() for _ in range(max(2, min(64, int(4e8 / (8 * n)))))]
        rx = capi.ReccIq(max_samples=n, time_kernels=True, max_bursts=4096)
        for i in range(6):
            rx.submit_dev(bufs[i % len(bufs)].data_ptr(), n, st)
        rx.collect()
        tiles = (n // 1600 + 2) // 3
        grid = 148 if 296 < tiles <= 740 else min(296, tiles)      # (rx_enqueue10: one CTA per SM between one and five tiles per SM)
        p = rx.debug_prof(grid).astype(np.int64)
        t0 = p[:, 0].min()
        rel = (p - t0) / 1e3
        ev = float(rx.front_times_ms(8)[-1]) * 1e3
        names = ["entry", "tile0", "warm", "seg", "flags", "search", "counted", "exit", "chan_done"]
        out = {"log2": lg, "grid": grid, "event_us": round(ev, 1), "span_us": round(float((p[:, 7].max() - t0) / 1e3), 1)}
        for k, nm in enumerate(names):
            col = rel[:, k][p[:, k] > 0]
            if len(col):
                out[nm] = [round(float(col.min()), 1), round(float(np.median(col)), 1), round(float(col.max()), 1)]
        print(json.dumps(out), flush=True)
        rx.close()
        del bufs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
