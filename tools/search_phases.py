#!/usr/bin/env python
"""Phase stamps of rx_search_kernel's CTA 0 (AMPS_RX_PROF=1; %globaltimer): entry | groups searched | completion counted |
selection done | exit (after the inline capture of a small call), relative to the entry, for a train of pipelined calls.
usage: python tools/search_phases.py [log2 size] [calls]"""
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
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n = 1 << lg
    period, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0)
    reps = int(np.ceil((calls + 2) * n / len(period))) + 1
    x = torch.from_numpy(np.tile(period, reps).view(np.float32).copy()).cuda()
    st = torch.cuda.current_stream().cuda_stream
    rx = capi.ReccIq(max_samples=n, max_bursts=4096)
    rows = []
    for i in range(calls):
        rx.submit_dev(x.data_ptr() + 8 * n * i, n, st)
        got = rx.collect()                                   # (synchronises: one call at a time, the stamps are this call's)
        p = rx.debug_prof(1).astype(np.int64)[0]
        s = p[9:14]
        rows.append([round(float(v - s[0]) / 1e3, 2) if v else None for v in s] + [round(float(s[0] - p[0]) / 1e3, 2), len(got)])
    for r in rows:
        print(json.dumps({"log2": lg, "groups_us": r[1], "counted_us": r[2], "selected_us": r[3], "exit_us": r[4],
                          "search_entry_after_front_entry_us": r[5], "bursts": r[6]}))


if __name__ == "__main__":
    main()
