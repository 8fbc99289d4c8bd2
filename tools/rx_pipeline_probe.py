"""Where does the step time of the RX pipeline go?  Runs K back-to-back device-resident steps of the bench workload and
prints step time and mean front-kernel time; run it under AMPS_RX_SERIAL=1 / AMPS_RX_FRONT_ONLY=1 / nothing.
usage: python tools/rx_pipeline_probe.py [periods=128] [steps=60] [sc16]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gr_amps_b200 import capi, synth  # noqa: E402

periods = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
x, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0)
sc16 = len(sys.argv) > 3 and sys.argv[3] == "sc16"
xd = torch.from_numpy(np.tile(x, periods).view(np.float32).copy()).cuda()
if sc16:
    xd = torch.clamp(torch.round(xd * 8192.0), -32768, 32767).to(torch.int16)
n = len(x) * periods
rx = capi.ReccIq(max_samples=n, time_kernels=True, max_bursts=4096, sc16=sc16, sc16_scale=1.0 / 8192.0 if sc16 else 0.0)
st = torch.cuda.current_stream()
for _ in range(5):
    rx.submit_dev(xd.data_ptr(), n, st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    rx.submit_dev(xd.data_ptr(), n, st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
ft = rx.front_times_ms(steps)
rx.close()
print(json.dumps({"mode": {k: os.environ.get(k) for k in ("AMPS_RX_SERIAL", "AMPS_RX_FRONT_ONLY", "AMPS_RX_DIAG", "AMPS_RX_DETECT_CTAS")}, "samples": n, "steps": steps,
                  "ms_per_step": e0.elapsed_time(e1) / steps, "front_ms_mean": float(np.mean(ft)), "front_ms_min": float(np.min(ft)),
                  "front_ms_max": float(np.max(ft)), "sc16": sc16, "GBps_step": (4.0 if sc16 else 8.0) * n / (e0.elapsed_time(e1) / steps) / 1e6}))
