#!/usr/bin/env python
"""What a pipelined device-resident call costs beyond its front kernel: whole-call time (events around a train of
back-to-back amps_recc_iq_submit_dev calls) with the per-kernel timing events on / off, with the side-stream kernels
(search, capture) on / off / in order on the caller's stream, next to the launch floor of the box (events around a
one-element fill).   usage: python tools/call_overheads.py [log2 sizes ...]   (run each variant in its own process: the
AMPS_RX_* switches are read at create)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "default_timed": ({}, True),
    "default": ({}, False),
    "front_only": ({"AMPS_RX_FRONT_ONLY": "1"}, False),
    "serial": ({"AMPS_RX_SERIAL": "1"}, False),
    "fused": ({"AMPS_RX_FUSED": "1"}, False),
}


def child(name, sizes):
    import numpy as np
    import torch
    from gr_amps_b200 import capi, synth
    timed = VARIANTS[name][1]
    period, _, _ = synth.config2_period(n_total=55 * 38400, snr_db=20.0)
    base = torch.from_numpy(period.view(np.float32).copy()).cuda()
    stream = torch.cuda.current_stream()
    one = torch.zeros(1, device="cuda")
    fl = []
    for _ in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); one.fill_(1.0); e1.record(stream)
        torch.cuda.synchronize()
        fl.append(e0.elapsed_time(e1) * 1e3)
    floor = float(np.median(fl[10:]))
    for lg in sizes:
        n = 1 << lg
        nbuf = min(max(2, int(np.ceil(3 * 126e6 / (8 * n)))), 1024)
        reps = int(np.ceil(n / len(period)))
        src = base.repeat(reps)[:2 * n].contiguous()
        bufs = [src.clone() for _ in range(nbuf)]
        rx = capi.ReccIq(max_samples=n, time_kernels=timed, max_bursts=4096)
        launches = max(40, min(nbuf, 200))
        for i in range(5):
            rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
        rx.peek()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(launches):
            rx.submit_dev(bufs[(5 + i) % nbuf].data_ptr(), n, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        rx.peek()
        torch.cuda.synchronize()
        out = {"variant": name, "log2": lg, "call_us": round(e0.elapsed_time(e1) * 1e3 / launches, 2), "launch_floor_us": round(floor, 2)}
        if timed:
            out["front_kernel_us"] = round(float(np.median(rx.front_times_ms(256)[-launches:])) * 1e3, 2)
        print(json.dumps(out), flush=True)
        rx.close()
        del bufs, src
        torch.cuda.empty_cache()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], [int(a) for a in sys.argv[3:]])
        return
    sizes = [a for a in sys.argv[1:]] or ["16", "21", "23", "24"]
    for name, (env, _) in VARIANTS.items():
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name] + sizes, env=e, check=True)


if __name__ == "__main__":
    main()
