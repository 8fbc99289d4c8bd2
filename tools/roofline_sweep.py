#!/usr/bin/env python
"""BASELINE config 5: HBM-roofline sweep of the fused RECC front kernel over buffer sizes 2^14 ... 2^24 complex samples
(plus 2^26 / 2^28 for reference and the batched case: 64 channels x 2^18 samples in one launch).

For each size the device-resident path (amps_recc_iq_submit_dev, any 2^k length: the remainder of a 1600-sample unit is
carried on the device) is timed over >= 20 launches after 3 warm-ups, rotating through enough distinct buffers that the
working set is > 3 x L2 (126 MB), so every launch streams from HBM.  One JSON object per size: front-kernel time (CUDA
events around the launch on its stream, AMPS_RX_TIME_KERNELS), whole-call time, algorithmic GB/s and the fraction of the
measured HBM peak.

  --quick : 3 warm-ups + 4 timed launches per size (for a run under ncu, see profiles/README.md)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PASS = 38400
L2_BYTES = 126e6
ALG = 8.0 + 1.0 / 500.0        # bytes per sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--min-log2", type=int, default=14)
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--extra", default="26,28", help="further log2 sizes, comma separated ('' = none)")
    ap.add_argument("--sc16", action="store_true")
    args = ap.parse_args()
    import torch
    from gr_amps_b200 import capi, synth
    from bench import load_peaks
    peak, src = load_peaks()
    period, _, _ = synth.config2_period(n_total=55 * PASS, snr_db=20.0)
    if args.sc16:
        base = torch.from_numpy(np.clip(np.round(period.view(np.float32) * 8192.0), -32768, 32767).astype(np.int16)).cuda()
        kw = dict(sc16=True, sc16_scale=1.0 / 8192.0)
        isz = 4
    else:
        base = torch.from_numpy(period.view(np.float32).copy()).cuda()
        kw = {}
        isz = 8
    stream = torch.cuda.current_stream()
    logs = list(range(args.min_log2, args.max_log2 + 1)) + [int(v) for v in args.extra.split(",") if v]
    warm = 3
    for lg in logs:
        n = 1 << lg
        nbuf = min(max(2, int(np.ceil(3 * L2_BYTES / (isz * n)))), 4096)
        reps = int(np.ceil(n / len(period)))
        one = base.repeat(reps)[:2 * n].contiguous()
        bufs = [one.clone() for _ in range(nbuf)]
        rx = capi.ReccIq(max_samples=n, time_kernels=True, max_bursts=4096, **kw)
        launches = 4 if args.quick else max(20, min(nbuf, 200))
        for i in range(warm):
            rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
        rx.peek()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(launches):
            rx.submit_dev(bufs[(warm + i) % nbuf].data_ptr(), n, stream.cuda_stream)
        e1.record(stream)
        _, _, first, count = rx.peek()
        torch.cuda.synchronize()
        rx.consume(count)
        call_ms = e0.elapsed_time(e1) / launches
        front = rx.front_times_ms(256)[-min(launches, 256):]
        fm = float(np.median(front))
        nproc = n // 1600 * 1600                      # a call processes whole units; the rest is carried (on average all of n)
        alg = (ALG - 8.0 + isz) * n
        st = rx.stats()
        rx.close()
        # the front kernel as a train: the same launches back to back with nothing else in the stream (no per-launch events,
        # no search / capture kernels next to it), two events around the whole train
        os.environ["AMPS_RX_FRONT_ONLY"] = "1"
        rx = capi.ReccIq(max_samples=n, max_bursts=4096, **kw)
        del os.environ["AMPS_RX_FRONT_ONLY"]
        for i in range(warm):
            rx.submit_dev(bufs[i % nbuf].data_ptr(), n, stream.cuda_stream)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for i in range(launches):
            rx.submit_dev(bufs[(warm + i) % nbuf].data_ptr(), n, stream.cuda_stream)
        t1.record(stream)
        torch.cuda.synchronize()
        train_ms = t0.elapsed_time(t1) / launches
        print(json.dumps({"case": "single", "samples": n, "log2": lg, "buffers": nbuf, "launches": launches,
                          "front_kernel_us": 1e3 * fm, "front_kernel_us_min": 1e3 * float(np.min(front)), "call_us": 1e3 * call_ms,
                          "front_GBps": alg / (fm * 1e-3) / 1e9, "front_frac_of_peak": alg / (fm * 1e-3) / 1e9 / peak,
                          "front_train_us": 1e3 * train_ms, "front_train_frac_of_peak": alg / (train_ms * 1e-3) / 1e9 / peak,
                          "call_frac_of_peak": alg / (call_ms * 1e-3) / 1e9 / peak,
                          "call_Msamples_s": n / (call_ms * 1e-3) / 1e6, "launches_per_call": st["kernel_launches"] / (warm + launches),
                          "bursts": int(count), "units_per_call": nproc // 1600, "input": "sc16" if args.sc16 else "fc32",
                          "peak_GBps": peak, "peak_source": src}), flush=True)
        rx.close()
        del bufs, one
        torch.cuda.empty_cache()

    # ---- batched: K channels x 2^18 samples, one front launch
    for K, lg in ((64, 18), (64, 14), (64, 15), (8, 21)):
        n = 1 << lg
        carriers = [-160e3 + 30e3 * (k % 8) for k in range(K)]
        hs = [capi.ReccIq(max_samples=n, center_freq=c, max_bursts=512, **kw) for c in carriers]
        b = capi.ReccIqBatch(hs, time_kernels=True)
        nset = min(max(2, int(np.ceil(3 * L2_BYTES / (isz * n * K)))), 64)
        reps = int(np.ceil(n / len(period)))
        one = base.repeat(reps)[:2 * n].contiguous()
        sets = [[one.clone() for _ in range(K)] for _ in range(nset)]
        prep = [b.prepare([t.data_ptr() for t in ts], n) for ts in sets]
        launches = 4 if args.quick else max(20, nset)
        for i in range(warm):
            b.submit_prepared(prep[i % nset], stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(launches):
            b.submit_prepared(prep[(warm + i) % nset], stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        call_ms = e0.elapsed_time(e1) / launches
        front = b.front_times_ms(256)[-launches:]
        fm = float(np.median(front))
        alg = (ALG - 8.0 + isz) * n * K
        nb = sum(len(h.collect()) for h in hs)
        lpc = b.stats()["kernel_launches"] / (warm + launches)
        b.close()
        for h in hs:
            h.close()
        os.environ["AMPS_RX_FRONT_ONLY"] = "1"
        hs = [capi.ReccIq(max_samples=n, center_freq=c, max_bursts=512, **kw) for c in carriers]
        del os.environ["AMPS_RX_FRONT_ONLY"]
        b = capi.ReccIqBatch(hs)
        prep = [b.prepare([t.data_ptr() for t in ts], n) for ts in sets]
        for i in range(warm):
            b.submit_prepared(prep[i % nset], stream.cuda_stream)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for i in range(launches):
            b.submit_prepared(prep[(warm + i) % nset], stream.cuda_stream)
        t1.record(stream)
        torch.cuda.synchronize()
        train_ms = t0.elapsed_time(t1) / launches
        print(json.dumps({"case": "batch", "channels": K, "samples_per_channel": n, "log2_total": float(np.log2(n * K)), "launches": launches,
                          "front_kernel_us": 1e3 * fm, "call_us": 1e3 * call_ms, "front_GBps": alg / (fm * 1e-3) / 1e9,
                          "front_frac_of_peak": alg / (fm * 1e-3) / 1e9 / peak,
                          "front_train_us": 1e3 * train_ms, "front_train_frac_of_peak": alg / (train_ms * 1e-3) / 1e9 / peak,
                          "call_frac_of_peak": alg / (call_ms * 1e-3) / 1e9 / peak, "call_Msamples_s": n * K / (call_ms * 1e-3) / 1e6,
                          "launches_per_call": lpc, "bursts": nb,
                          "input": "sc16" if args.sc16 else "fc32", "peak_GBps": peak}), flush=True)
        b.close()
        for h in hs:
            h.close()
        del sets, one
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
