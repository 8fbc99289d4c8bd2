#!/usr/bin/env python
"""Turn what tools/profile_r2.sh left in gpurun_out/ into the tracked summaries under profiles/:
  r2_config5_ncu_per_size.jsonl       per-size ncu sections of the config-5 sweep (mean over the profiled launches of a size)
  r2_rx_front_ncu_full_summary.json   the metrics DESIGN.md quotes from the one --set full capture of rx_front_kernel
  rx_front_traffic.json               its DRAM bytes per launch (bench.py's roofline.traffic)
  r2_launches_serial.csv              the launch list of a bench step (copied)
usage: python tools/summarize_r2.py   (here, after the gpurun call; needs ncu on PATH for the .ncu-rep)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
ALG = 8.0 + 1.0 / 500.0
FULL_KEYS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def num(s):
    return float(s.replace(",", ""))


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def per_size():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    from bench import load_peaks
    peak, _ = load_peaks()
    rows = list(csv.reader(l for l in open(os.path.join(G, "r2_config5_ncu.csv")) if l.startswith('"')))
    h = rows[0]
    ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
        d[r[ix["Metric Name"]]] = (num(r[ix["Metric Value"]]), r[ix["Metric Unit"]])
    sweep = [json.loads(l) for l in open(os.path.join(G, "r2_config5_sweep_under_ncu.jsonl"))]
    seq = list(launches.values())
    out = [{"_source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput...,sm__throughput... "
                       "--clock-control none -k regex:rx_front_kernel python tools/roofline_sweep.py --quick (per size: 3 warm-ups + 4 profiled launches "
                       "of the pipeline and 3 + 4 of the front-only train, all averaged except the warm-ups; serialised, cold cache)"}]
    pos = 0
    for s in sweep:
        per = 2 * (3 + 4)                                   # pipeline handle, then the front-only handle
        mine = seq[pos:pos + per]
        pos += per
        use = mine[3:7] + mine[10:14]
        if not use:
            break
        dur = sum(u["gpu__time_duration.sum"][0] * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}[u["gpu__time_duration.sum"][1]] for u in use) / len(use)
        rd = sum(to_bytes(*u["dram__bytes_read.sum"]) for u in use) / len(use)
        wr = sum(to_bytes(*u["dram__bytes_write.sum"]) for u in use) / len(use)
        n = s["samples"] if s["case"] == "single" else s["channels"] * s["samples_per_channel"]
        alg = ALG * n
        o = {"case": s["case"]}
        if s["case"] == "single":
            o["samples"] = n
        else:
            o["channels"] = s["channels"]; o["samples_per_channel"] = s["samples_per_channel"]
        o.update({"kernel": use[0]["kernel"], "grid": use[0]["grid"], "launches_averaged": len(use), "gpu_time_duration_us": round(dur, 2),
                  "dram_read_MB": round(rd / 1e6, 3), "dram_write_MB": round(wr / 1e6, 3), "dram_traffic_over_algorithmic": round((rd + wr) / alg, 3),
                  "algorithmic_GBps_at_ncu_duration": round(alg / dur / 1e3, 1), "frac_of_measured_peak_at_ncu_duration": round(alg / dur / 1e3 / peak, 3),
                  "dram_throughput_pct_of_peak": round(sum(u["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0] for u in use) / len(use), 2),
                  "sm_throughput_pct_of_peak": round(sum(u["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0] for u in use) / len(use), 2)})
        out.append(o)
    with open(os.path.join(P, "r2_config5_ncu_per_size.jsonl"), "w") as f:
        for o in out:
            f.write(json.dumps(o) + "\n")
    print("per-size:", len(out) - 1, "sizes,", pos, "of", len(seq), "launches used")


def full():
    rep = os.path.join(G, "r2_rx_front.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k in ["Kernel Name", "Block Size", "Grid Size"] + FULL_KEYS:
        if k in h:
            i = h.index(k)
            d[k] = [vals[i], units[i]]
    d["note"] = ("ncu --set full --clock-control none --import-source on -k regex:rx_front_kernel -s 4 -c 1, python tools/rx_pipeline_probe.py 128 3 "
                 "(one launch over 270336000 samples = 128 config-2 periods)")
    json.dump(d, open(os.path.join(P, "r2_rx_front_ncu_full_summary.json"), "w"), indent=1)
    rd = to_bytes(num(d["dram__bytes_read.sum"][0]), d["dram__bytes_read.sum"][1])
    wr = to_bytes(num(d["dram__bytes_write.sum"][0]), d["dram__bytes_write.sum"][1])
    n = 270336000
    json.dump({"kernel": "rx_front_kernel",
               "source": "ncu --set full --clock-control none --import-source on -k regex:rx_front_kernel -s 4 -c 1 (" + d["Kernel Name"][0].split("(")[0] + "), python tools/rx_pipeline_probe.py 128 3; "
                         "dram__bytes_read.sum + dram__bytes_write.sum of one launch over its 270336000 samples (profiles/r2_rx_front_ncu_full_summary.json)",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "samples": n, "dram_bytes_per_sample": (rd + wr) / n},
              open(os.path.join(P, "rx_front_traffic.json"), "w"), indent=1)
    print("full:", d["gpu__time_duration.sum"], d["launch__registers_per_thread"], d["smsp__inst_executed.sum"], (rd + wr) / n)


def main():
    sys.path.insert(0, ROOT)
    per_size()
    full()
    shutil.copy(os.path.join(G, "r2_launches_serial.csv"), os.path.join(P, "r2_launches_serial.csv"))


if __name__ == "__main__":
    main()
