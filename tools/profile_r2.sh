#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): config-5 sweep (timed, then per-size ncu sections), launch list of a bench
# step, one --set full capture of the front kernel.  Everything lands in gpurun_out/; tools/summarize_r2.py turns it into
# the files under profiles/.
set -u
mkdir -p gpurun_out
python tools/roofline_sweep.py > gpurun_out/r2_config5_sweep.jsonl 2> gpurun_out/r2_config5_sweep.err
AMPS_RX_FUSED=1 python tools/roofline_sweep.py --extra 28 > gpurun_out/r2_config5_sweep_fused.jsonl 2>> gpurun_out/r2_config5_sweep.err
python tools/roofline_sweep.py --sc16 --min-log2 20 --extra 28 > gpurun_out/r2_config5_sweep_sc16.jsonl 2>> gpurun_out/r2_config5_sweep.err
# per-size ncu (cold-cache, serialised: durations are upper bounds; the DRAM traffic per launch is the point)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    --clock-control none -k regex:rx_front_kernel --csv --log-file gpurun_out/r2_config5_ncu.csv \
    python tools/roofline_sweep.py --quick --extra "" > gpurun_out/r2_config5_sweep_under_ncu.jsonl 2>> gpurun_out/r2_config5_sweep.err
# launch list of the bench step (serialised)
AMPS_RX_SERIAL=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches_serial.csv python bench.py --steps 3 --warmup 3 --no-cpu --sustain 0 --shared-carriers 0 > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err
# the top kernel, once
ncu --set full --clock-control none --import-source on -k regex:rx_front_kernel -s 4 -c 1 -f -o gpurun_out/r2_rx_front \
    python tools/rx_pipeline_probe.py 128 3 > gpurun_out/r2_probe.log 2>&1

# bench lines (N=1), the reference arm, the forward workload
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_a.err
python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2>> gpurun_out/r2_bench_a.err
python bench.py --workload fwd > gpurun_out/r2_bench_fwd_n1.json 2>> gpurun_out/r2_bench_a.err
# where a launch / a call spends its time
python tools/front_phases.py 16 21 24 28 > gpurun_out/r2_front_phases.jsonl 2>&1
python tools/call_overheads.py 16 21 23 24 > gpurun_out/r2_call_overheads.jsonl 2>&1
python tools/search_phases.py 16 40 > gpurun_out/r2_search_phases.jsonl 2>&1
python tools/search_phases.py 21 8 >> gpurun_out/r2_search_phases.jsonl 2>&1
python tools/grid_probe.py 1048576,1500000,2097152,3000000,4194304 0,296,148 > gpurun_out/r2_grid_probe_after.jsonl 2>&1   # (0 = the rule, an explicit cap overrides it)
