#!/bin/bash
# compute-sanitizer passes over the C++ QA driver (no Python / torch in the process): memcheck, racecheck, synccheck
# on the closed-loop scenario (recc_iq front/detect/select/capture kernels, decode, focc/fvc byte kernels).
# usage (GPU box): bash tools/sanitize.sh  -> writes gpurun_out/sanitize_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/san
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from gr_amps_b200 import synth
period = 55 * 38400
msgs = [synth.origination_words(min10="2125550101"), synth.page_response_words(min10="2125550102")]
parts = [synth.burst_period(w, n_total=period, snr_db=20.0, seed=40 + i)[0] for i, w in enumerate(msgs)]
x = np.concatenate(parts + [np.zeros(38400, np.complex64)])
x.tofile('/tmp/san/iq.bin')
open('/tmp/san/n.txt', 'w').write(str(len(x)))
PY
N=$(cat /tmp/san/n.txt)
QA=gr_amps_b200/host/qa_blocks
if [ "${R2_ONLY:-0}" = "1" ]; then
  # round 2: the re-dealt front kernel + stand-alone search kernel (default), the fused search (per-pass shared-memory search,
  # boundary counters between CTAs, in-kernel selection), its sc16 instance, and the batched kernels (4 carriers, one upload)
  for tool in memcheck racecheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 262144 87970 /tmp/san/out > gpurun_out/sanitize_r2_split_$tool.log 2>&1
    echo "r2 split $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r2_split_$tool.log | tail -1)"
    AMPS_RX_FUSED=1 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 262144 87970 /tmp/san/outf > gpurun_out/sanitize_r2_fused_$tool.log 2>&1
    echo "r2 fused $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r2_fused_$tool.log | tail -1)"
    AMPS_RX_FUSED=1 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 99999 87970 /tmp/san/outf16 sc16 > gpurun_out/sanitize_r2_fused_sc16_$tool.log 2>&1
    echo "r2 fused sc16 $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r2_fused_sc16_$tool.log | tail -1)"
    timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 $QA batch /tmp/san/iq.bin $N 4 1000001 > gpurun_out/sanitize_r2_batch_$tool.log 2>&1
    echo "r2 batch $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r2_batch_$tool.log | tail -1)"
    AMPS_RX_FUSED=1 timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 $QA batch /tmp/san/iq.bin $N 4 1000001 > gpurun_out/sanitize_r2_batch_fused_$tool.log 2>&1
    echo "r2 batch fused $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_r2_batch_fused_$tool.log | tail -1)"
  done
  exit 0
fi
for tool in memcheck racecheck synccheck; do
  # the Manchester-bit fast path (fwd_bits_kernel) through the forward_iq composite block
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA txblock 6000 /tmp/san/tx.bin > gpurun_out/sanitize_txblock_$tool.log 2>&1
  echo "txblock $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_txblock_$tool.log | tail -1)"
  if [ "${FWD_ONLY:-0}" = "1" ]; then
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA fwd 20000 /tmp/san/fwd.bin > gpurun_out/sanitize_fwd_$tool.log 2>&1
    echo "fwd  $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_fwd_$tool.log | tail -1)"
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA fwd 20000 /tmp/san/fwdv.bin voice > gpurun_out/sanitize_voice_$tool.log 2>&1
    echo "voice $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_voice_$tool.log | tail -1)"
    continue
  fi
  if [ "${ONLY_NEW:-0}" != "1" ]; then
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 262144 87970 /tmp/san/out > gpurun_out/sanitize_$tool.log 2>&1
  echo "loop $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA fwd 20000 /tmp/san/fwd.bin > gpurun_out/sanitize_fwd_$tool.log 2>&1
  echo "fwd  $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_fwd_$tool.log | tail -1)"
  fi
  # round-1 additions: the M&M timing tail of recc_iq and the voice legs of the forward path
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 262144 87970 /tmp/san/outmm mm > gpurun_out/sanitize_mm_$tool.log 2>&1
  echo "mm   $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_mm_$tool.log | tail -1)"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA fwd 20000 /tmp/san/fwdv.bin voice > gpurun_out/sanitize_voice_$tool.log 2>&1
  echo "voice $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_voice_$tool.log | tail -1)"
  # the sc16 instance of the front kernel (int16 I,Q input, 3 CTAs/SM)
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 $QA loop /tmp/san/iq.bin $N 262144 87970 /tmp/san/out16 sc16 > gpurun_out/sanitize_sc16_$tool.log 2>&1
  echo "sc16 $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_sc16_$tool.log | tail -1)"
done
