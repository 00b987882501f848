#!/usr/bin/env bash
# One gpurun call that validates everything written after round 1's GPU budget ended and collects the numbers the next
# optimisation steps need.  Everything lands in gpurun_out/ (merged back by gpurun).  Each step has its own timeout so a hang
# in an experimental path costs minutes, not the call.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh'
set -u
OUT=gpurun_out
mkdir -p "$OUT"
step() { echo "=== $1" | tee -a "$OUT/checklist.log"; }

step "1. validated suite (everything but the late file)"
timeout 600 python -m pytest tests -m gpu -x -q --ignore=tests/test_zz_late_round1_gpu.py > "$OUT/pytest_validated.log" 2>&1
echo "exit $?" | tee -a "$OUT/checklist.log"; tail -3 "$OUT/pytest_validated.log" | tee -a "$OUT/checklist.log"

step "2. late file (bf16, micro-batcher, pcm16 fetch, PDL / autotune / fused-ResBlock bitwise, style encoder), no -x: see every failure"
timeout 900 python -m pytest tests/test_zz_late_round1_gpu.py -m gpu -q -rA > "$OUT/pytest_late.log" 2>&1
echo "exit $?" | tee -a "$OUT/checklist.log"; grep -E "^(PASSED|FAILED|ERROR)|passed|failed" "$OUT/pytest_late.log" | tail -60 | tee -a "$OUT/checklist.log"

step "3. bench (headline + experiments block: precisions, EV_PDL, EV_AUTOTUNE log, EV_FUSE_RES, style encoder)"
timeout 600 python bench.py --steps 30 --warmup 5 > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
echo "exit $?" | tee -a "$OUT/checklist.log"; head -c 1500 "$OUT/bench_default.json" | tee -a "$OUT/checklist.log"; echo | tee -a "$OUT/checklist.log"

step "4. batch sweep with and without the fused ResBlock layer (cfg3 / cfg4 points of tools/sweep.py)"
timeout 400 python tools/sweep.py --quick --precisions fp32,tf32,bf16 --out "$OUT/sweep_default.json" > "$OUT/sweep_default.log" 2>&1; echo "exit $?" | tee -a "$OUT/checklist.log"
EV_FUSE_RES=1 timeout 400 python tools/sweep.py --quick --precisions fp32,tf32,bf16 --out "$OUT/sweep_fuse_res.json" > "$OUT/sweep_fuse_res.log" 2>&1; echo "exit $?" | tee -a "$OUT/checklist.log"

step "5. launch list of one B=1 step, default vs EV_FUSE_RES=1 (ncu, per-launch durations)"
for mode in default fuse; do
  if [ "$mode" = fuse ]; then export EV_FUSE_RES=1; else unset EV_FUSE_RES; fi
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$OUT/launches_$mode.csv" \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-experiments > "$OUT/ncu_$mode.log" 2>&1
  echo "ncu $mode exit $?" | tee -a "$OUT/checklist.log"
done
unset EV_FUSE_RES

step "6. fused ResBlock layer vs the two launches, per vocoder stage shape of the bench utterance (and at batch 8)"
for args in "fp32 32 3 1 137472" "fp32 32 11 5 137472" "fp32 64 7 3 68736" "fp32 64 11 5 68736" "tf32 128 11 5 34368" "tf32 64 7 3 68736" \
            "bf16 32 11 5 137472" "fp32 32 11 5 137472 8" "fp32 64 7 3 68736 8" "tf32 128 7 3 34368 8"; do
  timeout 120 python tools/profile_resblock.py $args >> "$OUT/resblock_pairs.jsonl" 2>> "$OUT/resblock_pairs.err"; echo "$args -> exit $?" | tee -a "$OUT/checklist.log"
done
tail -12 "$OUT/resblock_pairs.jsonl" | tee -a "$OUT/checklist.log"
step "7. compute-sanitizer (memcheck) over the operator tests and one small forward (slow; bounded)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -m gpu -x -q > "$OUT/sanitizer_ops.log" 2>&1
echo "memcheck ops exit $?" | tee -a "$OUT/checklist.log"; tail -5 "$OUT/sanitizer_ops.log" | tee -a "$OUT/checklist.log"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/sanitizer_smoke.log" 2>&1
echo "memcheck smoke exit $?" | tee -a "$OUT/checklist.log"; tail -5 "$OUT/sanitizer_smoke.log" | tee -a "$OUT/checklist.log"
step "done"
