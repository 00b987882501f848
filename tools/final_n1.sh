# Round-end measurement pass on one B200 (everything lands in gpurun_out/r02_*; copied into profiles/ afterwards).  Every step under a timeout.
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_n1.json'))
print([round(x,2) for x in d["detail"]["step_ms_rank0"]])
print(d["line"]["value"], d["line"]["ms_per_step"], d["line"]["e2e"]["ms_per_step"], d["line"]["clocks"])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 480 --csv --log-file gpurun_out/r02_launches_b1_fp32.csv python tools/quick_fwd.py fp32 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/r02_launches_b32_bf16.csv python tools/quick_b32.py bf16 2 > /dev/null 2>&1
rm -f gpurun_out/r02_attention_tc_vs_ffma.jsonl; for a in "1 100" "1 537" "32 600" "32 1200"; do for m in 1 0; do timeout 100 python tools/profile_attn.py $a $m | tail -1 >> gpurun_out/r02_attention_tc_vs_ffma.jsonl; done; done
for p in fp32 bf16 tf32; do timeout 100 python tools/profile_dominant.py $p 511 | tail -1; done > gpurun_out/r02_dominant_launch.log 2>&1
timeout 300 python tools/sweep.py --quick --corner --no-cfg3 --precisions fp32,tf32,bf16 --out gpurun_out/r02_sweep_cfg4_with_corner.json > gpurun_out/r02_sweep.log 2>&1; tail -3 gpurun_out/r02_sweep.log | cut -c1-200
echo final_n1_done
