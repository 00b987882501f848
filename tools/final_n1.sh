timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_n1.json'))
print([round(x,2) for x in d["detail"]["step_ms_rank0"]])
print([round(x,2) for x in d["detail"]["e2e_step_ms_rank0"]])
print(d["line"]["value"], d["line"]["ms_per_step"], d["line"]["e2e"]["ms_per_step"], d["line"]["clocks"], d["line"]["b1"])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>/dev/null; tail -c 150 gpurun_out/r02_bench_reference_n1.json
echo final_n1_done
