# Round-end measurement pass on one B200 (everything lands in gpurun_out/r02_*; copied into profiles/ afterwards).
timeout 500 python tools/gp_check.py > gpurun_out/r02_gpcheck.log 2>&1; tail -1 gpurun_out/r02_gpcheck.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_reference_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 500 --csv --log-file gpurun_out/r02_launches_b1_fp32.csv python tools/quick_fwd.py fp32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/r02_launches_b32_bf16.csv python tools/quick_b32.py bf16 2 > /dev/null 2>&1
rm -f gpurun_out/r02_attention_tc_vs_ffma.jsonl; for a in "1 100" "1 537" "32 600" "32 1200"; do for m in 1 0; do timeout 100 python tools/profile_attn.py $a $m | tail -1 >> gpurun_out/r02_attention_tc_vs_ffma.jsonl; done; done
timeout 600 python tools/sweep.py --quick --corner --no-cfg3 --out gpurun_out/r02_sweep_cfg4_with_corner.json > /dev/null 2>&1; ls -la gpurun_out/r02_sweep_cfg4_with_corner.json
for p in fp32 bf16 tf32; do python tools/profile_dominant.py $p 511 | tail -1; done > gpurun_out/r02_dominant_launch.log 2>&1; cat gpurun_out/r02_dominant_launch.log
echo final_n1_done
