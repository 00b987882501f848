# Round-end measurement pass on one B200 (everything lands in gpurun_out/r02_*; copied into profiles/ afterwards).  Every step under a timeout.
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_full.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_full.log | tee gpurun_out/r02_pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"; tail -c 700 gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_reference_n1.json
echo final_n1_done
