timeout 300 python tools/gp_check.py > gpurun_out/c11_gpcheck.log 2>&1; tail -4 gpurun_out/c11_gpcheck.log
for pdl in 0 1 2; do echo "EV_PDL=$pdl"; EV_PDL=$pdl python tools/quick_fwd.py fp32 | tail -1; EV_PDL=$pdl python tools/quick_b32.py bf16 5 | tail -1; done 2>&1 | tee gpurun_out/c11_pdl.log
for w in none w all; do
  for shp in "gp:bf16x3 256 11 5 4296" "gp:bf16x3 256 3 1 4296" "gp:bf16x3 128 11 5 34368" "gp:bf16x3 64 3 1 68736" "gp:bf16x3 32 3 1 137472" "bf16x3 384:1536 3 1 537" "fp32 384:1152 1 1 100" "fp32 384:1536 3 1 100"; do
    WARM=$w KSPLIT=4 python tools/profile_conv.py $shp | tail -1
  done
done > gpurun_out/c11_warm.jsonl 2>&1
cat gpurun_out/c11_warm.jsonl | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
