timeout 400 python tools/gp_check.py > gpurun_out/c13_gpcheck.log 2>&1; tail -1 gpurun_out/c13_gpcheck.log
for fx in 1 0; do echo "EV_SPLITK_FIXUP=$fx"; EV_SPLITK_FIXUP=$fx timeout 200 python tools/quick_fwd.py fp32 | tail -1; EV_SPLITK_FIXUP=$fx timeout 200 python tools/quick_b32.py bf16 5 | tail -1; done 2>&1 | tee gpurun_out/c13_fixup.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python tools/profile_dominant.py fp32 511 | tail -1
python tools/profile_dominant.py bf16 511 | tail -1
python tools/profile_dominant.py tf32 511 | tail -1
ncu --set full --clock-control none --import-source on -k regex:conv1d_gp -s 2 -c 1 -o gpurun_out/c13_ncu_dominant_fp32 python tools/profile_dominant.py fp32 511 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 600 --csv --log-file gpurun_out/c13_launches_b1_fp32.csv python tools/quick_fwd.py fp32 > /dev/null 2>&1
