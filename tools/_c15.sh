timeout 500 python tools/gp_check.py > gpurun_out/c15_gpcheck.log 2>&1; grep -c '"ok": true' gpurun_out/c15_gpcheck.log; grep '"ok": false' gpurun_out/c15_gpcheck.log | head -5 | cut -c1-300; tail -1 gpurun_out/c15_gpcheck.log
for grp in 1 0; do echo "EV_VOC_GROUP=$grp"; EV_VOC_GROUP=$grp timeout 200 python tools/quick_fwd.py fp32 | tail -1; EV_VOC_GROUP=$grp timeout 200 python tools/quick_fwd.py tf32 | tail -1; done 2>&1 | tee gpurun_out/c15_group.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
