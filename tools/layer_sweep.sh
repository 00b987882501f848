#!/bin/bash
# Per-layer microbenchmarks of the vocoder's conv shapes: granule-planar kernel vs the round-1 time-major kernel.
#   bash tools/layer_sweep.sh [B] [F]  > gpurun_out/layer_sweep.jsonl        (default B=8, F=1024: the cfg4 point)
B=${1:-8}; F=${2:-1024}
for prec in tf32 fp32 bf16x3 bf16; do
  for spec in "256 3 1 $((F*8))" "256 11 5 $((F*8))" "128 3 1 $((F*64))" "128 11 5 $((F*64))" "64 3 1 $((F*128))" "64 11 5 $((F*128))" "32 3 1 $((F*256))" "32 7 3 $((F*256))" "32 11 5 $((F*256))"; do
    set -- $spec
    timeout 120 python tools/profile_conv.py gp:$prec $1 $2 $3 $4 $B 4
    [ "$prec" != bf16 ] && [ "$prec" != bf16x3 ] && timeout 120 python tools/profile_conv.py $prec $1 $2 $3 $4 $B 4
  done
  timeout 120 python tools/profile_conv.py gp:$prec 512:2048:8 3 1 $F $B 4
  timeout 120 python tools/profile_conv.py gp:$prec 128:128:2 3 1 $((F*64)) $B 4
done
