#!/bin/bash
# A/B of sim_stream_kernel builds and chunk widths: prints the similarity-kernel time per Fit.
# usage: tools/stream_sweep.sh <workload> <lib1> [lib2 ...]   (libs under recommend-sys_b200/)
wl=$1; shift
for lib in "$@"; do
  for jc in 128 256 512; do
    out=$(RS_KNN_LIB=$PWD/recommend-sys_b200/$lib RS_KNN_STREAM_JC=$jc python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
    echo "$wl $lib jc=$jc $(echo "$out" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("sim_ms=%.2f predict_ms=%.2f step_ms=%.2f e2e_ms=%.2f"%(j["kernel_ms"]["sim"],j["kernel_ms"]["predict"],j["ms_per_step"],j["e2e"]["ms_per_step"]))' 2>&1 | tail -1)"
  done
done
