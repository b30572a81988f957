#!/bin/bash
# A/B of pan_fast kernel variants: tools/probes/ab_pan.sh lib1.so lib2.so ...  (run on the GPU box; ROWS/ITERS from the env)
export ROWS=${ROWS:-131072} ITERS=${ITERS:-20} CHECKSUM=1
for lib in "$@"; do
  for rep in 1 2; do
    echo "== $lib"; OIP_B200_LIB=$lib python tools/profile_pan.py 2>&1 | tail -2
  done
done
