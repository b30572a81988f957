#!/bin/bash
# profiles of round 2, run under gpurun from the repo root:  bash tools/collect_profiles.sh
# every ncu run follows a plain run of the same command that exited 0; numbers printed under ncu are never bench values
set -x
O=gpurun_out
T=/tmp/oip_prof; mkdir -p $T
# ---- launch list of the bench command (short strip, no 2 s warm-up: ncu serialises every launch)
python bench.py --rows 131072 --steps 2 --warmup 3 --no-e2e --no-framed --no-parity > $T/plain.json 2>/dev/null || exit 1
OIP_BENCH_WARM_SECONDS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
   python bench.py --rows 131072 --steps 2 --warmup 3 --no-e2e --no-framed --no-parity > /dev/null 2>&1
# ---- top kernel of the bench: pan_fast_kernel on the C4 strip (one launch, full set)
ROWS=1048576 python tools/profile_pan.py > $O/r02_profile_pan_plain.txt 2>&1 || exit 1
ROWS=1048576 ncu --set full --clock-control none --import-source on -k regex:pan_fast -s 3 -c 1 -o $T/pan python tools/profile_pan.py > /dev/null 2>&1
python tools/ncu_summary.py $T/pan.ncu-rep 25769803776 > $O/r02_pan_fast_kernel_ncu_full.txt 2>&1
# ---- band alignment
python tools/profile_mss.py > $O/r02_profile_mss_plain.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mss_fast -s 3 -c 1 -o $T/mss python tools/profile_mss.py > /dev/null 2>&1
python tools/ncu_summary.py $T/mss.ncu-rep 201326592 > $O/r02_mss_fast_kernel_ncu_full.txt 2>&1
# ---- stage 1 (911 MB reference-geometry downlink)
FRAMES=24 python tools/profile_stage1.py > $O/r02_profile_stage1_plain.txt 2>&1 || exit 1
for k in aos_fused imtr_validate find_sig4; do
  FRAMES=24 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $T/$k python tools/profile_stage1.py > /dev/null 2>&1
  python tools/ncu_summary.py $T/$k.ncu-rep 911000000 > $O/r02_${k}_kernel_ncu_full.txt 2>&1
done
FRAMES=24 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_stage1.csv python tools/profile_stage1.py > /dev/null 2>&1
ls -la $O | tail -20
