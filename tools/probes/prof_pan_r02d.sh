set -x
O=gpurun_out
T=/tmp/oip_prof; mkdir -p $T
python bench.py --rows 131072 --steps 2 --warmup 3 --no-e2e --no-framed --no-parity > $T/plain.json 2>/dev/null || exit 1
OIP_BENCH_WARM_SECONDS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02d_launches_bench.csv \
   python bench.py --rows 131072 --steps 2 --warmup 3 --no-e2e --no-framed --no-parity > /dev/null 2>&1
ROWS=1048576 python tools/profile_pan.py > $O/r02d_profile_pan_plain.txt 2>&1 || exit 1
ROWS=1048576 ncu --set full --clock-control none --import-source on -k regex:pan_fast -s 3 -c 1 -o $T/pan python tools/profile_pan.py > /dev/null 2>&1
python tools/ncu_summary.py $T/pan.ncu-rep 25769803776 > $O/r02d_pan_fast_kernel_ncu_full.txt 2>&1
ls -la $O | tail -5
