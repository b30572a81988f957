#!/bin/bash
# C5 (BASELINE configs[4]) at N GPUs: the bench workload at several strip lengths, device-timed value + parity per size.
#   bash tools/sweep_multi.sh N   (under gpurun --gpus N)
N=${1:-8}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
P=29600
: > $O/r02_sweep_n$N.jsonl
for rows in 5208 65536 262144 1048576 1302528; do   # 256 MB, 3.2 GB, 12.9 GB, 51.5 GB (C4), 64 GB of input
  P=$((P+1))
  if [ "$N" = "1" ]; then python bench.py --rows $rows --steps 10 --warmup 3 --no-e2e --no-framed >> $O/r02_sweep_n$N.jsonl 2>> $O/r02_sweep_n$N.err
  else $TR --master-port $P bench.py --gpus $N --rows $rows --steps 10 --warmup 3 --no-e2e --no-framed >> $O/r02_sweep_n$N.jsonl 2>> $O/r02_sweep_n$N.err; fi
done
python - <<PY
import json
for l in open("$O/r02_sweep_n$N.jsonl"):
    d = json.loads(l)
    print(d["n_gpus"], d["config"]["total_rows"] if "total_rows" in d["config"] else "", round(d["value"], 1), "Gpx/s", round(d["ms_per_step"], 3), "ms", d["parity_ok"], round(d["roofline"]["frac"], 3))
PY
