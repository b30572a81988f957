#!/bin/bash
# where does the CLI's wall time go on this box's RAM disk?  (run under gpurun from the repo root)
set -e
D=/dev/shm/oip_sweep; mkdir -p $D; cd $D
CLI=$GRAFT_REPO_ROOT/opticalimageprocessor_b200/OpticalImageProcessor
python - <<'PY'
import numpy as np
rng=np.random.default_rng(1)
blk=rng.integers(0,4096,(8192,12288),dtype=np.uint16)
for n in ("L.RAW","R.RAW"):
    with open(n,"wb") as f:
        for i in range(16): f.write(blk.tobytes())
PY
ls -la
echo "--- startup (-v)"; ( time $CLI -v ) 2>&1 | grep -E "real|1\." 
echo "--- dd read one file (1 thread)"; ( time dd if=L.RAW of=/dev/null bs=8M 2>/dev/null ) 2>&1 | grep real
echo "--- dd copy one file (1 thread)"; ( time dd if=L.RAW of=X.RAW bs=8M 2>/dev/null ) 2>&1 | grep real; rm -f X.RAW
for win in 2048 2048 1024; do
  echo "--- window $win"; ( time OIP_TIMING=1 OIP_STITCH_WINDOW=$win $CLI stitch --image1 L.RAW --image2 R.RAW -c 200 -o OUT.RAW ) 2>&1 | grep -E "real|stitch:|pipeline thread"
done
echo "--- cp both inputs"; ( time sh -c "cp L.RAW A.RAW; cp R.RAW B.RAW" ) 2>&1 | grep real
rm -rf $D
