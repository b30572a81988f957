"""CLI end to end on a RAM disk: `stitch` (RAW, two 12288-px CCD files -> stitched RAW) and `auxsep` wall time against
the PCIe bound of the same bytes (tools/pcie_probe.py figure: ~46 GB/s per direction on this box).
    python tools/cli_stream_bench.py [lines]"""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "opticalimageprocessor_b200", "OpticalImageProcessor")
lines = int(sys.argv[1]) if len(sys.argv) > 1 else 170000
W, FOLD = 12288, 200
d = "/dev/shm/oip_cli_bench"
os.makedirs(d, exist_ok=True)
g = torch.Generator(device="cuda").manual_seed(1)
for name in ("L.RAW", "R.RAW"):
    with open(os.path.join(d, name), "wb") as f:
        for r in range(0, lines, 16384):
            n = min(16384, lines - r)
            f.write(torch.randint(0, 4096, (n, W), dtype=torch.int16, device="cuda", generator=g).cpu().numpy().tobytes())
in_bytes = 2 * lines * W * 2
out_w = 2 * (W - FOLD // 2)
out_bytes = lines * out_w * 2
res = {"lines": lines, "in_bytes": in_bytes, "out_bytes": out_bytes}
for rep in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([CLI, "stitch", "--image1", "L.RAW", "--image2", "R.RAW", "-c", str(FOLD), "-o", "OUT.RAW"], cwd=d, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    res[f"stitch_wall_s_{rep}"] = dt
pcie = 46e9
res["pcie_bound_s"] = max(in_bytes, out_bytes) / pcie
res["stitch_vs_pcie_bound"] = min(res["stitch_wall_s_0"], res["stitch_wall_s_1"]) / res["pcie_bound_s"]
res["stitch_gbs_in_plus_out"] = (in_bytes + out_bytes) / min(res["stitch_wall_s_0"], res["stitch_wall_s_1"]) / 1e9
# correctness: sampled row blocks of the product == L[:, :W-f] | R[:, f:]
L = np.memmap(os.path.join(d, "L.RAW"), np.uint16, "r").reshape(lines, W)
R = np.memmap(os.path.join(d, "R.RAW"), np.uint16, "r").reshape(lines, W)
O = np.memmap(os.path.join(d, "OUT.RAW"), np.uint16, "r").reshape(lines, out_w)
f = FOLD // 2
ok = True
for r0 in list(range(0, lines - 64, max(1, lines // 37))) + [lines - 64]:
    ok = ok and np.array_equal(O[r0:r0 + 64], np.concatenate([L[r0:r0 + 64, :W - f], R[r0:r0 + 64, f:]], axis=1))
res["stitch_output_ok"] = bool(ok)
# memcpy-speed reference of the RAM disk itself: cp of one input file
t0 = time.perf_counter(); subprocess.check_call(["cp", os.path.join(d, "L.RAW"), os.path.join(d, "L.copy")]); res["cp_one_file_s"] = time.perf_counter() - t0
print(json.dumps(res))
for n in os.listdir(d):
    os.remove(os.path.join(d, n))
