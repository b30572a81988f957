"""How far is the offset estimate (oip_phase_correlate_u16, cuFFT) from cv2.phaseCorrelate -- the reference's own call at
stitcher.h:180 -- and how often does that difference change the QUANTISED map of PreStitch (cvRound(float(x + dX) * 32),
ref stitcher.h:96-97 + cv::remap's 1/32 grid) for at least one of the 12288 columns / 30000 section rows?
    python tools/stt_tolerance.py [n_shifts] [rows] [cols]      -> one JSON line (profiles/r02_stt_tolerance.json)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2, torch
from opticalimageprocessor_b200 import ops
from test_phasecorr_cpu import _pair


def map_fixed(n, d):
    return np.rint((np.arange(n, dtype=np.float64) + d).astype(np.float32) * np.float32(32)).astype(np.int64)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    cols = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    ctx = ops.Context(0)
    rng = np.random.default_rng(2026)
    ddx, ddy, flips_x, flips_y, any_flip = [], [], [], [], 0
    for i in range(n):
        dx, dy = rng.uniform(-4, 4), rng.uniform(-6, 6)
        a, b = _pair(rows, cols, dx, dy, seed=10_000 + i)
        (cx, cy), cr = cv2.phaseCorrelate(a.astype(np.float32), b.astype(np.float32))
        gx, gy, gr = ops.phase_correlate(ctx, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
        ddx.append(abs(gx - cx)); ddy.append(abs(gy - cy))
        fx = int((map_fixed(12288, gx) != map_fixed(12288, cx)).sum())          # columns of a 12288-px line
        fy = int((map_fixed(30000, gy) != map_fixed(30000, cy)).sum())          # rows of a 30000-row section
        flips_x.append(fx); flips_y.append(fy); any_flip += int(fx + fy > 0)
    q = lambda v: [float(np.quantile(v, p)) for p in (0.5, 0.9, 0.99, 1.0)]
    print(json.dumps({"shifts": n, "slice": [rows, cols], "abs_diff_dx_px_q50_q90_q99_max": q(ddx), "abs_diff_dy_px_q50_q90_q99_max": q(ddy),
                      "runs_with_any_changed_map_entry": any_flip, "fraction_of_runs": any_flip / n,
                      "changed_columns_of_12288_mean_max": [float(np.mean(flips_x)), int(max(flips_x))],
                      "changed_rows_of_30000_mean_max": [float(np.mean(flips_y)), int(max(flips_y))]}))


if __name__ == "__main__":
    main()
