"""GPU parity: host-buffer entry point (H2D / kernels / D2H overlapped in row blocks) vs the oracle."""
import numpy as np
import pytest
import torch

from opticalimageprocessor_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,S,G,pinned", [(5000, 30000, 32767, True), (4500, 2100, 2200, False)])
def test_pan_pipeline_host_blocks(ctx, oracle_mod, rows, S, G, pinned):
    """> 2 row blocks of 2048 lines, with section edges and stale rows landing in different blocks"""
    from opticalimageprocessor_b200 import ops
    n, w, f = 3, 512, 20
    ccds = [synth.strip_dn(w, rows, 40 + i) for i in range(n)]
    kbs = [synth.rrc_coeffs(w, 50 + i) for i in range(n)]
    dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
    want = oracle_mod.pan_pipeline(ccds, kbs, dX, dY, f, S, G)
    host = [torch.from_numpy(c.byteswap()) for c in ccds]
    out = torch.empty((rows, ops.pan_out_width(n, w, f)), dtype=torch.uint16)
    if pinned:
        host = [h.pin_memory() for h in host]
        out = out.pin_memory()
    ops.pan_pipeline_host(ctx, host, [torch.from_numpy(k) for k in kbs], dX, dY, f, out, fmt=ops.FMT_BE16,
                          section_rows=S, row_guard=G)
    got = out.numpy()
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"
