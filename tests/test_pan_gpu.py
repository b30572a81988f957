"""GPU parity: fused PAN pipeline (and its stand-alone forms) vs the CPU oracle. Bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHIFTS = [(0.0, 0.0), (3.0, -2.0), (0.5, 0.5), (1.37, -2.61), (-1.984375, 4.015625), (0.015625, 0.0), (-5.3, 7.77),
          (-0.83, 3.19)]


@pytest.fixture(params=["fast", "generic"], autouse=True)
def pan_mode(request, ctx):
    """every test runs twice: planner splits work between pan_fast_kernel and pan_kernel / generic kernel only"""
    ctx.set_option("pan_fast", 1 if request.param == "fast" else 0)
    yield request.param
    ctx.set_option("pan_fast", 1)


def _rand_img(rng, h, w, full=True):
    hi = 65536 if full else 4096
    return rng.integers(0, hi, (h, w), dtype=np.uint16)


def _kb(rng, w):
    kb = np.empty((w, 2), np.float64)
    kb[:, 0] = 0.95 + 0.1 * rng.random(w)
    kb[:, 1] = 8.0 * rng.random(w)
    return kb


def _dev(a):
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("dX,dY", SHIFTS)
def test_shift_single_section(ctx, oracle_mod, dX, dY):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(7)
    src = _rand_img(rng, 300, 1024)
    want = oracle_mod.prestitch_shift(src, dX, dY)
    got = ops.prestitch_shift(ctx, _dev(src), dX, dY).cpu().numpy()
    assert np.array_equal(got, want), f"{int((got != want).sum())} px differ"


@pytest.mark.parametrize("dX,dY", [(1.37, -2.61), (-0.83, 3.19), (0.0, 0.0), (2.5, 40.25), (-3.0, -17.5)])
@pytest.mark.parametrize("rows", [1500, 1337, 1000, 953])
def test_shift_multi_section_with_stale_rows(ctx, oracle_mod, dX, dY, rows):
    """small section_rows/row_guard exercise section edges, ucut/bcut and the stale-row quirk"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(11)
    src = _rand_img(rng, rows, 512)
    want = oracle_mod.prestitch_shift(src, dX, dY, section_rows=400, row_guard=450)
    got = ops.prestitch_shift(ctx, _dev(src), dX, dY, section_rows=400, row_guard=450).cpu().numpy()
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


def test_rrc_inplace(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(3)
    w = 12288
    img = _rand_img(rng, 257, w)
    kb = _kb(rng, w)
    kb[5] = (1.0, -0.5)      # trunc toward zero
    kb[6] = (1.0, -70000.0)  # negative -> wraps mod 2^16
    kb[7] = (2.0, 0.0)       # > 65535 -> wraps
    kb[8] = (1.0, 4294967296.0)  # outside int32 -> 0
    want = oracle_mod.rrc(img, kb)
    got = ops.inplace_rrc(ctx, _dev(img), _dev(kb)).cpu().numpy()
    assert np.array_equal(got, want)


def test_concat(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(5)
    for n, w, f in [(2, 12288, 100), (3, 1024, 25), (1, 520, 0), (4, 256, 3)]:
        ccds = [_rand_img(rng, 70, w) for _ in range(n)]
        want = oracle_mod.stitch_concat(ccds, f)
        got = ops.stitch_big_raw(ctx, [_dev(c) for c in ccds], f).cpu().numpy()
        assert np.array_equal(got, want), (n, w, f)


@pytest.mark.parametrize("fmt", ["le", "be"])
@pytest.mark.parametrize("n,w,f,rows,S,G", [(2, 1536, 100, 700, 30000, 32767), (3, 1024, 100, 1100, 400, 450),
                                            (3, 8192, 100, 96, 30000, 32767)])
def test_fused_pipeline(ctx, oracle_mod, fmt, n, w, f, rows, S, G):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(13)
    ccds = [_rand_img(rng, rows, w, full=False) for _ in range(n)]
    kbs = [_kb(rng, w) for _ in range(n)]
    dX = [0.0, 1.37, -0.83, 2.2][:n]
    dY = [0.0, -2.61, 3.19, 0.4][:n]
    want = oracle_mod.pan_pipeline(ccds, kbs, dX, dY, f, S, G)
    dev = [_dev(c if fmt == "le" else c.byteswap()) for c in ccds]
    got = ops.pan_pipeline(ctx, dev, [_dev(k) for k in kbs], dX, dY, f,
                           fmt=ops.FMT_LE16 if fmt == "le" else ops.FMT_BE16, section_rows=S, row_guard=G)
    got = got.cpu().numpy()
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


def test_fused_pipeline_unaligned_generic_loader(ctx, oracle_mod):
    """w not a multiple of 8 -> generic loader path"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(17)
    n, w, f, rows = 2, 1001, 7, 333
    ccds = [_rand_img(rng, rows, w) for _ in range(n)]
    kbs = [_kb(rng, w) for _ in range(n)]
    want = oracle_mod.pan_pipeline(ccds, kbs, [0, 1.37], [0, -2.61], f)
    got = ops.pan_pipeline(ctx, [_dev(c) for c in ccds], [_dev(k) for k in kbs], [0, 1.37], [0, -2.61], f).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", ["negative_b", "wrap", "out_of_int32", "no_rrc"])
def test_fused_pipeline_rrc_modes(ctx, oracle_mod, case):
    """RRC corner cases inside the fused kernels: the fast kernel picks an exact per-warp mode (none / non-negative
    / general), ref imageop.h:134 semantics: truncation toward zero, wrap mod 2^16, x86 cvttsd2si out of range"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(23)
    n, w, f, rows = 2, 2048, 100, 300
    ccds = [_rand_img(rng, rows, w) for _ in range(n)]
    kbs = [_kb(rng, w) for _ in range(n)]
    for kb in kbs:
        if case == "negative_b":
            kb[::7, 1] = -300.5 * rng.random(len(kb[::7]))
            kb[3::11, 0] *= -1.0
        elif case == "wrap":
            kb[::5, 0] = 2.0 + rng.random(len(kb[::5]))
        elif case == "out_of_int32":
            kb[100:140, 1] = 4294967296.0
            kb[900:910, 1] = -4294967296.0
            kb[1500, 1] = np.nan
    dX, dY = [0.0, 1.37], [0.0, -2.61]
    use_kb = case != "no_rrc"
    ident = np.tile(np.array([1.0, 0.0]), (w, 1))            # k*s + b == s exactly
    want = oracle_mod.pan_pipeline(ccds, kbs if use_kb else [ident] * n, dX, dY, f)
    got = ops.pan_pipeline(ctx, [_dev(c) for c in ccds], [_dev(k) for k in kbs] if use_kb else None, dX, dY, f).cpu().numpy()
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} px differ, first {bad[:5].tolist()}"


def test_fused_pipeline_tall_tiles_and_stage_depths(ctx, oracle_mod, pan_mode):
    """warp-tile heights that are not multiples of the 4-row stage, every TMA stage depth"""
    from opticalimageprocessor_b200 import ops
    if pan_mode != "fast":
        pytest.skip("fast-kernel tunables")
    rng = np.random.default_rng(29)
    n, w, f, rows = 3, 1024, 26, 613
    ccds = [_rand_img(rng, rows, w) for _ in range(n)]
    kbs = [_kb(rng, w) for _ in range(n)]
    dX, dY = [0.0, 1.37, -0.83], [0.0, -2.61, 3.19]
    want = oracle_mod.pan_pipeline(ccds, kbs, dX, dY, f)
    try:
        for stages, th in [(2, 16), (3, 37), (5, 128), (8, 1000)]:
            ctx.set_option("pan_fast_stages", stages)
            ctx.set_option("pan_fast_rows", th)
            got = ops.pan_pipeline(ctx, [_dev(c) for c in ccds], [_dev(k) for k in kbs], dX, dY, f).cpu().numpy()
            bad = np.argwhere(got != want)
            assert bad.size == 0, f"stages={stages} rows={th}: {len(bad)} px differ, first {bad[:5].tolist()}"
    finally:
        ctx.set_option("pan_fast_stages", 4)
        ctx.set_option("pan_fast_rows", 128)
