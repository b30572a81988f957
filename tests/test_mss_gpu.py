"""GPU parity, MSS path: band split + RRC + polynomial bicubic remap + merge vs the CPU oracle. Bit-exact."""
import numpy as np
import pytest
import torch

from opticalimageprocessor_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["fast", "generic"], autouse=True)
def mss_mode(request, ctx):
    """every test runs twice: planner splits work between mss_fast_kernel and band_align_kernel / generic kernel only"""
    ctx.set_option("mss_fast", 1 if request.param == "fast" else 0)
    yield request.param
    ctx.set_option("mss_fast", 1)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _coeffs(scale=1.0):
    cX = [[0.8 + 0.1 * b, -1.5e-4 * (b + 1) * scale] for b in range(4)]
    cY = [[-3.2 + b, 2e-4 * (b + 1) * scale, -1e-8 * (b - 1.5) * scale] for b in range(4)]
    return cX, cY


def _oracle_full(oracle_mod, mixed, kbs, cX, cY, **kw):
    planes = oracle_mod.mss_split(mixed)                      # ref preproc.h:56-80
    if kbs is not None:
        planes = [oracle_mod.rrc(p, k) for p, k in zip(planes, kbs)]  # ref preproc.h:202-222
    return oracle_mod.band_align(planes, cX, cY, **kw)


@pytest.mark.parametrize("keep", [False, True])
@pytest.mark.parametrize("lines,wb,lps,overlap,off,scale", [(700, 96, 300, 40, 0, 10.0), (650, 512, 256, 32, 10, 1.0),
                                                            (300, 3072, 400, 20, 0, 1.0), (900, 1000, 333, 17, 5, 3.0)])
@pytest.mark.parametrize("fmt", ["le", "be"])
def test_band_align_merge(ctx, oracle_mod, keep, lines, wb, lps, overlap, off, scale, fmt):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(21)
    mixed = rng.integers(0, 4096, (lines, 4 * wb), dtype=np.uint16)
    kbs = [synth.rrc_coeffs(wb, 100 + b) for b in range(4)]
    cX, cY = _coeffs(scale)
    kw = dict(lines_per_section=lps, line_offset=off, overlap=overlap, keep_leading=keep, min_process_lines=64)
    n_w, want = _oracle_full(oracle_mod, mixed, kbs, cX, cY, **kw)
    d = _dev(mixed if fmt == "le" else mixed.byteswap())
    n_g, got = ops.band_align(ctx, d, wb, [_dev(k) for k in kbs], cX, cY,
                              fmt=ops.FMT_LE16 if fmt == "le" else ops.FMT_BE16, **kw)
    assert n_g == n_w
    got = got.cpu().numpy()
    bad = np.argwhere(got[:n_g] != want[:n_w])
    assert bad.size == 0, f"{len(bad)} samples differ, first {bad[:5].tolist()}"


def test_band_align_full_range_no_rrc(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(22)
    lines, wb = 400, 640
    mixed = rng.integers(0, 65536, (lines, 4 * wb), dtype=np.uint16)
    cX, cY = _coeffs(5.0)
    kw = dict(lines_per_section=20000, line_offset=0, overlap=52, keep_leading=False, min_process_lines=150)
    n_w, want = _oracle_full(oracle_mod, mixed, None, cX, cY, **kw)
    n_g, got = ops.band_align(ctx, _dev(mixed), wb, None, cX, cY, **kw)
    assert n_g == n_w and np.array_equal(got.cpu().numpy()[:n_g], want[:n_w])


def test_band_align_argument_errors(ctx):
    from opticalimageprocessor_b200 import ops
    from opticalimageprocessor_b200.capi import OipError
    mixed = torch.zeros((8000, 32), dtype=torch.uint16, device="cuda")
    z2, z3 = np.zeros((4, 2)), np.zeros((4, 3))
    # same conditions and wording as ref preproc.h:355-367
    for kw, msg in [(dict(overlap=3001), "exceeds maximum"), (dict(lines_per_section=32768), "OpenCV allowed"),
                    (dict(lines_per_section=1000), "too small"), (dict(line_offset=6600), "Too few")]:
        with pytest.raises(OipError, match=msg):
            ops.band_align(ctx, mixed, 8, None, z2, z3, **kw)


def test_stitch_tiff_geometry(ctx, oracle_mod):
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(4)
    for n, w, f, bm in [(2, 3072, 25, None), (2, 100, 7, [3, 2, 1, 4]), (3, 64, 5, [4, 3, 2, 1]), (1, 33, 0, None)]:
        imgs = [rng.integers(0, 65536, (17, w, 4), dtype=np.uint16) for _ in range(n)]
        want = oracle_mod.stitch_concat_c4(imgs, f, bm)
        got = ops.stitch_tiff_geometry(ctx, [_dev(i) for i in imgs], f, bm).cpu().numpy()
        assert np.array_equal(got, want)


def test_band_align_long_section_reference_geometry(ctx, oracle_mod):
    """reference geometry (3072 px bands, 20000-line sections, 520 overlap) across two sections and several
    power-of-two row crossings of the map; small tiles so that every fast / generic seam is exercised"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(31)
    lines, wb = 21600, 3072
    mixed = rng.integers(0, 4096, (lines, 4 * wb), dtype=np.uint16)
    kbs = [synth.rrc_coeffs(wb, 200 + b) for b in range(4)]
    kbs[2][::9, 1] = -20.5          # general RRC mode on one band
    cX, cY = _coeffs(1.0)
    n_w, want = _oracle_full(oracle_mod, mixed, kbs, cX, cY)
    ctx.set_option("mss_fast_rows", 61)
    try:
        n_g, got = ops.band_align(ctx, _dev(mixed), wb, [_dev(k) for k in kbs], cX, cY)
    finally:
        ctx.set_option("mss_fast_rows", 128)
    assert n_g == n_w
    got = got.cpu().numpy()
    bad = np.argwhere(got[:n_g] != want[:n_w])
    assert bad.size == 0, f"{len(bad)} samples differ, first {bad[:5].tolist()}"


def test_sections_computed_on_their_own_equal_the_whole_strip(ctx):
    """C3 sharding: the band alignment shards by section -- every rank's run of sections, computed in one call
    from only the source lines the rank holds, is bit-identical to the whole-strip call"""
    from opticalimageprocessor_b200 import ops, sharding
    lines, wb, lps, ov = 9000, 256, 2500, 300
    g = torch.Generator(device="cuda").manual_seed(3)
    mss = torch.randint(0, 4096, (lines, 4 * wb), device="cuda", dtype=torch.int32, generator=g).to(torch.uint16)
    kbs = [torch.from_numpy(synth.rrc_coeffs(wb, 60 + b)).cuda() for b in range(4)]
    cX, cY = _coeffs(2.0)
    for keep, off in [(False, 0), (True, 7)]:
        n, whole = ops.band_align(ctx, mss, wb, kbs, cX, cY, lines_per_section=lps, overlap=ov, keep_leading=keep, line_offset=off)
        secs = sharding.mss_sections(lines, lps, ov, off, keep, 1500)
        assert sum(s[4] for s in secs) == n and len(secs) >= 3
        for world in (2, 3):
            got = torch.zeros_like(whole)
            for rank in range(world):
                mine = sharding.mss_rank_sections(secs, world, rank)
                lo, hi = mine[0][0], mine[-1][0] + mine[-1][1]
                shard = mss[lo:hi].clone()                       # the rank holds only these source lines
                o0, no = mine[0][3], sum(s[4] for s in mine)
                k = ops.band_align_sections(ctx, shard, wb, kbs, cX, cY, mine, secs, got[o0:o0 + no], total_lines=lines,
                                            lines_per_section=lps, overlap=ov, line_offset=off, keep_leading=keep, src_row0=lo)
                assert k == no
            assert torch.equal(got[:n].view(torch.int16), whole[:n].view(torch.int16))


@pytest.mark.parametrize("cy1,cx1", [(0.5, -1.5e-4), (-0.9, 0.0), (2e-4, 0.6), (0.3, -0.4)])
def test_steep_polynomials_are_computed(ctx, oracle_mod, cy1, cx1):
    """a polynomial whose tap rows / columns spread beyond the staged window of a tile (round 1: OIP_E_UNSUPPORTED) is
    resampled tap by tap inside the generic kernel: still bit-identical to the oracle, and no error is reported"""
    from opticalimageprocessor_b200 import ops
    rng = np.random.default_rng(31)
    lines, wb = 420, 384
    mixed = rng.integers(0, 65536, (lines, 4 * wb), dtype=np.uint16)
    kbs = [synth.rrc_coeffs(wb, 120 + b) for b in range(4)]
    cX = [[0.8 + 0.1 * b, cx1] for b in range(4)]
    cY = [[-3.2 + b, cy1, -1e-8] for b in range(4)]
    kw = dict(lines_per_section=200, line_offset=0, overlap=24, keep_leading=True, min_process_lines=64)
    n_w, want = _oracle_full(oracle_mod, mixed, kbs, cX, cY, **kw)
    n_g, got = ops.band_align(ctx, _dev(mixed), wb, [_dev(k) for k in kbs], cX, cY, **kw)
    ctx.sync()
    assert n_g == n_w
    bad = np.argwhere(got.cpu().numpy()[:n_g] != want[:n_w])
    assert bad.size == 0, f"{len(bad)} samples differ, first {bad[:5].tolist()}"
