"""host/tiff_io.hpp against libtiff as shipped in cv2 (the library the reference writes its TIFFs with,
ref preproc.h:167-185, imageop.h:444): files we write are read back by cv2.imread with the same pixels, and uncompressed
files cv2 writes are read by our reader; compressed input is refused with a clear message."""
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tiff") / "tiff_io_host_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", out, os.path.join(HERE, "native", "tiff_io_host_test.cpp")])
    return out


@pytest.mark.parametrize("h,w,spp", [(37, 53, 1), (1200, 3072, 4), (5, 70000, 1), (3000, 1500, 4)])
def test_written_files_are_read_by_libtiff(exe, tmp_path, h, w, spp):
    px = np.random.default_rng(h + w).integers(0, 65536, (h, w, spp), dtype=np.uint16)
    raw, tif = str(tmp_path / "in.raw"), str(tmp_path / "out.TIFF")
    px.tofile(raw)
    subprocess.check_call([exe, "write", tif, str(w), str(h), str(spp), raw])
    img = cv2.imread(tif, cv2.IMREAD_UNCHANGED)
    assert img is not None and img.dtype == np.uint16
    if spp == 1:
        assert np.array_equal(img, px[:, :, 0])
    else:  # cv::imread hands a 4-sample RGBA file back as BGRA: samples 0 and 2 swapped
        assert np.array_equal(img, px[:, :, [2, 1, 0, 3]])
    # and our own reader
    back = str(tmp_path / "back.raw")
    dims = subprocess.check_output([exe, "read", tif, back], text=True).split()
    assert [int(v) for v in dims] == [w, h, spp]
    assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, spp), px)


@pytest.mark.parametrize("shape", [(300, 200, 4), (257, 333, 1), (64, 5000, 4), (2000, 96, 1)])
def test_reads_libtiff_files_uncompressed_and_lzw(exe, tmp_path, shape):
    """cv::imwrite defaults (LZW + horizontal predictor, many strips) = what the reference's own TIFF products look like"""
    rng = np.random.default_rng(shape[0])
    smooth = (np.cumsum(rng.integers(-3, 4, shape), axis=1) + 2000).astype(np.uint16)   # compressible, exercises long LZW strings
    noise = rng.integers(0, 65536, shape, dtype=np.uint16)
    for img in (smooth, noise, np.zeros(shape, np.uint16)):
        im = img[:, :, 0] if shape[2] == 1 else img
        plain, lzw, back = str(tmp_path / "plain.TIFF"), str(tmp_path / "lzw.TIFF"), str(tmp_path / "back.raw")
        assert cv2.imwrite(plain, im, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
        assert cv2.imwrite(lzw, im)  # cv::imwrite default: LZW, what the reference produces
        want = img[:, :, [2, 1, 0, 3]] if shape[2] == 4 else img  # file order = RGBA
        for path in (plain, lzw):
            subprocess.check_call([exe, "read", path, back])
            assert np.array_equal(np.fromfile(back, np.uint16).reshape(shape), want), path


@pytest.mark.parametrize("h,w,spp,kind", [(37, 53, 1, "noise"), (1200, 3072, 4, "smooth"), (5, 70000, 1, "smooth"), (3000, 1500, 4, "noise"),
                                          (700, 3072, 4, "zeros"), (4096, 1024, 1, "dn12")])
def test_lzw_predictor2_files_are_read_by_libtiff(exe, tmp_path, h, w, spp, kind):
    """the reference's product options COMPRESS=LZW + PREDICTOR=2 (ref imageop.h:470-474, cv::imwrite's TIFF default): what
    we write, libtiff (cv2.imread) decodes to the same pixels, the tags say LZW / predictor 2, and compressible data shrinks"""
    rng = np.random.default_rng(h * 3 + w)
    if kind == "noise":
        px = rng.integers(0, 65536, (h, w, spp), dtype=np.uint16)
    elif kind == "smooth":
        px = (np.cumsum(rng.integers(-3, 4, (h, w, spp)), axis=1) + 2000).astype(np.uint16)
    elif kind == "zeros":
        px = np.zeros((h, w, spp), np.uint16)
    else:
        px = rng.integers(64, 4032, (h, w, spp), dtype=np.uint16)
    raw, tif = str(tmp_path / "in.raw"), str(tmp_path / "out.TIFF")
    px.tofile(raw)
    subprocess.check_call([exe, "write", tif, str(w), str(h), str(spp), raw, "lzw"])
    img = cv2.imread(tif, cv2.IMREAD_UNCHANGED)
    assert img is not None and img.dtype == np.uint16
    assert np.array_equal(img, px[:, :, 0] if spp == 1 else px[:, :, [2, 1, 0, 3]])
    back = str(tmp_path / "back.raw")
    dims = subprocess.check_output([exe, "read", tif, back], text=True).split()
    assert [int(v) for v in dims] == [w, h, spp]
    assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, spp), px)
    data = open(tif, "rb").read()
    assert data[:2] == b"II"
    if kind in ("smooth", "zeros"):
        assert len(data) < px.nbytes * 0.6
    # tags 259 (Compression) = 5 and 317 (Predictor) = 2, via our own parser's view
    import struct
    ifd = struct.unpack("<I", data[4:8])[0]
    n = struct.unpack("<H", data[ifd:ifd + 2])[0]
    tags = {}
    for k in range(n):
        t, ty, cnt, val = struct.unpack("<HHII", data[ifd + 2 + 12 * k: ifd + 14 + 12 * k])
        tags[t] = val & 0xFFFF if ty == 3 else val
    assert tags[259] == 5 and tags[317] == 2


def test_refuses_other_compressions(exe, tmp_path):
    img = np.random.default_rng(2).integers(0, 65536, (64, 64, 4), dtype=np.uint16)
    z, back = str(tmp_path / "deflate.TIFF"), str(tmp_path / "back.raw")
    assert cv2.imwrite(z, img, [cv2.IMWRITE_TIFF_COMPRESSION, 8])  # Adobe deflate
    r = subprocess.run([exe, "read", z, back], capture_output=True, text=True)
    assert r.returncode == 1 and "compressed TIFF" in r.stderr


def test_lzw_round_trips_with_libtiff_fuzz(exe, tmp_path):
    """200 random rasters (1..400 x 1..900, 1 or 4 samples; noise, constants, random walks, 16-pixel runs, four-level data --
    short strips, strings that fill and reset the LZW table): what we write libtiff reads, what libtiff writes we read"""
    rng = np.random.default_rng(11)
    raw, tif, back = str(tmp_path / "in.raw"), str(tmp_path / "o.TIFF"), str(tmp_path / "b.raw")
    for it in range(200):
        h, w, spp = int(rng.integers(1, 400)), int(rng.integers(1, 900)), int(rng.choice([1, 4]))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            px = rng.integers(0, 65536, (h, w, spp), dtype=np.uint16)
        elif kind == 1:
            px = np.full((h, w, spp), int(rng.integers(0, 65536)), np.uint16)
        elif kind == 2:
            px = (np.cumsum(rng.integers(-2, 3, (h, w, spp)), axis=1) + 3000).astype(np.uint16)
        elif kind == 3:
            px = np.repeat(rng.integers(0, 65536, (h, (w + 15) // 16, spp), dtype=np.uint16), 16, axis=1)[:, :w]
        else:
            px = (rng.integers(0, 4, (h, w, spp)) * 257).astype(np.uint16)
        px = np.ascontiguousarray(px)
        px.tofile(raw)
        subprocess.check_call([exe, "write", tif, str(w), str(h), str(spp), raw, "lzw"])
        want = px[:, :, 0] if spp == 1 else px[:, :, [2, 1, 0, 3]]
        img = cv2.imread(tif, cv2.IMREAD_UNCHANGED)
        assert img is not None and img.shape == want.shape and np.array_equal(img, want), (it, h, w, spp, kind)
        subprocess.check_output([exe, "read", tif, back])
        assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, spp), px), (it, h, w, spp, kind)
        assert cv2.imwrite(tif, want)                     # libtiff's own LZW + predictor file of the same pixels
        subprocess.check_output([exe, "read", tif, back])
        assert np.array_equal(np.fromfile(back, np.uint16).reshape(h, w, spp), px), (it, h, w, spp, kind)
