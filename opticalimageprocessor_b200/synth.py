"""Synthetic downlink data: the INVERSE of the reference's three de-framing levels
(ref aux_separator.h:29-138), plus strip / coefficient generators for the
benchmarks.  The reference ships no sample data, so every test and bench input comes from here.

Pure numpy (host side); nothing in here is on the timed path.
"""
from __future__ import annotations

import numpy as np

AOS_SYNC = bytes([0x1A, 0xCF, 0xFC, 0x1D])
IMTR_SIG = bytes([0x49, 0x54, 0xCE, 0x1F])
IMTR_END = bytes([0x2E, 0xE9, 0xC8, 0xFD])
IMG_SIG = bytes([0xEB, 0x90, 0xE1, 0x4D])

_CRC_TAB = None


def _crc_tab() -> np.ndarray:
    global _CRC_TAB
    if _CRC_TAB is None:
        t = np.zeros(256, np.uint16)
        for i in range(256):
            r = i << 8
            for _ in range(8):
                r = ((r << 1) ^ 0x1021) & 0xFFFF if r & 0x8000 else (r << 1) & 0xFFFF
            t[i] = r
        _CRC_TAB = t
    return _CRC_TAB


def crc16_rows(rows: np.ndarray) -> np.ndarray:
    """CRC-16/CCITT-FALSE of every row of a 2-D uint8 array (vectorised over rows)."""
    tab = _crc_tab()
    rows = np.ascontiguousarray(rows, np.uint8)
    crc = np.full(rows.shape[0], 0xFFFF, np.uint16)
    for j in range(rows.shape[1]):
        idx = ((crc >> 8) ^ rows[:, j]).astype(np.uint8)
        crc = ((crc << 8) & 0xFFFF) ^ tab[idx]
    return crc


# ------------------------------------------------------------------------------------ image frames
def make_image_frame(seq: int, aux: np.ndarray, tiles_be: np.ndarray, tile_cols: int, tile_lines: int,
                     z_ratio: int = 0, sub_dwords=None, image_dwords=None) -> np.ndarray:
    """aux: 192*tile_lines bytes; tiles_be: (40, tile_lines*tile_cols*2) bytes (big-endian samples)."""
    tile_bytes = tile_lines * tile_cols * 2
    assert aux.size == 192 * tile_lines and tiles_be.shape == (40, tile_bytes)
    assert tile_bytes % 4 == 0
    tr = np.zeros(172, np.uint8)
    tr[0:4] = np.frombuffer(IMG_SIG, np.uint8)
    tr[4] = z_ratio & 0x3F
    tr[5] = 0
    tr[6] = (seq >> 8) & 0xFF
    tr[7] = seq & 0xFF
    sd = np.full(40, tile_bytes // 4, np.uint32) if sub_dwords is None else np.asarray(sub_dwords, np.uint32)
    idw = int(sd.sum()) if image_dwords is None else image_dwords
    tr[8:12] = np.frombuffer(np.array([idw], ">u4").tobytes(), np.uint8)
    tr[12:172] = np.frombuffer(sd.astype(">u4").tobytes(), np.uint8)
    return np.concatenate([aux.astype(np.uint8), tiles_be.reshape(-1), tr])


def pan_mss_to_tiles(pan: np.ndarray, mss: np.ndarray, tile_cols: int, tile_lines: int) -> np.ndarray:
    """pan: (4*tile_lines, 8*tile_cols) u16, mss: (tile_lines, 8*tile_cols) u16 -> (40, tile bytes) BE"""
    tiles = []
    for r in range(5):
        src = pan[r * tile_lines:(r + 1) * tile_lines] if r < 4 else mss
        for c in range(8):
            t = src[:, c * tile_cols:(c + 1) * tile_cols]
            tiles.append(np.frombuffer(np.ascontiguousarray(t).astype(">u2").tobytes(), np.uint8))
    return np.stack(tiles)


def make_imdt(n_frames: int, tile_cols: int, tile_lines: int, seed: int = 0, skip_seqs=(), junk_prefix: int = 0,
              max_dn: int = 4096):
    """returns (imdt bytes, dict with the expected aux/pan/mss per sequence number)"""
    rng = np.random.default_rng(seed)
    W = 8 * tile_cols
    parts = []
    truth = {}
    if junk_prefix:
        parts.append(rng.integers(0, 0xE0, junk_prefix, dtype=np.uint8))  # no 0xEB.. signature bytes
    for s in range(1, n_frames + 1):
        if s in skip_seqs:
            continue
        aux = rng.integers(0, 256, 192 * tile_lines, dtype=np.uint8)
        aux[aux == 0xEB] = 0
        pan = rng.integers(0, max_dn, (4 * tile_lines, W), dtype=np.uint16)
        mss = rng.integers(0, max_dn, (tile_lines, W), dtype=np.uint16)
        truth[s] = (aux, pan, mss)
        parts.append(make_image_frame(s, aux, pan_mss_to_tiles(pan, mss, tile_cols, tile_lines), tile_cols, tile_lines))
    return np.concatenate(parts), truth


# ------------------------------------------------------------------------------------ IMTR
def imtr_frames(imdt: np.ndarray, chid: int = 0x11, seq_start: int = 1, dtmark: int = 0x22) -> np.ndarray:
    """cut the IMDT stream into 866-byte bodies (zero padded) and wrap them: (n, 882) uint8"""
    n = (imdt.size + 865) // 866
    body = np.zeros((n, 866), np.uint8)
    body.reshape(-1)[:imdt.size] = imdt
    f = np.zeros((n, 882), np.uint8)
    f[:, 0:4] = np.frombuffer(IMTR_SIG, np.uint8)
    seq = (np.arange(n, dtype=np.uint64) + seq_start).astype(">u4")
    f[:, 4:8] = np.frombuffer(seq.tobytes(), np.uint8).reshape(n, 4)
    f[:, 8] = chid
    f[:, 9] = dtmark
    f[:, 10:876] = body
    crc = crc16_rows(f[:, :876])
    f[:, 876] = crc >> 8
    f[:, 877] = crc & 0xFF
    f[:, 878:882] = np.frombuffer(IMTR_END, np.uint8)
    return f


def refresh_imtr_crc(f: np.ndarray) -> None:
    crc = crc16_rows(f[:, :876])
    f[:, 876] = crc >> 8
    f[:, 877] = crc & 0xFF


# ------------------------------------------------------------------------------------ AOS
def aos_frames(stream: np.ndarray, vcid: int = 0x01, seq_start: int = 0, ldpc_seed: int = 1) -> np.ndarray:
    """wrap a byte stream into 1024-byte AOS frames (payload 880, zero padded): (n, 1024) uint8"""
    n = (stream.size + 879) // 880
    pay = np.zeros((n, 880), np.uint8)
    pay.reshape(-1)[:stream.size] = stream.reshape(-1)
    f = np.zeros((n, 1024), np.uint8)
    f[:, 0:4] = np.frombuffer(AOS_SYNC, np.uint8)
    f[:, 4] = 0x40
    f[:, 5] = vcid & 0x3F
    seq = (np.arange(n) + seq_start) & 0xFFFFFF
    f[:, 6] = seq >> 16
    f[:, 7] = (seq >> 8) & 0xFF
    f[:, 8] = seq & 0xFF
    f[:, 9] = 0
    f[:, 10:14] = 0  # inject word: valid
    f[:, 14:894] = pay
    crc = crc16_rows(f[:, 4:894])
    f[:, 894] = crc >> 8
    f[:, 895] = crc & 0xFF
    rng = np.random.default_rng(ldpc_seed)
    ld = rng.integers(0, 256, (n, 128), dtype=np.uint8)
    ld[ld == 0x1A] = 0x1B  # keep the parity field free of sync bytes
    f[:, 896:1024] = ld
    return f


def aos_empty_frame() -> np.ndarray:
    f = np.zeros(1024, np.uint8)
    f[0:4] = np.frombuffer(AOS_SYNC, np.uint8)
    f[5] = 0x3F
    f[10:14] = 0xAA
    f[14:894:2] = 0x55
    f[15:894:2] = 0xAA
    return f


def build_aos_file(frames: np.ndarray, empty_every: int = 0, bad_crc_at=(), bad_inject_at=(), prefix: bytes = b"",
                   suffix: bytes = b"") -> np.ndarray:
    """serialise frames, inserting an empty frame before every `empty_every`-th frame and corrupting
    the listed frame indices (corrupted frames are dropped by the reference, so callers normally
    corrupt DUPLICATES to keep the 882-byte cadence intact)."""
    out = [np.frombuffer(prefix, np.uint8)] if prefix else []
    emp = aos_empty_frame()
    for i in range(frames.shape[0]):
        if empty_every and i % empty_every == 0:
            out.append(emp)
        fr = frames[i]
        if i in bad_crc_at:
            bad = fr.copy()
            bad[500] ^= 0x5A
            out.append(bad)  # corrupted duplicate first, then the good frame
        if i in bad_inject_at:
            bad = fr.copy()
            bad[10:14] = [0x12, 0x34, 0x56, 0x78]
            out.append(bad)
        out.append(fr)
    if suffix:
        out.append(np.frombuffer(suffix, np.uint8))
    return np.concatenate(out)


# ------------------------------------------------------------------------------------ strips
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def strip_dn(w: int, rows: int, seed: int, row0: int = 0) -> np.ndarray:
    """DN(x,y) = 64 + ((A(x) + B(y) + N(x,y)) mod 3968), 12-bit range [64, 4031]"""
    with np.errstate(over="ignore"):
        x = np.arange(w, dtype=np.uint64)[None, :]
        y = (np.arange(rows, dtype=np.uint64) + np.uint64(row0))[:, None]
        A = (x * np.uint64(37)) % np.uint64(1500)
        B = (y // np.uint64(8)) % np.uint64(1200)
        N = splitmix64(np.uint64(seed) ^ (y * np.uint64(w) + x)) & np.uint64(0xFF)
        return (np.uint64(64) + (A + B + N) % np.uint64(3968)).astype(np.uint16)


def rrc_coeffs(w: int, seed: int) -> np.ndarray:
    """k = 0.95 + 0.1 u1, b = 8 u2, rounded through the CSV text form so CPU and GPU hold identical doubles"""
    with np.errstate(over="ignore"):
        i = np.arange(w, dtype=np.uint64)
        u1 = splitmix64(np.uint64(seed) ^ (i * np.uint64(2))).astype(np.float64) / 2.0 ** 64
        u2 = splitmix64(np.uint64(seed) ^ (i * np.uint64(2) + np.uint64(1))).astype(np.float64) / 2.0 ** 64
    kb = np.empty((w, 2), np.float64)
    kb[:, 0] = [float("%.12f" % v) for v in (0.95 + 0.1 * u1)]
    kb[:, 1] = [float("%.12f" % v) for v in (8.0 * u2)]
    return kb


def write_rrc_csv(path: str, kb: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write("1\n%d\n0\n" % kb.shape[0])
        for k, b in kb:
            f.write("%.12f , %.12f\n" % (k, b))


def pack_bits(img: np.ndarray, bits: int) -> np.ndarray:
    """MSB-first big-endian bitstream per line (extension formats); returns (rows, ceil(w*bits/8)) u8"""
    rows, w = img.shape
    b = ((img[:, :, None].astype(np.uint32) >> np.arange(bits - 1, -1, -1, dtype=np.uint32)) & 1).astype(np.uint8)
    b = b.reshape(rows, w * bits)
    pad = (-b.shape[1]) % 8
    if pad:
        b = np.concatenate([b, np.zeros((rows, pad), np.uint8)], axis=1)
    return np.packbits(b, axis=1)
