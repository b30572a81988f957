"""ctypes mirror of include/oip_b200.h.  Loads the in-tree liboip_b200.so and fails loudly if it is
missing -- there is no Python/CPU fallback for any operation."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboip_b200.so")

OIP_OK = 0
OIP_E_INVALID, OIP_E_CUDA, OIP_E_NOMEM, OIP_E_RANGE, OIP_E_UNSUPPORTED, OIP_E_IO = -1, -2, -3, -4, -5, -6
FMT_LE16, FMT_BE16, FMT_PACK12, FMT_PACK10, FMT_BE16_TILES = 0, 1, 2, 3, 4
MAX_SEG = 4


class OipError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oip status {code}: {msg}")
        self.code = code


class RowSeg(C.Structure):
    _fields_ = [("base", C.c_void_p), ("row0", C.c_int64), ("n_rows", C.c_int64), ("pitch_bytes", C.c_int64)]


class CcdSrc(C.Structure):
    _fields_ = [("fmt", C.c_int), ("n_seg", C.c_int), ("seg", RowSeg * MAX_SEG), ("d_kb", C.c_void_p),
                ("shifted", C.c_int), ("dX", C.c_double), ("dY", C.c_double), ("d_tile_off", C.c_void_p),
                ("tile_cols", C.c_int), ("tile_lines", C.c_int), ("h_tile_off", C.c_void_p)]


class PanDesc(C.Structure):
    _fields_ = [("n_ccd", C.c_int), ("w", C.c_int), ("total_rows", C.c_int64), ("row0", C.c_int64),
                ("n_rows", C.c_int64), ("fold_half", C.c_int), ("section_rows", C.c_int), ("row_guard", C.c_int),
                ("ccd", CcdSrc * 8), ("d_out", C.c_void_p), ("out_pitch_px", C.c_int64)]


class FrameGeom(C.Structure):
    _fields_ = [("tile_cols", C.c_int), ("tile_lines", C.c_int)]


class FrameEntry(C.Structure):
    _fields_ = [("frame_off", C.c_int64), ("tile_off", C.c_int64 * 40), ("seq", C.c_int32), ("z_ratio", C.c_int32)]


class MssDesc(C.Structure):
    _fields_ = [("fmt", C.c_int), ("wb", C.c_int), ("lines", C.c_int64), ("pitch_px", C.c_int64),
                ("d_kb", C.c_void_p * 4), ("cX", C.c_double * 8), ("cY", C.c_double * 12),
                ("lines_per_section", C.c_int), ("line_offset", C.c_int64), ("overlap", C.c_int),
                ("keep_leading", C.c_int), ("min_process_lines", C.c_int), ("sec_first", C.c_int), ("sec_count", C.c_int),
                ("src_row0", C.c_int64)]


class DownlinkSrc(C.Structure):
    _fields_ = [("d_file", C.c_void_p), ("n_bytes", C.c_size_t), ("d_kb", C.c_void_p), ("shifted", C.c_int), ("dX", C.c_double),
                ("dY", C.c_double)]


class DownlinkDesc(C.Structure):
    _fields_ = [("n_ccd", C.c_int), ("geom", FrameGeom), ("fold_half", C.c_int), ("section_rows", C.c_int), ("row_guard", C.c_int),
                ("ccd", DownlinkSrc * 8), ("d_out", C.c_void_p), ("out_pitch_px", C.c_int64), ("out_rows_cap", C.c_int64),
                ("d_aux", C.c_void_p * 8), ("d_mss", C.c_void_p * 8)]


class DownlinkStats(C.Structure):
    _fields_ = [("aos", C.c_int64 * 3), ("imtr", C.c_int64 * 9), ("frames", C.c_int64 * 4), ("imdt_bytes", C.c_int64)]


class SttConfig(C.Structure):
    _fields_ = [("sections", C.c_int32), ("lines_per_section", C.c_int32), ("overlap_cols", C.c_int32),
                ("edge_cols", C.c_int32), ("threshold", C.c_double), ("max_delta_y", C.c_double)]


class SttSection(C.Structure):
    _fields_ = [("line_offset", C.c_int64), ("dx", C.c_double), ("dy", C.c_double), ("response", C.c_double),
                ("valid", C.c_int32), ("pad", C.c_int32)]


class IbcConfig(C.Structure):
    _fields_ = [("slices", C.c_int32), ("sections", C.c_int32), ("threshold", C.c_double), ("correlation_lines", C.c_int32),
                ("min_slices", C.c_int32), ("min_count", C.c_int32), ("pad", C.c_int32)]


class IbcShift(C.Structure):
    _fields_ = [("dx", C.c_double), ("dy", C.c_double), ("rs", C.c_double), ("cx", C.c_int32), ("pad", C.c_int32)]


# every symbol include/oip_b200.h declares: name -> (restype, argtypes)
_VP, _I, _I64, _SZ, _D = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double
SYMBOLS = {
    "oip_ctx_create": (_I, [_I, _VP, _I, C.POINTER(_VP)]),
    "oip_ctx_destroy": (None, [_VP]),
    "oip_ctx_sync": (_I, [_VP]),
    "oip_ctx_stream": (_VP, [_VP]),
    "oip_last_error": (C.c_char_p, []),
    "oip_abi_version": (_I, []),
    "oip_ctx_launch_count": (_I64, [_VP]),
    "oip_ctx_set_option": (_I, [_VP, C.c_char_p, _I64]),
    "oip_dev_alloc": (_I, [_VP, _SZ, C.POINTER(_VP)]),
    "oip_dev_free": (_I, [_VP, _VP]),
    "oip_host_alloc_pinned": (_I, [_SZ, C.POINTER(_VP)]),
    "oip_host_free_pinned": (_I, [_VP]),
    "oip_copy_h2d": (_I, [_VP, _VP, _VP, _SZ]),
    "oip_copy_d2h": (_I, [_VP, _VP, _VP, _SZ]),
    "oip_memset_d": (_I, [_VP, _VP, _I, _SZ]),
    "oip_ipc_export": (_I, [_VP, _VP, C.c_char_p]),
    "oip_ipc_open": (_I, [_VP, C.c_char_p, C.POINTER(_VP)]),
    "oip_ipc_close": (_I, [_VP, _VP]),
    "oip_crc16_batch": (_I, [_VP, _VP, _VP, _I64, _I, _VP]),
    "oip_aos_scan": (_I, [_VP, _VP, _SZ, _VP, _SZ, C.POINTER(_I64)]),
    "oip_aos_scan_shard": (_I, [_VP, _VP, _SZ, _SZ, _SZ, _VP, _SZ, C.POINTER(_I64), C.POINTER(_I64)]),
    "oip_imtr_deframe": (_I, [_VP, _VP, _VP, _I64, _VP, _SZ, C.POINTER(_I64), C.POINTER(_I64)]),
    "oip_imtr_deframe_shard": (_I, [_VP, _VP, _VP, _I64, _I, _I64, _I64, _VP, _SZ, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "oip_image_frames_index": (_I, [_VP, _VP, _SZ, C.POINTER(FrameGeom), C.POINTER(FrameEntry), _I64, C.POINTER(_I64)]),
    "oip_image_frames_hits": (_I, [_VP, _VP, _SZ, _VP, _VP, _I64, C.POINTER(_I64)]),
    "oip_image_frames_chain": (_I, [_VP, _VP, _I64, _SZ, C.POINTER(FrameGeom), C.POINTER(FrameEntry), _I64, C.POINTER(_I64)]),
    "oip_unpack_frames": (_I, [_VP, _VP, _SZ, C.POINTER(FrameGeom), C.POINTER(FrameEntry), _I64, _VP, _VP, _VP]),
    "oip_rrc_u16": (_I, [_VP, _VP, _I, _I64, _I64, _VP]),
    "oip_load_rrc_csv": (_I, [C.c_char_p, _I, _VP]),
    "oip_pan_pipeline": (_I, [_VP, C.POINTER(PanDesc)]),
    "oip_pan_plan_coverage": (_I, [C.POINTER(PanDesc), _I, _I, _VP, C.POINTER(_I64)]),
    "oip_mss_plan_coverage": (_I, [C.POINTER(MssDesc), _I, _I, _VP, C.POINTER(_I64)]),
    "oip_pan_out_width": (_I, [_I, _I, _I]),
    "oip_pan_check_error": (_I, [_VP]),
    "oip_cubic_tab": (None, [_VP]),
    "oip_pan_rows_needed": (_I, [C.POINTER(PanDesc), _I, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "oip_pan_row_ranges": (_I, [C.POINTER(PanDesc), _I, C.POINTER(_I64), _I, C.POINTER(_I)]),
    "oip_shift_cubic_u16": (_I, [_VP, _VP, _VP, _I, _I64, _D, _D, _I, _I]),
    "oip_stitch_concat_u16": (_I, [_VP, C.POINTER(_VP), _I, _I, _I64, _I, _VP]),
    "oip_band_align_merge": (_I, [_VP, _VP, C.POINTER(MssDesc), _VP, C.POINTER(_I64)]),
    "oip_stitch_concat_c4": (_I, [_VP, C.POINTER(_VP), _I, _I, _I64, _I, C.POINTER(_I), _VP]),
    "oip_unpack_lines": (_I, [_VP, _VP, _I, _I, _I64, _I64, _VP]),
    "oip_pan_pipeline_host": (_I, [_VP, C.POINTER(PanDesc)]),
    "oip_downlink_to_stitched": (_I, [_VP, C.POINTER(DownlinkDesc), C.POINTER(_I64), C.POINTER(DownlinkStats)]),
    "oip_synth_strip_dn": (_I, [_VP, _VP, _I, _I64, _I64, _I64, C.c_uint64, _I]),
    "oip_phase_correlate_u16": (_I, [_VP, _VP, _I64, _VP, _I64, _I, _I, C.POINTER(_D)]),
    "oip_inter_band_correlation": (_I, [_VP, _VP, _I, _I64, _I64, _VP, _I64, _I64, C.POINTER(IbcConfig), C.POINTER(IbcShift),
                                        C.POINTER(_D), C.POINTER(_D)]),
    "oip_stt_parameters": (_I, [_VP, _VP, _VP, _I, _I64, _I64, _I64, _I64, C.POINTER(SttConfig), C.POINTER(SttSection),
                                C.POINTER(_D)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m opticalimageprocessor_b200.build` "
                "(there is no CPU fallback)")
        L = C.CDLL(os.environ.get("OIP_B200_LIB", LIB_PATH))  # override: kernel experiments (tools/probes/build_variant.py)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return load().oip_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OIP_OK:
        raise OipError(rc, last_error())
