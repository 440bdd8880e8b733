"""`.4spl` volume-video container of the reference's `th3cs` exporter (th3cs.cu:17-62, 1132-1240) over the
C-ABI, plus a reader that follows the reference's own (viewer.html:67-96).  What of the byte layout is
pinned by the reference and what is this project's choice is spelled out in csrc/splat4.cu."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, declare

_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_palette = declare("tau_4spl_thermal_palette", [_f32, C.c_int], None)
_write = declare("tau_4spl_write", [C.c_char_p] + [C.c_int] * 5 + [C.c_uint, _f32, _u8])
_info = declare("tau_4spl_info", [C.c_char_p, C.POINTER(C.c_int)])
_thr = declare("tau_4spl_index_thresholds", [_f32], None)

FLAGS_F32_IDX8 = 0x0004   # "Float32 Precision (0x04) and 8-bit Index Width (0x00)", th3cs.cu:1226


def thermal_palette(p_size: int = 256) -> np.ndarray:
    """(p_size, 12) float32 Splat4D entries: the exporter's black-red-yellow-white map (:1136-1144)."""
    pal = np.zeros((p_size, 12), np.float32)
    _palette(pal.ravel(), p_size)
    return pal


def index_thresholds() -> np.ndarray:
    """thr[k-1] = smallest norm in [0, 1] whose palette index (int)(powf(norm, 0.65f)*255) is >= k"""
    t = np.zeros(255, np.float32)
    _thr(t)
    return t


def write(path: str, indices: np.ndarray, palette: np.ndarray | None = None, flags: int = FLAGS_F32_IDX8) -> None:
    """indices: (frames, depth, height, width) uint8"""
    idx = np.ascontiguousarray(indices, np.uint8)
    if idx.ndim != 4:
        raise ValueError("indices must be (frames, depth, height, width)")
    pal = thermal_palette() if palette is None else np.ascontiguousarray(palette, np.float32).reshape(-1, 12)
    f, d, h, w = idx.shape
    check(_write(path.encode(), w, h, d, f, pal.shape[0], flags, pal.ravel(), idx.ravel()))


def info(path: str) -> dict:
    v = (C.c_int * 6)()
    check(_info(path.encode(), v))
    return dict(zip(("width", "height", "depth", "frames", "pSize", "flags"), [int(x) for x in v]))


def parse(buffer: bytes) -> dict:
    """parse4Splat (viewer.html:67-96), field for field: what the reference's player sees in a file."""
    dv = memoryview(buffer)
    u32 = lambda off: int.from_bytes(dv[off:off + 4], "little")  # noqa: E731
    width, height, depth, frames, p_size = u32(8), u32(12), u32(16), u32(20), u32(24)
    pal = np.frombuffer(buffer, "<f4", count=12 * p_size, offset=32).reshape(p_size, 12)
    colors = pal[:, 8:11].copy()                       # r, g, b at +32/+36/+40 of each 48-byte entry
    indices_offset = 32 + p_size * 48
    voxels = width * height * depth
    indices = np.frombuffer(buffer, np.uint8, count=voxels * frames, offset=indices_offset)
    return dict(width=width, height=height, depth=depth, frames=frames, palette=colors,
                indices=indices.reshape(frames, depth, height, width), voxelsPerFrame=voxels)
