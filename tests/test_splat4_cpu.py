"""`.4spl` export, the parts that need no GPU: the palette and the container against
the reference's own reader (viewer.html:67-96, restated in fluid_sims_b200.splat4.parse), and the device
quantiser's step table against the reference's host formula (th3cs.cu:1215-1218, oracle/splat4_oracle.c)."""
import numpy as np
import pytest

import oracle
from fluid_sims_b200 import TauError, splat4


def test_thermal_palette_equals_oracle_and_formula():
    pal = splat4.thermal_palette(256)
    assert np.array_equal(pal, oracle.splat4_palette(256))
    assert np.array_equal(pal[:, :8], np.tile(np.array([0, 1] * 4, np.float32), (256, 1))) and np.all(pal[:, 11] == 1)
    assert tuple(pal[0, 8:11]) == (0, 0, 0) and tuple(pal[255, 8:11]) == (1, 1, 1)      # black ... white
    assert tuple(pal[102, 8:11]) == (1.0, np.float32(np.float32(102 / 255) * np.float32(2.5)) - np.float32(0.5), 0.0)


def test_index_steps_reproduce_the_host_formula_exactly():
    """index(norm) = #{k : thr[k-1] <= norm} must equal (int)(powf(norm, 0.65f) * 255.0f) for every float in
    [0, 1].  The formula is monotone over all 1 065 353 217 such floats (oracle_4spl_index_monotone_violations,
    run once: 0 violations, 11 s), so agreement at every step and its neighbours is agreement everywhere."""
    thr = splat4.index_thresholds()
    assert thr.shape == (255,) and np.all(np.diff(thr) > 0) and thr[-1] == 1.0
    bits = thr.view(np.uint32)
    for k in range(1, 256):
        below = np.array([bits[k - 1] - 1], np.uint32).view(np.float32)[0]
        assert oracle.splat4_index(float(thr[k - 1])) == k and oracle.splat4_index(float(below)) == k - 1
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(20000, dtype=np.float32), np.float32([0, 1, 1e-30, 0.5])])
    want = np.array([oracle.splat4_index(float(v)) for v in x])
    assert np.array_equal(np.searchsorted(thr, x, side="right"), want)
    # a window of the exhaustive check, so that the property is exercised on every run
    L = oracle.lib
    import ctypes as C
    L.oracle_4spl_index_monotone_violations.argtypes = [C.c_uint32, C.c_uint32]
    L.oracle_4spl_index_monotone_violations.restype = C.c_long
    assert L.oracle_4spl_index_monotone_violations(0x3f000000, 0x3f000000 + 2_000_000) == 0


def test_frame_quantisation_oracle_basics():
    v = np.array([[3.0, 1.0], [2.0, 1.0]], np.float32)
    idx, mm = oracle.splat4_frame_indices(v)
    assert mm == (1.0, 3.0) and idx[0, 0] == 255 and idx[0, 1] == 0 and idx[1, 0] == int(0.5 ** 0.65 * 255)
    idx, mm = oracle.splat4_frame_indices(np.zeros(7, np.float32))      # flat field: range floor 1e-12
    assert not idx.any() and mm == (0.0, 0.0)


def test_container_is_what_the_reference_viewer_reads(tmp_path):
    rng = np.random.default_rng(1)
    idx = rng.integers(0, 256, size=(3, 4, 5, 6), dtype=np.uint8)       # frames, depth, height, width
    p = str(tmp_path / "v.4spl")
    splat4.write(p, idx)
    raw = open(p, "rb").read()
    assert len(raw) == 32 + 256 * 48 + idx.size + 16 and raw[:4] == b"4SPL" and raw[-4:] == b"LPS4"
    assert splat4.info(p) == dict(width=6, height=5, depth=4, frames=3, pSize=256, flags=4)
    import zlib
    assert zlib.crc32(raw[:-16]) == int.from_bytes(raw[-16:-12], "little")       # the footer's checksum is plain CRC-32
    assert int.from_bytes(raw[-12:-4], "little") == 32 + 256 * 48                # idxoffset
    d = splat4.parse(raw)                                                # viewer.html parse4Splat
    assert (d["width"], d["height"], d["depth"], d["frames"], d["voxelsPerFrame"]) == (6, 5, 4, 3, 120)
    assert np.array_equal(d["indices"], idx) and np.array_equal(d["palette"], splat4.thermal_palette()[:, 8:11])
    # the viewer addresses voxel (x, y, z) of frame f at f*voxelsPerFrame + (z*height + y)*width + x (th3cs.cu:1211)
    assert d["indices"].ravel()[2 * 120 + (3 * 5 + 4) * 6 + 5] == idx[2, 3, 4, 5]
    # integrity: a flipped byte or a truncated file is reported
    bad = bytearray(raw)
    bad[40000 % len(raw)] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(TauError, match="checksum"):
        splat4.info(p)
    open(p, "wb").write(raw[:-20])
    with pytest.raises(TauError, match="truncated"):
        splat4.info(p)
    with pytest.raises(ValueError):
        splat4.write(p, idx[0])
