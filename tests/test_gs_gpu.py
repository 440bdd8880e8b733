"""GPU parity: product Gray-Scott path (C-ABI) vs the CPU oracle and vs the reference's own
step_kernel (oracle/_ref/libref_gs.so) on the same device.  Bit-exact is the bar: the arithmetic is
the reference's expression tree compiled with the reference's math flags."""
import numpy as np
import pytest

import oracle
from fluid_sims_b200.gray_scott import GrayScott, Params, init_pattern

pytestmark = pytest.mark.gpu


def run_product(u, v, steps, **kw):
    ny, nx = u.shape
    g = GrayScott(Params(nx=nx, ny=ny, **kw))
    g.upload(u, v)
    g.step(steps)
    out = g.download()
    g.close()
    return out


@pytest.mark.parametrize("nx,ny,steps", [(128, 128, 100), (96, 64, 200), (200, 120, 40),
                                         (4, 4, 9), (8, 3, 9), (132, 33, 17), (1024, 40, 12)])
def test_tma_path_bit_exact_vs_oracle(nx, ny, steps):
    u0, v0 = oracle.gs_init_pattern(nx, ny)
    rng = np.random.default_rng(nx * 1000 + ny)
    u0 = (u0 * rng.uniform(0.9, 1.0, u0.shape)).astype(np.float32)
    v0 = (v0 + rng.uniform(0.0, 0.05, v0.shape)).astype(np.float32)
    eu, ev = oracle.gs_run(u0, v0, steps)
    pu, pv = run_product(u0, v0, steps)
    assert np.array_equal(pu.view(np.uint32), eu.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), ev.view(np.uint32))


@pytest.mark.parametrize("nx,ny,steps", [(130, 67, 25), (1, 1, 3), (3, 5, 7), (237, 61, 30)])
def test_generic_path_bit_exact_vs_oracle(nx, ny, steps):
    rng = np.random.default_rng(nx + ny)
    u0 = rng.random((ny, nx), dtype=np.float32)
    v0 = (rng.random((ny, nx), dtype=np.float32) * 0.4).astype(np.float32)
    eu, ev = oracle.gs_run(u0, v0, steps)
    pu, pv = run_product(u0, v0, steps)
    assert np.array_equal(pu.view(np.uint32), eu.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), ev.view(np.uint32))


def test_non_default_coefficients_vs_oracle():
    rng = np.random.default_rng(3)
    u0 = rng.random((72, 256), dtype=np.float32)
    v0 = (rng.random((72, 256), dtype=np.float32) * 0.5).astype(np.float32)
    kw = dict(Du=0.16, Dv=0.08, dt=0.25, dx=0.5, feed=0.0367, kill=0.0649)
    eu, ev = oracle.gs_run(u0, v0, 33, **kw)
    pu, pv = run_product(u0, v0, 33, **kw)
    assert np.isfinite(eu).all()
    assert np.array_equal(pu.view(np.uint32), eu.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), ev.view(np.uint32))


def test_flush_to_zero_regime_matches():
    u0, v0 = oracle.gs_init_pattern(96, 64)
    v0 = (v0 * np.float32(1e-30)).astype(np.float32)
    eu, ev = oracle.gs_run(u0, v0, 400)
    pu, pv = run_product(u0, v0, 400)
    assert np.array_equal(pu.view(np.uint32), eu.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), ev.view(np.uint32))


@pytest.mark.skipif(not oracle.has_ref("ref_gs"), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,steps", [(512, 300), (2048, 60)])
def test_bit_exact_vs_reference_kernel(n, steps):
    u0, v0 = oracle.ref_gs_init_pattern(n, n, 1337)
    pu0, pv0 = init_pattern(n, n, 1337)
    assert np.array_equal(u0, pu0) and np.array_equal(v0, pv0)
    ru, rv = oracle.ref_gs_run(u0, v0, steps)
    g = GrayScott(Params(nx=n, ny=n)).init()
    g.step(steps)
    pu, pv = g.download()
    assert np.array_equal(pu.view(np.uint32), ru.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), rv.view(np.uint32))


@pytest.mark.skipif(not oracle.has_ref("ref_gs"), reason="oracle/_ref not built")
def test_full_size_8192_vs_reference_kernel():
    """BASELINE config 3 (8192x8192): 20 steps, every cell compared with the reference kernel."""
    n = 8192
    u0, v0 = init_pattern(n, n, 1337)
    ru, rv = oracle.ref_gs_run(u0, v0, 20)
    g = GrayScott(Params(nx=n, ny=n)).init()
    g.step(20)
    pu, pv = g.download()
    assert np.array_equal(pu.view(np.uint32), ru.view(np.uint32))
    assert np.array_equal(pv.view(np.uint32), rv.view(np.uint32))
    # size-independent property: mass of u+v only changes through feed/kill terms -> finite, bounded
    assert np.isfinite(pu).all() and 0.0 <= pv.min() and pu.max() <= 1.0 + 1e-6


def test_slab_handles_match_single_domain():
    """Two slab handles on one GPU with host-mediated ghost exchange == one full-domain handle."""
    nx, ny, steps = 256, 96, 20
    u0, v0 = oracle.gs_init_pattern(nx, ny)
    full = GrayScott(Params(nx=nx, ny=ny)).upload(u0, v0)
    full.step(steps)
    fu, fv = full.download()
    half = ny // 2
    import ctypes as C
    import torch
    a = GrayScott(Params(nx=nx, ny=ny), y_begin=0, ny_local=half).upload(u0[:half], v0[:half])
    b = GrayScott(Params(nx=nx, ny=ny), y_begin=half, ny_local=ny - half).upload(u0[half:], v0[half:])
    from fluid_sims_b200.slab import wrap_plane
    for _ in range(steps):
        ta = [wrap_plane(p, (half + 2, nx), torch.float32) for p in a.device_planes()]
        tb = [wrap_plane(p, (ny - half + 2, nx), torch.float32) for p in b.device_planes()]
        a.sync(); b.sync()
        for pa, pb in zip(ta, tb):
            pa[0].copy_(pb[-2])      # a's top ghost   <- b's last row (periodic)
            pa[-1].copy_(pb[1])      # a's bottom ghost <- b's first row
            pb[0].copy_(pa[-2])
            pb[-1].copy_(pa[1])
        torch.cuda.synchronize()
        a.step(1); b.step(1)
    au, av = a.download()
    bu, bv = b.download()
    assert np.array_equal(np.vstack([au, bu]), fu) and np.array_equal(np.vstack([av, bv]), fv)


def test_errors_are_loud():
    from fluid_sims_b200 import TauError
    with pytest.raises(TauError):
        GrayScott(Params(nx=0, ny=8))
    with pytest.raises(TauError):
        GrayScott(Params(nx=8, ny=8), y_begin=4, ny_local=8)
