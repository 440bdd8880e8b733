"""oracle — the CHECKER.  Test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; the product (fluid_sims_b200/) never does.

  * `oracle.lib`      — liboracle.so: plain-C restatements of the reference algorithms
                        (oracle/*_oracle.c, each function citing the reference file:line).
  * `oracle.ref(name)`— oracle/_ref/lib<name>.so: the reference's OWN sources compiled from
                        /root/reference by oracle/Makefile (prebuilt files travel to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and oracle/_ref (when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "-j8", "liboracle.so"] + (["ref"] if ref else []),
                   check=True, stdout=subprocess.DEVNULL)


def _load_lib() -> C.CDLL:
    if not os.path.exists(_LIB):
        build(ref=False)
    return C.CDLL(_LIB)


lib = _load_lib()

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def has_ref(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"lib{name}.so"))


_ref_cache: dict = {}


_host_refs = False


class host_refs:
    """with oracle.host_refs(): every ref_*() wrapper drives the reference's kernels on the CPU emulator
    (oracle/_ref/lib<name>_host.so, built by the HOST_REF_RULE of oracle/Makefile) instead of on a GPU."""

    def __enter__(self):
        global _host_refs
        self._old, _host_refs = _host_refs, True

    def __exit__(self, *a):
        global _host_refs
        _host_refs = self._old


def has_host_ref(name: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"lib{name}_host.so"))


def ref(name: str) -> C.CDLL:
    """Load oracle/_ref/lib<name>.so (the compiled reference)."""
    if _host_refs and not name.endswith("_host") and has_host_ref(name):
        name += "_host"
    if name not in _ref_cache:
        path = os.path.join(REF_DIR, f"lib{name}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run `make -C oracle ref` where "
                                    "/root/reference exists)")
        _ref_cache[name] = C.CDLL(path)
    return _ref_cache[name]


# ------------------------------------------------------------------------------------------------
# Gray-Scott
# ------------------------------------------------------------------------------------------------
lib.oracle_gs_run.argtypes = [f32p, f32p, C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int]
lib.oracle_gs_run.restype = C.c_int
lib.oracle_gs_init_pattern.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_uint]
lib.oracle_gs_init_pattern.restype = None


def gs_init_pattern(nx, ny, seed=1337):
    u = np.empty((ny, nx), np.float32)
    v = np.empty((ny, nx), np.float32)
    lib.oracle_gs_init_pattern(u, v, nx, ny, seed)
    return u, v


def gs_run(u, v, steps, Du=0.2, Dv=0.1, dt=1.0, dx=1.0, feed=0.03, kill=0.06):
    """CPU oracle: `steps` Gray-Scott steps; returns new (u, v)."""
    u = np.array(u, np.float32, order="C", copy=True)
    v = np.array(v, np.float32, order="C", copy=True)
    ny, nx = u.shape
    rc = lib.oracle_gs_run(u, v, nx, ny, Du, Dv, dt, dx, feed, kill, steps)
    assert rc == 0
    return u, v


def ref_gs_run(u, v, steps, Du=0.2, Dv=0.1, dt=1.0, dx=1.0, feed=0.03, kill=0.06):
    """The reference's own step_kernel on the GPU (oracle/_ref/libref_gs.so)."""
    r = ref("ref_gs")
    r.ref_gs_run.argtypes = [f32p, f32p, C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_int]
    r.ref_gs_run.restype = C.c_int
    u = np.array(u, np.float32, order="C", copy=True)
    v = np.array(v, np.float32, order="C", copy=True)
    ny, nx = u.shape
    rc = r.ref_gs_run(u, v, nx, ny, Du, Dv, dt, dx, feed, kill, steps)
    if rc != 0:
        raise RuntimeError(f"reference Gray-Scott run failed with cudaError {rc}")
    return u, v


def ref_gs_init_pattern(nx, ny, seed=1337):
    r = ref("ref_gs")
    r.ref_gs_init_pattern.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_uint]
    r.ref_gs_init_pattern.restype = None
    u = np.empty((ny, nx), np.float32)
    v = np.empty((ny, nx), np.float32)
    r.ref_gs_init_pattern(u, v, nx, ny, seed)
    return u, v


# ------------------------------------------------------------------------------------------------
# 2-D hypersonic (tau_hypersonic_cuda.cu), fp64
# ------------------------------------------------------------------------------------------------
class Hyp2dCfg(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gamma", "cfl", "visc_nu", "visc_rho", "visc_e",
                                          "inflow_mach", "geom_x0", "geom_cy", "geom_Rb",
                                          "geom_Rn", "geom_theta")] + [("W", C.c_int), ("H", C.c_int)]

    def as11(self):
        return np.array([self.gamma, self.cfl, self.visc_nu, self.visc_rho, self.visc_e,
                         self.inflow_mach, self.geom_x0, self.geom_cy, self.geom_Rb, self.geom_Rn,
                         self.geom_theta], np.float64)


_cfgp = C.POINTER(Hyp2dCfg)
lib.oracle_hyp2d_default_cfg.argtypes = [_cfgp, C.c_int, C.c_int]
lib.oracle_hyp2d_default_cfg.restype = None
lib.oracle_hyp2d_init.argtypes = [_cfgp, f64p, f64p, f64p, f64p, u8p]
lib.oracle_hyp2d_init.restype = None
lib.oracle_hyp2d_step.argtypes = [_cfgp, f64p, f64p, f64p, f64p, u8p]
lib.oracle_hyp2d_step.restype = C.c_double
lib.oracle_hyp2d_run.argtypes = [_cfgp, f64p, f64p, f64p, f64p, u8p, C.c_int, C.c_void_p]
lib.oracle_hyp2d_run.restype = C.c_double
lib.oracle_hyp2d_max_wavespeed.argtypes = [_cfgp, f64p, f64p, f64p, f64p, u8p]
lib.oracle_hyp2d_max_wavespeed.restype = C.c_double
lib.oracle_hyp2d_dt.argtypes = [_cfgp, C.c_double]
lib.oracle_hyp2d_dt.restype = C.c_double
lib.oracle_hyp2d_snapshot.argtypes = [_cfgp, C.c_int, f64p, f64p, f64p, f64p, u8p, f64p]
lib.oracle_hyp2d_snapshot.restype = None
lib.oracle_hyp2d_sdf.argtypes = [C.c_double] * 5
lib.oracle_hyp2d_sdf.restype = C.c_double


def hyp2d_cfg(W, H, **over) -> Hyp2dCfg:
    c = Hyp2dCfg()
    lib.oracle_hyp2d_default_cfg(C.byref(c), W, H)
    for k, v in over.items():
        setattr(c, k, v)
    return c


def hyp2d_init(cfg: Hyp2dCfg):
    N = cfg.W * cfg.H
    planes = [np.empty(N, np.float64) for _ in range(4)]
    mask = np.empty(N, np.uint8)
    lib.oracle_hyp2d_init(C.byref(cfg), *planes, mask)
    return planes, mask


def hyp2d_run(cfg: Hyp2dCfg, planes, mask, steps):
    """CPU oracle: returns (new planes, sim_t, dts)."""
    planes = [np.array(p, np.float64, order="C", copy=True).ravel() for p in planes]
    mask = np.ascontiguousarray(mask, np.uint8).ravel()
    dts = np.zeros(max(steps, 1), np.float64)
    t = lib.oracle_hyp2d_run(C.byref(cfg), *planes, mask, steps, dts.ctypes.data_as(C.c_void_p))
    return planes, float(t), dts[:steps]


def hyp2d_snapshot(cfg: Hyp2dCfg, steps, planes, mask):
    out = np.zeros(12, np.float64)
    lib.oracle_hyp2d_snapshot(C.byref(cfg), steps, *[np.ascontiguousarray(p, np.float64).ravel()
                                                      for p in planes],
                              np.ascontiguousarray(mask, np.uint8).ravel(), out)
    return out


def ref_hyp2d_run(W, H, cfg11, steps, planes=None, mask=None, tile=(32, 8), host=False, f32=False):
    """The reference's own kernels on the GPU (oracle/_ref/libref_hyp2d_<W>x<H>.so) or, host=True, the same
    driver and kernels executed by the CPU emulator of tests/hostemu (libref_hyp2d_host_<W>x<H>.so), or, f32=True,
    the float-typed scratch copy of the reference (libref_hyp2d_f32_<W>x<H>.so, oracle/gen_f32_src.py): the
    reference's algorithm evaluated in fp32 (planes still cross this interface as float64).
    planes=None -> start from k_init.  Returns (planes, mask, sim_t, dts, ms)."""
    r = ref(f"ref_hyp2d_f32_{W}x{H}" if f32 else (f"ref_hyp2d_host_{W}x{H}" if host else f"ref_hyp2d_{W}x{H}"))
    r.ref_hyp2d_run.argtypes = [f64p, C.c_int, C.c_int, C.c_int, C.c_int, f64p, f64p, f64p, f64p,
                                u8p, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_float)]
    r.ref_hyp2d_run.restype = C.c_int
    N = W * H
    do_init = planes is None
    if do_init:
        planes = [np.zeros(N, np.float64) for _ in range(4)]
        mask = np.zeros(N, np.uint8)
    else:
        planes = [np.array(p, np.float64, order="C", copy=True).ravel() for p in planes]
        mask = np.array(mask, np.uint8, order="C", copy=True).ravel()
    t = C.c_double()
    ms = C.c_float()
    dts = np.zeros(max(steps, 1), np.float64)
    rc = r.ref_hyp2d_run(np.ascontiguousarray(cfg11, np.float64), steps, tile[0], tile[1],
                         1 if do_init else 0, *planes, mask, C.byref(t),
                         dts.ctypes.data_as(C.c_void_p), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference hyp2d run failed with cudaError {rc}")
    return planes, mask, float(t.value), dts[:steps], float(ms.value)


def hyp2d_render(cfg: Hyp2dCfg, planes, mask, view_mode):
    """CPU oracle of the render pass: (rgba (H, W, 4) uint8, vals (H, W) f64, (min, max))."""
    lib.oracle_hyp2d_render.argtypes = [C.POINTER(Hyp2dCfg), f64p, f64p, f64p, f64p, u8p, C.c_int, u8p,
                                        f64p, f64p]
    lib.oracle_hyp2d_render.restype = None
    planes = [np.ascontiguousarray(p, np.float64).ravel() for p in planes]
    mask = np.ascontiguousarray(mask, np.uint8).ravel()
    rgba = np.zeros(cfg.W * cfg.H * 4, np.uint8)
    vals = np.zeros(cfg.W * cfg.H, np.float64)
    mm = np.zeros(2, np.float64)
    lib.oracle_hyp2d_render(C.byref(cfg), *planes, mask, view_mode, rgba, vals, mm)
    return rgba.reshape(cfg.H, cfg.W, 4), vals.reshape(cfg.H, cfg.W), (float(mm[0]), float(mm[1]))


def ref_hyp2d_render(W, H, cfg11, planes, mask, view_mode):
    """The reference's own render kernels (k_render_vals .. k_render_pixels) on the GPU."""
    r = ref(f"ref_hyp2d_{W}x{H}")
    r.ref_hyp2d_render.argtypes = [f64p, C.c_int, f64p, f64p, f64p, f64p, u8p, u8p, f64p, f64p]
    r.ref_hyp2d_render.restype = C.c_int
    planes = [np.ascontiguousarray(p, np.float64).ravel() for p in planes]
    mask = np.ascontiguousarray(mask, np.uint8).ravel()
    rgba = np.zeros(W * H * 4, np.uint8)
    vals = np.zeros(W * H, np.float64)
    mm = np.zeros(2, np.float64)
    rc = r.ref_hyp2d_render(np.ascontiguousarray(cfg11, np.float64), view_mode, *planes, mask, rgba, vals, mm)
    if rc != 0:
        raise RuntimeError(f"reference hyp2d render failed with cudaError {rc}")
    return rgba.reshape(H, W, 4), vals.reshape(H, W), (float(mm[0]), float(mm[1]))


def ref_hyp2d_eval(W, H, cfg11, kind, vecs):
    r = ref(f"ref_hyp2d_{W}x{H}")
    r.ref_hyp2d_eval.argtypes = [f64p, C.c_int, C.c_int, f64p, f64p]
    r.ref_hyp2d_eval.restype = C.c_int
    vecs = np.ascontiguousarray(vecs, np.float64)
    n = vecs.shape[0]
    out = np.zeros((n, 8), np.float64)
    rc = r.ref_hyp2d_eval(np.ascontiguousarray(cfg11, np.float64), kind, n, vecs, out)
    if rc != 0:
        raise RuntimeError(f"reference hyp2d eval failed with cudaError {rc}")
    return out


# ------------------------------------------------------------------------------------------------
# CPU hypersonic reference (tau_hypersonic.c / tau_hypersonic_simd.c) — timing baseline
# ------------------------------------------------------------------------------------------------
class RefHypCpu:
    """The reference CPU solver compiled at a fixed WxH (file-static state: one instance per
    process per library)."""

    def __init__(self, W=256, H=256, simd=False):
        self.W, self.H = W, H
        self.lib = ref(f"ref_{'hypsimd' if simd else 'hypcpu'}_{W}x{H}")
        self.lib.ref_hypcpu_steps.argtypes = [C.c_int]
        self.lib.ref_hypcpu_steps.restype = C.c_double
        self.lib.ref_hypcpu_time.restype = C.c_double
        self.lib.ref_hypcpu_get.argtypes = [f64p, f64p, f64p, f64p, u8p]
        self.lib.ref_hypcpu_set.argtypes = [f64p, f64p, f64p, f64p, u8p]

    def init(self):
        self.lib.ref_hypcpu_init()

    def steps(self, n) -> float:
        return float(self.lib.ref_hypcpu_steps(n))

    @property
    def sim_t(self) -> float:
        return float(self.lib.ref_hypcpu_time())

    def get(self):
        N = self.W * self.H
        planes = [np.empty(N, np.float64) for _ in range(4)]
        mask = np.empty(N, np.uint8)
        self.lib.ref_hypcpu_get(*planes, mask)
        return planes, mask


# plain-C restatement of tau_hypersonic.c (oracle/hypcpu_oracle.c), run-time grid extents
lib.hypcpu_init.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, f64p, u8p]
lib.hypcpu_init.restype = None
lib.hypcpu_step.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, f64p, u8p, C.c_int, C.POINTER(C.c_double), f64p]
lib.hypcpu_step.restype = None
lib.hypcpu_render.argtypes = [C.c_int, C.c_int, f64p, f64p, f64p, f64p, u8p, C.c_int, u8p, f64p, f64p]
lib.hypcpu_render.restype = None


def hypcpu_init(W, H):
    """init_sim (tau_hypersonic.c:450): 4 planes (rho, mx, my, E) of W*H doubles + mask."""
    planes = [np.empty(W * H, np.float64) for _ in range(4)]
    mask = np.empty(W * H, np.uint8)
    lib.hypcpu_init(W, H, *planes, mask)
    return planes, mask


def hypcpu_run(W, H, planes, mask, steps, sim_t=0.0):
    """`steps` x step_physics (tau_hypersonic.c:500-674) -> (planes, sim_t, dts)."""
    planes = [np.ascontiguousarray(p, np.float64).ravel().copy() for p in planes]
    t = C.c_double(sim_t)
    dts = np.zeros(max(steps, 1), np.float64)
    lib.hypcpu_step(W, H, *planes, np.ascontiguousarray(mask, np.uint8).ravel(), steps, C.byref(t), dts)
    return planes, float(t.value), dts[:steps]


def hypcpu_render(W, H, planes, mask, view_mode=2):
    """main()'s render loop (tau_hypersonic.c:713-786): (rgba[H, W, 4], (min, max), values[H, W])."""
    planes = [np.ascontiguousarray(p, np.float64).ravel() for p in planes]
    rgba = np.zeros(W * H * 4, np.uint8)
    mm = np.zeros(2, np.float64)
    vals = np.zeros(W * H, np.float64)
    lib.hypcpu_render(W, H, *planes, np.ascontiguousarray(mask, np.uint8).ravel(), view_mode, rgba, mm, vals)
    return rgba.reshape(H, W, 4), (float(mm[0]), float(mm[1])), vals.reshape(H, W)


# ------------------------------------------------------------------------------------------------
# 3-D hypersonic (tau_hypersonic_3d_cuda.cu), fp32
# ------------------------------------------------------------------------------------------------
_H3_FIELDS = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)] + \
             [(n, C.c_float) for n in ("dx", "dy", "dz", "cfl", "u_ref", "R", "gamma_floor", "Twall",
                                       "tau_vib", "theta_v", "sdf_cx", "sdf_cy", "sdf_cz", "sdf_r",
                                       "inflow_r", "inflow_p", "inflow_u", "inflow_v", "inflow_w")] + \
             [("sponge_n", C.c_int), ("sponge_strength", C.c_float), ("sponge_out_n", C.c_int),
              ("sponge_out_strength", C.c_float)]


class Hyp3dParams(C.Structure):
    _fields_ = _H3_FIELDS

    def as23(self):
        """the 23 floats oracle/ref_drivers/ref_hyp3d.cu expects (everything but nx, ny, nz)"""
        return np.array([float(getattr(self, f[0])) for f in _H3_FIELDS[3:]], np.float32)


_h3p = C.POINTER(Hyp3dParams)
_pp6 = C.POINTER(C.c_void_p)
lib.oracle_hyp3d_default_params.argtypes = [_h3p, C.c_int, C.c_int, C.c_int]
lib.oracle_hyp3d_default_params.restype = None
lib.oracle_hyp3d_build_solid.argtypes = [_h3p, u8p]
lib.oracle_hyp3d_build_solid.restype = None
lib.oracle_hyp3d_init.argtypes = [_h3p] + [f32p] * 6 + [u8p]
lib.oracle_hyp3d_init.restype = None
lib.oracle_hyp3d_run.argtypes = [_h3p, _pp6, u8p, C.c_int, f32p, C.c_void_p, C.c_void_p]
lib.oracle_hyp3d_run.restype = None


def hyp3d_params(nx, ny, nz, **over) -> Hyp3dParams:
    p = Hyp3dParams()
    lib.oracle_hyp3d_default_params(C.byref(p), nx, ny, nz)
    for k, v in over.items():
        setattr(p, k, v)
    return p


def hyp3d_init(p: Hyp3dParams):
    N = p.nx * p.ny * p.nz
    solid = np.empty(N, np.uint8)
    lib.oracle_hyp3d_build_solid(C.byref(p), solid)
    planes = [np.empty(N, np.float32) for _ in range(6)]
    lib.oracle_hyp3d_init(C.byref(p), *planes, solid)
    return planes, solid


def hyp3d_run(p: Hyp3dParams, planes, solid, steps, clock=(1e-5, 1e-3)):
    """CPU oracle: returns (planes, clock(t, d_tau), dt_hist, maxs_hist)."""
    planes = [np.array(a, np.float32, order="C", copy=True).ravel() for a in planes]
    solid = np.ascontiguousarray(solid, np.uint8).ravel()
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in planes])
    ck = np.array(clock, np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    ms = np.zeros(max(steps, 1), np.float32)
    lib.oracle_hyp3d_run(C.byref(p), ptrs, solid, steps, ck, dts.ctypes.data_as(C.c_void_p),
                         ms.ctypes.data_as(C.c_void_p))
    return planes, (float(ck[0]), float(ck[1])), dts[:steps], ms[:steps]


def ref_hyp3d_run(p: Hyp3dParams, steps, planes=None, clock=(1e-5, 1e-3)):
    """The reference's own k_step on the GPU (oracle/_ref/libref_hyp3d.so).
    Returns (planes, solid, clock, dt_hist, maxs_hist, ms)."""
    r = ref("ref_hyp3d")
    r.ref_hyp3d_run.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _pp6, u8p, f32p,
                                C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    r.ref_hyp3d_run.restype = C.c_int
    N = p.nx * p.ny * p.nz
    do_init = planes is None
    planes = [np.zeros(N, np.float32) for _ in range(6)] if do_init else \
        [np.array(a, np.float32, order="C", copy=True).ravel() for a in planes]
    solid = np.zeros(N, np.uint8)
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in planes])
    ck = np.array(clock, np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    mx = np.zeros(max(steps, 1), np.float32)
    ms = C.c_float()
    rc = r.ref_hyp3d_run(p.as23(), p.nx, p.ny, p.nz, steps, 1 if do_init else 0, ptrs, solid, ck,
                         dts.ctypes.data_as(C.c_void_p), mx.ctypes.data_as(C.c_void_p), C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference hyp3d run failed with cudaError {rc}")
    return planes, solid, (float(ck[0]), float(ck[1])), dts[:steps], mx[:steps], float(ms.value)


# ------------------------------------------------------------------------------------------------
# SPH (tau_sph.cu), fp32
# ------------------------------------------------------------------------------------------------
_SPH_FIELDS = [("N", C.c_int), ("boxX", C.c_float), ("boxY", C.c_float), ("dTau", C.c_float),
               ("t0", C.c_float), ("CFL", C.c_float), ("rho0", C.c_float), ("c0", C.c_float),
               ("gammaEOS", C.c_float), ("hMul", C.c_float), ("viscAlpha", C.c_float),
               ("gravity", C.c_float), ("rain", C.c_int), ("useVisc", C.c_int), ("useGrav", C.c_int),
               ("viscSub", C.c_int), ("useXSPH", C.c_int), ("xsphEps", C.c_float), ("seed", C.c_int)]


class SphParams(C.Structure):
    _fields_ = _SPH_FIELDS

    def as19(self):
        return np.array([float(getattr(self, f[0])) for f in _SPH_FIELDS], np.float32)


class SphClock(C.Structure):
    _fields_ = [("t", C.c_float), ("tau", C.c_float), ("rain_carry", C.c_float), ("step", C.c_longlong)]


u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_sphp = C.POINTER(SphParams)
lib.oracle_sph_run.argtypes = [_sphp, f32p, f32p, f32p, f32p, f32p, C.c_int, C.POINTER(SphClock)]
lib.oracle_sph_run.restype = None
lib.oracle_sph_cell_sort.argtypes = [_sphp, f32p, u32p, u32p, i32p]
lib.oracle_sph_cell_sort.restype = None
lib.oracle_sph_derived.argtypes = [_sphp] + [C.POINTER(C.c_float)] * 3 + [C.POINTER(C.c_int)] * 2
lib.oracle_sph_derived.restype = None


def sph_params(N=1 << 16, **over) -> SphParams:
    p = SphParams(N, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 2.0, 0.25, 9.81, 1, 1, 1, 1, 0, 0.25, 69420)
    for k, v in over.items():
        setattr(p, k, v)
    return p


def sph_derived(p: SphParams):
    m, h, c = C.c_float(), C.c_float(), C.c_float()
    gx, gy = C.c_int(), C.c_int()
    lib.oracle_sph_derived(C.byref(p), C.byref(m), C.byref(h), C.byref(c), C.byref(gx), C.byref(gy))
    return dict(mass=m.value, h=h.value, cell=c.value, Gx=gx.value, Gy=gy.value)


def sph_cell_sort(p: SphParams, pos):
    """(sorted keys, stable permutation, cellStart[M+1]) for the given positions."""
    d = sph_derived(p)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1)
    k, v = np.empty(p.N, np.uint32), np.empty(p.N, np.uint32)
    cs = np.empty(d["Gx"] * d["Gy"] + 1, np.int32)
    lib.oracle_sph_cell_sort(C.byref(p), pos, k, v, cs)
    return k, v, cs


def sph_run(p: SphParams, pos, vel, nframes, clock=None):
    """CPU oracle: returns (pos, vel, acc, s, press, clock)."""
    pos = np.array(pos, np.float32, order="C", copy=True).reshape(-1)
    vel = np.array(vel, np.float32, order="C", copy=True).reshape(-1)
    acc = np.zeros(2 * p.N, np.float32)
    s, pr = np.zeros(p.N, np.float32), np.zeros(p.N, np.float32)
    ck = clock or SphClock(p.t0, 0.0, 0.0, 0)
    lib.oracle_sph_run(C.byref(p), pos, vel, acc, s, pr, nframes, C.byref(ck))
    return pos.reshape(-1, 2), vel.reshape(-1, 2), acc.reshape(-1, 2), s, pr, ck


def ref_sph_reset_particles(p: SphParams):
    r = ref("ref_sph")
    r.ref_sph_reset_particles.argtypes = [f32p, f32p, f32p]
    r.ref_sph_reset_particles.restype = None
    pos, vel = np.empty(2 * p.N, np.float32), np.empty(2 * p.N, np.float32)
    r.ref_sph_reset_particles(p.as19(), pos, vel)
    return pos.reshape(-1, 2), vel.reshape(-1, 2)


def ref_sph_run(p: SphParams, pos, vel, nframes, clock=None):
    """The reference's own kernels on the GPU (oracle/_ref/libref_sph.so).
    Returns (pos, vel, acc, s, press, clock[t, tau, rain_carry, step], ms)."""
    r = ref("ref_sph")
    r.ref_sph_run.argtypes = [f32p] * 6 + [C.c_int, f32p, C.POINTER(C.c_float)]
    r.ref_sph_run.restype = C.c_int
    pos = np.array(pos, np.float32, order="C", copy=True).reshape(-1)
    vel = np.array(vel, np.float32, order="C", copy=True).reshape(-1)
    acc = np.zeros(2 * p.N, np.float32)
    s, pr = np.zeros(p.N, np.float32), np.zeros(p.N, np.float32)
    ck = np.array(clock if clock is not None else [p.t0, 0.0, 0.0, 0.0], np.float32)
    ms = C.c_float()
    rc = r.ref_sph_run(p.as19(), pos, vel, acc, s, pr, nframes, ck, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference SPH run failed with cudaError {rc}")
    return pos.reshape(-1, 2), vel.reshape(-1, 2), acc.reshape(-1, 2), s, pr, ck, float(ms.value)


def sph_rasterize(pos, W, H, boxX=1.0, boxY=1.0):
    """CPU oracle of k_rasterize (tau_sph.cu:363-374): (2H, W) int32 counts."""
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    lib.oracle_sph_rasterize.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, i32p]
    lib.oracle_sph_rasterize.restype = None
    pos = np.ascontiguousarray(pos, np.float32)
    g = np.zeros((2 * H, W), np.int32)
    lib.oracle_sph_rasterize(pos.ravel(), pos.shape[0], W, H, boxX, boxY, g)
    return g


def ref_sph_rasterize(pos, W, H, boxX=1.0, boxY=1.0):
    """The reference's own k_clear_grid + k_rasterize on the GPU."""
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    r = ref("ref_sph")
    r.ref_sph_rasterize.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, i32p]
    r.ref_sph_rasterize.restype = C.c_int
    pos = np.ascontiguousarray(pos, np.float32)
    g = np.zeros((2 * H, W), np.int32)
    rc = r.ref_sph_rasterize(pos.ravel(), pos.shape[0], W, H, boxX, boxY, g)
    if rc != 0:
        raise RuntimeError(f"reference sph rasterize failed with cudaError {rc}")
    return g


def hyp3d_vis(p: Hyp3dParams, planes, solid, mode):
    """CPU oracle of k_vis (tau_hypersonic_3d_cuda.cu:800-905): (nz, ny, nx) float32."""
    lib.oracle_hyp3d_vis.argtypes = [C.POINTER(Hyp3dParams), _pp6, u8p, C.c_int, f32p]
    lib.oracle_hyp3d_vis.restype = None
    planes = [np.ascontiguousarray(a, np.float32).ravel() for a in planes]
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in planes])
    out = np.zeros(p.nx * p.ny * p.nz, np.float32)
    lib.oracle_hyp3d_vis(C.byref(p), ptrs, np.ascontiguousarray(solid, np.uint8).ravel(), mode, out)
    return out.reshape(p.nz, p.ny, p.nx)


def ref_hyp3d_vis(p: Hyp3dParams, planes, mode):
    """The reference's own k_vis on the GPU."""
    r = ref("ref_hyp3d")
    r.ref_hyp3d_vis.argtypes = [f32p, C.c_int, C.c_int, C.c_int, _pp6, C.c_int, f32p]
    r.ref_hyp3d_vis.restype = C.c_int
    planes = [np.ascontiguousarray(a, np.float32).ravel() for a in planes]
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in planes])
    out = np.zeros(p.nx * p.ny * p.nz, np.float32)
    rc = r.ref_hyp3d_vis(p.as23(), p.nx, p.ny, p.nz, ptrs, mode, out)
    if rc != 0:
        raise RuntimeError(f"reference hyp3d vis failed with cudaError {rc}")
    return out.reshape(p.nz, p.ny, p.nx)


# ------------------------------------------------------------------------------------------------
# Burgers
# ------------------------------------------------------------------------------------------------
class BurgersParams(C.Structure):
    """simulation fields of `struct Params` tau_burgers.cu:53-90"""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int)] + [(k, C.c_float) for k in (
        "dx", "dy", "nu", "u0", "amp", "bsig", "swirl", "rc", "offx", "offy", "asym", "CFL", "tau0", "t0",
        "dtau")] + [("muscl", C.c_int), ("visc_substeps", C.c_int), ("colehopf", C.c_int), ("ck", C.c_int),
                    ("ca", C.c_float)]

    def as22(self):
        return np.array([getattr(self, f[0]) for f in self._fields_], np.float32)

    @property
    def shape(self):
        return (1 if self.colehopf else self.ny, self.nx)


def burgers_params(**over) -> BurgersParams:
    p = BurgersParams()
    lib.oracle_burgers_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def burgers_init(p: BurgersParams):
    lib.oracle_burgers_init.argtypes = [C.POINTER(BurgersParams), f32p, f32p]
    lib.oracle_burgers_init.restype = None
    n = p.shape[0] * p.shape[1]
    u, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.oracle_burgers_init(C.byref(p), u, v)
    return u.reshape(p.shape), v.reshape(p.shape)


def burgers_run(p: BurgersParams, phi_u, phi_v, steps, clock=None, skip_visc=False):
    """CPU oracle: (phi_u, phi_v, (t, tau), dts)."""
    lib.oracle_burgers_run.argtypes = [C.POINTER(BurgersParams), f32p, f32p, C.c_int, f32p, f32p, C.c_int]
    lib.oracle_burgers_run.restype = None
    u = np.array(phi_u, np.float32, order="C", copy=True).ravel()
    v = np.array(phi_v, np.float32, order="C", copy=True).ravel()
    ck = np.array(clock if clock is not None else (p.t0, p.tau0), np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    lib.oracle_burgers_run(C.byref(p), u, v, steps, ck, dts, 1 if skip_visc else 0)
    return u.reshape(p.shape), v.reshape(p.shape), (float(ck[0]), float(ck[1])), dts[:steps]


def burgers_colehopf_error(p: BurgersParams, phi_u, t_now):
    lib.oracle_burgers_colehopf_error.argtypes = [C.POINTER(BurgersParams), f32p, C.c_float]
    lib.oracle_burgers_colehopf_error.restype = C.c_double
    return float(lib.oracle_burgers_colehopf_error(C.byref(p), np.ascontiguousarray(phi_u, np.float32).ravel(),
                                                   t_now))


def ref_burgers_init(p: BurgersParams):
    r = ref("ref_burgers")
    r.ref_burgers_init.argtypes = [f32p, f32p, f32p]
    r.ref_burgers_init.restype = None
    n = p.shape[0] * p.shape[1]
    u, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    r.ref_burgers_init(p.as22(), u, v)
    return u.reshape(p.shape), v.reshape(p.shape)


def ref_burgers_run(p: BurgersParams, phi_u, phi_v, steps, clock=None, skip_visc=False):
    """The reference's own kernels on the GPU: (phi_u, phi_v, (t, tau), dts, ms)."""
    r = ref("ref_burgers")
    r.ref_burgers_run.argtypes = [f32p, f32p, f32p, C.c_int, f32p, f32p, C.c_int, C.POINTER(C.c_float)]
    r.ref_burgers_run.restype = C.c_int
    u = np.array(phi_u, np.float32, order="C", copy=True).ravel()
    v = np.array(phi_v, np.float32, order="C", copy=True).ravel()
    ck = np.array(clock if clock is not None else (p.t0, p.tau0), np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    ms = C.c_float()
    rc = r.ref_burgers_run(p.as22(), u, v, steps, ck, dts, 1 if skip_visc else 0, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference burgers run failed with cudaError {rc}")
    return u.reshape(p.shape), v.reshape(p.shape), (float(ck[0]), float(ck[1])), dts[:steps], float(ms.value)


# ------------------------------------------------------------------------------------------------
# Shallow water
# ------------------------------------------------------------------------------------------------
class SwParams(C.Structure):
    """simulation fields of `struct Params` tau_shallow_water.cu:52-89"""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int)] + [(k, C.c_float) for k in (
        "dx", "dy", "g", "f0", "nu", "H0", "bumpAmp", "bumpSigma", "CFL", "offx", "offy", "asym", "swirl",
        "swirlRc", "tau0", "t0", "dtau")]

    def as19(self):
        return np.array([getattr(self, f[0]) for f in self._fields_], np.float32)

    @property
    def shape(self):
        return (self.ny, self.nx)


def sw_params(**over) -> SwParams:
    p = SwParams()
    lib.oracle_sw_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _sw3(p, a, b, c):
    return tuple(np.array(x, np.float32, order="C", copy=True).ravel() for x in (a, b, c))


def sw_init(p: SwParams):
    lib.oracle_sw_init.argtypes = [C.POINTER(SwParams), f32p, f32p, f32p]
    lib.oracle_sw_init.restype = None
    n = p.nx * p.ny
    s, u, v = (np.zeros(n, np.float32) for _ in range(3))
    lib.oracle_sw_init(C.byref(p), s, u, v)
    return s.reshape(p.shape), u.reshape(p.shape), v.reshape(p.shape)


def sw_run(p: SwParams, sigma, u, v, steps, clock=None):
    """CPU oracle: (sigma, u, v, (t, tau), dts)."""
    lib.oracle_sw_run.argtypes = [C.POINTER(SwParams), f32p, f32p, f32p, C.c_int, f32p, f32p]
    lib.oracle_sw_run.restype = None
    s, u, v = _sw3(p, sigma, u, v)
    ck = np.array(clock if clock is not None else (p.t0, p.tau0), np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    lib.oracle_sw_run(C.byref(p), s, u, v, steps, ck, dts)
    return s.reshape(p.shape), u.reshape(p.shape), v.reshape(p.shape), (float(ck[0]), float(ck[1])), dts[:steps]


def sw_cmax(p: SwParams, sigma, u, v):
    lib.oracle_sw_cmax.argtypes = [C.POINTER(SwParams), f32p, f32p, f32p]
    lib.oracle_sw_cmax.restype = C.c_float
    return float(lib.oracle_sw_cmax(C.byref(p), *_sw3(p, sigma, u, v)))


def _ref_sw_init(libname, fn, p):
    r = ref(libname)
    f = getattr(r, fn)
    f.argtypes = [f32p, f32p, f32p, f32p]
    f.restype = None
    n = p.nx * p.ny
    s, u, v = (np.zeros(n, np.float32) for _ in range(3))
    f(p.as19(), s, u, v)
    return s.reshape(p.shape), u.reshape(p.shape), v.reshape(p.shape)


def ref_sw_host_init(p: SwParams):
    """the reference's initialize_host, compiled by g++ (no GPU needed)"""
    return _ref_sw_init("ref_sw_host", "ref_sw_host_init", p)


def ref_sw_host_run(p: SwParams, sigma, u, v, steps, clock=None, skip_visc=False):
    """The reference's own kernel bodies emulated thread by thread on the CPU (oracle/ref_drivers/
    ref_sw_host.cpp): (sigma, u, v, (t, tau), dts)."""
    r = ref("ref_sw_host")
    r.ref_sw_host_run.argtypes = [f32p, f32p, f32p, f32p, C.c_int, f32p, f32p, C.c_int]
    r.ref_sw_host_run.restype = C.c_int
    s, u, v = _sw3(p, sigma, u, v)
    ck = np.array(clock if clock is not None else (p.t0, p.tau0), np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    r.ref_sw_host_run(p.as19(), s, u, v, steps, ck, dts, 1 if skip_visc else 0)
    return s.reshape(p.shape), u.reshape(p.shape), v.reshape(p.shape), (float(ck[0]), float(ck[1])), dts[:steps]


def ref_sw_init(p: SwParams):
    return _ref_sw_init("ref_sw", "ref_sw_init", p)


def ref_sw_run(p: SwParams, sigma, u, v, steps, clock=None, skip_visc=False):
    """The reference's own kernels on the GPU: (sigma, u, v, (t, tau), dts, ms)."""
    r = ref("ref_sw")
    r.ref_sw_run.argtypes = [f32p, f32p, f32p, f32p, C.c_int, f32p, f32p, C.c_int, C.POINTER(C.c_float)]
    r.ref_sw_run.restype = C.c_int
    s, u, v = _sw3(p, sigma, u, v)
    ck = np.array(clock if clock is not None else (p.t0, p.tau0), np.float32)
    dts = np.zeros(max(steps, 1), np.float32)
    ms = C.c_float()
    rc = r.ref_sw_run(p.as19(), s, u, v, steps, ck, dts, 1 if skip_visc else 0, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"reference shallow-water run failed with cudaError {rc}")
    return (s.reshape(p.shape), u.reshape(p.shape), v.reshape(p.shape), (float(ck[0]), float(ck[1])), dts[:steps],
            float(ms.value))


# ------------------------------------------------------------------------------------------------
# `.4spl` export: host-side quantisation + palette of th3cs.cu main
# ------------------------------------------------------------------------------------------------
def splat4_palette(p_size=256):
    lib.oracle_4spl_palette.argtypes = [f32p, C.c_int]
    lib.oracle_4spl_palette.restype = None
    pal = np.zeros(12 * p_size, np.float32)
    lib.oracle_4spl_palette(pal, p_size)
    return pal.reshape(p_size, 12)


def splat4_index(norm: float) -> int:
    lib.oracle_4spl_index.argtypes = [C.c_float]
    lib.oracle_4spl_index.restype = C.c_int
    return int(lib.oracle_4spl_index(norm))


def splat4_frame_indices(sch):
    """th3cs.cu:1199-1222 on a schlieren volume: (uint8 indices of the same shape, (min, max))"""
    lib.oracle_4spl_frame_indices.argtypes = [f32p, C.c_long, u8p, f32p]
    lib.oracle_4spl_frame_indices.restype = None
    a = np.ascontiguousarray(sch, np.float32)
    out = np.zeros(a.size, np.uint8)
    mm = np.zeros(2, np.float32)
    lib.oracle_4spl_frame_indices(a.ravel(), a.size, out, mm)
    return out.reshape(a.shape), (float(mm[0]), float(mm[1]))


def ref_th3cs_host_run(n, frames):
    """The reference's `.4spl` exporter (th3cs.cu's whole main()) run on the CPU emulator for an n^3 grid:
    (header dict, palette (pSize, 12), indices (frames, n, n, n) uint8).  See oracle/ref_drivers/ref_th3cs_host.cpp."""
    r = ref("ref_th3cs_host")
    u32p_ = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    r.ref_th3cs_host_run.argtypes = [C.c_int, C.c_int, u32p_, f32p, u8p]
    r.ref_th3cs_host_run.restype = C.c_int
    hdr = np.zeros(6, np.uint32)
    pal = np.zeros(256 * 12, np.float32)
    idx = np.zeros(frames * n ** 3, np.uint8)
    rc = r.ref_th3cs_host_run(n, frames, hdr, pal, idx)
    if rc != 0:
        raise RuntimeError(f"emulated th3cs main() failed with {rc}")
    keys = ("width", "height", "depth", "frames", "pSize", "flags")
    return dict(zip(keys, [int(x) for x in hdr])), pal.reshape(256, 12), idx.reshape(frames, n, n, n)
