"""ctypes binding of libpairs_b200.so (include/pairs_b200.h): the thin C-ABI shim between the Python host
code and the hand-written sm_100a kernels.  No PyTorch, no Triton, no CPU fallback: if the library or a CUDA
device is missing, construction fails loudly.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "lib", "libpairs_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["ctx.cu", "binning.cu", "neighbor.cu", "md_kernels.cu", "comm.cu", "comm_nccl.cu", "migrate.cu", "setup.cu", "md_run.cu", "dem_kernels.cu", "jit.cu", "props.cu", "tile_lists.cu"]

# --fmad=false: fp64 multiplies and adds are never contracted, so per-operation results equal the reference CPU
# build compiled with -ffp-contract=off (the parity contract).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


class BackendError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into pairs_b200/lib/libpairs_b200.so (in-tree).  One object per source, compiled
    in parallel and only when the source or a header is newer, then linked."""
    from concurrent.futures import ThreadPoolExecutor
    headers = [os.path.join(CSRC, "ctx.cuh"), os.path.join(CSRC, "dem_math.h"), os.path.join(CSRC, "md_math.h"), os.path.join(CSRC, "dem_force_kernel.cuh"),
               os.path.join(INCLUDE, "pairs_b200.h")]
    objdir = os.path.join(os.path.dirname(LIB_PATH), "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-I" + objdir]
    # texts of the headers NVRTC needs at run time for user-defined DEM contact models (csrc/jit.cu), as raw string literals
    emb = os.path.join(objdir, "embedded_sources.inc")
    text = "".join(f'static const char *{sym} = R"PBSRC({open(os.path.join(CSRC, f)).read()})PBSRC";\n'
                   for sym, f in (("PB_SRC_DEM_MATH", "dem_math.h"), ("PB_SRC_DEM_FORCE_KERNEL", "dem_force_kernel.cuh")))
    if not os.path.exists(emb) or open(emb).read() != text:
        with open(emb, "w") as f:
            f.write(text)

    def stale(target, deps):
        return force or not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)

    def compile_one(name):
        src, obj = os.path.join(CSRC, name), os.path.join(objdir, name[:-3] + ".o")
        if not stale(obj, [src] + headers + [__file__]):
            return obj, 0, ""
        r = subprocess.run([nvcc, *flags, "-I" + INCLUDE, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return obj, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for obj, rc, out in results:
        if verbose or rc != 0:
            print(out)
        if rc != 0:
            raise BackendError(f"nvcc failed compiling {os.path.basename(obj)[:-2]}.cu for libpairs_b200.so")
    objs = [obj for obj, _, _ in results]
    if stale(LIB_PATH, objs):
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-o", LIB_PATH, *objs, "-ldl"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout)
        if r.returncode != 0:
            raise BackendError("nvcc failed linking libpairs_b200.so")
    return LIB_PATH


class MdParams(ctypes.Structure):
    _fields_ = [("dt", ctypes.c_double), ("cutoff_force", ctypes.c_double), ("cutoff_lists", ctypes.c_double),
                ("cell_spacing", ctypes.c_double), ("reneighbor_every", ctypes.c_int), ("thermo_every", ctypes.c_int)]


_P = ctypes.c_void_p
_I = ctypes.c_int
_D = ctypes.c_double
_IP = ctypes.POINTER(ctypes.c_int)
_DP = ctypes.POINTER(ctypes.c_double)
_S = ctypes.c_char_p

# every symbol include/pairs_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pb_create": (_I, [ctypes.POINTER(_P), _I]),
    "pb_destroy": (None, [_P]),
    "pb_last_error": (_S, [_P]),
    "pb_version": (_S, []),
    "pb_init_domain": (_I, [_P, _DP, _IP, _I, _I, _I]),
    "pb_get_decomposition": (_I, [_P, _IP, _IP, _IP, _DP]),
    "pb_rank_grid": (_I, [_I, _DP, _I, _IP]),
    "pb_reserve": (_I, [_P, _I, _I]),
    "pb_copper_fcc_lattice": (_I, [_P, _I, _I, _I, _D, _I, _IP]),
    "pb_adjust_thermo": (_I, [_P, _D]),
    "pb_upload_particles": (_I, [_P, _I, _DP, _DP, _DP, _IP, _IP, _IP, _IP]),
    "pb_counts": (_I, [_P, _IP, _IP]),
    "pb_download_real": (_I, [_P, _S, _DP, _I]),
    "pb_download_int": (_I, [_P, _S, _IP, _I]),
    "pb_download_neighbors": (_I, [_P, _IP, _I]),
    "pb_neighbor_capacity": (_I, [_P]),
    "pb_max_neighbors": (_I, [_P]),
    "pb_download_ghost_map": (_I, [_P, _IP, _IP]),
    "pb_setup_cells": (_I, [_P, _D]),
    "pb_get_cells": (_I, [_P, _IP, _IP, _IP]),
    "pb_build_cell_lists": (_I, [_P]),
    "pb_download_cell_lists": (_I, [_P, _IP, _IP]),
    "pb_build_neighbor_lists": (_I, [_P, _D]),
    "pb_set_lj_params": (_I, [_P, _I, _DP, _DP]),
    "pb_reset_volatile": (_I, [_P]),
    "pb_lennard_jones": (_I, [_P, _D]),
    "pb_initial_integrate": (_I, [_P, _D]),
    "pb_final_integrate": (_I, [_P, _D]),
    "pb_lj_legacy": (_I, [_P, _D, _D, _D]),
    "pb_euler_legacy": (_I, [_P, _D]),
    "pb_lj_energy_virial": (_I, [_P, _D, _DP, _DP]),
    "pb_compute_thermo": (_I, [_P, _DP, _DP]),
    "pb_thermo_partial": (_I, [_P, _DP, _IP]),
    "pb_exchange": (_I, [_P]),
    "pb_borders": (_I, [_P]),
    "pb_synchronize": (_I, [_P]),
    "pb_add_property": (_I, [_P, _S, _I, _I, _DP, _IP]),
    "pb_property_count": (_I, [_P]),
    "pb_property_info": (_I, [_P, _I, _IP, _IP, _IP]),
    "pb_upload_property": (_I, [_P, _I, _I, _DP]),
    "pb_download_property": (_I, [_P, _I, _DP, _I]),
    "pb_dem_enable": (_I, [_P, _I]),
    "pb_dem_enable_ex": (_I, [_P, _I, _I, _DP]),
    "pb_dem_contact_extras": (_I, [_P, _I, _DP, _I]),
    "pb_dem_contact_extra_lanes": (_I, [_P]),
    "pb_dem_set_params": (_I, [_P, _D, _D, _D, _D, _D, _D, _D, _D, _I, _DP, _DP]),
    "pb_dem_sc_grid": (_I, [_P, _D, _D, _D, _D, _D, _D, _D, _D, _D, _I, _I, _IP, _IP, _DP, _DP, _DP, _DP, _IP]),
    "pb_dem_upload_real": (_I, [_P, _S, _I, _I, _DP]),
    "pb_dem_download_real": (_I, [_P, _S, _I, _I, _DP]),
    "pb_set_counts": (_I, [_P, _I, _I]),
    "pb_dem_upload_contacts": (_I, [_P, _I, _IP, _IP, _IP, _DP, _DP]),
    "pb_dem_download_contacts": (_I, [_P, _I, _IP, _IP, _IP, _IP, _DP, _DP]),
    "pb_dem_update_mass_and_inertia": (_I, [_P]),
    "pb_dem_reset_contact_usage": (_I, [_P]),
    "pb_dem_clear_unused_contacts": (_I, [_P]),
    "pb_dem_gravity": (_I, [_P]),
    "pb_dem_linear_spring_dashpot": (_I, [_P]),
    "pb_dem_euler": (_I, [_P]),
    "pb_dem_contact_overflow": (_I, [_P]),
    "pb_dem_check_contacts": (_I, [_P]),
    "pb_dem_contact_capacity": (_I, [_P]),
    "pb_dem_run": (_I, [_P, _D, _I, _I]),
    "pb_nccl_unique_id": (_I, [_P]),
    "pb_nccl_init": (_I, [_P, _P]),
    "pb_md_run": (_I, [_P, ctypes.POINTER(MdParams), _I, _I, _DP, _I, _IP]),
    "pb_md_run_from_host": (_I, [_P, ctypes.POINTER(MdParams), _I, _DP, _DP, _DP, _IP, _IP, _IP, _IP, _I, _I, _DP, _I, _IP]),
    "pb_set_option": (_I, [_P, _S, _I]),
    "pb_board_selftest": (_I, [_S, _I, _I, _I]),
    "pb_board_unlink": (_I, [_S]),
    "pb_jit_prelude": (_S, []),
    "pb_jit_check": (_I, [_S, ctypes.c_char_p, _I]),
    "pb_jit_compile": (_I, [_P, _S, _S, _IP]),
    "pb_jit_launch": (_I, [_P, _I, _I, _D]),
    "pb_jit_check_dem_model": (_I, [_S, _S, ctypes.c_char_p, _I]),
    "pb_jit_set_dem_model": (_I, [_P, _S, _S]),
    "pb_synchronize_device": (_I, [_P]),
    "pb_timers_enable": (_I, [_P, _I]),
    "pb_timers_get": (_I, [_P, _S, _DP, ctypes.POINTER(ctypes.c_long)]),
    "pb_timers_reset": (_I, [_P]),
    "pb_stream_timer_start": (_I, [_P]),
    "pb_stream_timer_stop": (_I, [_P, _DP]),
    "pb_host_register": (_I, [_P, _P, ctypes.c_size_t]),
    "pb_host_unregister": (_I, [_P, _P]),
    "pb_kernel_launches": (ctypes.c_long, [_P]),
}

_LIB = None


def load():
    """dlopen the shim and type every entry point.  Raises BackendError if the library is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise BackendError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(_DP)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_IP)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def rank_grid(world_size, grid, partitioner=0):
    out = np.zeros(3, np.int32)
    g = _f64(grid)
    load().pb_rank_grid(world_size, _dp(g), partitioner, _ip(out))
    return tuple(int(x) for x in out)


class Context:
    """One GPU's particle store + the per-stage entry points (one per reference module)."""

    def __init__(self, device=0):
        self.lib = load()
        h = _P()
        rc = self.lib.pb_create(ctypes.byref(h), device)
        if rc != 0:
            raise BackendError(self.lib.pb_last_error(None).decode())
        self._h = h
        self.device = device

    @property
    def h(self):
        """The library handle; a closed context raises instead of handing a null pointer to the library."""
        if self._h is None:
            raise BackendError("this Context has been closed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.pb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise BackendError(self.lib.pb_last_error(self.h).decode())
        return rc

    # ---- domain ----
    def init_domain(self, grid, pbc=(1, 1, 1), partitioner=0, world_size=1, rank=0):
        g = _f64(grid)
        p = _i32([1 if b else 0 for b in pbc])
        self._ck(self.lib.pb_init_domain(self.h, _dp(g), _ip(p), partitioner, world_size, rank))

    def decomposition(self):
        nr = np.zeros(3, np.int32)
        nb = np.zeros(6, np.int32)
        pbc = np.zeros(6, np.int32)
        sub = np.zeros(6, np.float64)
        self._ck(self.lib.pb_get_decomposition(self.h, _ip(nr), _ip(nb), _ip(pbc), _dp(sub)))
        return {"nranks": nr, "neighbor_ranks": nb, "pbc": pbc, "subdom": sub}

    def reserve(self, particle_capacity=0, neighbor_capacity=0):
        self._ck(self.lib.pb_reserve(self.h, particle_capacity, neighbor_capacity))

    # ---- set-up ----
    def copper_fcc_lattice(self, nx, ny, nz, rho, ntypes):
        n = _I(0)
        self._ck(self.lib.pb_copper_fcc_lattice(self.h, nx, ny, nz, rho, ntypes, ctypes.byref(n)))
        return n.value

    def adjust_thermo(self, temp):
        self._ck(self.lib.pb_adjust_thermo(self.h, temp))

    def upload(self, position, velocity=None, mass=None, type_=None, flags=None, uid=None, shape=None):
        pos = _f64(position)
        n = pos.size // 3
        vel, m = _f64(velocity), _f64(mass)
        t, f, u, s = _i32(type_), _i32(flags), _i32(uid), _i32(shape)
        self._ck(self.lib.pb_upload_particles(self.h, n, _dp(pos), _dp(vel), _dp(m), _ip(t), _ip(f), _ip(u), _ip(s)))

    # ---- download ----
    def counts(self):
        a, b = _I(0), _I(0)
        self.lib.pb_counts(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def real(self, name, with_ghosts=False):
        nl, ng = self.counts()
        n = nl + (ng if with_ghosts else 0)
        w = 1 if name == "mass" else 3
        out = np.zeros(n * w, np.float64)
        self._ck(self.lib.pb_download_real(self.h, name.encode(), _dp(out), 1 if with_ghosts else 0))
        return out.reshape(n, 3) if w == 3 else out

    def ints(self, name, with_ghosts=False):
        nl, ng = self.counts()
        n = nl + (ng if with_ghosts else 0)
        if name == "numneighs":
            n = nl
        out = np.zeros(n, np.int32)
        self._ck(self.lib.pb_download_int(self.h, name.encode(), _ip(out), 1 if with_ghosts else 0))
        return out

    def neighbors(self):
        nl, _ = self.counts()
        cap = max(self.lib.pb_max_neighbors(self.h), 1)
        out = np.zeros(nl * cap, np.int32)
        self._ck(self.lib.pb_download_neighbors(self.h, _ip(out), cap))
        return out.reshape(nl, cap)

    def ghost_map(self):
        _, ng = self.counts()
        src = np.zeros(ng, np.int32)
        mult = np.zeros(ng * 3, np.int32)
        self._ck(self.lib.pb_download_ghost_map(self.h, _ip(src), _ip(mult)))
        return src, mult.reshape(ng, 3)

    # ---- stages ----
    def setup_cells(self, spacing):
        self._ck(self.lib.pb_setup_cells(self.h, spacing))

    def cells(self):
        dc = np.zeros(3, np.int32)
        nc = _I(0)
        st = np.zeros(27, np.int32)
        self._ck(self.lib.pb_get_cells(self.h, _ip(dc), ctypes.byref(nc), _ip(st)))
        return dc, nc.value, st

    def build_cell_lists(self):
        self._ck(self.lib.pb_build_cell_lists(self.h))

    def cell_lists(self):
        _, nc, _ = self.cells()
        nl, ng = self.counts()
        cs = np.zeros(nc + 1, np.int32)
        cl = np.zeros(nl + ng, np.int32)
        self._ck(self.lib.pb_download_cell_lists(self.h, _ip(cs), _ip(cl)))
        return cs, cl

    def build_neighbor_lists(self, cutoff):
        self._ck(self.lib.pb_build_neighbor_lists(self.h, cutoff))

    def set_lj_params(self, ntypes, epsilon, sigma6):
        e, s = _f64(epsilon), _f64(sigma6)
        self._ck(self.lib.pb_set_lj_params(self.h, ntypes, _dp(e), _dp(s)))

    def reset_volatile(self):
        self._ck(self.lib.pb_reset_volatile(self.h))

    def lennard_jones(self, cutoff):
        self._ck(self.lib.pb_lennard_jones(self.h, cutoff))

    def initial_integrate(self, dt):
        self._ck(self.lib.pb_initial_integrate(self.h, dt))

    def final_integrate(self, dt):
        self._ck(self.lib.pb_final_integrate(self.h, dt))

    def lj_legacy(self, cutoff, epsilon, sigma6):
        self._ck(self.lib.pb_lj_legacy(self.h, cutoff, epsilon, sigma6))

    def euler_legacy(self, dt):
        self._ck(self.lib.pb_euler_legacy(self.h, dt))

    def lj_energy_virial(self, cutoff):
        e, w = _D(0.0), _D(0.0)
        self._ck(self.lib.pb_lj_energy_virial(self.h, cutoff, ctypes.byref(e), ctypes.byref(w)))
        return e.value, w.value

    def compute_thermo(self):
        t, p = _D(0.0), _D(0.0)
        self._ck(self.lib.pb_compute_thermo(self.h, ctypes.byref(t), ctypes.byref(p)))
        return t.value, p.value

    def exchange(self):
        self._ck(self.lib.pb_exchange(self.h))

    def borders(self):
        self._ck(self.lib.pb_borders(self.h))

    def synchronize(self):
        self._ck(self.lib.pb_synchronize(self.h))

    # ---- DEM ----
    DEM_WIDTH = {"radius": 1, "angular_velocity": 3, "torque": 3, "normal": 3, "inv_inertia": 9, "rotation_matrix": 9, "rotation_quat": 4,
                 "force": 3, "mass": 1, "linear_velocity": 3}

    def dem_enable(self, contact_capacity=20, extra_defaults=()):
        """extra_defaults: one default per further double lane of a contact (contact properties beyond dem.py's three)"""
        d = _f64(list(extra_defaults))
        self._ck(self.lib.pb_dem_enable_ex(self.h, contact_capacity, len(d), _dp(d) if len(d) else None))
        self._dem = True

    @property
    def contact_extra_lanes(self):
        return int(self.lib.pb_dem_contact_extra_lanes(self.h)) if getattr(self, "_dem", False) else 0

    def dem_upload_contact_extras(self, values):
        """values[n][contact_capacity][extra lanes]"""
        a = _f64(values)
        n = a.size // max(1, self.contact_capacity * self.contact_extra_lanes)
        self._ck(self.lib.pb_dem_contact_extras(self.h, n, _dp(a), 1))

    def dem_download_contact_extras(self, n=None):
        nl, _ = self.counts()
        n = nl if n is None else n
        out = np.zeros((n, self.contact_capacity, self.contact_extra_lanes))
        if out.size:
            self._ck(self.lib.pb_dem_contact_extras(self.h, n, _dp(out), 0))
        return out

    @property
    def contact_capacity(self):
        """slots per contact row (0: not a DEM context); grows when a row comes close to full (pb_dem_check_contacts)"""
        return int(self.lib.pb_dem_contact_capacity(self.h)) if getattr(self, "_dem", False) else 0

    def dem_check_contacts(self):
        self._ck(self.lib.pb_dem_check_contacts(self.h))

    def dem_set_params(self, dt, pi, kappa, ln_dry_res_coeff, collision_time, density_particle, density_fluid, gravity, ntypes,
                       friction_static, friction_dynamic):
        fs, fd = _f64(friction_static), _f64(friction_dynamic)
        self._ck(self.lib.pb_dem_set_params(self.h, dt, pi, kappa, ln_dry_res_coeff, collision_time, density_particle, density_fluid,
                                            gravity, ntypes, _dp(fs), _dp(fd)))

    def dem_sc_grid(self, xmax, ymax, zmax, spacing, diameter, min_diameter, max_diameter, initial_velocity, particle_density, ntypes):
        n = _I(0)
        args = (xmax, ymax, zmax, spacing, diameter, min_diameter, max_diameter, initial_velocity, particle_density, ntypes)
        self._ck(self.lib.pb_dem_sc_grid(self.h, *args, 0, None, None, None, None, None, None, ctypes.byref(n)))
        c = n.value
        out = {"uid": np.zeros(c, np.int32), "type": np.zeros(c, np.int32), "mass": np.zeros(c), "radius": np.zeros(c),
               "position": np.zeros((c, 3)), "linear_velocity": np.zeros((c, 3))}
        self._ck(self.lib.pb_dem_sc_grid(self.h, *args, c, _ip(out["uid"]), _ip(out["type"]), _dp(out["mass"]), _dp(out["radius"]),
                                         _dp(out["position"]), _dp(out["linear_velocity"]), ctypes.byref(n)))
        return out

    def dem_upload(self, name, data, first=0):
        a = _f64(data)
        n = a.size // self.DEM_WIDTH[name]
        self._ck(self.lib.pb_dem_upload_real(self.h, name.encode(), first, n, _dp(a)))

    def dem_download(self, name, n=None, first=0):
        nl, ng = self.counts()
        n = nl if n is None else n
        w = self.DEM_WIDTH[name]
        out = np.zeros(n * w)
        self._ck(self.lib.pb_dem_download_real(self.h, name.encode(), first, n, _dp(out)))
        return out.reshape(n, w) if w > 1 else out

    def set_counts(self, nlocal, nghost):
        self._ck(self.lib.pb_set_counts(self.h, nlocal, nghost))

    def dem_upload_contacts(self, num, uid, sticking, tsd, ivm):
        num, uid, st = _i32(num), _i32(uid), _i32(sticking)
        t, v = _f64(tsd), _f64(ivm)
        self._ck(self.lib.pb_dem_upload_contacts(self.h, len(num), _ip(num), _ip(uid), _ip(st), _dp(t), _dp(v)))

    def dem_download_contacts(self, n=None):
        nl, _ = self.counts()
        n = nl if n is None else n
        C = self.contact_capacity
        num = np.zeros(n, np.int32)
        uid, used, st = (np.zeros((n, C), np.int32) for _ in range(3))
        tsd, ivm = np.zeros((n, C, 3)), np.zeros((n, C))
        self._ck(self.lib.pb_dem_download_contacts(self.h, n, _ip(num), _ip(uid), _ip(used), _ip(st), _dp(tsd), _dp(ivm)))
        return {"num_contacts": num, "contact_lists": uid, "contact_used": used, "is_sticking": st,
                "tangential_spring_displacement": tsd, "impact_velocity_magnitude": ivm}

    def dem_stage(self, name):
        self._ck(getattr(self.lib, "pb_dem_" + name)(self.h))

    def dem_run(self, cell_spacing, ts_begin, ts_end):
        self._ck(self.lib.pb_dem_run(self.h, cell_spacing, ts_begin, ts_end))

    def nccl_init(self, id128: bytes):
        buf = ctypes.create_string_buffer(id128, 128)
        self._ck(self.lib.pb_nccl_init(self.h, ctypes.cast(buf, _P)))

    def md_run(self, ts_begin, ts_end, dt, cutoff_force, cutoff_lists, cell_spacing, reneighbor_every, thermo_every):
        p = MdParams(dt, cutoff_force, cutoff_lists, cell_spacing, reneighbor_every, thermo_every)
        cap = max(8, (ts_end - ts_begin) // max(thermo_every, 1) + 4) if thermo_every > 0 else 1
        out = np.zeros(cap * 3, np.float64)
        n = _I(0)
        self._ck(self.lib.pb_md_run(self.h, ctypes.byref(p), ts_begin, ts_end, _dp(out), cap, ctypes.byref(n)))
        return out[: min(n.value, cap) * 3].reshape(-1, 3)

    def md_run_from_host(self, position, velocity, mass, type_, ts_begin, ts_end, dt, cutoff_force, cutoff_lists, cell_spacing,
                         reneighbor_every, thermo_every, flags=None, uid=None, shape=None):
        """upload() + md_run() in one call: velocities and masses are copied while the first list build runs (pb_md_run_from_host)"""
        pos = _f64(position)
        vel, m = _f64(velocity), _f64(mass)
        t, f, u, s = _i32(type_), _i32(flags), _i32(uid), _i32(shape)
        p = MdParams(dt, cutoff_force, cutoff_lists, cell_spacing, reneighbor_every, thermo_every)
        cap = max(8, (ts_end - ts_begin) // max(thermo_every, 1) + 4) if thermo_every > 0 else 1
        out = np.zeros(cap * 3, np.float64)
        n = _I(0)
        self._ck(self.lib.pb_md_run_from_host(self.h, ctypes.byref(p), pos.size // 3, _dp(pos), _dp(vel), _dp(m), _ip(t), _ip(f), _ip(u), _ip(s),
                                              ts_begin, ts_end, _dp(out), cap, ctypes.byref(n)))
        return out[: min(n.value, cap) * 3].reshape(-1, 3)

    # ---- user-defined properties (csrc/props.cu) ----
    def add_property(self, name, ncomps, volatile=False, defaults=None):
        """-> (property id, first row of the property in PbJitArgs.xdata)"""
        d = np.zeros(ncomps) if defaults is None else np.ascontiguousarray(np.broadcast_to(np.asarray(defaults, np.float64), (ncomps,)))
        pid, row0 = _I(0), _I(0)
        self._ck(self.lib.pb_add_property(self.h, name.encode(), ncomps, 1 if volatile else 0, _dp(d), ctypes.byref(pid)))
        self._ck(self.lib.pb_property_info(self.h, pid.value, None, ctypes.byref(row0), None))
        self._xprops = getattr(self, "_xprops", {})
        self._xprops[name] = (pid.value, ncomps)
        return pid.value, row0.value

    def upload_property(self, name, values):
        pid, ncomps = self._xprops[name]
        v = np.ascontiguousarray(values, np.float64)
        n = v.shape[0]
        assert v.size == n * ncomps, f"{name}: expected [n][{ncomps}] values"
        self._ck(self.lib.pb_upload_property(self.h, pid, n, _dp(v)))

    def download_property(self, name, with_ghosts=False):
        pid, ncomps = self._xprops[name]
        nl, ng = self.counts()
        n = nl + (ng if with_ghosts else 0)
        out = np.zeros((n, ncomps) if ncomps > 1 else n, np.float64)
        self._ck(self.lib.pb_download_property(self.h, pid, _dp(out), 1 if with_ghosts else 0))
        return out

    def jit_compile(self, source, kernel_name):
        h = _I(0)
        self._ck(self.lib.pb_jit_compile(self.h, source.encode(), kernel_name.encode(), ctypes.byref(h)))
        return h.value

    def jit_set_dem_model(self, model_source, model_name):
        """Installs a generated contact model into the DEM contact kernel (None: back to examples/dem.py's)."""
        self._ck(self.lib.pb_jit_set_dem_model(self.h, None if model_source is None else model_source.encode(),
                                               None if model_name is None else model_name.encode()))

    def jit_launch(self, handle, kind, cutoff=0.0):
        self._ck(self.lib.pb_jit_launch(self.h, handle, kind, cutoff))

    def set_option(self, name, value):
        self._ck(self.lib.pb_set_option(self.h, name.encode(), int(value)))

    def sync(self):
        self._ck(self.lib.pb_synchronize_device(self.h))

    def timers_enable(self, on=True):
        self.lib.pb_timers_enable(self.h, 1 if on else 0)

    def timers_reset(self):
        self.lib.pb_timers_reset(self.h)

    def timer(self, name):
        ms = _D(0.0)
        calls = ctypes.c_long(0)
        self.lib.pb_timers_get(self.h, name.encode(), ctypes.byref(ms), ctypes.byref(calls))
        return ms.value, calls.value

    def stream_timer_start(self):
        self._ck(self.lib.pb_stream_timer_start(self.h))

    def stream_timer_stop(self):
        ms = _D(0.0)
        self._ck(self.lib.pb_stream_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def host_register(self, arr):
        self._ck(self.lib.pb_host_register(self.h, arr.ctypes.data_as(_P), arr.nbytes))

    def host_unregister(self, arr):
        self._ck(self.lib.pb_host_unregister(self.h, arr.ctypes.data_as(_P)))

    def real_into(self, name, out, with_ghosts=False):
        self._ck(self.lib.pb_download_real(self.h, name.encode(), _dp(out), 1 if with_ghosts else 0))
        return out

    def kernel_launches(self):
        return int(self.lib.pb_kernel_launches(self.h))


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    if load().pb_nccl_unique_id(ctypes.cast(buf, _P)) != 0:
        raise BackendError("ncclGetUniqueId failed")
    return buf.raw


def jit_prelude():
    """Source text every generated kernel starts with (struct PbJitArgs etc., csrc/jit.cu)."""
    return load().pb_jit_prelude().decode()


def jit_check_dem_model(model_source, model_name):
    """Compile-only check of the DEM contact kernel built around a generated contact model (no GPU needed)."""
    log = ctypes.create_string_buffer(16384)
    n = load().pb_jit_check_dem_model(None if model_source is None else model_source.encode(),
                                      None if model_name is None else model_name.encode(), log, len(log))
    if n < 0:
        raise BackendError(log.value.decode(errors="replace"))
    return n


def jit_check(source):
    """Compile-only check with NVRTC (no GPU needed): returns the cubin size, raises BackendError with the compiler log."""
    log = ctypes.create_string_buffer(16384)
    n = load().pb_jit_check(source.encode(), log, len(log))
    if n < 0:
        raise BackendError(log.value.decode(errors="replace"))
    return n
