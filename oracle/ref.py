"""ctypes access to oracle/_ref/libref_<variant>.so (the reference's own generated serial C++,
built by oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY -- importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from pairs_b200/.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

HOOK = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p)


def available(variant):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class RefProgram:
    """One generated reference program (md, md_t1, ...)."""

    def __init__(self, variant):
        path = os.path.join(REF_DIR, f"libref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        self.variant = variant
        self.lib = ctypes.CDLL(path)
        self.lib.ref_run.argtypes = [HOOK, ctypes.c_void_p, ctypes.c_int]
        self.lib.ref_run.restype = ctypes.c_int
        self.lib.ref_property_ptr.argtypes = [ctypes.c_char_p]
        self.lib.ref_property_ptr.restype = ctypes.c_void_p
        self.lib.ref_array_ptr.argtypes = [ctypes.c_char_p]
        self.lib.ref_array_ptr.restype = ctypes.c_void_p
        self.lib.ref_array_size.argtypes = [ctypes.c_char_p]
        self.lib.ref_array_size.restype = ctypes.c_long
        self.lib.ref_contact_property_ptr.argtypes = [ctypes.c_char_p]
        self.lib.ref_contact_property_ptr.restype = ctypes.c_void_p

    # ---- access to live state (valid only inside a hook) ----
    def prop(self, name, n, width=1, dtype=np.float64):
        p = self.lib.ref_property_ptr(name.encode())
        if not p:
            raise KeyError(name)
        ct = ctypes.c_double if dtype == np.float64 else ctypes.c_int
        buf = (ct * (n * width)).from_address(p)
        a = np.frombuffer(buf, dtype=dtype).copy()
        return a.reshape(n, width) if width > 1 else a

    def contact_prop(self, name, n, width=1, dtype=np.float64):
        p = self.lib.ref_contact_property_ptr(name.encode())
        if not p:
            raise KeyError(name)
        ct = ctypes.c_double if dtype == np.float64 else ctypes.c_int
        buf = (ct * (n * width)).from_address(p)
        return np.frombuffer(buf, dtype=dtype).copy()

    DEM_REAL = (("position", 3), ("linear_velocity", 3), ("angular_velocity", 3), ("force", 3), ("torque", 3), ("mass", 1), ("radius", 1),
                ("normal", 3), ("inv_inertia", 9), ("rotation_matrix", 9), ("rotation_quat", 4))
    DEM_INT = ("uid", "type", "flags", "shape")

    def dem_state(self, n, ncap=20):
        """Full DEM particle + contact-history state of the first n particles (valid inside a hook)."""
        s = {}
        for name, w in self.DEM_REAL:
            s[name] = self.prop(name, n, w)
        for name in self.DEM_INT:
            s[name] = self.prop(name, n, 1, np.int32)
        s["num_contacts"] = self.array("num_contacts", np.int32, n)
        s["contact_lists"] = self.array("contact_lists", np.int32, n * ncap).reshape(n, ncap)
        s["contact_used"] = self.array("contact_used", np.int32, n * ncap).reshape(n, ncap)
        s["is_sticking"] = self.contact_prop("is_sticking", n * ncap, 1, np.int32).reshape(n, ncap)
        s["tangential_spring_displacement"] = self.contact_prop("tangential_spring_displacement", n * ncap, 3).reshape(n, ncap, 3)
        s["impact_velocity_magnitude"] = self.contact_prop("impact_velocity_magnitude", n * ncap, 1).reshape(n, ncap)
        return s

    def array(self, name, dtype=np.int32, count=None):
        p = self.lib.ref_array_ptr(name.encode())
        if not p:
            raise KeyError(name)
        nbytes = self.lib.ref_array_size(name.encode())
        item = np.dtype(dtype).itemsize
        total = nbytes // item
        count = total if count is None else min(count, total)
        ct = ctypes.c_double if dtype == np.float64 else ctypes.c_int
        buf = (ct * count).from_address(p)
        return np.frombuffer(buf, dtype=dtype).copy()

    def run(self, on_event, quiet=True):
        """Run the generated main(); on_event(event:str, a:int) is called from its hooks."""
        def tramp(ev, a, _user):
            on_event(ev.decode(), a)
        cb = HOOK(tramp)
        return self.lib.ref_run(cb, None, 1 if quiet else 0)

    def run_collect_thermo(self, props=(("position", 3), ("linear_velocity", 3), ("force", 3), ("mass", 1)),
                           int_props=("type", "flags"), with_ghosts=False, steps=None, with_lists=True):
        """Full run; returns one snapshot dict per compute_thermo call (i.e. per thermo step).

        Snapshots are taken right after final_integrate of that step, so `force` is that step's
        freshly computed pair force and `position` the positions it was computed from."""
        snaps = []

        def on_event(ev, a):
            if ev != "thermo":
                return
            if steps is not None and len(snaps) >= steps:
                return
            nlocal = a
            nrecv = self.array("nrecv", np.int32, 6)
            nghost = int(nrecv.sum())
            n = nlocal + (nghost if with_ghosts else 0)
            s = {"nlocal": nlocal, "nghost": nghost}
            for name, w in props:
                s[name] = self.prop(name, n, w)
            for name in int_props:
                s[name] = self.prop(name, n, 1, np.int32)
            if not snaps and with_lists:        # the lists of the first iteration (built from the initial positions)
                cap = self.array("neighborlists", np.int32).size // max(self.array("numneighs", np.int32).size, 1)
                s["numneighs"] = self.array("numneighs", np.int32, nlocal)
                s["neighborlists"] = self.array("neighborlists", np.int32, nlocal * cap).reshape(nlocal, cap)
            snaps.append(s)

        self.run(on_event)
        return snaps

    # ---- direct module calls (md programs; variants built with -DREF_LJ_MODULE export only lennard_jones) ----
    def lennard_jones(self, neighbor_capacity, nlocal, numneighs, neighborlists, flags, position, type_, force, sigma6, epsilon):
        self.lib.ref_md_lennard_jones(ctypes.c_int(neighbor_capacity), ctypes.c_int(nlocal), _ip(numneighs), _ip(neighborlists),
                                      _ip(flags), _dp(position), _ip(type_), _dp(force), _dp(sigma6), _dp(epsilon))

    def initial_integrate(self, nlocal, flags, force, mass, vel, pos):
        self.lib.ref_md_initial_integrate(ctypes.c_int(nlocal), _ip(flags), _dp(force), _dp(mass), _dp(vel), _dp(pos))

    def final_integrate(self, nlocal, flags, force, mass, vel):
        self.lib.ref_md_final_integrate(ctypes.c_int(nlocal), _ip(flags), _dp(force), _dp(mass), _dp(vel))

    # ---- generic module access (variants built with "modules:..." in oracle/build_ref.py) ----
    def call_module(self, module, **kw):
        """Calls a generated module with arguments given BY NAME (numpy arrays for pointers, numbers for scalars); the order the
        generator printed them in is read from oracle/_ref/modules_<variant>.json."""
        import json
        if not hasattr(self, "_modules"):
            with open(os.path.join(REF_DIR, f"modules_{self.variant}.json")) as f:
                self._modules = json.load(f)
        sig = self._modules[module]
        missing = [n for _, n in sig if n not in kw]
        assert not missing and len(kw) == len(sig), f"{module}: arguments are {[n for _, n in sig]}"
        keep, ptrs = [], (ctypes.c_void_p * len(sig))()
        for k, (ctype, name) in enumerate(sig):
            v = kw[name]
            if ctype.endswith("*"):
                want = np.int32 if ctype.startswith("int") else np.float64
                assert isinstance(v, np.ndarray) and v.dtype == want and v.flags.c_contiguous, f"{module}: {name} must be a contiguous {want.__name__} array"
                ptrs[k] = v.ctypes.data
            else:
                c = ctypes.c_int(int(v)) if ctype == "int" else ctypes.c_double(float(v))
                keep.append(c)
                ptrs[k] = ctypes.addressof(c)
        fn = getattr(self.lib, f"ref_mod_{module}")
        fn.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        fn.restype = None
        fn(ptrs)

    def cell_stencil(self, subdom, ncells_capacity=1 << 30):
        ncells = ctypes.c_int(0)
        nstencil = ctypes.c_int(0)
        shapes_buffer = np.zeros(4, np.int32)
        dim_cells = np.zeros(3, np.int32)
        resizes = np.zeros(3, np.int32)
        stencil = np.zeros(28, np.int32)
        subdom = np.ascontiguousarray(subdom, np.float64)
        self.lib.ref_md_build_cell_lists_stencil(ctypes.c_int(ncells_capacity), ctypes.byref(ncells), ctypes.byref(nstencil),
                                                 _ip(shapes_buffer), _dp(subdom), _ip(dim_cells), _ip(resizes), _ip(stencil[1:]))
        return ncells.value, dim_cells, stencil[1:28].copy()

    def build_lists(self, position, flags, shape, nlocal, nghost, subdom, cell_capacity=64, neighbor_capacity=100):
        """build_cell_lists + partition_cell_lists + build_neighbor_lists on caller data
        (the module sequence of sim/simulation.py:392-400).  Returns dict of the reference arrays."""
        ncells, dim_cells, stencil = self.cell_stencil(subdom)
        n = nlocal + nghost
        position = np.ascontiguousarray(position, np.float64)
        flags = np.ascontiguousarray(flags, np.int32)
        shape = np.ascontiguousarray(shape, np.int32)
        subdom = np.ascontiguousarray(subdom, np.float64)
        shapes_buffer = np.array([2], np.int32)
        while True:
            cell_sizes = np.zeros(ncells, np.int32)
            particle_cell = np.zeros(n, np.int32)
            resizes = np.zeros(3, np.int32)
            cell_particles = np.zeros(ncells * cell_capacity, np.int32)
            self.lib.ref_md_build_cell_lists(ctypes.c_int(ncells), ctypes.c_int(nlocal), ctypes.c_int(nghost),
                                             ctypes.c_int(cell_capacity), _ip(cell_sizes), _dp(subdom), _ip(dim_cells),
                                             _ip(particle_cell), _ip(resizes), _ip(cell_particles), _ip(flags), _dp(position))
            if resizes[0] > 0:
                cell_capacity = int(resizes[0]) * 2
                continue
            break
        nshapes = np.zeros(ncells, np.int32)
        self.lib.ref_md_partition_cell_lists(ctypes.c_int(cell_capacity), ctypes.c_int(ncells), _ip(cell_sizes), _ip(nshapes),
                                             _ip(shapes_buffer), _ip(cell_particles), _ip(shape))
        st = np.zeros(28, np.int32)          # stencil[-1] is read (value unused) by the generated code
        st[1:] = stencil
        while True:
            numneighs = np.zeros(nlocal, np.int32)
            neighborlists = np.zeros(nlocal * neighbor_capacity, np.int32)
            resizes = np.zeros(3, np.int32)
            self.lib.ref_md_build_neighbor_lists(ctypes.c_int(nlocal), ctypes.c_int(ncells), ctypes.c_int(cell_capacity),
                                                 ctypes.c_int(neighbor_capacity), ctypes.c_int(27), _ip(numneighs),
                                                 _ip(particle_cell), _ip(st[1:]), _ip(nshapes), _ip(cell_particles),
                                                 _ip(neighborlists), _ip(resizes), _ip(flags), _dp(position))
            if resizes[0] > 0:
                neighbor_capacity = int(resizes[0]) * 2
                continue
            break
        return {"ncells": ncells, "dim_cells": dim_cells, "stencil": stencil, "particle_cell": particle_cell,
                "cell_sizes": cell_sizes, "cell_particles": cell_particles.reshape(ncells, cell_capacity),
                "numneighs": numneighs, "neighborlists": neighborlists.reshape(nlocal, neighbor_capacity),
                "cell_capacity": cell_capacity, "neighbor_capacity": neighbor_capacity}
