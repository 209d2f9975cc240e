"""The `pairs` DSL surface (src/pairs/__init__.py:9-67, Simulation methods src/pairs/sim/simulation.py:37-452) on top of
the B200 backend: examples/md.py runs unchanged, but generate() plans the per-timestep procedure list
(sim/simulation.py:387-417) and RUNS it on the GPU through the C-ABI shim instead of printing C++/CUDA source.

User kernels are ordinary Python functions (examples/md.py:5-17).  The reference lowers them through an IR
(mapping/funcs.py:285-334); this backend RECOGNISES them: the function's AST is unified with the templates of the
hand-written kernel families below (names of properties / symbols are free, structure and literals must match) and
bound to the corresponding CUDA kernel.  A body that matches no family is translated to CUDA and compiled at run time
(kernelgen.py + csrc/jit.cu, for the properties the MD path stores); what cannot be translated fails loudly -- there is
no CPU fallback.
"""
import ast
import inspect
import math
import os
import sys
import textwrap
import time

import numpy as np

_builtin_float = float     # the façade defines pairs.float() below


# ---- enums of the façade (src/pairs/ir/types.py, sim/shapes.py, sim/domain_partitioners.py, code_gen/target.py) --------
class Types:
    Int32, Float, Double, Real, Vector, Matrix, Quaternion = range(7)


class Shapes:
    Sphere, Halfspace, PointMass = 0, 1, 2


class DomainPartitioners:
    Regular, RegularXY = 0, 1


class Target:
    def __init__(self, gpu, parallel=False):
        self.gpu = gpu
        self.parallel = parallel

    def is_gpu(self):
        return self.gpu


class DslError(RuntimeError):
    pass


# ---- kernel recognition -------------------------------------------------------------------------------------------------
# Templates: role names starting with `R_` unify with any user name (bound consistently); `P_i` / `P_j` are the particle
# parameters; DSL keywords (squared_distance, delta, apply) and literals must match exactly.
TEMPLATES = {
    # examples/md.py:5-8
    "lennard_jones": """
def K(P_i, P_j):
    L_sr2 = 1.0 / squared_distance(P_i, P_j)
    L_sr6 = L_sr2 * L_sr2 * L_sr2 * R_sigma6[P_i, P_j]
    apply(R_force, delta(P_i, P_j) * (48.0 * L_sr6 * (L_sr6 - 0.5) * L_sr2 * R_epsilon[P_i, P_j]))
""",
    # examples/md.py:11-13
    "initial_integrate": """
def K(P_i):
    R_velocity[P_i] += (R_dt * 0.5) * R_force[P_i] / R_mass[P_i]
    R_position[P_i] += R_dt * R_velocity[P_i]
""",
    # examples/md.py:16-17
    "final_integrate": """
def K(P_i):
    R_velocity[P_i] += (R_dt * 0.5) * R_force[P_i] / R_mass[P_i]
""",
    # examples/lj_onetype.py:5-8 (older API: bare rsq / delta symbols, scalar sigma6 / epsilon, force[i] += ...)
    "lj_legacy": """
def K(P_i, P_j):
    L_sr2 = 1.0 / rsq
    L_sr6 = L_sr2 * L_sr2 * L_sr2 * R_sigma6
    R_force[P_i] += delta * 48.0 * L_sr6 * (L_sr6 - 0.5) * L_sr2 * R_epsilon
""",
    # examples/lj_onetype.py:11-13
    "euler_legacy": """
def K(P_i):
    R_velocity[P_i] += R_dt * R_force[P_i] / R_mass[P_i]
    R_position[P_i] += R_dt * R_velocity[P_i]
""",
    # examples/dem.py:6-15 (a setup() function)
    "update_mass_and_inertia": """
def K(P_i):
    R_rotation_matrix[P_i] = diagonal_matrix(1.0)
    R_rotation_quat[P_i] = default_quaternion()

    if is_sphere(P_i):
        R_inv_inertia[P_i] = inversed(diagonal_matrix(0.4 * R_mass[P_i] * R_radius[P_i] * R_radius[P_i]))

    else:
        R_mass[P_i] = R_infinity
        R_inv_inertia[P_i] = 0.0
""",
    # examples/dem.py:88-90
    "gravity": """
def K(P_i):
    L_volume = (4.0 / 3.0) * R_pi * R_radius[P_i] * R_radius[P_i] * R_radius[P_i]
    R_force[P_i][2] = R_force[P_i][2] - (R_density_particle - R_density_fluid) * L_volume * R_gravity
""",
    # examples/dem.py:77-85
    "euler": """
def K(P_i):
    L_inv_mass = 1.0 / R_mass[P_i]
    R_position[P_i] += 0.5 * L_inv_mass * R_force[P_i] * R_dt * R_dt + R_velocity[P_i] * R_dt
    R_velocity[P_i] += L_inv_mass * R_force[P_i] * R_dt
    L_wdot = R_rotation_matrix[P_i] * (R_inv_inertia[P_i] * R_torque[P_i]) * transposed(R_rotation_matrix[P_i])
    L_phi = R_angular_velocity[P_i] * R_dt + 0.5 * L_wdot * R_dt * R_dt
    R_rotation_quat[P_i] = quaternion(L_phi, length(L_phi)) * R_rotation_quat[P_i]
    R_rotation_matrix[P_i] = quaternion_to_rotation_matrix(R_rotation_quat[P_i])
    R_angular_velocity[P_i] += L_wdot * R_dt
""",
    # examples/dem.py:18-74: linear spring-dashpot with tangential history and Coulomb friction
    "linear_spring_dashpot": """
def K(P_i, P_j):
    L_delta = -penetration_depth(P_i, P_j)
    skip_when(L_delta < 0.0)

    L_meff = 1.0 / ((1.0 / R_mass[P_i]) + (1.0 / R_mass[P_j]))
    L_kn = L_meff * (R_pi * R_pi + R_ln_coeff * R_ln_coeff) / (R_collision_time * R_collision_time)
    L_kt = R_kappa * L_kn
    L_dn = -2.0 * L_meff * R_ln_coeff / R_collision_time
    L_dtan = sqrt(R_kappa) * L_dn

    L_vwi = R_velocity[P_i] + cross(R_angular_velocity[P_i], contact_point(P_i, P_j) - R_position[P_i])
    L_vwj = R_velocity[P_j] + cross(R_angular_velocity[P_j], contact_point(P_i, P_j) - R_position[P_j])

    L_rel = -(L_vwi - L_vwj)
    L_reln = dot(L_rel, contact_normal(P_i, P_j)) * contact_normal(P_i, P_j)
    L_relt = L_rel - L_reln
    L_fN = L_kn * L_delta * contact_normal(P_i, P_j) + L_dn * L_reln

    L_tsd = R_tsd[P_i, P_j]
    L_ivm = R_ivm[P_i, P_j]
    L_imp = select(L_ivm > 0.0, L_ivm, length(L_rel))
    L_stick = R_sticking[P_i, P_j]

    L_rot = L_tsd - contact_normal(P_i, P_j) * dot(L_tsd, contact_normal(P_i, P_j))
    L_rot2 = squared_length(L_rot)
    L_ntsd = R_dt * L_relt + select(L_rot2 <= 0.0, zero_vector(), L_rot * sqrt(squared_length(L_tsd) / L_rot2))

    L_fTLS = L_kt * L_ntsd + L_dtan * L_relt
    L_fTLS_len = length(L_fTLS)
    L_t = normalized(L_fTLS)

    L_fs = R_friction_static[P_i, P_j] * length(L_fN)
    L_fd = R_friction_dynamic[P_i, P_j] * length(L_fN)
    L_thr = 1e-8

    L_c1 = L_stick == 1 and length(L_relt) < L_thr and L_fTLS_len < L_fs
    L_c2 = L_stick == 1 and L_fTLS_len < L_fd
    L_fabs = select(L_c1, L_fs, L_fd)
    L_nstick = select(L_c1 or L_c2 or L_fTLS_len < L_fd, 1, 0)
    R_tsd[P_i, P_j] = select(not L_c1 and not L_c2 and L_kt > 0.0, (L_fabs * L_t - L_dtan * L_relt) / L_kt, L_ntsd)

    R_ivm[P_i, P_j] = L_imp
    R_sticking[P_i, P_j] = L_nstick

    L_fTabs = min(L_fTLS_len, L_fabs)
    L_fT = L_fTabs * L_t
    L_pf = L_fN + L_fT

    apply(R_force, L_pf)
    apply(R_torque, cross(contact_point(P_i, P_j) - R_position, L_pf))
""",
}


def _unify(t, u, binding):
    """Structural unification of template AST `t` with user AST `u`."""
    if isinstance(t, ast.Name):
        if not isinstance(u, ast.Name):
            return False
        if t.id[:2] in ("R_", "P_", "L_"):
            if t.id in binding:
                return binding[t.id] == u.id
            if t.id[:2] != "R_" and u.id in binding.values():
                return False
            binding[t.id] = u.id
            return True
        return t.id == u.id
    if type(t) is not type(u):
        return False
    if isinstance(t, ast.Constant):
        return type(t.value) is type(u.value) and t.value == u.value
    if isinstance(t, ast.arg):
        binding.setdefault(t.arg, u.arg)
        return binding[t.arg] == u.arg
    for field in t._fields:
        if field in ("ctx", "lineno", "col_offset", "end_lineno", "end_col_offset", "type_comment", "kind", "decorator_list",
                     "returns", "name", "type_params"):
            continue
        a, b = getattr(t, field, None), getattr(u, field, None)
        if isinstance(a, list):
            if not isinstance(b, list) or len(a) != len(b):
                return False
            if not all(_unify(x, y, binding) for x, y in zip(a, b)):
                return False
        elif isinstance(a, ast.AST):
            if not isinstance(b, ast.AST) or not _unify(a, b, binding):
                return False
        elif a != b:
            return False
    return True


FORCE_GENERIC = False        # tests: send every kernel through the generic (NVRTC) path, also the recognised ones
EXACT_ARITHMETIC = False     # tests: the hand-written pair kernel with the reference's expression tree (no fused multiply-adds, IEEE division:
                             # option "lj_fma" = 0), the arithmetic the generated kernels always use -- for bit-for-bit comparisons
FORCE_GENERIC_NAMES = set()  # tests: kernels with these function names are generated even if they would be recognised
FORCE_GENERIC_CONTACT_MODEL = False      # tests: DEM scripts get their contact model generated even if it is examples/dem.py's


def recognise(func):
    """-> (family, {role: user name}) or raises DslError."""
    src = textwrap.dedent(inspect.getsource(func))
    tree = ast.parse(src).body[0]
    if not isinstance(tree, ast.FunctionDef):
        raise DslError(f"{func.__name__}: not a plain function")
    for family, tsrc in TEMPLATES.items():
        ttree = ast.parse(textwrap.dedent(tsrc)).body[0]
        binding = {}
        if _unify(ttree.args, tree.args, binding) and _unify_body(ttree.body, tree.body, binding):
            return family, {k[2:]: v for k, v in binding.items() if k.startswith("R_")}
    raise DslError(
        f"kernel '{func.__name__}' is not one of the kernel families this backend implements in CUDA "
        f"({', '.join(TEMPLATES)})")


def _unify_body(tb, ub, binding):
    ub = [s for s in ub if not (isinstance(s, ast.Expr) and isinstance(s.value, ast.Constant))]   # docstrings
    return len(tb) == len(ub) and all(_unify(a, b, binding) for a, b in zip(tb, ub))


# ---- the Simulation object ----------------------------------------------------------------------------------------------
_CSV_WIDTH = {Types.Vector: 3, Types.Matrix: 9, Types.Quaternion: 4}      # columns per property, runtime/read_from_file.hpp:48-104


class _Prop:
    def __init__(self, name, ptype, value, volatile):
        self.name, self.type, self.value, self.volatile = name, ptype, value, volatile


def _fmt(x):
    """std::ostream default formatting of a double (precision 6, %g-like) as used by runtime/thermo.hpp:47."""
    return f"{x:.6g}"


class Simulation:
    def __init__(self, ref, shapes=None, dims=3, timesteps=100, double_prec=False, use_contact_history=False,
                 particle_capacity=800000, neighbor_capacity=100, debug=False):
        if dims != 3:
            raise DslError("only 3-D simulations are supported")
        self.ref = ref
        self.shapes = list(shapes) if shapes is not None else [Shapes.PointMass]     # legacy API: shapes optional
        self.ntimesteps = timesteps
        self.double_prec = double_prec
        self.use_contact_history = use_contact_history
        self.particle_capacity = particle_capacity
        self.neighbor_capacity = neighbor_capacity
        self.debug = debug
        self.props = {}
        self.position_name = None
        self.features = {}
        self.feature_props = {}
        self.contact_props = {}
        self._target = None
        self._partitioner = DomainPartitioners.Regular
        self._compute_half = False
        self.vtk_file = None
        self.vtk_frequency = 0
        self.ckpt_file = None
        self.ckpt_frequency = 0
        self._pbc = [True, True, True]
        self.grid = None
        self.reneighbor_frequency = 1            # sim/simulation.py:89
        self._compute_thermo = 0
        self._cell_spacing = None
        self.neighbor_cutoff = None
        self.setups = []
        self.setup_functions = []
        self.pre_step = []
        self.functions = []
        self.vtk_file = None
        self.ctx = None
        self.thermo_log = []
        for n in ("uid", "shape", "flags"):     # implicit properties, sim/simulation.py:63-65
            self.props[n] = _Prop(n, Types.Int32, 0, False)

    # -- configuration (names, defaults and call-order tolerance of the reference) --
    def target(self, t):
        self._target = t

    def set_domain_partitioner(self, p):
        if p not in (DomainPartitioners.Regular, DomainPartitioners.RegularXY):
            raise Exception("Invalid domain partitioner.")
        self._partitioner = p

    def partitioner(self):
        return self._partitioner

    def pbc(self, cfg):
        assert len(cfg) == 3, "PBC must be specified for each dimension."
        self._pbc = list(cfg)

    # -- queries of the reference's Simulation object (sim/simulation.py:116-128, 159-201, 373-377) --
    def enable_profiler(self):
        """sim/simulation.py:116-117 switches on LIKWID marker regions around compute() kernels; here every stage is always
        bracketed by CUDA-event timers (ctx.timer(name)) and, with this call, also by an NVTX range of the same name."""
        self._enable_profiler = True

    def use_double_precision(self):
        return self.double_prec

    def get_shape_id(self, shape):
        return self.shapes[shape]

    def max_shapes(self):
        return len(self.shapes)

    def ndims(self):
        return 3

    def property(self, name):
        return self.props.get(name)

    def position(self):
        return self.props.get(self.position_name)

    def feature(self, name):
        return self.features.get(name)

    def feature_property(self, name):
        return self.feature_props.get(name)

    def contact_property(self, name):
        return self.contact_props.get(name)

    def cell_spacing(self):
        return self._cell_spacing

    def compute_half(self):
        """sim/simulation.py:119-120: half neighbour lists, each pair term applied to both partners (ir/apply.py:111-125)."""
        self._compute_half = True

    def add_property(self, name, ptype, value=0.0, volatile=False):
        assert name not in self.props, f"Property already defined: {name}"
        self.props[name] = _Prop(name, ptype, value, volatile)
        return self.props[name]

    def add_position(self, name, value=(0.0, 0.0, 0.0), volatile=False, layout=None):
        assert name not in self.props, f"Property already defined: {name}"
        self.position_name = name
        self.props[name] = _Prop(name, Types.Vector, value, volatile)
        return self.props[name]

    def add_feature(self, name, nkinds):
        assert name not in self.features, f"Feature already defined: {name}"
        self.features[name] = nkinds
        self.props.setdefault(name, _Prop(name, Types.Int32, 0, False))     # e.g. 'type': one int per particle

    def add_feature_property(self, feature, name, ptype, data):
        assert feature in self.features, f"Feature not found: {feature}"
        nk = self.features[feature]
        assert len(data) == nk * nk
        self.feature_props[name] = (feature, [_builtin_float(x) for x in data])

    def add_contact_property(self, name, ptype, default, layout=None):
        self.contact_props[name] = (ptype, default)

    # legacy API used by examples/lj_onetype.py (absent from the reference)
    def add_real_property(self, name, value=0.0, vol=False):
        return self.add_property(name, Types.Real, value, vol)

    def add_vector_property(self, name, value=(0.0, 0.0, 0.0), vol=False):
        return self.add_property(name, Types.Vector, value, vol)

    def set_domain(self, grid):
        self.grid = [_builtin_float(g) for g in grid]       # xmin, ymin, zmin, xmax, ymax, zmax (sim/simulation.py:228)

    def reneighbor_every(self, n):
        self.reneighbor_frequency = n

    def compute_thermo(self, every=0):
        self._compute_thermo = every

    def vtk_output(self, filename, frequency=0):
        """sim/simulation.py:365-367: <filename>_local_<ts>.vtk and <filename>_ghost_<ts>.vtk after every iteration whose
        number is a multiple of `frequency` (every iteration for 0), written by vtk_write (runtime/vtk.hpp:11-86)."""
        self.vtk_file = filename
        self.vtk_frequency = frequency

    def checkpoint_output(self, filename, frequency=0):
        """Checkpoint dump of the per-uid state (the reference reads such files -- runtime/read_from_file.hpp:33-115
        -- but cannot write them): after every iteration whose number is a multiple of `frequency` each rank writes
          <filename>_<ts>[_r<rank>].csv    one row per local particle in the reference's CSV layout (Appendix A.3): every non-volatile
                                           property in the column order the manifest lists (vectors 3 columns, matrices 9,
                                           quaternions 4), 17 significant digits, i.e. every double survives the round trip
          <filename>_<ts>[_r<rank>].contacts.csv   DEM: uid_i, uid_j, is_sticking, tangential_spring_displacement x 3,
                                           impact_velocity_magnitude (and the lanes of any further contact property) for
                                           every live contact
          <filename>_<ts>.json             the manifest (columns, iteration, ranks, domain)
        read_checkpoint(filename, ts) in a later script continues from it, on any number of ranks."""
        self.ckpt_file = filename
        self.ckpt_frequency = frequency

    def read_checkpoint(self, filename, ts):
        self.setups.append(("read_checkpoint", (filename, ts)))

    def _ckpt_due(self, ts):
        return self.ckpt_file is not None and (self.ckpt_frequency == 0 or ts % self.ckpt_frequency == 0)

    def _vtk_only_due(self, ts):
        return self.vtk_file is not None and (self.vtk_frequency == 0 or ts % self.vtk_frequency == 0)

    def _dumps(self):
        return self.vtk_file is not None or self.ckpt_file is not None

    def _vtk_due(self, ts):          # "something is written after iteration ts": VTK files and / or a checkpoint
        return self._vtk_only_due(ts) or self._ckpt_due(ts)

    _CKPT_DEM = (("angvel", "angular_velocity"), ("radius", "radius"), ("inv_inertia", "inv_inertia"), ("rotmat", "rotation_matrix"),
                 ("quat", "rotation_quat"))

    def _checkpoint_columns(self, ctx, dem):
        """-> [(property name, host array of the locals)] of everything that is not volatile, in a fixed order"""
        cols = [("uid", ctx.ints("uid")), ("type", ctx.ints("type")), ("flags", ctx.ints("flags")), ("shape", ctx.ints("shape"))]
        storage = self._dem_storage() if dem else self._device_storage()
        nl = ctx.counts()[0]
        for name, st in storage.items():
            if name in ("uid", "flags", "shape") or name in self.features:
                continue
            if name in self.props and self.props[name].volatile:
                continue
            if st == "pos":
                cols.append((name, ctx.real("position")))
            elif st == "vel":
                cols.append((name, ctx.real("linear_velocity")))
            elif st == "mass":
                cols.append((name, ctx.real("mass")))
            elif isinstance(st, tuple):
                cols.append((name, ctx.download_property(name)))
            elif dem and st in dict(self._CKPT_DEM):
                cols.append((name, ctx.dem_download(dict(self._CKPT_DEM)[st], nl)))
        if dem and "normal" in self.props:
            cols.append(("normal", ctx.dem_download("normal", nl)))
        return cols

    def _checkpoint_write(self, ctx, ts, rank, world):
        import json
        dem = bool(getattr(ctx, "contact_capacity", 0))
        cols = self._checkpoint_columns(ctx, dem)
        infix = f"_r{rank}" if world > 1 else ""
        table = np.column_stack([np.asarray(a, np.float64).reshape(len(a), -1) for _, a in cols]) if cols[0][1].size else np.zeros((0, 1))
        np.savetxt(f"{self.ckpt_file}_{ts}{infix}.csv", table, delimiter=",", fmt="%.17g")
        if dem:
            c = ctx.dem_download_contacts(ctx.counts()[0])
            cx = ctx.dem_download_contact_extras(ctx.counts()[0])         # further contact properties: columns after the seven
            uid = ctx.ints("uid")
            rows = []
            for i in np.nonzero(c["num_contacts"])[0]:
                for k in range(int(c["num_contacts"][i])):
                    rows.append([uid[i], c["contact_lists"][i, k], c["is_sticking"][i, k], *c["tangential_spring_displacement"][i, k],
                                 c["impact_velocity_magnitude"][i, k], *cx[i, k]])
            np.savetxt(f"{self.ckpt_file}_{ts}{infix}.contacts.csv", np.asarray(rows, np.float64).reshape(len(rows), 7 + cx.shape[2]),
                       delimiter=",", fmt="%.17g")
        if rank == 0:
            manifest = {"iteration": ts, "ranks": world, "dem": dem, "domain": list(self.grid),
                        "columns": [[n, int(np.asarray(a).reshape(len(a), -1).shape[1]) if len(a) else 1, "int" if np.asarray(a).dtype.kind == "i" else "real"]
                                    for n, a in cols]}
            with open(f"{self.ckpt_file}_{ts}.json", "w") as f:
                json.dump(manifest, f)

    def _checkpoint_read(self, ctx, filename, ts):
        """-> dict of host arrays (this rank's rows) + 'contacts' (DEM) of the checkpoint written after iteration ts"""
        import json
        with open(f"{filename}_{ts}.json") as f:
            man = json.load(f)
        files = [f"{filename}_{ts}"] if man["ranks"] == 1 else [f"{filename}_{ts}_r{r}" for r in range(man["ranks"])]
        tables = [np.loadtxt(fn + ".csv", delimiter=",", ndmin=2) for fn in files]
        tables = [t for t in tables if t.size]
        data = np.concatenate(tables) if tables else np.zeros((0, sum(w for _, w, _ in man["columns"])))
        part, k = {}, 0
        for name, w, kind in man["columns"]:
            col = data[:, k:k + w] if w > 1 else data[:, k]
            part[name] = col.astype(np.int32) if kind == "int" else np.ascontiguousarray(col)
            k += w
        if self.position_name not in part:
            raise DslError(f"checkpoint {filename}_{ts}: no column for the position property '{self.position_name}'")
        part["position"] = part[self.position_name]
        part = self._keep_own(ctx, part)
        if man["dem"]:
            rows = [np.loadtxt(fn + ".contacts.csv", delimiter=",", ndmin=2) for fn in files if os.path.exists(fn + ".contacts.csv")]
            rows = [r for r in rows if r.size]
            part["contacts"] = np.concatenate(rows) if rows else np.zeros((0, 7))
        return part

    def _vtk_write(self, ctx, ts, rank, world):
        if self._ckpt_due(ts):
            self._checkpoint_write(ctx, ts, rank, world)
        if not self._vtk_only_due(ts):
            return
        nl, ng = ctx.counts()
        pos, mass, flags = ctx.real("position", True), ctx.real("mass", True), ctx.ints("flags", True)
        infix = f"r{rank}_" if world > 1 else ""
        vtk_write(f"{self.vtk_file}_local_{infix}{ts}.vtk", pos[:nl], mass[:nl], flags[:nl])
        vtk_write(f"{self.vtk_file}_ghost_{infix}{ts}.vtk", pos[nl:nl + ng], mass[nl:nl + ng], flags[nl:nl + ng])

    def copper_fcc_lattice(self, nx, ny, nz, rho, temperature, ntypes):
        lattice = pow((4.0 / rho), (1.0 / 3.0))    # sim/copper_fcc_lattice.py:19-23
        self.set_domain([0.0, 0.0, 0.0, nx * lattice, ny * lattice, nz * lattice])
        self.setups.append(("copper_fcc_lattice", (nx, ny, nz, rho, temperature, ntypes)))

    def read_particle_data(self, filename, prop_names, shape_id):
        self.setups.append(("read_particle_data", (filename, list(prop_names), shape_id)))

    def from_file(self, filename, prop_names):      # legacy API of examples/lj_onetype.py
        """miniMD set-up files `minimd_setup_<nx>x<ny>x<nz>*.input` (rows: mass, position, velocity).  The box is not stored in
        the file: it is nx * a with the FCC lattice constant a = (4 / 0.8442)^(1/3) of the generator that wrote those files.  If
        the file is absent (the 32x32x32 one is not shipped, .MISSING_LARGE_BLOBS) the same system is generated synthetically."""
        import re
        m = re.search(r"(\d+)x(\d+)x(\d+)", os.path.basename(filename))
        if m is None:
            raise DslError("from_file: cannot derive the box (expected a name like minimd_setup_32x32x32.input)")
        nx, ny, nz = (int(g) for g in m.groups())
        lattice = pow((4.0 / 0.8442), (1.0 / 3.0))
        self.set_domain([0.0, 0.0, 0.0, nx * lattice, ny * lattice, nz * lattice])
        if os.path.exists(filename):
            self.read_particle_data(filename, prop_names, Shapes.PointMass)
        else:
            print(f"from_file: {filename} not found -- generating the same FCC system synthetically ({4 * nx * ny * nz} atoms)")
            self.setups.append(("copper_fcc_lattice", (nx, ny, nz, 0.8442, 1.44, 1)))

    def dem_sc_grid(self, xmax, ymax, zmax, spacing, diameter, min_diameter, max_diameter, initial_velocity, particle_density, ntypes):
        self.setups.append(("dem_sc_grid", (xmax, ymax, zmax, spacing, diameter, min_diameter, max_diameter, initial_velocity,
                                            particle_density, ntypes)))

    def setup(self, func, symbols={}):
        """sim/simulation.py:266-267: a per-particle function run once over the locals after the set-up statements."""
        try:
            if func.__name__ in FORCE_GENERIC_NAMES:
                raise DslError("generic path forced")
            family, roles = recognise(func)
            if family != "update_mass_and_inertia":
                raise DslError("not a set-up family")
        except DslError:
            if len(inspect.signature(func).parameters) != 1:
                raise DslError(f"setup(): '{func.__name__}' must take one particle argument") from None
            family, roles = "generic_setup", {}       # generic path: kernelgen -> NVRTC, no FIXED filter
        self.setup_functions.append({"name": func.__name__, "family": family, "roles": roles, "symbols": dict(symbols),
                                     "globals": func.__globals__, "func": func})

    def build_cell_lists(self, spacing, store_neighbors_per_cell=False):
        # store_neighbors_per_cell (sim/cell_lists.py:174-206): the reference then walks one pre-concatenated list per cell
        # -- cell 0, then the stencil cells in stencil order, i.e. the SAME candidate sequence as its direct stencil walk
        # (its generated dem.cpp gives identical bits either way, tests/test_oracle_pin.py).  The CUDA kernels traverse the
        # CSR cell lists directly in that order, so there is nothing to materialise: the flag is accepted and recorded.
        self._store_neighbors_per_cell = bool(store_neighbors_per_cell)
        self._cell_spacing = spacing

    def build_neighbor_lists(self, spacing):
        assert not getattr(self, "_store_neighbors_per_cell", False), \
            "Using neighbor-lists with store_neighbors_per_cell option is invalid."      # sim/simulation.py:256-257
        self._cell_spacing = spacing                 # sim/simulation.py:255-261: cells and lists share the spacing
        self.neighbor_cutoff = spacing

    def compute(self, func, cutoff_radius=None, symbols={}, pre_step=False, skip_first=False):
        try:
            if FORCE_GENERIC or func.__name__ in FORCE_GENERIC_NAMES or (FORCE_GENERIC_CONTACT_MODEL and self.use_contact_history and len(inspect.signature(func).parameters) == 2):
                raise DslError("generic path forced")
            family, roles = recognise(func)
        except DslError:
            # not one of the hand-written families: CUDA is generated for the body and compiled at run time (kernelgen.py,
            # csrc/jit.cu) -- the reference's code generator handles arbitrary bodies too (mapping/funcs.py:39-334)
            nparams = len(inspect.signature(func).parameters)
            if nparams not in (1, 2):
                raise
            family, roles = ("generic_pair" if nparams == 2 else "generic_particle"), {}
        entry = {"name": func.__name__, "family": family, "roles": roles, "cutoff": cutoff_radius, "symbols": dict(symbols),
                 "skip_first": skip_first, "globals": func.__globals__, "func": func}
        (self.pre_step if pre_step else self.functions).append(entry)

    # -- helpers --
    def _symbol(self, entry, role):
        name = entry["roles"][role]
        if name in entry["symbols"]:
            return _builtin_float(entry["symbols"][name])
        if name in entry["globals"] and isinstance(entry["globals"][name], (int, float)):
            return _builtin_float(entry["globals"][name])
        raise DslError(f"{entry['name']}: symbol '{name}' has no value (pass it in symbols={{...}})")

    def _check_prop(self, entry, role, expected):
        name = entry["roles"][role]
        if name not in self.props and name not in self.feature_props:
            raise DslError(f"{entry['name']}: '{name}' is not a registered property")
        return name

    def _roles_fit_the_storage(self, e):
        """A recognised kernel runs as hand-written CUDA on FIXED arrays.  Recognition is structural (names are free), so the roles
        it bound must be exactly the properties those arrays hold -- e.g. a body that integrates some other vector property with the
        text of initial_integrate, or that multiplies torque * inv_inertia instead of inv_inertia * torque, matches the template but
        is a different computation.  Such kernels are generated instead."""
        roles = e["roles"]
        if self.use_contact_history:
            fixed = {"position": self.position_name, "velocity": "linear_velocity", "angular_velocity": "angular_velocity", "force": "force",
                     "torque": "torque", "mass": "mass", "radius": "radius", "inv_inertia": "inv_inertia", "rotation_matrix": "rotation_matrix",
                     "rotation_quat": "rotation_quat"}
            types = {"position": Types.Vector, "velocity": Types.Vector, "angular_velocity": Types.Vector, "force": Types.Vector,
                     "torque": Types.Vector, "mass": Types.Real, "radius": Types.Real, "inv_inertia": Types.Matrix,
                     "rotation_matrix": Types.Matrix, "rotation_quat": Types.Quaternion}
            contact = {"sticking": Types.Int32, "tsd": Types.Vector, "ivm": Types.Real}
            for role, name in roles.items():
                if role in fixed:
                    if name != fixed[role] or name not in self.props or self.props[name].type != types[role]:
                        return False
                elif role in contact:
                    if name not in self.contact_props or self.contact_props[name][0] != contact[role]:
                        return False
                elif role in ("friction_static", "friction_dynamic"):
                    if name not in self.feature_props:
                        return False
                elif name in self.props or name in self.contact_props or name in self.feature_props:
                    return False                      # a symbol role (dt, pi, kappa, ...) bound to a property
            return True
        storage = self._device_storage()
        slots = {"position": "pos", "velocity": "vel", "force": "force", "mass": "mass"}
        for role, name in roles.items():
            if role in slots:
                if storage.get(name) != slots[role]:
                    return False
            elif role in ("epsilon", "sigma6") and e["family"] == "lennard_jones":
                if name not in self.feature_props:
                    return False
            elif name in self.props or name in self.feature_props:
                return False
        return True

    def _reclassify(self):
        """Recognised kernels whose roles do not fit the fixed arrays go through the generic path (see _roles_fit_the_storage)."""
        for group, generic in ((self.pre_step, None), (self.functions, None), (self.setup_functions, "generic_setup")):
            for e in group:
                if e["family"].startswith("generic") or self._roles_fit_the_storage(e):
                    continue
                nparams = len(inspect.signature(e["func"]).parameters)
                e["family"], e["roles"] = generic or ("generic_pair" if nparams == 2 else "generic_particle"), {}

    # -- generate(): plan + run on the GPU(s) --
    def generate(self):
        assert self._target is not None, "Target not specified!"
        self._reclassify()
        if not self._target.is_gpu():
            raise DslError("this backend executes on B200 GPUs only: use pairs.target_gpu() (there is no CPU path; the "
                           "reference's own CPU build is available as the parity oracle under oracle/)")
        # real_t is double throughout the reference runtime (runtime/pairs_common.hpp:7, MPI_DOUBLE hard-coded): all
        # arithmetic here is fp64 whatever double_prec says
        from . import backend
        rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
        local = int(os.environ.get("LOCAL_RANK", 0))
        if self.grid is None:
            raise DslError("no domain: call set_domain() or copper_fcc_lattice()")
        g = self.grid
        grid = [g[0], g[3], g[1], g[4], g[2], g[5]]
        ctx = backend.Context(local)
        if EXACT_ARITHMETIC:
            ctx.set_option("lj_fma", 0)
        self.ctx = ctx
        ctx.init_domain(grid, self._pbc, self._partitioner, world, rank)
        if getattr(self, "_enable_profiler", False):
            ctx.set_option("profiler", 1)
        if world > 1:
            ctx.nccl_init(_broadcast_nccl_id(backend, rank, world))
        if self._cell_spacing is None:
            raise DslError("build_cell_lists() / build_neighbor_lists() was not called")
        families = [e["family"] for e in self.pre_step + self.functions]
        if "linear_spring_dashpot" in families or (self.use_contact_history and "generic_pair" in families):
            if self._compute_half:
                raise DslError("compute_half() is implemented for the neighbour-list (md.py) path only")
            return self._generate_dem(ctx, rank, world)
        if self._compute_half:
            if "lj_legacy" in families:
                raise DslError("compute_half() is not available for the legacy lj kernel")
            ctx.set_option("compute_half", 1)
        ctx.reserve(0, self.neighbor_capacity)
        for k, (name, comps, volatile, dflt) in enumerate(self._user_props()):
            _, row0 = ctx.add_property(name, comps, volatile, dflt)
            assert row0 == self._device_storage()[name][1]

        # ---- set-up ----
        nlocal = 0
        for kind, args in self.setups:
            if kind == "copper_fcc_lattice":
                nx, ny, nz, rho, temp, ntypes = args
                nlocal = ctx.copper_fcc_lattice(nx, ny, nz, rho, ntypes)
                ctx.adjust_thermo(temp)
            elif kind == "read_particle_data":
                nlocal = self._read_particle_data(ctx, *args)
            elif kind == "read_checkpoint":
                part = self._checkpoint_read(ctx, *args)
                storage = self._device_storage()
                vel_name = next((n for n in part if storage.get(n) == "vel"), None)
                mass_name = next((n for n in part if storage.get(n) == "mass"), None)
                ctx.upload(part["position"], part.get(vel_name), part.get(mass_name), part.get("type"), part.get("flags"), part.get("uid"),
                           part.get("shape"))
                for name, st in storage.items():
                    if isinstance(st, tuple) and name in part:
                        ctx.upload_property(name, part[name])
                nlocal = len(part["position"])
        ctx.setup_cells(self._cell_spacing)
        for e in self.setup_functions:          # setup() functions: once over the locals, right after the set-up statements
            if e["family"] != "generic_setup":
                raise DslError(f"setup(): '{e['name']}' belongs to the DEM path")
            self._bind_generic(ctx, e, skip_fixed=False)["call"]()

        # ---- bind kernels ----
        plan_pre, plan_fn = [self._bind(ctx, e) for e in self.pre_step], [self._bind(ctx, e) for e in self.functions]
        native = self._native_md_params(plan_pre, plan_fn)

        # ---- timestep loop (sim/timestep.py:9-72) ----
        ctx.timers_enable(True)
        ctx.sync()
        t0 = time.perf_counter()
        nsteps = self.ntimesteps + 1
        if native is not None:
            # one native call per stretch between two VTK dumps (chunked calls are bit-identical to one call)
            cuts = [ts + 1 for ts in range(nsteps) if self._vtk_due(ts)] if self._dumps() else []
            begin = 0
            for end in cuts + ([nsteps] if not cuts or cuts[-1] != nsteps else []):
                th = ctx.md_run(begin, end, *native)
                for row in th:
                    self._thermo_line(rank, int(row[0]), row[1], row[2])
                if self._vtk_due(end - 1):
                    self._vtk_write(ctx, end - 1, rank, world)
                begin = end
        else:
            self._python_loop(ctx, plan_pre, plan_fn, nsteps, rank)
        ctx.sync()
        all_ms = (time.perf_counter() - t0) * 1e3
        self._print_summary(ctx, all_ms, rank)
        return ctx

    # -- DEM (examples/dem.py): spheres + half-spaces, contact history, cell-list traversal, reneighbouring every step --
    def _generate_dem(self, ctx, rank, world):
        fams = [e["family"] for e in self.functions]
        pair_at = [k for k, f in enumerate(fams) if f in ("linear_spring_dashpot", "generic_pair")]
        if self.pre_step or len(pair_at) != 1 or any(f not in ("gravity", "euler", "generic_particle", "linear_spring_dashpot", "generic_pair")
                                                     for f in fams) \
                or fams.count("gravity") > 1 or fams.count("euler") > 1 \
                or not self.use_contact_history or self.neighbor_cutoff is not None:
            raise DslError("DEM: a procedure list like examples/dem.py's is implemented -- per-particle kernels (gravity, euler or any "
                           "other body) and ONE contact model (any body) over cell lists with contact history")
        lsd = self.functions[pair_at[0]]
        grav = next((e for e in self.functions if e["family"] == "gravity"), None)
        eul = next((e for e in self.functions if e["family"] == "euler"), None)
        # the native loop (pb_dem_run) is the generated loop of examples/dem.py: the standard list, reneighbouring every iteration
        standard = fams in (["gravity", "linear_spring_dashpot", "euler"], ["gravity", "generic_pair", "euler"]) and self.reneighbor_frequency == 1
        ctx.dem_enable(self.neighbor_capacity, self._contact_layout()[2])
        for name, comps, volatile, dflt in self._dem_user_props():
            _, row0 = ctx.add_property(name, comps, volatile, dflt)
            assert row0 == self._dem_storage()[name][1]
        dt = self._symbol(eul, "dt") if eul is not None else 0.0
        g_pi = self._symbol(grav, "pi") if grav is not None else math.pi
        g_par = [self._symbol(grav, r) for r in ("density_particle", "density_fluid", "gravity")] if grav is not None else [0.0, 0.0, 0.0]
        if lsd["family"] == "generic_pair":
            # a contact model other than examples/dem.py's: CUDA is generated for the body and the contact kernel is compiled
            # around it at run time (kernelgen.translate_dem_model, csrc/jit.cu pb_jit_set_dem_model); feature properties and
            # symbols become literals of the generated function, so the built-in model's parameters are not needed
            name, src, nk = self._translate_dem_model(lsd)
            ctx.jit_set_dem_model(src, name)
            ctx.dem_set_params(dt, g_pi, 0.0, 0.0, 1.0, *g_par, nk, [0.0] * (nk * nk), [0.0] * (nk * nk))
        else:
            fs_name, fd_name = lsd["roles"]["friction_static"], lsd["roles"]["friction_dynamic"]
            for nme in (fs_name, fd_name):
                if nme not in self.feature_props:
                    raise DslError(f"{lsd['name']}: '{nme}' must be a feature property")
            nk = self.features[self.feature_props[fs_name][0]]
            if eul is not None and dt != self._symbol(lsd, "dt"):
                raise DslError("DEM: contact kernel and integrator use different dt")
            ctx.dem_set_params(self._symbol(lsd, "dt"), self._symbol(lsd, "pi"), self._symbol(lsd, "kappa"), self._symbol(lsd, "ln_coeff"),
                               self._symbol(lsd, "collision_time"), *g_par, nk, self.feature_props[fs_name][1], self.feature_props[fd_name][1])
        # ---- set-up: particles are appended in the order of the setup statements (sim/simulation.py:238-247) ----
        parts = []
        for kind, args in self.setups:
            if kind == "dem_sc_grid":
                g = ctx.dem_sc_grid(*args)
                g["shape"] = np.zeros(len(g["uid"]), np.int32)
                g["flags"] = np.zeros(len(g["uid"]), np.int32)
                parts.append(g)
                if rank == 0:
                    self._dem_banner(args, len(g["uid"]))
            elif kind == "read_particle_data":
                parts.append(self._keep_own(ctx, self._read_csv(*args)))
            elif kind == "read_checkpoint":
                if parts:
                    raise DslError("DEM: read_checkpoint() must be the only set-up statement (it restores every particle)")
                parts.append(self._checkpoint_read(ctx, *args))
            else:
                raise DslError(f"DEM: unsupported set-up statement {kind}")
        n = sum(len(p["position"]) for p in parts)

        def cat(name, width, dtype, default=0):
            out = np.full((n, width) if width > 1 else n, default, dtype)
            k = 0
            for p in parts:
                m = len(p["position"])
                if name in p:
                    out[k:k + m] = p[name]
                k += m
            return out
        # untouched slots are zero in the reference (add_property defaults are not applied at run time)
        ctx.upload(cat("position", 3, np.float64), cat("linear_velocity", 3, np.float64), cat("mass", 1, np.float64),
                   cat("type", 1, np.int32), cat("flags", 1, np.int32), cat("uid", 1, np.int32), cat("shape", 1, np.int32))
        ctx.dem_upload("radius", cat("radius", 1, np.float64))
        ctx.dem_upload("normal", cat("normal", 3, np.float64))
        restored = parts[0] if len(parts) == 1 and "contacts" in parts[0] else None
        if restored is not None:            # a checkpoint: the rigid-body state and the contact history come back as well
            for _, canon in self._CKPT_DEM:
                if canon in restored and canon != "radius":
                    ctx.dem_upload(canon, restored[canon])
            for name, st in self._dem_storage().items():
                if isinstance(st, tuple) and name in restored:
                    ctx.upload_property(name, restored[name])
            C = ctx.contact_capacity
            num = np.zeros(n, np.int32)
            cuid, stick = np.zeros((n, C), np.int32), np.zeros((n, C), np.int32)
            tsd, ivm = np.zeros((n, C, 3)), np.zeros((n, C))
            cx = np.zeros((n, C, ctx.contact_extra_lanes))
            row_of = {int(u): k for k, u in enumerate(restored["uid"])}
            for r in restored["contacts"]:
                i = row_of.get(int(r[0]))
                if i is None:            # the owner of this contact row lives on another rank now
                    continue
                k = num[i]
                if k >= C:
                    raise DslError("read_checkpoint: more contacts per particle than the contact capacity (neighbor_capacity)")
                cuid[i, k], stick[i, k], tsd[i, k], ivm[i, k] = int(r[1]), int(r[2]), r[3:6], r[6]
                cx[i, k] = r[7:7 + cx.shape[2]]
                num[i] = k + 1
            ctx.dem_upload_contacts(num, cuid, stick, tsd, ivm)
            if cx.size:
                ctx.dem_upload_contact_extras(cx)
        for f in self.setup_functions:
            if f["family"] == "update_mass_and_inertia":
                ctx.dem_stage("update_mass_and_inertia")
                continue
            from . import backend, kernelgen              # any other set-up body: generated, run once over all locals
            try:
                _, kname, src = kernelgen.translate(f["func"], self._dem_storage(), {}, 1, f["symbols"], backend.jit_prelude(), skip_fixed=False)
            except kernelgen.KernelGenError as err:
                raise DslError(f"setup() function '{f['name']}': {err}") from None
            ctx.jit_launch(ctx.jit_compile(src, kname), 1)
        ctx.timers_enable(True)
        ctx.sync()
        t0 = time.perf_counter()
        nsteps = self.ntimesteps + 1
        if standard:
            cuts = [ts + 1 for ts in range(nsteps) if self._vtk_due(ts)] if self._dumps() else []
            begin = 0
            for end in cuts + ([nsteps] if not cuts or cuts[-1] != nsteps else []):
                ctx.dem_run(self._cell_spacing, begin, end)
                if self._vtk_due(end - 1):
                    self._vtk_write(ctx, end - 1, rank, world)
                begin = end
        else:
            self._dem_staged_loop(ctx, nsteps, rank, world)
        ctx.sync()
        all_ms = (time.perf_counter() - t0) * 1e3
        self._print_summary(ctx, all_ms, rank)
        return ctx

    def _dem_storage(self):
        """Property names of a DEM script -> device arrays for generated per-particle kernels (the property set of examples/dem.py,
        by its names)."""
        storage = {self.position_name: "pos"}
        for name, slot in (("linear_velocity", "vel"), ("angular_velocity", "angvel"), ("mass", "mass"), ("radius", "radius"),
                           ("force", "force"), ("torque", "torque"), ("uid", "uid"), ("shape", "shape"), ("flags", "flags"),
                           ("inv_inertia", "inv_inertia"), ("rotation_matrix", "rotmat"), ("rotation_quat", "quat")):
            if name in self.props:
                storage[name] = slot
        for name in self.features:
            storage[name] = "type"
        # whatever else the script declares (reals, vectors, integers): rows of the user-property block, as on the md.py path
        row = 0
        for name, p in self.props.items():
            if name in storage or name == "normal" or name in self.feature_props:
                continue
            if p.type in (Types.Real, Types.Vector):
                comps = 3 if p.type == Types.Vector else 1
                storage[name] = ("x", row, comps)
                row += comps
            elif p.type == Types.Int32:
                storage[name] = ("x", row, 1, "i")
                row += 1
        return storage

    def _dem_user_props(self):
        out = []
        for name, st in self._dem_storage().items():
            if isinstance(st, tuple):
                p = self.props[name]
                v = p.value if isinstance(p.value, (list, tuple)) else [p.value] * st[2]
                out.append((name, st[2], p.volatile, [_builtin_float(x) for x in v]))
        return out

    def _dem_staged_loop(self, ctx, nsteps, rank, world):
        """A DEM procedure list with further per-particle kernels (or without gravity / euler): the modules of the generated
        loop one by one (sim/simulation.py:387-417 with contact history), user bodies through the generic path."""
        from . import backend, kernelgen
        storage = self._dem_storage()
        calls = []
        for e in self.functions:
            if e["family"] in ("gravity", "euler", "linear_spring_dashpot", "generic_pair"):
                stage = "linear_spring_dashpot" if e["family"] in ("linear_spring_dashpot", "generic_pair") else e["family"]
                calls.append(lambda stage=stage: ctx.dem_stage(stage))
                continue
            try:
                kind, kname, src = kernelgen.translate(e["func"], storage, {}, 1, e["symbols"], backend.jit_prelude())
            except kernelgen.KernelGenError as err:
                raise DslError(f"kernel '{e['name']}': {err}") from None
            handle = ctx.jit_compile(src, kname)
            calls.append(lambda handle=handle: ctx.jit_launch(handle, 1))
        ctx.setup_cells(self._cell_spacing)
        for ts in range(nsteps):
            if ((ts + 1) % self.reneighbor_frequency) == 0 or ts == 0:      # sim/timestep.py:36-61, as on the md.py path
                ctx.exchange()
                ctx.borders()
                ctx.build_cell_lists()
            else:
                ctx.synchronize()            # positions, linear and angular velocities of the ghosts (sim/comm.py:45-54)
            ctx.dem_stage("reset_contact_usage")
            ctx.reset_volatile()
            for e, call in zip(self.functions, calls):
                if ts > 0 or not e["skip_first"]:
                    call()
            ctx.dem_stage("clear_unused_contacts")
            ctx.dem_check_contacts()          # contact rows nearly full: the capacity grows ahead of need (csrc/dem_kernels.cu)
            if self._vtk_due(ts):
                self._vtk_write(ctx, ts, rank, world)

    def _contact_layout(self):
        """Where each contact property lives: the first integer, vector and real ones take the three columns examples/dem.py declares
        (is_sticking, tangential_spring_displacement, impact_velocity_magnitude); any further one -- the reference gives every
        add_contact_property() its own array, mapping/funcs.py:230-263 -- takes lanes of the contact's extra doubles
        (pb_dem_enable_ex).  -> ({name: kind}, {kind: default} of the three, [default per extra lane])."""
        contact, defaults, extra = {}, {}, []
        for name, (ptype, default) in self.contact_props.items():
            kind = {Types.Int32: "c_stick", Types.Vector: "c_tsd", Types.Real: "c_ivm"}.get(ptype)
            if kind is None:
                raise DslError(f"contact property '{name}': contact properties are integers, reals or vectors")
            if kind in contact.values():
                ctype, width = {"c_stick": ("int", 1), "c_tsd": ("vec", 3), "c_ivm": ("real", 1)}[kind]
                if len(extra) + width > 16:
                    raise DslError(f"contact property '{name}': at most 16 doubles per contact beyond the first integer, vector and real")
                contact[name] = f"cx:{ctype}:{len(extra)}"
                d = list(default) if isinstance(default, (list, tuple)) else [default] * width
                extra += [np.float64(int(x) if ctype == "int" else x).item() for x in d[:width]]        # (float is the DSL type here)
                continue
            contact[name] = kind
            defaults[kind] = default
        return contact, defaults, extra

    def _translate_dem_model(self, e):
        """-> (function name, CUDA source, number of types) of a user-defined contact model (kernelgen.translate_dem_model).  The
        kernel hands a model the DEM property set of examples/dem.py by these names; contact properties are told apart by type."""
        from . import kernelgen
        storage = {self.position_name: "pos"}
        for name, slot in (("linear_velocity", "vel"), ("angular_velocity", "angvel"), ("mass", "mass"), ("radius", "radius"),
                           ("force", "force"), ("torque", "torque")):
            if name in self.props:
                storage[name] = slot
        for name in self.props:
            storage.setdefault(name, name)               # anything else: named in the error message if the body touches it
        contact, defaults, extra_defaults = self._contact_layout()
        nk, tables = 1, {}
        for name, (feat, data) in self.feature_props.items():
            tables[name] = data
            nk = self.features[feat]
        try:
            name, src = kernelgen.translate_dem_model(e["func"], storage, contact, tables, e["symbols"], contact_defaults=defaults,
                                                      extra_lanes=len(extra_defaults))
        except kernelgen.KernelGenError as err:
            raise DslError(f"contact model '{e['name']}': {err}") from None
        return name, src, nk

    _FLAGS_KEEP = 1 | 4 | 8      # infinite | fixed | global (runtime/pairs_common.hpp flags; PB_FLAG_* in include/pairs_b200.h)

    @staticmethod
    def _keep_own(ctx, part):
        """A rank keeps the rows inside its subdomain plus every infinite / fixed / global body (runtime/read_from_file.hpp:44-106;
        the subdomain test is the partitioner's isWithinSubdomain, runtime/domain/regular_6d_stencil.cpp)."""
        dec = ctx.decomposition()
        if int(np.prod(dec["nranks"])) == 1:
            return part
        sd, x = dec["subdom"], np.asarray(part["position"], np.float64)
        keep = np.ones(len(x), bool)
        for d in range(3):
            keep &= (x[:, d] >= sd[2 * d]) & (x[:, d] < sd[2 * d + 1] - 0.00001)
        if "flags" in part:
            keep |= (np.asarray(part["flags"]) & Simulation._FLAGS_KEEP) != 0
        return {k: (np.asarray(v)[keep] if len(np.atleast_1d(v)) == len(x) else v) for k, v in part.items()}

    def _dem_banner(self, a, count):
        # runtime/dem_sc_grid.hpp:159-169
        print("DEM Simple-Cubic Grid")
        print(f"Domain size: <{_fmt(a[0])}, {_fmt(a[1])}, {_fmt(a[2])}>")
        print(f"Spacing: {_fmt(a[3])}")
        print(f"Diameter: {_fmt(a[4])} (min = {_fmt(a[5])}, max = {_fmt(a[6])})")
        print(f"Initial velocity: {_fmt(a[7])}")
        print(f"Particle density: {_fmt(a[8])}")
        print(f"Number of types: {a[9]}")
        print(f"Number of particles: {count}")

    def _read_csv(self, filename, prop_names, shape_id):
        """runtime/read_from_file.hpp:33-115 -> dict of host arrays (vectors = 3 columns, matrices 9, quaternions 4, in the order of
        prop_names)."""
        path = filename
        if not os.path.exists(path):
            alt = os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])), "..", filename)
            if not os.path.exists(alt):
                raise DslError(f"read_particle_data: {filename} not found")
            path = alt
        data = np.loadtxt(path, delimiter=",", ndmin=2)
        out, k = {}, 0
        for nme in prop_names:
            w = _CSV_WIDTH.get(self.props[nme].type, 1)
            col = data[:, k:k + w] if w > 1 else data[:, k]
            out[nme] = col.astype(np.int32) if self.props[nme].type == Types.Int32 else col
            k += w
        out["shape"] = np.full(len(data), shape_id, np.int32)
        if "position" not in out:
            out["position"] = out[self.position_name]
        return out

    def _bind(self, ctx, e):
        fam = e["family"]
        if fam in ("lennard_jones", "lj_legacy") and self.neighbor_cutoff is None:
            # build_cell_lists() only: the pair kernel walks the cell lists (sim/interaction.py:92-118) -- generated code
            return self._bind_generic(ctx, dict(e, family="generic_pair"))
        if fam == "lennard_jones":
            eps_name, sig_name = e["roles"]["epsilon"], e["roles"]["sigma6"]
            for n in (eps_name, sig_name):
                if n not in self.feature_props:
                    raise DslError(f"{e['name']}: '{n}' must be a feature property (add_feature_property)")
            feat = self.feature_props[eps_name][0]
            nk = self.features[feat]
            ctx.set_lj_params(nk, self.feature_props[eps_name][1], self.feature_props[sig_name][1])
            self._check_prop(e, "force", Types.Vector)
            cutoff = _builtin_float(e["cutoff"])
            return dict(e, call=lambda: ctx.lennard_jones(cutoff), cutoff_value=cutoff)
        if fam == "lj_legacy":
            eps, sig6 = self._symbol(e, "epsilon"), self._symbol(e, "sigma6")
            cutoff = _builtin_float(e["cutoff"])
            return dict(e, call=lambda: ctx.lj_legacy(cutoff, eps, sig6))
        if fam == "euler_legacy":
            dt = self._symbol(e, "dt")
            return dict(e, call=lambda: ctx.euler_legacy(dt))
        if fam in ("initial_integrate", "final_integrate"):
            dt = self._symbol(e, "dt")
            for role in ("velocity", "force", "mass"):
                self._check_prop(e, role, None)
            if fam == "initial_integrate":
                if e["roles"]["position"] != self.position_name:
                    raise DslError(f"{e['name']}: '{e['roles']['position']}' is not the position property")
                return dict(e, call=lambda: ctx.initial_integrate(dt), dt=dt)
            return dict(e, call=lambda: ctx.final_integrate(dt), dt=dt)
        if fam in ("generic_pair", "generic_particle"):
            return self._bind_generic(ctx, e)
        raise DslError(f"unbound kernel family {fam}")

    def _device_storage(self):
        """Property names -> device storage.  The MD path keeps dedicated arrays (csrc/ctx.cuh) for the position, ONE
        non-volatile vector (the velocity: 'linear_velocity' if declared, else the first one), ONE volatile vector (the force:
        'force' if declared, else the first one) and ONE non-volatile real ('mass', else the first one).  Every further real /
        vector / integer property is user-defined storage: ('x', first row, components[, 'i']) in the row block of csrc/props.cu,
        rows numbered in declaration order."""
        def pick(preferred, ptype, volatile):
            names = [n for n, p in self.props.items() if p.type == ptype and p.volatile == volatile and n != self.position_name
                     and n not in self.feature_props]
            return preferred if preferred in names else (names[0] if names else None)
        builtin = {self.position_name: "pos", pick("linear_velocity", Types.Vector, False): "vel",
                   pick("force", Types.Vector, True): "force", pick("mass", Types.Real, False): "mass"}
        m, row = {}, 0
        for name, p in self.props.items():
            if name in builtin:
                m[name] = builtin[name]
            elif name in ("uid", "shape", "flags"):           # the implicit integer properties (sim/simulation.py:63-65): readable
                m[name] = name
            elif name in self.features:                       # the feature index of a particle ('type')
                m[name] = "type"
            elif p.type in (Types.Real, Types.Vector) and name not in self.feature_props:
                comps = 3 if p.type == Types.Vector else 1
                m[name] = ("x", row, comps)
                row += comps
            elif p.type == Types.Int32:                       # integers live in a double row too (exact below 2^53; the reference's
                m[name] = ("x", row, 1, "i")                  # wire format carries them as doubles as well, sim/comm.py:328-329)
                row += 1
        return m

    def _user_props(self):
        """[(name, components, volatile, defaults)] of the user-defined properties, in row order."""
        out = []
        for name, st in self._device_storage().items():
            if isinstance(st, tuple):
                p = self.props[name]
                v = p.value if isinstance(p.value, (list, tuple)) else [p.value] * st[2]     # add_property's default is the scalar 0.0
                d = [_builtin_float(x) for x in v]
                out.append((name, st[2], p.volatile, d))
        return out

    def _bind_generic(self, ctx, e, skip_fixed=True):
        from . import backend, kernelgen
        if self._compute_half and self.neighbor_cutoff is None:
            raise DslError("compute_half() needs neighbour lists (build_neighbor_lists)")
        nk = 1
        tables = {}
        for name, (feat, data) in self.feature_props.items():
            tables[name] = data
            nk = self.features[feat]
        try:
            kind, kname, src = kernelgen.translate(e["func"], self._device_storage(), tables, nk, e["symbols"], backend.jit_prelude(),
                                                   skip_fixed=skip_fixed, traversal="lists" if self.neighbor_cutoff is not None else "cells",
                                                   half=self._compute_half)
        except kernelgen.KernelGenError as err:
            raise DslError(f"kernel '{e['name']}': {err}") from None
        handle = ctx.jit_compile(src, kname)
        if kind == "pair":
            if e.get("cutoff") is None:
                raise DslError(f"kernel '{e['name']}': pair kernels need a cutoff_radius")
            cutoff = _builtin_float(e["cutoff"])
            # over the neighbour lists (full / half: compute_half()) or over the cell lists
            launch_kind = (3 if self._compute_half else 0) if self.neighbor_cutoff is not None else 2
            return dict(e, call=lambda: ctx.jit_launch(handle, launch_kind, cutoff), source=src)
        return dict(e, call=lambda: ctx.jit_launch(handle, 1), source=src)

    def _native_md_params(self, pre, fn):
        """The standard md.py procedure list runs in the native loop (pb_md_run); anything else in the Python loop."""
        if [p["family"] for p in pre] == ["initial_integrate"] and [f["family"] for f in fn] == ["lennard_jones", "final_integrate"] \
                and pre[0]["skip_first"] and fn[1]["skip_first"] and not fn[0]["skip_first"] and pre[0]["dt"] == fn[1]["dt"]:
            return (pre[0]["dt"], fn[0]["cutoff_value"], self.neighbor_cutoff, self._cell_spacing, self.reneighbor_frequency,
                    self._compute_thermo)
        return None

    def _python_loop(self, ctx, pre, fn, nsteps, rank):
        for ts in range(nsteps):
            for p in pre:
                if ts > 0 or not p["skip_first"]:
                    p["call"]()
            if ((ts + 1) % self.reneighbor_frequency) == 0 or ts == 0:
                ctx.exchange()
                ctx.borders()
                ctx.build_cell_lists()
                if self.neighbor_cutoff is not None:
                    ctx.build_neighbor_lists(self.neighbor_cutoff)
            else:
                ctx.synchronize()
            ctx.reset_volatile()
            for f in fn:
                if ts > 0 or not f["skip_first"]:
                    f["call"]()
            if self._compute_thermo > 0 and (((ts + 1) % self._compute_thermo) == 0 or ts == 0):
                t, p = ctx.compute_thermo()
                self._thermo_line(rank, ts, t, p)
            if self._vtk_due(ts):
                self._vtk_write(ctx, ts, rank, int(os.environ.get("WORLD_SIZE", 1)))

    def _thermo_line(self, rank, ts, t, p):
        self.thermo_log.append((ts, t, p))
        if rank == 0:
            print(f"{_fmt(t)}\t{_fmt(p)}")       # runtime/thermo.hpp:45-48

    def _print_summary(self, ctx, all_ms, rank):
        """stdout contract of the reference: all + timer categories + particle counts."""
        cats = {"communication": ("exchange", "borders", "synchronize"), "neighbors": ("build_cell_lists", "build_neighbor_lists")}
        lines = [("all", all_ms)]
        for e in self.pre_step + self.functions:
            timer = {"generic_pair": "linear_spring_dashpot" if self.use_contact_history else f"user_{e['name']}",
                     "generic_particle": f"user_{e['name']}"}.get(e["family"], e["family"])
            lines.append((e["name"], ctx.timer(timer)[0]))
        if self.use_contact_history:
            cats["contact_history"] = ("reset_contact_history_usage_status", "clear_unused_contact_history")
        for cat, names in cats.items():
            lines.append((cat, sum(ctx.timer(n)[0] for n in names)))
        nl, ng = ctx.counts()
        if rank == 0:
            for name, ms in lines:
                print(f"{name}: {_fmt(ms)}")
            print(f"Number of local particles: {nl} / {nl}")
            print(f"Number of ghost particles: {ng} / {ng}")

    def _read_particle_data(self, ctx, filename, prop_names, shape_id):
        """runtime/read_from_file.hpp:33-115: CSV rows, columns in the order of prop_names (vectors = 3 columns)."""
        path = filename
        if not os.path.exists(path):
            alt = os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])), "..", filename)
            if os.path.exists(alt):
                path = alt
            else:
                raise DslError(f"read_particle_data: {filename} not found")
        cols = []
        for n in prop_names:
            w = _CSV_WIDTH.get(self.props[n].type, 1)
            cols.append((n, w))
        data = np.loadtxt(path, delimiter=",", ndmin=2)
        k = 0
        arrays = {}
        for n, w in cols:
            arrays[n] = data[:, k:k + w] if w > 1 else data[:, k]
            k += w
        # every rank reads the whole file and keeps the rows inside its sub-box (runtime/read_from_file.hpp:106, isWithinSubdomain):
        # without this every rank would start with the whole system
        part = dict(arrays)
        part["position"] = arrays[self.position_name]
        part = self._keep_own(ctx, part)
        if self.position_name != "position":
            part[self.position_name] = part.pop("position")
        arrays = {k: v for k, v in part.items() if k in arrays}
        pos = arrays[self.position_name]
        storage = self._device_storage()
        vel_name = next((n for n in arrays if storage.get(n) == "vel"), None)
        mass_name = next((n for n in arrays if storage.get(n) == "mass"), None)
        ctx.upload(pos, arrays.get(vel_name), arrays.get(mass_name), arrays.get("type"), arrays.get("flags"), arrays.get("uid"),
                   np.full(len(pos), shape_id, np.int32))
        for name, st in storage.items():          # columns of user-defined properties
            if isinstance(st, tuple) and name in arrays:
                ctx.upload_property(name, arrays[name])
        return len(pos)


def vtk_write(path, position, mass, flags):
    """runtime/vtk.hpp:11-86, byte for byte: legacy ASCII unstructured grid of vertices, fixed notation with 8 decimals, INFINITE
    particles left out -- while the CELLS section keeps the index a particle has in the written range (vtk.hpp:57-61).  If the
    file cannot be opened nothing is written (vtk.hpp:42), e.g. when the output directory does not exist."""
    keep = (np.asarray(flags) & 1) == 0
    n = int(keep.sum())
    try:
        f = open(path, "w")
    except OSError:
        return False
    with f:
        f.write("# vtk DataFile Version 2.0\nParticle data\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {n} double\n")
        f.write("".join("%.8f %.8f %.8f\n" % (p[0], p[1], p[2]) for p in np.asarray(position)[keep]))
        f.write("\n\n")
        f.write(f"CELLS {n} {n * 2}\n")
        f.write("".join(f"1 {i}\n" for i in np.nonzero(keep)[0]))
        f.write("\n\n")
        f.write(f"CELL_TYPES {n}\n")
        f.write("1\n" * n)
        f.write("\n\n")
        f.write(f"POINT_DATA {n}\nSCALARS mass double\nLOOKUP_TABLE default\n")
        f.write("".join("%.8f\n" % m for m in np.asarray(mass)[keep]))
        f.write("\n\n")
    return True


_ID_STORE, _ID_CALLS = None, 0


def _broadcast_nccl_id(backend, rank, world):
    """ncclUniqueId from rank 0 to all ranks through a TCP store next to the launcher's (torchrun env: MASTER_ADDR / MASTER_PORT + 17).
    The store lives as long as the process -- a reader may arrive after rank 0 has moved on -- and every generate() of the process
    uses a key of its own (all ranks call generate() the same number of times)."""
    global _ID_STORE, _ID_CALLS
    from torch.distributed import TCPStore
    from datetime import timedelta
    if _ID_STORE is None:
        _ID_STORE = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", 29500)) + 17, world,
                             is_master=(rank == 0), timeout=timedelta(seconds=120))
    key = f"pairs_b200_nccl_id_{_ID_CALLS}"
    _ID_CALLS += 1
    if rank == 0:
        _ID_STORE.set(key, backend.nccl_unique_id())
    return bytes(_ID_STORE.get(key))


# ---- module-level factory functions (src/pairs/__init__.py:9-67) -------------------------------------------------------
def simulation(ref, shapes=None, dims=3, timesteps=100, double_prec=False, use_contact_history=False, particle_capacity=800000,
               neighbor_capacity=100, debug=False):
    return Simulation(ref, shapes, dims, timesteps, double_prec, use_contact_history, particle_capacity, neighbor_capacity, debug)


def target_cpu(parallel=False):
    return Target(False, parallel)


def target_gpu():
    return Target(True)


def int32():
    return Types.Int32


def float():
    return Types.Float


def double():
    return Types.Double


def real():
    return Types.Real


def vector():
    return Types.Vector


def matrix():
    return Types.Matrix


def quaternion():
    return Types.Quaternion


def point_mass():
    return Shapes.PointMass


def sphere():
    return Shapes.Sphere


def halfspace():
    return Shapes.Halfspace


def regular_domain_partitioner():
    return DomainPartitioners.Regular


def regular_domain_partitioner_xy():
    return DomainPartitioners.RegularXY
