"""Generic kernels (pairs_b200/kernelgen.py): translation of kernel bodies outside the hand-written families and compilation of the
result with NVRTC for sm_100a -- neither needs a GPU.  Execution parity is in tests/test_gpu_md.py."""
import os
import sys

import pytest

from pairs_b200 import backend, dsl, kernelgen

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))


def test_custom_kernels_take_the_generic_path_and_compile_for_sm100a():
    import custom_script
    psim = custom_script.build("gpu", 8, 10, 20, 1)
    assert [e["family"] for e in psim.pre_step] == ["generic_particle"]
    assert [e["family"] for e in psim.functions] == ["generic_pair", "generic_particle"]
    storage = psim._device_storage()
    assert storage == {"position": "pos", "mass": "mass", "linear_velocity": "vel", "force": "force", "uid": "uid", "shape": "shape", "flags": "flags", "type": "type"}
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    kind, name, src = kernelgen.translate(custom_script.lennard_jones, storage, tables, 4, {"kspring": 3.5, "rsoft": 1.05}, backend.jit_prelude())
    assert kind == "pair" and name == "user_lennard_jones"
    # one statement per operation, symbols as literals, select with both arms evaluated, table lookup by type pair
    assert "sqrt(rsq)" in src and "3.5 *" in src and "1.05 -" in src and "fp_epsilon[ti + tj]" in src and "? (" in src
    assert "fp_epsilon[16] = {1.0, 1.05, 1.1," in src
    assert backend.jit_check(src) > 1000          # cubin bytes
    kind, name, src = kernelgen.translate(custom_script.initial_integrate, storage, tables, 4, {"dt": 0.005, "gamma": 0.05}, backend.jit_prelude())
    assert kind == "particle" and "0.005 * 0.5" in src and "a.pos_w[i] = pi;" in src
    assert backend.jit_check(src) > 1000


def test_recognised_kernels_can_be_forced_through_the_generic_path():
    import lj_script
    dsl.FORCE_GENERIC = True
    try:
        psim = lj_script.build("gpu", 8, 10, 20, 1)
    finally:
        dsl.FORCE_GENERIC = False
    assert [e["family"] for e in psim.functions] == ["generic_pair", "generic_particle"]
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    _, _, src = kernelgen.translate(lj_script.lennard_jones, psim._device_storage(), tables, 4, {}, backend.jit_prelude())
    assert backend.jit_check(src) > 1000


def test_errors_name_the_offending_construct():
    def uses_unknown_property(i, j):
        apply(force, delta(i, j) * charge[i])

    def bad_statement(i):
        while True:
            pass

    storage = {"position": "pos", "force": "force"}
    with pytest.raises(kernelgen.KernelGenError, match="charge"):
        kernelgen.translate(uses_unknown_property, storage, {}, 1, {}, "")
    with pytest.raises(kernelgen.KernelGenError, match="unsupported statement"):
        kernelgen.translate(bad_statement, storage, {}, 1, {}, "")
    with pytest.raises(backend.BackendError, match="error"):
        backend.jit_check("this is not CUDA")


def test_vocabulary_translates_and_compiles():
    """Every keyword the generic path knows, in one pair kernel and one particle kernel; operation order is Python's."""
    def pair(i, j):
        d = delta(i, j)
        r2 = squared_distance(i, j)
        u = normalized(d)
        w = linear_velocity[i] - linear_velocity[j]
        vn = dot(w, u)
        s = select(vn < 0.0, -vn, 0.0)
        a = min(r2, rcap) + max(length(d), 0.5) + abs(vn) + squared_length(w) + sqrt(r2)
        apply(force, u * (kn * s / a) + zero_vector() - vector(0.0, 0.0, g) * mass[i] * mass[j])

    def particle(i):
        linear_velocity[i] = linear_velocity[i] * damp + force[i] * (dt / mass[i])
        position[i] += linear_velocity[i] * dt
        force[i] -= force[i]

    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass"}
    kind, name, src = kernelgen.translate(pair, storage, {}, 1, {"rcap": 4.0, "kn": 10.0, "g": 9.81}, backend.jit_prelude())
    assert kind == "pair"
    for piece in ("sqrt(", "fabs(", "? (", "4.0", "9.81", "a.vel[", "a.mass[j]", "acc_force_2 = acc_force_2 +"):
        assert piece in src, piece
    # u * (kn * s / a): the scalar is formed first ((kn * s) / a), then multiplied into each component
    assert src.index("10.0 *") < src.index("acc_force_0 = acc_force_0 +")
    assert backend.jit_check(src) > 1000
    kind, name, src = kernelgen.translate(particle, storage, {}, 1, {"damp": 0.999, "dt": 0.005}, backend.jit_prelude())
    assert kind == "particle" and src.count("a.pos_w[i] = pi;") == 1 and "a.force[2 * (size_t) a.cap + i] =" in src
    assert backend.jit_check(src) > 1000


def test_integer_division_is_c_division_as_in_the_reference():
    """ir/scalars.py:66-91 types an operation by its operands and the generator prints C: int / int truncates."""
    def k(i):
        mass[i] = mass[i] * (7 / 2) + (uid[i] / 2) * 1.0 + 7.0 / 2

    src = kernelgen.translate(k, {"mass": "mass", "uid": "uid"}, {}, 1, {}, "")[2]
    assert "const int t2 = 7 / 2;" in src and "7.0 / 2" in src and "(double)" not in src


def test_legacy_accumulation_in_a_pair_kernel_is_apply():
    """`force[i] += expr` inside a pair kernel (the older API of examples/lj_onetype.py) generates what apply(force, expr) does."""
    def new_style(i, j):
        sr2 = 1.0 / squared_distance(i, j)
        apply(force, delta(i, j) * (k * sr2))

    def old_style(i, j):
        sr2 = 1.0 / rsq
        force[i] += delta * (k * sr2)

    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass"}
    a = kernelgen.translate(new_style, storage, {}, 1, {"k": 2.0}, "")[2].replace("user_new_style", "K")
    b = kernelgen.translate(old_style, storage, {}, 1, {"k": 2.0}, "")[2].replace("user_old_style", "K")
    assert a == b and "acc_force_2 = acc_force_2 +" in a


def test_unsupported_constructs_are_rejected_with_a_reason():
    storage = {"position": "pos", "force": "force", "linear_velocity": "vel", "mass": "mass"}

    def vec_times_vec(i, j):
        apply(force, delta(i, j) * delta(i, j))

    def scalar_apply(i, j):
        apply(force, squared_distance(i, j))

    def three_args(i, j, k):
        pass

    def assigns_partner(i):
        mass[j] = 1.0

    for fn, msg in ((vec_times_vec, "dot"), (scalar_apply, "vector expression"), (assigns_partner, "unsupported statement")):
        with pytest.raises(kernelgen.KernelGenError, match=msg):
            kernelgen.translate(fn, storage, {}, 1, {}, "")
    with pytest.raises(kernelgen.KernelGenError, match=r"\(i\) or \(i, j\)"):
        kernelgen.translate(three_args, storage, {}, 1, {}, "")
    ns = {}
    exec("def made_at_run_time(i):\n    mass[i] = 1.0\n", ns)
    with pytest.raises(kernelgen.KernelGenError, match="source text"):
        kernelgen.translate(ns["made_at_run_time"], storage, {}, 1, {}, "")


def test_if_statements_and_local_updates():
    """mapping/funcs.py:179-195 (Filter / Branch) and augmented assignment of locals."""
    def pair(i, j):
        r2 = squared_distance(i, j)
        f = 1.0 / r2
        f *= 2.0
        if r2 < rin * rin:
            g = f * kin
            apply(force, delta(i, j) * g)
        else:
            apply(force, delta(i, j) * f)

    def particle(i):
        if mass[i] > 2.0:
            linear_velocity[i] = linear_velocity[i] * 0.5
        position[i] += linear_velocity[i] * dt

    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass"}
    _, _, src = kernelgen.translate(pair, storage, {}, 1, {"rin": 1.1, "kin": 3.0}, backend.jit_prelude())
    body = src[src.index("if(rsq < a.cutsq)"):]
    assert body.count("} else {") == 1 and body.count("acc_force_0 = acc_force_0 +") == 2
    assert backend.jit_check(src) > 1000
    _, _, src = kernelgen.translate(particle, storage, {}, 1, {"dt": 0.005}, backend.jit_prelude())
    # the velocity is read again after the conditional store
    after = src[src.rindex("}", 0, src.index("a.pos_w[i] = pi;")):]
    assert "a.vel[i]" in after
    assert backend.jit_check(src) > 1000

    def leaks(i, j):
        if squared_distance(i, j) < 1.0:
            g = 2.0
        apply(force, delta(i, j) * g)

    with pytest.raises(kernelgen.KernelGenError, match="'g'"):
        kernelgen.translate(leaks, storage, {}, 1, {}, "")


# ---- executing generated kernels on the host: bit-for-bit against the oracle --------------------------------------------------
def _host_kernel(tmp_path, name, src):
    """Compiles a generated kernel for the HOST (tests/host/jit_host_emulation.h stands in for the CUDA bits) together with a
    driver that runs it for every particle; returns run(n, nslots, cap, cutsq, pos4, vel, force, mass, flags, numneigh, neigh,
    xdata=None, uid=None, shape=None, radius=None, angvel=None, torque=None, inv_inertia=None, rotmat=None, quat=None, cells=None) taking array addresses (cells = (particle_cell, cell_start, cell_list, ncells, dim1,
    dim2) for pair kernels over cell lists)."""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text('#include "jit_host_emulation.h"\n' + src + f'''
extern "C" void run(int n, int nslots, int cap, double cutsq, double4 *pos, double *vel, double *force, double *mass, int *flags,
                    int *numneigh, int *neigh, double *xdata, int *uid, int *shape, double *radius, double *angvel, double *torque,
                    double *inv_inertia, double *rotmat, double *quat, int *particle_cell, int *cell_start, int *cell_list, int ncells, int dim1,
                    int dim2) {{
    PbJitArgs a;
    a.nlocal = n; a.nslots = nslots; a.cap = cap; a.pad = 0; a.cutsq = cutsq; a.pos = pos; a.pos_w = pos; a.vel = vel; a.force = force;
    a.mass = mass; a.flags = flags; a.numneigh = numneigh; a.neigh = neigh; a.xdata = xdata; a.uid = uid; a.shape = shape; a.radius = radius; a.angvel = angvel; a.torque = torque;
    a.inv_inertia = inv_inertia; a.rotmat = rotmat; a.quat = quat;
    a.particle_cell = particle_cell; a.cell_start = cell_start; a.cell_list = cell_list; a.ncells = ncells; a.dim1 = dim1; a.dim2 = dim2;
    blockDim.x = 128;
    for(int i = 0; i < n; i++) {{ blockIdx.x = i / 128; threadIdx.x = i % 128; {name}(a); }}
}}
''')
    so = tmp_path / f"{name}.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(here, "host"), str(cpp), "-o", str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    P = ctypes.c_void_p
    lib.run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [P] * 19 + [ctypes.c_int] * 3

    def run(n, nslots, cap, cutsq, pos, vel, force, mass, flags, numneigh, neigh, xdata=None, uid=None, shape=None, radius=None, angvel=None,
            torque=None, inv_inertia=None, rotmat=None, quat=None, cells=None):
        particle_cell, cell_start, cell_list, ncells, dim1, dim2 = cells if cells is not None else (None, None, None, 0, 0, 0)
        lib.run(n, nslots, cap, cutsq, pos, vel, force, mass, flags, numneigh, neigh, xdata, uid, shape, radius, angvel, torque, inv_inertia,
                rotmat, quat, particle_cell, cell_start, cell_list, ncells, dim1, dim2)
    return run


def _ptr(a):
    return a.ctypes.data


def test_generated_md_kernels_equal_the_oracle_bit_for_bit_on_the_host(tmp_path):
    """examples/md.py's three kernels through kernelgen, compiled for the host and run on the oracle's own state and lists
    (list order = the reference's, so even the summation order is the same): forces, velocities and positions are identical bits
    to the restatement of the reference's generated C++ -- per operation, per particle."""
    import numpy as np
    import lj_script
    from oracle import port
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    rng = np.random.default_rng(3)
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.05 * (rng.random((n, 3)) - 0.5)
    sim.step(0)                                   # exchange, borders, lists, force at ts = 0
    tot = n + r.nghost
    psim = lj_script.build("gpu", nx, 10, 20, 1)
    storage = psim._device_storage()
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    src = {}
    for fn, sym in ((lj_script.lennard_jones, {}), (lj_script.initial_integrate, {"dt": 0.005}), (lj_script.final_integrate, {"dt": 0.005})):
        _, name, code = kernelgen.translate(fn, storage, tables, 4, sym, backend.jit_prelude())
        src[name] = _host_kernel(tmp_path, name, code)
    # device-layout copies of the oracle's state
    cap = tot
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)          # type bits in w
    vel = np.ascontiguousarray(r.real("linear_velocity", tot).T)
    mass = r.real("mass", tot).copy()
    flags = r.ints("flags", tot).copy()
    nn, nl = r.neighbor_sets()
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    force = np.zeros((3, cap))
    src["user_lennard_jones"](n, nslots, cap, 2.5 * 2.5, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
    f_oracle = r.real("force")
    assert np.abs(f_oracle).max() > 1.0 and np.array_equal(force[:, :n].T, f_oracle)
    # the integrators of iteration 1 on the same state
    sim.initial_integrate()
    src["user_initial_integrate"](n, nslots, cap, 0.0, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
    assert np.array_equal(pos4[:n, :3], r.real("position")) and np.array_equal(vel[:, :n].T, r.real("linear_velocity"))
    sim.final_integrate()
    src["user_final_integrate"](n, nslots, cap, 0.0, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
    assert np.array_equal(vel[:, :n].T, r.real("linear_velocity"))


def test_generated_custom_kernel_equals_the_reference_generators_module_on_the_host(tmp_path):
    """The pair kernel of tests/scripts/custom_script.py: kernelgen's CUDA compiled for the host against the module the
    REFERENCE's code generator printed for the same text (oracle/_ref variant md_custom_t1, called directly on the same arrays and
    lists): identical bits, including the summation order."""
    import numpy as np
    import custom_script
    from oracle import port, ref
    if not ref.available("md_custom_t1"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    prog = ref.RefProgram("md_custom_t1")
    if not hasattr(prog.lib, "ref_md_lennard_jones"):
        pytest.skip("oracle/_ref/libref_md_custom_t1.so predates the module export")
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    rng = np.random.default_rng(9)
    n = r.nlocal
    # a liquid-like configuration with some pairs inside the softened core (r < 1.05)
    r.real("position", n, view=True)[:] += 0.25 * (rng.random((n, 3)) - 0.5)
    sim.step(0)
    tot = n + r.nghost
    nn, nl = r.neighbor_sets()
    ntypes = 4
    eps = np.array([1.0 + 0.05 * ((i % ntypes) + (i // ntypes)) for i in range(ntypes * ntypes)])
    sig6 = np.ones(ntypes * ntypes)
    # the reference's generated module on AoS arrays
    f_ref = np.zeros((n, 3))
    prog.lennard_jones(r.neighbor_capacity, n, nn.astype(np.int32), np.ascontiguousarray(nl, np.int32), r.ints("flags", tot),
                       r.real("position", tot), r.ints("type", tot), f_ref, sig6, eps)
    # kernelgen's kernel on the device layout
    psim = custom_script.build("gpu", nx, 10, 20, 1)
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    assert np.array_equal(np.array(tables["epsilon"]), eps)
    _, name, code = kernelgen.translate(custom_script.lennard_jones, psim._device_storage(), tables, ntypes, {"kspring": 3.5, "rsoft": 1.05},
                                        backend.jit_prelude())
    run = _host_kernel(tmp_path, name, code)
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)
    vel, mass, flags = np.zeros((3, tot)), np.ones(tot), r.ints("flags", tot).copy()
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    force = np.zeros((3, tot))
    run(n, nslots, tot, 2.5 * 2.5, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
    d = r.real("position", tot)
    assert np.abs(f_ref).max() > 10.0                         # pairs inside the softened core exist
    assert np.array_equal(force[:, :n].T, f_ref)


def test_kernels_on_user_defined_properties_equal_the_reference_generators_modules_on_the_host(tmp_path):
    """tests/scripts/props_script.py declares six properties beyond the MD set (reals, vectors, one volatile, an integer) and uses them in a
    setup() function, the pair kernel and both integrators.  kernelgen's CUDA for the four kernels, compiled for the host and run
    on the row layout of csrc/props.cu, against the modules the REFERENCE's generator printed for the same text (oracle/_ref
    variant md_props_t1) on its own AoS arrays: every property is identical bit for bit after set-up, force evaluation and the
    two integrator halves."""
    import numpy as np
    import props_script
    from oracle import port, ref
    if not ref.available("md_props_t1"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    prog = ref.RefProgram("md_props_t1")
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    rng = np.random.default_rng(21)
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.05 * (rng.random((n, 3)) - 0.5)
    sim.step(0)                                   # lists and ghosts of the oracle at ts = 0
    tot = n + r.nghost
    nn, nl = r.neighbor_sets()
    ntypes = 4
    psim = props_script.build("gpu", nx, 10, 20, 1)
    storage = psim._device_storage()
    assert storage == {"position": "pos", "mass": "mass", "linear_velocity": "vel", "force": "force", "scale": ("x", 0, 1),
                       "heat": ("x", 1, 1), "work": ("x", 2, 1), "path": ("x", 3, 3), "pull": ("x", 6, 3), "ups": ("x", 9, 1, "i"), "uid": "uid", "shape": "shape", "flags": "flags", "type": "type"}
    assert [e["family"] for e in psim.setup_functions] == ["generic_setup"]
    assert psim._user_props() == [("scale", 1, False, [1.0]), ("heat", 1, False, [0.0]), ("work", 1, False, [0.0]),
                                  ("path", 3, False, [0.0, 0.0, 0.0]), ("pull", 3, True, [0.0, 0.0, 0.0]), ("ups", 1, False, [0.0])]
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    kern = {}
    for fn, sym, skip_fixed in ((props_script.init_scale, {"xlen": props_script.XLEN}, False), (props_script.lennard_jones, {}, True),
                                (props_script.initial_integrate, {"dt": 0.005}, True), (props_script.final_integrate, {"dt": 0.005}, True)):
        _, name, code = kernelgen.translate(fn, storage, tables, ntypes, sym, backend.jit_prelude(), skip_fixed=skip_fixed)
        assert backend.jit_check(code) > 1000
        kern[name] = _host_kernel(tmp_path, name, code)

    # ---- the reference's arrays (AoS) and ours (device layout), same initial state; some FIXED particles ----
    pos_r = r.real("position", tot).copy()
    typ = r.ints("type", tot).copy()
    flags = r.ints("flags", tot).copy()
    flags[:n:17] |= 4
    vel_r = r.real("linear_velocity", tot).copy()
    mass = r.real("mass", tot).copy()
    scale_r, heat_r, work_r = np.full(tot, 1.0), np.zeros(tot), np.zeros(tot)
    ups_r = np.zeros(tot, np.int32)
    uid = rng.integers(0, 1000, tot).astype(np.int32)
    path_r, pull_r, force_r = np.zeros((tot, 3)), np.zeros((tot, 3)), np.zeros((tot, 3))
    cap = tot
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = pos_r
    pos4[:, 3] = typ.astype(np.int64).view(np.float64)
    vel = np.ascontiguousarray(vel_r.T)
    force = np.zeros((3, cap))
    xdata = np.zeros((10, cap))
    xdata[0] = 1.0
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    shape = np.full(tot, 2, np.int32)
    args = (_ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh), _ptr(xdata), _ptr(uid), _ptr(shape))

    def same():
        return (np.array_equal(pos4[:n, :3], pos_r[:n]) and np.array_equal(vel[:, :n].T, vel_r[:n]) and np.array_equal(force[:, :n].T, force_r[:n])
                and np.array_equal(xdata[0, :n], scale_r[:n]) and np.array_equal(xdata[1, :n], heat_r[:n]) and np.array_equal(xdata[2, :n], work_r[:n])
                and np.array_equal(xdata[3:6, :n].T, path_r[:n]) and np.array_equal(xdata[6:9, :n].T, pull_r[:n])
                and np.array_equal(xdata[9, :n], ups_r[:n].astype(np.float64)))

    # set-up function: every local, FIXED ones included
    prog.call_module("init_scale", nlocal=n, position=pos_r, scale=scale_r)
    kern["user_init_scale"](n, nslots, cap, 0.0, *args)
    assert same() and scale_r[:n].min() < 1.05 and scale_r[:n].max() > 1.15 and scale_r[0] != 1.0
    # ghosts carry the value of their source in this backend (the reference leaves them undefined; the kernels read scale[i] only)
    sigma6, epsilon = np.ones(ntypes * ntypes), np.ones(ntypes * ntypes)
    for it in range(2):
        prog.call_module("lennard_jones", neighbor_capacity=r.neighbor_capacity, nlocal=n, numneighs=nn.astype(np.int32),
                         neighborlists=np.ascontiguousarray(nl, np.int32), flags=flags, position=pos_r, type=typ, scale=scale_r, pull=pull_r,
                         force=force_r, sigma6=sigma6, epsilon=epsilon)
        kern["user_lennard_jones"](n, nslots, cap, 2.5 * 2.5, *args)
        assert same() and np.abs(force_r).max() > 1.0 and np.abs(pull_r).max() > 0.1
        assert not force_r[:n:17].any()               # FIXED particles are skipped by compute() kernels
        prog.call_module("final_integrate", nlocal=n, flags=flags, force=force_r, mass=mass, linear_velocity=vel_r, work=work_r, pull=pull_r,
                         uid=uid, ups=ups_r)
        kern["user_final_integrate"](n, nslots, cap, 0.0, *args)
        assert same() and np.abs(work_r).max() > 0.0 and ups_r[:n].min() == 0 and ups_r[:n].max() == 2 * (it + 1)
        prog.call_module("initial_integrate", nlocal=n, flags=flags, force=force_r, mass=mass, linear_velocity=vel_r, position=pos_r,
                         path=path_r, heat=heat_r)
        kern["user_initial_integrate"](n, nslots, cap, 0.0, *args)
        assert same() and np.abs(path_r).max() > 0.0 and np.abs(heat_r).max() > 0.0
        # reset_volatile_properties: force and the volatile user property
        force_r[:] = 0.0
        pull_r[:] = 0.0
        force[:] = 0.0
        xdata[6:9] = 0.0


def test_rest_of_the_vocabulary_equals_the_reference_generators_modules_on_the_host(tmp_path):
    """tests/scripts/vocab_script.py: skip_when, cross, is_point_mass, integer properties (uid, shape, the feature) with
    % & | ^ ~, and / or / not, n-ary min / max, normalized (zero vectors included), length, squared_length, dot, zero_vector.
    kernelgen's pair and particle kernel, compiled for the host, against the modules the REFERENCE's generator printed for the same
    text (oracle/_ref variant md_vocab_t1), on the same arrays: identical bits."""
    import numpy as np
    import vocab_script
    from oracle import port, ref
    if not ref.available("md_vocab_t1"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    prog = ref.RefProgram("md_vocab_t1")
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    rng = np.random.default_rng(33)
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.05 * (rng.random((n, 3)) - 0.5)
    sim.step(0)
    tot = n + r.nghost
    nn, nl = r.neighbor_sets()
    psim = vocab_script.build("gpu", nx, 10, 20, 1)
    storage = psim._device_storage()
    kern = {}
    for fn, sym in ((vocab_script.lennard_jones, {}), (vocab_script.final_integrate, {"dt": 0.005})):
        _, name, code = kernelgen.translate(fn, storage, {}, 4, sym, backend.jit_prelude())
        assert backend.jit_check(code) > 1000
        kern[name] = _host_kernel(tmp_path, name, code)
    pos_r = r.real("position", tot).copy()
    typ = r.ints("type", tot).copy()
    flags = r.ints("flags", tot).copy()
    flags[:n:13] |= 4
    uid = rng.integers(0, 1000, tot).astype(np.int32)
    shape = np.where(rng.random(tot) < 0.7, 2, rng.integers(0, 2, tot)).astype(np.int32)      # point masses, some spheres / half-spaces
    vel_r = rng.standard_normal((tot, 3))
    vel_r[5:tot:11] = vel_r[0]                    # partners with the SAME velocity as particle 0: normalized() of a zero vector
    mass = np.where(rng.random(tot) < 0.5, 1.0, 0.25)
    force_r = np.zeros((tot, 3))
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = pos_r
    pos4[:, 3] = typ.astype(np.int64).view(np.float64)
    vel = np.ascontiguousarray(vel_r.T)
    force = np.zeros((3, tot))
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    args = (_ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh), None, _ptr(uid), _ptr(shape))
    prog.call_module("lennard_jones", neighbor_capacity=r.neighbor_capacity, nlocal=n, numneighs=nn.astype(np.int32),
                     neighborlists=np.ascontiguousarray(nl, np.int32), flags=flags, position=pos_r, linear_velocity=vel_r, uid=uid, type=typ,
                     shape=shape, force=force_r)
    kern["user_lennard_jones"](n, nslots, tot, 2.5 * 2.5, *args)
    assert np.isfinite(force_r).all() and np.abs(force_r).max() > 10.0 and np.array_equal(force[:, :n].T, force_r[:n])
    v_before = vel_r.copy()
    prog.call_module("final_integrate", nlocal=n, flags=flags, shape=shape, uid=uid, mass=mass, force=force_r, linear_velocity=vel_r)
    kern["user_final_integrate"](n, nslots, tot, 0.0, *args)
    changed = np.any(vel_r[:n] != v_before[:n], axis=1)
    assert 0.1 * n < changed.sum() < 0.5 * n and np.array_equal(vel[:, :n].T, vel_r[:n])


# ---- DEM contact models -----------------------------------------------------------------------------------------------------------
DEM_STORAGE = {"position": "pos", "linear_velocity": "vel", "angular_velocity": "angvel", "mass": "mass", "radius": "radius",
               "force": "force", "torque": "torque"}
DEM_CONTACT = {"is_sticking": "c_stick", "tangential_spring_displacement": "c_tsd", "impact_velocity_magnitude": "c_ivm"}


def _model_vs_builtin(tmp_path, name, code, params, more=False):
    """Compiles a generated contact model for the host next to dem_math.h; returns run(n, seed) -> number of random pairs for which
    the model's outputs (F, T, tsd, ivm, sticking) differ in any bit from pb_dem_pair_force's, and the largest |F| seen.
    more=True: the model of dem_script.contact_model_with_more_properties -- five further lanes cx[] (copy of the new tangential
    displacement, an age that counts up by 1.0, an integer that counts up by 2) are handed in and checked as well."""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    cpp = tmp_path / f"{name}.cpp"
    cpp.write_text('#include "jit_host_emulation.h"\n#include "dem_math.h"\n#include <random>\n' + code + f'''
extern "C" int run(int n, unsigned seed, double *fmax) {{
    PbDemParams P;
    P.dt = {params["dt"]!r}; P.pi = {params["pi"]!r}; P.kappa = {params["kappa"]!r}; P.ln_coeff = {params["ln"]!r}; P.ct = {params["ct"]!r};
    P.c_sum = P.pi * P.pi + P.ln_coeff * P.ln_coeff; P.ct2 = P.ct * P.ct; P.sqrt_kappa = sqrt(P.kappa);
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    int bad = 0;
    *fmax = 0.0;
    for(int k = 0; k < n; k++) {{
        double xi[3], vi[3], wi[3], xj[3], vj[3], wj[3], nn[3], cp[3];
        for(int d = 0; d < 3; d++) {{ xi[d] = U(g); vi[d] = U(g); wi[d] = 10.0 * U(g); vj[d] = U(g); wj[d] = 10.0 * U(g); nn[d] = U(g); }}
        const double len = sqrt((nn[0] * nn[0] + nn[1] * nn[1]) + nn[2] * nn[2]);
        for(int d = 0; d < 3; d++) {{ nn[d] /= len; }}
        const double ri = 0.002 + 0.001 * (U(g) + 1.0), rj = 0.002 + 0.001 * (U(g) + 1.0), delta = 1e-4 * (U(g) + 1.0);
        for(int d = 0; d < 3; d++) {{ xj[d] = xi[d] - nn[d] * (ri + rj - delta); cp[d] = xj[d] + nn[d] * (rj - 0.5 * delta); }}
        const double mi = 1e-4 * (U(g) + 1.5), mj = (k % 7 == 0) ? INFINITY : 1e-4 * (U(g) + 1.5);     // half-spaces have infinite mass
        double tsd_a[3], tsd_b[3];
        const int mode = k % 4;        // fresh contact / live history / history parallel to the normal / sticking
        for(int d = 0; d < 3; d++) {{ tsd_a[d] = (mode == 0) ? 0.0 : ((mode == 2) ? 1e-5 * nn[d] : 1e-5 * U(g)); tsd_b[d] = tsd_a[d]; }}
        double ivm_a = (mode == 0) ? 0.0 : 0.3 * (U(g) + 1.0), ivm_b = ivm_a;
        int st_a = (mode == 3) ? 1 : 0, st_b = st_a;
        const double fs = (k % 3 == 0) ? 0.0 : 0.6, fd = 0.5;
        static const double fs_tab[2] = {{0.0, 0.6}};
        double Fa[3], Ta[3], Fb[3], Tb[3];
        pb_dem_pair_force(P, xi, vi, wi, 1.0 / mi, xj, vj, wj, mj, nn, cp, delta, fs, fd, tsd_a, &ivm_a, &st_a, Fa, Ta);
        const double age0 = (double) (k % 11) - 1.0;
        const int hits0 = 3 + 2 * (k % 5);
        double cxv[5] = {{9.0, 9.0, 9.0, age0, (double) hits0}};
        const bool kept = {name}(xi, vi, wi, mi, ri, xj, vj, wj, mj, rj, nn, cp, delta, (k % 3 == 0) ? 0 : 1, tsd_b, &ivm_b, &st_b,
                                 {"cxv" if more else "nullptr"}, Fb, Tb);
        (void) fs_tab;
        bool same = kept && st_a == st_b && memcmp(&ivm_a, &ivm_b, 8) == 0;
        if({"true" if more else "false"}) {{
            same = same && memcmp(cxv, tsd_b, 24) == 0 && cxv[3] == age0 + 1.0 && cxv[4] == (double) (hits0 + 2);
        }}
        for(int d = 0; d < 3; d++) {{
            same = same && Fa[d] == Fb[d] && Ta[d] == Tb[d] && memcmp(&tsd_a[d], &tsd_b[d], 8) == 0;
            if(fabs(Fa[d]) > *fmax) {{ *fmax = fabs(Fa[d]); }}
        }}
        bad += same ? 0 : 1;
    }}
    return bad;
}}
''')
    so = tmp_path / f"{name}.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(here, "host"),
                    "-I" + os.path.join(os.path.dirname(here), "pairs_b200", "csrc"), str(cpp), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.run.argtypes = [ctypes.c_int, ctypes.c_uint, ctypes.POINTER(ctypes.c_double)]

    def run(n, seed):
        fmax = ctypes.c_double(0.0)
        return lib.run(n, seed, ctypes.byref(fmax)), fmax.value
    return run


def test_generated_dem_contact_model_equals_the_hand_written_one_bit_for_bit(tmp_path):
    """examples/dem.py's linear_spring_dashpot (tests/scripts/dem_script.py holds it verbatim) through kernelgen.translate_dem_model,
    compiled for the host, against pb_dem_pair_force of csrc/dem_math.h -- the function the CUDA contact kernel calls and that
    tests/test_dem_host.py pins to the reference's generated module: forces, torques and the three contact properties are the same
    bits for 20000 random touching pairs (fresh and live contacts, history parallel to the normal, sticking, infinite partner mass,
    zero static friction).  The whole contact kernel built around the generated model compiles with NVRTC for sm_100a."""
    import math
    import dem_script
    params = {"dt": 5e-5, "pi": math.pi, "kappa": 2.0 * (1.0 - 0.22) / (2.0 - 0.22), "ln": -0.1053605156578263, "ct": 4.0 * 5e-5}
    symbols = {"dt": params["dt"], "pi": params["pi"], "kappa": params["kappa"], "lnDryResCoeff": params["ln"], "collisionTime_SI": params["ct"]}
    tables = {"friction_static": [0.0, 0.6], "friction_dynamic": [0.5, 0.5]}
    name, code = kernelgen.translate_dem_model(dem_script.linear_spring_dashpot, DEM_STORAGE, DEM_CONTACT, tables, symbols)
    assert name == "user_model_linear_spring_dashpot" and "return false;" in code and "fp_friction_static[tij]" in code
    run = _model_vs_builtin(tmp_path, name, code, params)
    bad, fmax = run(20000, 7)
    assert bad == 0 and fmax > 0.1
    assert backend.jit_check_dem_model(code, name) > 10000          # prelude + dem_math.h + model + dem_force_kernel.cuh, both variants
    assert backend.jit_check_dem_model(None, None) > 10000          # the same kernel around the hand-written model (option "dem_force_maxreg")


def test_contact_model_with_further_contact_properties(tmp_path):
    """contact tables are not fixed to examples/dem.py's three.  The model of
    dem_script.contact_model_with_more_properties (dem.py's body plus a second vector, a second real and a second integer contact
    property) keeps the further properties in lanes of cx[]: its forces, torques and first three properties are still
    pb_dem_pair_force's bits, the lanes hold what the three extra statements say, and the contact kernel compiles around it with the
    lanes in registers (PB_DEM_NX)."""
    import math
    import dem_script
    params = {"dt": 5e-5, "pi": math.pi, "kappa": 2.0 * (1.0 - 0.22) / (2.0 - 0.22), "ln": -0.1053605156578263, "ct": 4.0 * 5e-5}
    symbols = {"dt": params["dt"], "pi": params["pi"], "kappa": params["kappa"], "lnDryResCoeff": params["ln"], "collisionTime_SI": params["ct"]}
    tables = {"friction_static": [0.0, 0.6], "friction_dynamic": [0.5, 0.5]}
    contact = dict(DEM_CONTACT, tsd_seen="cx:vec:0", contact_age="cx:real:3", hits="cx:int:4")
    name, code = kernelgen.translate_dem_model(dem_script.contact_model_with_more_properties(), DEM_STORAGE, contact, tables, symbols,
                                               extra_lanes=5)
    assert name == "user_model_spring_dashpot_more" and code.startswith("#define PB_DEM_NX 5\n") and "cx[4] = (double) (int) (" in code
    bad, fmax = _model_vs_builtin(tmp_path, name, code, params, more=True)(20000, 11)
    assert bad == 0 and fmax > 0.1
    assert backend.jit_check_dem_model(code, name) > 10000

    def wrong_shape(i, j):
        contact_age[i, j] = contact_normal(i, j)

    with pytest.raises(kernelgen.KernelGenError, match="scalar contact property"):
        kernelgen.translate_dem_model(wrong_shape, DEM_STORAGE, contact, {}, {}, extra_lanes=5)


def test_a_different_contact_model_translates_and_compiles():
    """A body that is NOT examples/dem.py's: Hertz-like normal force with viscous damping, no tangential spring, the impact speed
    remembered in a contact property, skip for slow approaches; errors name what a contact model cannot touch."""
    def hertz(i, j):
        d = -penetration_depth(i, j)
        skip_when(d < 1e-9)
        reff = 1.0 / (1.0 / radius[i] + 1.0 / radius[j])
        rel = linear_velocity[i] - linear_velocity[j]
        vn = dot(rel, contact_normal(i, j))
        first = select(impact_velocity_magnitude[i, j] > 0.0, impact_velocity_magnitude[i, j], abs(vn))
        impact_velocity_magnitude[i, j] = first
        is_sticking[i, j] = select(abs(vn) < 1e-6, 1, 0)
        fn = kn * sqrt(reff) * d * sqrt(d) - gamma_n * vn
        if fn > 0.0:
            apply(force, fn * contact_normal(i, j))
            apply(torque, cross(contact_point(i, j) - position, fn * contact_normal(i, j)) * friction_dynamic[i, j])

    name, code = kernelgen.translate_dem_model(hertz, DEM_STORAGE, DEM_CONTACT, {"friction_dynamic": [0.5]}, {"kn": 1e6, "gamma_n": 0.2})
    assert "F[0] = F[0] +" in code and "*sticking = (int) (" in code and "*ivm =" in code and "ri" in code
    assert backend.jit_check_dem_model(code, name) > 10000

    def touches_user_state(i, j):
        apply(force, contact_normal(i, j) * inv_inertia[i])

    def applies_elsewhere(i, j):
        apply(linear_velocity, contact_normal(i, j))

    for fn, msg in ((touches_user_state, "inv_inertia"), (applies_elsewhere, "force and the torque")):
        with pytest.raises(kernelgen.KernelGenError, match=msg):
            kernelgen.translate_dem_model(fn, dict(DEM_STORAGE, inv_inertia="inv_inertia"), DEM_CONTACT, {}, {})


def test_generated_dem_per_particle_kernel_on_the_host(tmp_path):
    """examples/dem.py's gravity through the generic path (DEM scripts may carry any per-particle kernel next to the contact model):
    component assignment force[i][2] = ..., radius as a device array; the same bits as numpy doing the same IEEE operations."""
    import math
    import numpy as np
    import dem_script
    storage = dict(DEM_STORAGE, uid="uid", shape="shape", flags="flags")
    sym = {"densityParticle_SI": 2550, "densityFluid_SI": 1000, "gravity_SI": 9.81, "pi": math.pi}
    kind, name, code = kernelgen.translate(dem_script.gravity, storage, {}, 1, sym, backend.jit_prelude())
    assert kind == "particle" and "a.force[2 * (size_t) a.cap + i] =" in code and "a.radius[i]" in code
    assert "a.force[i] =" not in code and "a.force[1 * (size_t) a.cap + i] =" not in code          # only the z component is stored
    assert backend.jit_check(code) > 1000
    run = _host_kernel(tmp_path, name, code)
    rng = np.random.default_rng(2)
    n = 500
    pos4, vel, mass = np.zeros((n, 4)), np.zeros((3, n)), np.ones(n)
    force = rng.standard_normal((3, n))
    radius = 0.001 + 0.001 * rng.random(n)
    flags = np.zeros(n, np.int32)
    flags[::9] = 4
    f0 = force.copy()
    run(n, 0, n, 0.0, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), None, None, None, None, None, _ptr(radius), None, None)
    volume = (4.0 / 3.0) * math.pi * radius * radius * radius
    expect = f0[2] - (2550 - 1000) * volume * 9.81
    expect[::9] = f0[2, ::9]                       # FIXED particles are skipped
    assert np.array_equal(force[2], expect) and np.array_equal(force[:2], f0[:2])


def test_generated_dem_integrator_and_setup_equal_dem_math_on_the_host(tmp_path):
    """examples/dem.py's euler and update_mass_and_inertia through the generic path: matrix / quaternion properties and the
    keywords transposed, inversed, diagonal_matrix, default_quaternion, quaternion, quaternion_to_rotation_matrix, the matrix *
    vector / vector * matrix / quaternion * quaternion products, `mass[i] = infinity`, a scalar assigned to a matrix.  Compiled for
    the host and compared with pb_dem_euler / pb_dem_sphere_inv_inertia of csrc/dem_math.h (pinned to the reference's generated code
    by tests/test_dem_host.py): every array identical, bit for bit."""
    import ctypes
    import math
    import subprocess
    import numpy as np
    import dem_script
    here = os.path.dirname(os.path.abspath(__file__))
    storage = dict(DEM_STORAGE, uid="uid", shape="shape", flags="flags", inv_inertia="inv_inertia", rotation_matrix="rotmat",
                   rotation_quat="quat")
    dt = 5e-5
    _, n_eul, c_eul = kernelgen.translate(dem_script.euler, storage, {}, 1, {"dt": dt}, backend.jit_prelude())
    _, n_upd, c_upd = kernelgen.translate(dem_script.update_mass_and_inertia, storage, {}, 1,
                                          {"densityParticle_SI": 2550, "pi": math.pi, "infinity": math.inf}, backend.jit_prelude(), skip_fixed=False)
    assert "PB_INFINITY" in c_upd and "sin(" in c_eul and "cos(" in c_eul
    assert backend.jit_check(c_eul) > 1000 and backend.jit_check(c_upd) > 1000
    run_eul, run_upd = _host_kernel(tmp_path, n_eul, c_eul), _host_kernel(tmp_path, n_upd, c_upd)
    # the hand-written side: tests/host/dem_host.cpp (dem_math.h on AoS arrays)
    so = tmp_path / "dem_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(os.path.dirname(here), "pairs_b200", "csrc"),
                    os.path.join(here, "host", "dem_host.cpp"), "-o", str(so)], check=True)
    host = ctypes.CDLL(str(so))
    rng = np.random.default_rng(11)
    n = 400
    flags = np.zeros(n, np.int32)
    flags[::7] = 4
    shape = np.zeros(n, np.int32)
    shape[3::50] = 1
    mass = 1e-4 * (1.0 + rng.random(n))
    radius = 0.001 * (1.0 + rng.random(n))
    # ---- update_mass_and_inertia (a setup() function: FIXED particles included) ----
    pos4 = np.zeros((n, 4))
    m_g = mass.copy()
    Iinv_g, R_g, q_g = (np.full((k, n), 7.0) for k in (9, 9, 4))
    z3 = np.zeros((3, n))
    run_upd(n, 0, n, 0.0, _ptr(pos4), _ptr(z3), _ptr(z3.copy()), _ptr(m_g), _ptr(flags), None, None, None, None, _ptr(shape), _ptr(radius), None, None,
            _ptr(Iinv_g), _ptr(R_g), _ptr(q_g))
    assert np.array_equal(R_g.T, np.tile(np.eye(3).ravel(), (n, 1))) and np.array_equal(q_g.T, np.tile([1.0, 0.0, 0.0, 0.0], (n, 1)))
    assert np.all(np.isinf(m_g[shape == 1])) and np.array_equal(m_g[shape == 0], mass[shape == 0]) and not Iinv_g[:, shape == 1].any()
    host.host_dem_sphere_inv_inertia.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    want = np.zeros(9)
    for i in np.flatnonzero(shape == 0)[:60]:
        host.host_dem_sphere_inv_inertia(mass[i], radius[i], _ptr(want))
        assert np.array_equal(Iinv_g[:, i], want), i
    # ---- euler: two steps on random states (tiny and large rotations) ----
    pos = rng.random((n, 3))
    vel, f, tau = rng.standard_normal((n, 3)), 1e-3 * rng.standard_normal((n, 3)), 1e-7 * rng.standard_normal((n, 3))
    w = 50.0 * rng.standard_normal((n, 3))
    w[::5] = 0.0
    tau[::5] = 0.0                                         # no rotation at all: quaternion() takes its zero branch
    Iinv = np.ascontiguousarray(Iinv_g.T.copy())
    Iinv[shape == 1] = 0.0
    q = np.tile([1.0, 0.0, 0.0, 0.0], (n, 1))
    R = np.tile(np.eye(3).ravel(), (n, 1))
    pos4[:, :3] = pos
    vel_g, f_g, tau_g, w_g = (np.ascontiguousarray(a.T) for a in (vel, f, tau, w))
    Iinv_g2, q_g2, R_g2 = (np.ascontiguousarray(a.T) for a in (Iinv, q, R))
    mass_e = mass.copy()
    P = (ctypes.c_double * 16)()
    host.host_dem_params.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 8
    host.host_dem_params(P, dt, math.pi, 0.8, -0.1, 2e-4, 2550.0, 1000.0, 9.81)
    host.host_dem_euler.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 10
    for step in range(2):
        host.host_dem_euler(P, n, _ptr(flags), _ptr(mass_e), _ptr(f), _ptr(tau), _ptr(Iinv), _ptr(pos), _ptr(vel), _ptr(w), _ptr(q), _ptr(R))
        run_eul(n, 0, n, 0.0, _ptr(pos4), _ptr(vel_g), _ptr(f_g), _ptr(mass_e), _ptr(flags), None, None, None, None, _ptr(shape), _ptr(radius), _ptr(w_g),
                _ptr(tau_g), _ptr(Iinv_g2), _ptr(R_g2), _ptr(q_g2))
        assert np.array_equal(pos4[:, :3], pos) and np.array_equal(vel_g.T, vel) and np.array_equal(w_g.T, w)
        assert np.array_equal(q_g2.T, q) and np.array_equal(R_g2.T, R)
    assert np.abs(q[1] - [1.0, 0.0, 0.0, 0.0]).max() > 1e-4 and np.array_equal(q[0], [1.0, 0.0, 0.0, 0.0])     # rotated / FIXED


def test_generated_pair_kernel_over_cell_lists_on_the_host(tmp_path):
    """A script that builds cell lists only (build_cell_lists, no build_neighbor_lists) has its pair kernels walk cell 0 and the 27
    stencil cells of the particle's cell (sim/interaction.py:92-118).  md.py's lennard_jones generated with traversal="cells",
    compiled for the host and run on the oracle's cell lists (as a CSR, in the oracle's in-cell order): the same pairs pass the
    cutoff as through the oracle's neighbour lists, met in the same order -- identical bits."""
    import numpy as np
    import lj_script
    from oracle import port
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    rng = np.random.default_rng(5)
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.05 * (rng.random((n, 3)) - 0.5)
    sim.step(0)
    tot = n + r.nghost
    ncells, ccap = r.ncells, r.cell_capacity
    sizes = r.ints("cell_sizes", ncells)
    cp = r.ints("cell_particles", ncells * ccap).reshape(ncells, ccap)
    cell_start = np.zeros(ncells + 1, np.int32)
    cell_start[1:] = np.cumsum(sizes)
    cell_list = np.concatenate([cp[c, :sizes[c]] for c in range(ncells)]).astype(np.int32)
    assert len(cell_list) == tot and sizes[0] == 0
    particle_cell = r.ints("particle_cell", tot).copy()
    dc = r.decomposition()["dim_cells"]
    psim = lj_script.build("gpu", nx, 10, 20, 1)
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    _, name, code = kernelgen.translate(lj_script.lennard_jones, psim._device_storage(), tables, 4, {}, backend.jit_prelude(), traversal="cells")
    assert "a.cell_start[c_lo]" in code and "a.numneigh" not in code.split('extern "C"')[1] and backend.jit_check(code) > 1000
    run = _host_kernel(tmp_path, name, code)
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)
    vel, mass, flags = np.zeros((3, tot)), np.ones(tot), r.ints("flags", tot).copy()
    force = np.zeros((3, tot))
    run(n, 0, tot, 2.5 * 2.5, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), None, None,
        cells=(_ptr(particle_cell), _ptr(cell_start), _ptr(cell_list), ncells, int(dc[1]), int(dc[2])))
    f_oracle = r.real("force")
    assert np.abs(f_oracle).max() > 1.0
    assert np.abs(force[:, :n].T - f_oracle).max() <= 1e-13 * np.abs(f_oracle).max()
    assert np.array_equal(force[:, :n].T, f_oracle)


def test_generated_pair_kernel_for_half_lists_on_the_host(tmp_path):
    """compute_half() with a generated pair kernel: every pair once, the partner gets the opposite term by an atomic add unless it
    is a ghost or FIXED (ir/apply.py:111-125).  Run one "thread" after the other on the oracle's half lists this is exactly the
    serial order of the reference's generated code -- identical bits with the oracle's force (the restatement of the reference run
    with compute_half() enabled, pinned by tests/test_oracle_pin.py)."""
    import numpy as np
    import lj_script
    from oracle import port
    nx = 6
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    sim.compute_half(True)
    r = sim.ranks[0]
    rng = np.random.default_rng(6)
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.05 * (rng.random((n, 3)) - 0.5)
    r.ints("flags", n, view=True)[::19] |= 4              # some FIXED particles: no own sum, no partner update
    sim.step(0)
    tot = n + r.nghost
    nn, nl = r.neighbor_sets()
    assert nn.mean() < 60                                  # half lists (full ones hold ~76; pairs with ghosts stay at the local end)
    psim = lj_script.build("gpu", nx, 10, 20, 1)
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    _, name, code = kernelgen.translate(lj_script.lennard_jones, psim._device_storage(), tables, 4, {}, backend.jit_prelude(), half=True)
    assert code.count("atomicAdd(") == 6 and "j < a.nlocal" in code and backend.jit_check(code) > 1000
    run = _host_kernel(tmp_path, name, code)
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)
    vel, mass, flags = np.zeros((3, tot)), np.ones(tot), r.ints("flags", tot).copy()
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    force = np.zeros((3, tot))
    run(n, nslots, tot, 2.5 * 2.5, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
    f_oracle = r.real("force")
    fixed = (flags[:n] & 4) != 0
    assert np.abs(f_oracle).max() > 1.0 and fixed.sum() > 20 and not f_oracle[fixed].any() and not force[:, n:].any()
    assert np.array_equal(force[:, :n].T, f_oracle)


def test_random_expressions_generated_code_equals_python_arithmetic(tmp_path):
    """Differential test of the translator: 40 random per-particle kernel bodies (nested arithmetic with Python's precedence, unary
    minus, locals, augmented assignment, select / min / max / sqrt / abs, if / else, vector algebra with dot / cross / length and
    component access) are translated, compiled for the host and run on random inputs; CPython evaluates the very same source with
    IEEE doubles and the reference's keyword semantics.  One operation per statement and no contraction: every result must be the
    same bits."""
    import importlib.util
    import math
    import random
    import numpy as np
    rnd = random.Random(1234)

    def scalar(depth):
        if depth <= 0 or rnd.random() < 0.2:
            return rnd.choice(["mass[i]", "position[i][0]", "position[i][1]", "linear_velocity[i][2]", "1.5", "0.25", "c3", "2"] + names["s"])
        k = rnd.random()
        if k < 0.45:
            return f"({scalar(depth - 1)} {rnd.choice('+-*')} {scalar(depth - 1)})"
        if k < 0.55:
            return f"({scalar(depth - 1)} / (1.0 + abs({scalar(depth - 1)})))"
        if k < 0.62:
            return f"-{scalar(depth - 1)}"
        if k < 0.72:
            return f"select({scalar(depth - 1)} < {scalar(depth - 1)}, {scalar(depth - 1)}, {scalar(depth - 1)})"
        if k < 0.80:
            return f"{rnd.choice(['min', 'max'])}({scalar(depth - 1)}, {scalar(depth - 1)}, {scalar(depth - 1)})"
        if k < 0.86:
            return f"sqrt(abs({scalar(depth - 1)}))"
        if k < 0.90:
            return f"dot({vector(depth - 1)}, {vector(depth - 1)})"
        if k < 0.94:
            return f"{rnd.choice(['squared_length', 'length'])}({vector(depth - 1)})"
        return f"{vector(depth - 1)}[{rnd.randrange(3)}]"

    def vector(depth):
        if depth <= 0 or rnd.random() < 0.3:
            return rnd.choice(["position[i]", "linear_velocity[i]", "force[i]"] + names["v"])
        k = rnd.random()
        if k < 0.4:
            return f"({vector(depth - 1)} {rnd.choice('+-')} {vector(depth - 1)})"
        if k < 0.7:
            return f"({vector(depth - 1)} * {scalar(depth - 1)})"
        if k < 0.85:
            return f"cross({vector(depth - 1)}, {vector(depth - 1)})"
        return f"normalized({vector(depth - 1)})"

    bodies = []
    names = {"s": [], "v": []}                  # locals defined so far
    for n in range(40):
        names["s"], names["v"] = [], []
        lines = [f"def k{n}(i):", f"    a = {scalar(2)}"]
        names["s"].append("a")
        lines.append(f"    u = {vector(1)}")
        names["v"].append("u")
        lines.append(f"    b = {scalar(2)}")
        names["s"].append("b")
        lines.append(f"    b *= {scalar(1)}")
        lines += [f"    if {scalar(1)} > {scalar(1)}:", f"        a = {scalar(2)}", f"        mass[i] = a + b", "    else:", f"        mass[i] = {scalar(3)}"]
        lines += [f"    force[i] = {vector(2)}", f"    linear_velocity[i] += {vector(2)}"]
        if n % 2 == 0:      # position read, assigned, read again: the first read keeps its value (snapshots, as for every property)
            lines += ["    old = position[i]", f"    position[i] = position[i] + {vector(1)} * 0.5", "    linear_velocity[i] = (position[i] - old) * 2.0"]
        bodies.append("\n".join(lines))
    text = "\n\n\n".join(bodies) + "\n"
    mod_path = tmp_path / "fuzz_kernels.py"
    mod_path.write_text(text)
    spec = importlib.util.spec_from_file_location("fuzz_kernels", mod_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    # ---- CPython as the reference: IEEE doubles, keyword semantics of mapping/keywords.py ----
    class V:
        def __init__(self, c):
            self.c = [float(x) for x in c]

        def __add__(self, o):
            return V([x + y for x, y in zip(self.c, o.c)])

        def __sub__(self, o):
            return V([x - y for x, y in zip(self.c, o.c)])

        def __mul__(self, s):
            return V([x * s for x in self.c])

        def __rmul__(self, s):
            return V([s * x for x in self.c])

        def __neg__(self):
            return V([-x for x in self.c])

        def __getitem__(self, k):
            return self.c[k]

    def fold(better):
        def f(*args):
            e = args[0]
            for x in args[1:]:
                e = x if better(x, e) else e
            return e
        return f

    def dot(p, q):
        return (p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]

    def cross(p, q):
        return V([p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]])

    def normalized(p):
        ln = math.sqrt(dot(p, p))
        inv = 1.0 / ln if ln != 0.0 else math.inf
        return V([(x * inv) if ln > 0.0 else 0.0 for x in p.c])

    class Prop:
        def __init__(self, rows, vec):
            self.rows, self.vec = rows, vec

        def __getitem__(self, i):
            return V(self.rows[:, i]) if self.vec else float(self.rows[i])

        def __setitem__(self, i, v):
            if self.vec:
                self.rows[:, i] = v.c
            else:
                self.rows[i] = v

    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass"}
    npart = 64
    rng = np.random.default_rng(99)
    checked = 0
    for n in range(40):
        fn = getattr(mod, f"k{n}")
        try:
            _, name, code = kernelgen.translate(fn, storage, {}, 1, {"c3": 0.75}, backend.jit_prelude())
        except kernelgen.KernelGenError:
            continue                                     # (the random generator may produce vector +- scalar etc.: not the subject here)
        run = _host_kernel(tmp_path, name, code)
        pos4 = np.zeros((npart, 4))
        pos4[:, :3] = rng.standard_normal((npart, 3))
        vel, force, mass = rng.standard_normal((3, npart)), rng.standard_normal((3, npart)), 0.5 + rng.random(npart)
        flags = np.zeros(npart, np.int32)
        g_pos, g_vel, g_force, g_mass = pos4.copy(), vel.copy(), force.copy(), mass.copy()
        run(npart, 0, npart, 0.0, _ptr(g_pos), _ptr(g_vel), _ptr(g_force), _ptr(g_mass), _ptr(flags), None, None)
        prow = np.ascontiguousarray(pos4[:, :3].T)
        env = {"position": Prop(prow, True), "linear_velocity": Prop(vel, True), "force": Prop(force, True),
               "mass": Prop(mass, False), "c3": 0.75, "select": lambda c, x, y: x if c else y, "min": fold(lambda x, e: x < e),
               "max": fold(lambda x, e: x > e), "sqrt": math.sqrt, "abs": abs, "dot": dot, "cross": cross, "normalized": normalized,
               "squared_length": lambda p: dot(p, p), "length": lambda p: math.sqrt(dot(p, p))}
        fn.__globals__.update(env)
        with np.errstate(all="ignore"):
            for i in range(npart):
                fn(i)
        for got, want in ((g_mass, mass), (g_force, force), (g_vel, vel), (np.ascontiguousarray(g_pos[:, :3].T), prow)):
            assert np.array_equal(got.view(np.int64), np.asarray(want).view(np.int64)) or \
                np.array_equal(np.nan_to_num(got, nan=7.0), np.nan_to_num(want, nan=7.0)), (n, text.split("\n\n\n")[n])
        checked += 1
    assert checked >= 25


@pytest.mark.parametrize("half", [False, True])
def test_random_pair_kernels_generated_code_equals_python_arithmetic(tmp_path, half):
    """(half = True: generated for compute_half() -- the partner receives the opposite term unless it is a ghost or FIXED; one thread
    after the other on the host, so the "atomic" updates happen in the order Python performs them.)
    The same differential test for pair kernels: 20 random bodies over delta / squared_distance, properties of both partners, a
    feature table, locals, if / else, skip_when and one or two apply() -- the generated neighbour loop (hoisted loads of i, cached
    loads of j, register accumulators, cutoff, FIXED filter) against CPython walking the same lists."""
    import importlib.util
    import math
    import random
    import numpy as np
    rnd = random.Random(77)
    names = {"s": [], "v": []}

    def scalar(depth):
        if depth <= 0 or rnd.random() < 0.25:
            return rnd.choice(["mass[i]", "mass[j]", "squared_distance(i, j)", "eps[i, j]", "linear_velocity[j][1]", "0.5", "kk", "3"] + names["s"])
        k = rnd.random()
        if k < 0.5:
            return f"({scalar(depth - 1)} {rnd.choice('+-*')} {scalar(depth - 1)})"
        if k < 0.6:
            return f"(1.0 / (0.5 + abs({scalar(depth - 1)})))"
        if k < 0.7:
            return f"select({scalar(depth - 1)} < {scalar(depth - 1)}, {scalar(depth - 1)}, {scalar(depth - 1)})"
        if k < 0.8:
            return f"dot({vector(depth - 1)}, {vector(depth - 1)})"
        if k < 0.9:
            return f"sqrt(squared_distance(i, j) + abs({scalar(depth - 1)}))"
        return f"-{scalar(depth - 1)}"

    def vector(depth):
        if depth <= 0 or rnd.random() < 0.35:
            return rnd.choice(["delta(i, j)", "linear_velocity[i]", "linear_velocity[j]", "position[j]"] + names["v"])
        k = rnd.random()
        if k < 0.4:
            return f"({vector(depth - 1)} {rnd.choice('+-')} {vector(depth - 1)})"
        if k < 0.8:
            return f"({vector(depth - 1)} * {scalar(depth - 1)})"
        return f"cross({vector(depth - 1)}, {vector(depth - 1)})"

    bodies = []
    for n in range(20):
        names["s"], names["v"] = [], []
        lines = [f"def p{n}(i, j):", f"    skip_when({scalar(1)} > 2.5)", f"    a = {scalar(2)}"]
        names["s"].append("a")
        lines.append(f"    w = {vector(1)}")
        names["v"].append("w")
        lines += [f"    if {scalar(1)} < {scalar(1)}:", f"        a = {scalar(2)}", f"        apply(force, {vector(2)})", "    else:", f"        w = {vector(1)}"]
        lines.append(f"    apply(force, w * a + {vector(2)})")
        bodies.append("\n".join(lines))
    text = "\n\n\n".join(bodies) + "\n"
    mod_path = tmp_path / "fuzz_pairs.py"
    mod_path.write_text(text)
    spec = importlib.util.spec_from_file_location("fuzz_pairs", mod_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class V:
        def __init__(self, c):
            self.c = [float(x) for x in c]

        def __add__(self, o):
            return V([x + y for x, y in zip(self.c, o.c)])

        def __sub__(self, o):
            return V([x - y for x, y in zip(self.c, o.c)])

        def __mul__(self, s):
            return V([x * s for x in self.c])

        def __rmul__(self, s):
            return V([s * x for x in self.c])

        def __getitem__(self, k):
            return self.c[k]

    class Skip(Exception):
        pass

    # a small random system with explicit lists (some partners beyond the cutoff, some FIXED particles)
    rng = np.random.default_rng(3)
    n, nt, ntypes = 48, 64, 2
    pos = rng.random((nt, 3)) * 3.0
    typ = rng.integers(0, ntypes, nt)
    vel = rng.standard_normal((3, nt))
    mass = 0.5 + rng.random(nt)
    flags = np.zeros(nt, np.int32)
    flags[5:n:11] = 4
    eps = [1.0, 1.5, 0.75, 2.0]
    lists = [sorted(rng.choice([j for j in range(nt) if j != i], size=rng.integers(3, 12), replace=False)) for i in range(n)]
    nslots = max(len(x) for x in lists)
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    numneigh = np.zeros(nt, np.int32)
    for i, lst in enumerate(lists):
        neigh[i // 32, :len(lst), i % 32] = lst
        numneigh[i] = len(lst)
    cutsq = 2.0 * 2.0
    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass"}
    checked = 0
    for k in range(20):
        fn = getattr(mod, f"p{k}")
        try:
            _, name, code = kernelgen.translate(fn, storage, {"eps": eps}, ntypes, {"kk": 1.25}, backend.jit_prelude(), half=half)
        except kernelgen.KernelGenError:
            continue
        run = _host_kernel(tmp_path, name, code)
        pos4 = np.zeros((nt, 4))
        pos4[:, :3] = pos
        pos4[:, 3] = typ.astype(np.int64).view(np.float64)
        g_vel, g_force, g_mass, g_flags = vel.copy(), np.zeros((3, nt)), mass.copy(), flags.copy()
        run(n, nslots, nt, cutsq, _ptr(pos4), _ptr(g_vel), _ptr(g_force), _ptr(g_mass), _ptr(g_flags), _ptr(numneigh), _ptr(neigh))
        want = np.zeros((3, nt))
        state = {}

        def apply_(_prop, v):
            state["acc"] = [x + y for x, y in zip(state["acc"], v.c)]
            jj = state["j"]
            if half and jj < n and not (flags[jj] & 4):
                for d in range(3):
                    want[d, jj] = want[d, jj] + -(v.c[d])

        def skip_when(c):
            if c:
                raise Skip()

        class P:
            def __init__(self, rows):
                self.rows = rows

            def __getitem__(self, idx):
                return V(self.rows[:, idx]) if self.rows.ndim == 2 else float(self.rows[idx])

        class FP:
            def __getitem__(self, ij):
                return eps[typ[ij[0]] * ntypes + typ[ij[1]]]

        def dot(p, q):
            return (p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]

        env = {"position": P(np.ascontiguousarray(pos.T)), "linear_velocity": P(vel), "mass": P(mass), "force": "force", "eps": FP(), "kk": 1.25,
               "apply": apply_, "skip_when": skip_when, "select": lambda c, x, y: x if c else y, "sqrt": math.sqrt, "abs": abs, "dot": dot,
               "cross": lambda p, q: V([p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]]),
               "delta": lambda i, j: V(pos[i]) - V(pos[j]),
               "squared_distance": lambda i, j: dot(V(pos[i]) - V(pos[j]), V(pos[i]) - V(pos[j]))}
        fn.__globals__.update(env)
        for i in range(n):
            if flags[i] & 4:
                continue
            state["acc"] = [0.0, 0.0, 0.0]
            for j in lists[i]:
                if env["squared_distance"](i, j) < cutsq:
                    state["j"] = int(j)
                    try:
                        fn(i, int(j))
                    except Skip:
                        pass
            want[:, i] = [want[d, i] + x for d, x in enumerate(state["acc"])]
        assert np.array_equal(g_force, want), (k, bodies[k])
        checked += 1
    assert checked >= 10


def test_partner_prefetch_does_not_change_a_bit(tmp_path):
    """Generated pair kernels fetch partners four at a time (kernelgen.PAIR_PREFETCH); the plain loop gives the same bits."""
    import numpy as np
    import custom_script
    from oracle import port
    nx = 5
    sim = port.md_example(nx, reneigh_every=20, particle_capacity=60000, send_capacity=60000)
    r = sim.ranks[0]
    n = r.nlocal
    r.real("position", n, view=True)[:] += 0.2 * (np.random.default_rng(1).random((n, 3)) - 0.5)
    sim.step(0)
    tot = n + r.nghost
    nn, nl = r.neighbor_sets()
    psim = custom_script.build("gpu", nx, 10, 20, 1)
    tables = {k: v[1] for k, v in psim.feature_props.items()}
    pos4 = np.zeros((tot, 4))
    pos4[:, :3] = r.real("position", tot)
    pos4[:, 3] = r.ints("type", tot).astype(np.int64).view(np.float64)
    vel, mass, flags = np.zeros((3, tot)), np.ones(tot), r.ints("flags", tot).copy()
    nslots = int(nn.max())
    neigh = np.zeros(((n + 31) // 32, nslots, 32), np.int32)
    for i in range(n):
        neigh[i // 32, :nn[i], i % 32] = nl[i, :nn[i]]
    numneigh = np.zeros(tot, np.int32)
    numneigh[:n] = nn
    out = {}
    saved = kernelgen.PAIR_PREFETCH
    try:
        for u in (1, 3, 4):
            kernelgen.PAIR_PREFETCH = u
            _, name, code = kernelgen.translate(custom_script.lennard_jones, psim._device_storage(), tables, 4, {"kspring": 3.5, "rsoft": 1.05},
                                                backend.jit_prelude())
            assert ("pp[u]" in code) == (u > 1)
            force = np.zeros((3, tot))
            sub = tmp_path / f"u{u}"                     # same kernel name for every depth: one build directory each
            sub.mkdir()
            run = _host_kernel(sub, name, code)
            run(n, nslots, tot, 2.5 * 2.5, _ptr(pos4), _ptr(vel), _ptr(force), _ptr(mass), _ptr(flags), _ptr(numneigh), _ptr(neigh))
            out[u] = force
    finally:
        kernelgen.PAIR_PREFETCH = saved
    assert np.abs(out[1]).max() > 1.0 and np.array_equal(out[1], out[3]) and np.array_equal(out[1], out[4])


def test_random_contact_models_generated_code_equals_python_arithmetic(tmp_path):
    """Differential test of translate_dem_model: 15 random contact-model bodies (pair geometry symbols, properties of both partners,
    the three contact properties read and assigned in statement order, skip_when, if / else, one or two apply() to force and torque)
    compiled for the host against CPython evaluating the same source on the same random pairs: F, T, the contact properties and the
    keep / skip decision are the same bits."""
    import ctypes
    import importlib.util
    import math
    import random
    import subprocess
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    rnd = random.Random(5)
    names = {"s": [], "v": []}

    def scalar(depth):
        if depth <= 0 or rnd.random() < 0.25:
            return rnd.choice(["mass[i]", "mass[j]", "radius[i]", "radius[j]", "penetration_depth(i, j)", "impact_velocity_magnitude[i, j]",
                               "fric[i, j]", "0.5", "kn", "2"] + names["s"])
        k = rnd.random()
        if k < 0.5:
            return f"({scalar(depth - 1)} {rnd.choice('+-*')} {scalar(depth - 1)})"
        if k < 0.6:
            return f"(1.0 / (0.5 + abs({scalar(depth - 1)})))"
        if k < 0.7:
            return f"select({scalar(depth - 1)} < {scalar(depth - 1)}, {scalar(depth - 1)}, {scalar(depth - 1)})"
        if k < 0.85:
            return f"dot({vector(depth - 1)}, {vector(depth - 1)})"
        return f"length({vector(depth - 1)})"

    def vector(depth):
        if depth <= 0 or rnd.random() < 0.35:
            return rnd.choice(["contact_normal(i, j)", "contact_point(i, j)", "position[i]", "position[j]", "linear_velocity[i]", "linear_velocity[j]",
                               "angular_velocity[i]", "angular_velocity[j]", "tangential_spring_displacement[i, j]"] + names["v"])
        k = rnd.random()
        if k < 0.4:
            return f"({vector(depth - 1)} {rnd.choice('+-')} {vector(depth - 1)})"
        if k < 0.8:
            return f"({vector(depth - 1)} * {scalar(depth - 1)})"
        return f"cross({vector(depth - 1)}, {vector(depth - 1)})"

    bodies = []
    for n in range(15):
        names["s"], names["v"] = [], []
        lines = [f"def m{n}(i, j):", f"    skip_when({scalar(1)} > 1.5)", f"    a = {scalar(2)}"]
        names["s"].append("a")
        lines.append(f"    w = {vector(1)}")
        names["v"].append("w")
        lines += [f"    tangential_spring_displacement[i, j] = {vector(2)}", f"    if is_sticking[i, j] == 1:", f"        a = {scalar(2)}",
                  f"        impact_velocity_magnitude[i, j] = {scalar(2)}", "    else:", f"        apply(torque, {vector(2)})",
                  f"    is_sticking[i, j] = select({scalar(1)} < {scalar(1)}, 1, 0)", f"    apply(force, w * a + tangential_spring_displacement[i, j])",
                  f"    apply(torque, cross(contact_point(i, j) - position, {vector(1)}))"]
        bodies.append("\n".join(lines))
    mod_path = tmp_path / "fuzz_models.py"
    mod_path.write_text("\n\n\n".join(bodies) + "\n")
    spec = importlib.util.spec_from_file_location("fuzz_models", mod_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class V:
        def __init__(self, c):
            self.c = [float(x) for x in c]

        def __add__(self, o):
            return V([x + y for x, y in zip(self.c, o.c)])

        def __sub__(self, o):
            if not isinstance(o, V):
                return NotImplemented                     # a bare property (apply context): Side.__rsub__
            return V([x - y for x, y in zip(self.c, o.c)])

        def __mul__(self, s):
            return V([x * s for x in self.c])

        def __rmul__(self, s):
            return V([s * x for x in self.c])

        def __getitem__(self, k):
            return self.c[k]

    class Skip(Exception):
        pass

    def dot(p, q):
        return (p[0] * q[0] + p[1] * q[1]) + p[2] * q[2]

    rng = np.random.default_rng(8)
    npairs = 200
    X = {k: rng.standard_normal((npairs, 3)) for k in ("xi", "vi", "wi", "xj", "vj", "wj", "n", "cp", "tsd")}
    S = {k: 0.5 + rng.random(npairs) for k in ("mi", "ri", "mj", "rj", "delta", "ivm")}
    stick = rng.integers(0, 2, npairs).astype(np.int32)
    tij = rng.integers(0, 4, npairs).astype(np.int32)
    fric = [0.1, 0.5, 0.25, 0.75]
    checked = 0
    for n in range(15):
        fn = getattr(mod, f"m{n}")
        try:
            name, code = kernelgen.translate_dem_model(fn, DEM_STORAGE, DEM_CONTACT, {"fric": fric}, {"kn": 3.5})
        except kernelgen.KernelGenError:
            continue
        assert backend.jit_check_dem_model(code, name) > 10000
        cpp = tmp_path / f"{name}.cpp"
        cpp.write_text('#include "jit_host_emulation.h"\n' + code + f'''
extern "C" void run(int np_, const double *xi, const double *vi, const double *wi, const double *mi, const double *ri, const double *xj,
                    const double *vj, const double *wj, const double *mj, const double *rj, const double *nn, const double *cp, const double *delta,
                    const int *tij, double *tsd, double *ivm, int *stick, double *F, double *T, int *kept) {{
    for(int k = 0; k < np_; k++) {{
        kept[k] = {name}(xi + 3 * k, vi + 3 * k, wi + 3 * k, mi[k], ri[k], xj + 3 * k, vj + 3 * k, wj + 3 * k, mj[k], rj[k], nn + 3 * k, cp + 3 * k,
                         delta[k], tij[k], tsd + 3 * k, ivm + k, stick + k, nullptr, F + 3 * k, T + 3 * k) ? 1 : 0;
    }}
}}
''')
        so = tmp_path / f"{name}.so"
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17", "-I" + os.path.join(here, "host"), str(cpp), "-o", str(so)],
                       check=True)
        lib = ctypes.CDLL(str(so))
        lib.run.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 20
        g_tsd, g_ivm, g_stick = X["tsd"].copy(), S["ivm"].copy(), stick.copy()
        F, T, kept = np.zeros((npairs, 3)), np.zeros((npairs, 3)), np.zeros(npairs, np.int32)
        lib.run(npairs, _ptr(X["xi"]), _ptr(X["vi"]), _ptr(X["wi"]), _ptr(S["mi"]), _ptr(S["ri"]), _ptr(X["xj"]), _ptr(X["vj"]), _ptr(X["wj"]),
                _ptr(S["mj"]), _ptr(S["rj"]), _ptr(X["n"]), _ptr(X["cp"]), _ptr(S["delta"]), _ptr(tij), _ptr(g_tsd), _ptr(g_ivm), _ptr(g_stick), _ptr(F),
                _ptr(T), _ptr(kept))
        # ---- CPython on the same pairs ----
        p_tsd, p_ivm, p_stick = X["tsd"].copy(), S["ivm"].copy(), stick.copy()
        pF, pT, p_kept = np.zeros((npairs, 3)), np.zeros((npairs, 3)), np.zeros(npairs, np.int32)
        for k in range(npairs):
            acc = {"force": [0.0] * 3, "torque": [0.0] * 3, "first": {"force": True, "torque": True}}

            class Side:
                def __init__(self, a_i, a_j, vec):
                    self.a_i, self.a_j, self.vec = a_i, a_j, vec

                def __getitem__(self, who):
                    a = self.a_i if who == "i" else self.a_j
                    return V(a[k]) if self.vec else float(a[k])

                # a bare property inside apply() means prop[i]
                def __rsub__(self, other):
                    return other - V(self.a_i[k])

            class Contact:
                def __init__(self, kind):
                    self.kind = kind

                def __getitem__(self, _ij):
                    return V(p_tsd[k]) if self.kind == "tsd" else (float(p_ivm[k]) if self.kind == "ivm" else int(p_stick[k]))

                def __setitem__(self, _ij, v):
                    if self.kind == "tsd":
                        p_tsd[k] = v.c
                    elif self.kind == "ivm":
                        p_ivm[k] = v
                    else:
                        p_stick[k] = int(v)

            class Fric:
                def __getitem__(self, _ij):
                    return fric[tij[k]]

            def apply_(target, v):
                if acc["first"][target]:
                    acc[target] = list(v.c)
                    acc["first"][target] = False
                else:
                    acc[target] = [x + y for x, y in zip(acc[target], v.c)]

            def skip_when(c):
                if c:
                    raise Skip()

            env = {"i": "i", "j": "j", "position": Side(X["xi"], X["xj"], True), "linear_velocity": Side(X["vi"], X["vj"], True),
                   "angular_velocity": Side(X["wi"], X["wj"], True), "mass": Side(S["mi"], S["mj"], False), "radius": Side(S["ri"], S["rj"], False),
                   "tangential_spring_displacement": Contact("tsd"), "impact_velocity_magnitude": Contact("ivm"), "is_sticking": Contact("stick"),
                   "fric": Fric(), "kn": 3.5, "force": "force", "torque": "torque", "apply": apply_, "skip_when": skip_when,
                   "select": lambda c, x, y: x if c else y, "abs": abs, "dot": dot, "length": lambda p: math.sqrt(dot(p, p)),
                   "cross": lambda p, q: V([p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]]),
                   "contact_normal": lambda i, j: V(X["n"][k]), "contact_point": lambda i, j: V(X["cp"][k]),
                   "penetration_depth": lambda i, j: -float(S["delta"][k])}
            fn.__globals__.update(env)
            try:
                fn("i", "j")
                p_kept[k] = 1
            except Skip:
                p_kept[k] = 0
            pF[k], pT[k] = acc["force"], acc["torque"]
        assert np.array_equal(kept, p_kept), (n, bodies[n])
        kept_total = kept_total + int(kept.sum()) if "kept_total" in dir() else int(kept.sum())
        live = kept == 1
        for got, want in ((F[live], pF[live]), (T[live], pT[live]), (g_tsd, p_tsd), (g_ivm, p_ivm), (g_stick, p_stick)):
            assert np.array_equal(got, want), (n, bodies[n])
        checked += 1
    assert checked >= 8 and kept_total > 500


def _ints_kernel(i):
    k = (uid[i] % 5) + (uid[i] & 6) * 2 - (uid[i] | 9) + (uid[i] ^ 3) + (~uid[i] & 7)
    m = select(k > 3, k, 0 - k)
    mass[i] = mass[i] * m + abs(0 - k) + min(k, 2, uid[i]) + max(k, 7) + sqrt(uid[i] + 1)
    if uid[i] % 2 == 0 and not (k == 4) or shape[i] == 2:
        linear_velocity[i] = linear_velocity[i] * (1 + k)
    else:
        m = m * 3
    force[i] = force[i] * m


def test_integer_expressions_generated_code_equals_python_arithmetic(tmp_path):
    """Integer properties and operators (% & | ^ ~, comparisons, and / or / not, select / min / max / abs on integers, promotion
    to real): the generated kernel against CPython running the same function on non-negative uids (where C and Python agree on %)."""
    import math
    import numpy as np
    storage = {"position": "pos", "linear_velocity": "vel", "force": "force", "mass": "mass", "uid": "uid", "shape": "shape"}
    _, name, code = kernelgen.translate(_ints_kernel, storage, {}, 1, {}, backend.jit_prelude())
    assert backend.jit_check(code) > 1000
    run = _host_kernel(tmp_path, name, code)
    rng = np.random.default_rng(4)
    n = 300
    pos4 = np.zeros((n, 4))
    vel, force, mass = rng.standard_normal((3, n)), rng.standard_normal((3, n)), 0.5 + rng.random(n)
    uid = rng.integers(0, 5000, n).astype(np.int32)
    shape = rng.integers(0, 3, n).astype(np.int32)
    flags = np.zeros(n, np.int32)
    g_vel, g_force, g_mass = vel.copy(), force.copy(), mass.copy()
    run(n, 0, n, 0.0, _ptr(pos4), _ptr(g_vel), _ptr(g_force), _ptr(g_mass), _ptr(flags), None, None, None, _ptr(uid), _ptr(shape))

    class Vec:
        def __init__(self, rows, k):
            self.rows, self.k = rows, k

        def __mul__(self, s):
            return [float(x) * s for x in self.rows[:, self.k]]

    class VProp:
        def __init__(self, rows):
            self.rows = rows

        def __getitem__(self, k):
            return Vec(self.rows, k)

        def __setitem__(self, k, v):
            self.rows[:, k] = v

    def fold(better):
        def f(*args):
            e = args[0]
            for x in args[1:]:
                e = x if better(x, e) else e
            return e
        return f

    env = {"uid": [int(x) for x in uid], "shape": [int(x) for x in shape], "mass": mass, "linear_velocity": VProp(vel), "force": VProp(force),
           "select": lambda c, x, y: x if c else y, "min": fold(lambda x, e: x < e), "max": fold(lambda x, e: x > e), "abs": abs, "sqrt": math.sqrt}
    _ints_kernel.__globals__.update(env)
    for i in range(n):
        _ints_kernel(i)
    assert np.array_equal(g_mass, mass) and np.array_equal(g_vel, vel) and np.array_equal(g_force, force)
