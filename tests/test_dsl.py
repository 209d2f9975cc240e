"""The `pairs` DSL surface (CPU part): API names/defaults of the reference façade, kernel recognition, error behaviour."""
import ast
import importlib.util
import os
import sys

import numpy as np
import pytest

import pairs
from pairs_b200 import dsl
from tests.conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))


def test_facade_names_of_the_reference_api():
    # src/pairs/__init__.py:9-67
    for name in ("simulation", "target_cpu", "target_gpu", "int32", "float", "double", "real", "vector", "matrix", "quaternion",
                 "point_mass", "sphere", "halfspace", "regular_domain_partitioner", "regular_domain_partitioner_xy"):
        assert callable(getattr(pairs, name)), name
    sim = pairs.simulation("md", [pairs.point_mass()], timesteps=200, double_prec=True)
    # Simulation methods, src/pairs/sim/simulation.py
    for m in ("target", "add_position", "add_property", "add_feature", "add_feature_property", "add_contact_property", "set_domain",
              "set_domain_partitioner", "pbc", "copper_fcc_lattice", "dem_sc_grid", "read_particle_data", "setup", "build_cell_lists",
              "build_neighbor_lists", "reneighbor_every", "compute_half", "compute_thermo", "vtk_output", "compute", "generate",
              "add_real_property", "add_vector_property", "from_file"):
        assert callable(getattr(sim, m)), m
    assert sim.reneighbor_frequency == 1 and sim.particle_capacity == 800000 and sim.neighbor_capacity == 100
    assert set(sim.props) == {"uid", "shape", "flags"}
    assert (pairs.sphere(), pairs.halfspace(), pairs.point_mass()) == (0, 1, 2)          # sim/shapes.py


def test_recognises_the_md_kernels_with_renamed_properties():
    import lj_script
    assert dsl.recognise(lj_script.lennard_jones)[0] == "lennard_jones"
    fam, roles = dsl.recognise(lj_script.initial_integrate)
    assert fam == "initial_integrate" and roles["velocity"] == "linear_velocity" and roles["dt"] == "dt"
    assert dsl.recognise(lj_script.final_integrate)[0] == "final_integrate"

    def lj2(a, b):
        s2 = 1.0 / squared_distance(a, b)                                   # noqa: F821
        s6 = s2 * s2 * s2 * sg[a, b]                                        # noqa: F821
        apply(f, delta(a, b) * (48.0 * s6 * (s6 - 0.5) * s2 * ep[a, b]))    # noqa: F821
    fam, roles = dsl.recognise(lj2)
    assert fam == "lennard_jones" and roles == {"sigma6": "sg", "force": "f", "epsilon": "ep"}


def test_unknown_kernels_fail_loudly():
    def morse(i, j):
        apply(force, delta(i, j) * 2.0)      # noqa: F821

    def lj_wrong_constant(i, j):
        sr2 = 1.0 / squared_distance(i, j)                                               # noqa: F821
        sr6 = sr2 * sr2 * sr2 * sigma6[i, j]                                             # noqa: F821
        apply(force, delta(i, j) * (24.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j]))     # noqa: F821
    for k in (morse, lj_wrong_constant):
        with pytest.raises(dsl.DslError, match="not one of the kernel families"):
            dsl.recognise(k)


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference checkout not present")
def test_recognises_the_reference_example_functions_verbatim(tmp_path):
    src = open("/root/reference/examples/md.py").read()
    got = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef):
            p = tmp_path / f"k_{node.name}.py"
            p.write_text(ast.get_source_segment(src, node) + "\n")
            spec = importlib.util.spec_from_file_location(node.name, p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            got[node.name] = dsl.recognise(getattr(mod, node.name))[0]
    assert got == {"lennard_jones": "lennard_jones", "initial_integrate": "initial_integrate", "final_integrate": "final_integrate"}


def test_plan_and_error_behaviour():
    import lj_script
    psim = lj_script.build("gpu", 8, 100, 20, 10)
    assert [e["family"] for e in psim.pre_step] == ["initial_integrate"]
    assert [e["family"] for e in psim.functions] == ["lennard_jones", "final_integrate"]
    assert psim.cell_spacing() == 2.5 + 0.3 and psim.neighbor_cutoff == 2.5 + 0.3
    # the reference's query methods (sim/simulation.py:116-128, 159-201)
    assert psim.use_double_precision() and psim.ndims() == 3 and psim.max_shapes() == 1 and psim.get_shape_id(0) == pairs.point_mass()
    assert psim.position() is psim.property("position") and psim.property("mass").value == 1.0 and psim.property("nope") is None
    assert psim.feature("type") == 4 and psim.feature_property("epsilon")[0] == "type" and psim.contact_property("x") is None
    psim.enable_profiler()
    L = 8 * pow(4.0 / 0.8442, 1.0 / 3.0)
    assert psim.grid == [0.0, 0.0, 0.0, L, L, L]
    cpu = lj_script.build("cpu", 8, 10, 20, 10)
    with pytest.raises(dsl.DslError, match="B200 GPUs only"):
        cpu.generate()
    assert psim._compute_half is False
    half = lj_script.build("gpu", 8, 10, 20, 10)
    half.compute_half()                      # sim/simulation.py:119-120
    assert half._compute_half is True
    assert dsl._fmt(1.44) == "1.44" and dsl._fmt(1.215643219) == "1.21564" and dsl._fmt(0.6928049) == "0.692805"
    if not HAS_GPU:
        from pairs_b200.backend import BackendError
        with pytest.raises(BackendError):
            psim.generate()


def test_recognises_the_dem_kernels():
    import dem_script
    assert dsl.recognise(dem_script.update_mass_and_inertia)[0] == "update_mass_and_inertia"
    assert dsl.recognise(dem_script.gravity)[0] == "gravity"
    assert dsl.recognise(dem_script.euler)[0] == "euler"
    fam, roles = dsl.recognise(dem_script.linear_spring_dashpot)
    assert fam == "linear_spring_dashpot" and roles["tsd"] == "tangential_spring_displacement" and roles["ln_coeff"] == "lnDryResCoeff"
    psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10)
    assert [e["family"] for e in psim.functions] == ["gravity", "linear_spring_dashpot", "euler"]
    assert psim.setup_functions[0]["family"] == "update_mass_and_inertia" and psim.use_contact_history and psim._pbc == [True, True, False]


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="reference checkout not present")
def test_recognises_the_reference_dem_example_verbatim(tmp_path):
    src = open("/root/reference/examples/dem.py").read()
    got = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef):
            p = tmp_path / f"k_{node.name}.py"
            p.write_text(ast.get_source_segment(src, node) + "\n")
            spec = importlib.util.spec_from_file_location(node.name, p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            got[node.name] = dsl.recognise(getattr(mod, node.name))[0]
    assert got == {k: k for k in ("update_mass_and_inertia", "linear_spring_dashpot", "euler", "gravity")}


def test_legacy_api_of_lj_onetype():
    import lj_legacy_script
    assert dsl.recognise(lj_legacy_script.lj)[0] == "lj_legacy" and dsl.recognise(lj_legacy_script.euler)[0] == "euler_legacy"
    psim = lj_legacy_script.build("gpu", 6, 5)
    assert psim.shapes == [pairs.point_mass()] and psim.reneighbor_frequency == 1 and psim._target.is_gpu()
    a = pow(4.0 / 0.8442, 1.0 / 3.0)
    assert psim.grid == [0.0, 0.0, 0.0, 6 * a, 6 * a, 6 * a] and psim.setups[0][0] == "copper_fcc_lattice"
    assert psim.functions[0]["symbols"] == {"sigma6": 1.0, "epsilon": 1.0}


def test_vtk_writer_reproduces_reference_file_bytes(tmp_path):
    """vtk_write against a file runtime/vtk.hpp wrote (tests/golden/dem_vtk_t1_local_0.vtk, reference run of examples/dem.py):
    parse the golden back into arrays, write them again, compare the bytes -- format, precision, INFINITE filtering, CELLS ids."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dem_vtk_t1_local_0.vtk")
    lines = open(gold).read().split("\n")
    n = int(lines[4].split()[1])
    pos = np.array([[float(x) for x in ln.split()] for ln in lines[5:5 + n]])
    k = lines.index(f"CELLS {n} {2 * n}")
    ids = [int(ln.split()[1]) for ln in lines[k + 1:k + 1 + n]]
    m0 = lines.index("LOOKUP_TABLE default")
    mass = np.array([float(x) for x in lines[m0 + 1:m0 + 1 + n]])
    # the reference's range held 422 particles: 420 spheres followed by the two INFINITE planes, which it skips
    assert n == 420 and ids == list(range(420))
    position = np.vstack([pos, [[0.0, 0.0, 0.0], [0.8, 0.015, 0.2]]])
    masses = np.concatenate([mass, [1.0, 1.0]])
    flags = np.array([0] * 420 + [13, 13], np.int32)
    out = tmp_path / "again.vtk"
    assert dsl.vtk_write(str(out), position, masses, flags)
    assert open(out, "rb").read() == open(gold, "rb").read()
    # INFINITE particles in the middle keep their slot number in CELLS (runtime/vtk.hpp:57-61)
    assert dsl.vtk_write(str(out), position[[0, 420, 1]], masses[[0, 420, 1]], flags[[0, 420, 1]])
    txt = open(out).read()
    assert "POINTS 2 double" in txt and "CELLS 2 4\n1 0\n1 2\n" in txt
    assert not dsl.vtk_write(str(tmp_path / "missing_dir" / "x.vtk"), position, masses, flags)


def test_multi_rank_read_particle_data_keeps_own_rows_and_global_bodies():
    """runtime/read_from_file.hpp:44-106: a rank keeps the rows inside its sub-box (x < max - 1e-5) plus infinite / fixed / global bodies."""
    class FakeCtx:
        def decomposition(self):
            return {"nranks": np.array([2, 1, 1]), "subdom": np.array([0.0, 0.5, 0.0, 1.0, 0.0, 1.0])}

    part = {"position": np.array([[0.1, 0.5, 0.5], [0.5 - 1e-6, 0.5, 0.5], [0.7, 0.5, 0.5], [0.9, 0.2, 0.2]]),
            "uid": np.array([1, 2, 3, 4], np.int32), "flags": np.array([0, 0, 0, 13], np.int32), "shape": np.array([0, 0, 0, 1], np.int32)}
    kept = dsl.Simulation._keep_own(FakeCtx(), part)
    assert list(kept["uid"]) == [1, 4]            # row 2 sits inside the 1e-5 exclusion band at the upper face, row 4 is a global plane
    assert kept["position"].shape == (2, 3) and list(kept["shape"]) == [0, 1]

    class OneRank(FakeCtx):
        def decomposition(self):
            return {"nranks": np.array([1, 1, 1]), "subdom": np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0])}
    assert dsl.Simulation._keep_own(OneRank(), part) is part


def test_md_read_particle_data_uploads_only_the_rows_of_the_own_sub_box(tmp_path):
    """The MD path of read_particle_data / from_file (examples/lj_onetype.py) with two ranks: each rank uploads the rows inside its
    sub-box only (runtime/read_from_file.hpp:106) -- without the filter every rank would start with the whole system and the
    particles would be duplicated `world` times."""
    import pairs
    rows = np.array([[1.0, 0.2, 0.3, 0.4, 0.01, 0.02, 0.03], [1.0, 3.1, 0.3, 0.4, 0.0, 0.0, 0.0], [1.0, 6.0, 5.0, 1.0, 0.1, 0.0, 0.0],
                     [1.0, 6.6, 6.6, 6.6, 0.0, 0.0, 0.2]])
    path = tmp_path / "minimd_setup_4x4x4_test.input"
    np.savetxt(path, rows, delimiter=",")
    psim = pairs.simulation("lj", debug=True, timesteps=1)
    psim.add_real_property('mass', 1.0)
    psim.add_position('position')
    psim.add_vector_property('velocity')
    psim.add_vector_property('force', vol=True)
    psim.from_file(str(path), ['mass', 'position', 'velocity'])
    L = 4 * pow(4.0 / 0.8442, 1.0 / 3.0)
    uploads = []

    class FakeCtx:
        def __init__(self, lo, hi):
            self.lo, self.hi = lo, hi

        def decomposition(self):
            return {"nranks": np.array([2, 1, 1]), "subdom": np.array([self.lo, self.hi, 0.0, L, 0.0, L])}

        def upload(self, pos, vel, mass, *rest):
            uploads.append((np.array(pos), np.array(vel), np.array(mass)))

    kind, args = psim.setups[-1]
    assert kind == "read_particle_data"
    n0 = psim._read_particle_data(FakeCtx(0.0, L / 2), *args)
    n1 = psim._read_particle_data(FakeCtx(L / 2, L), *args)
    assert (n0, n1) == (2, 2) and np.array_equal(uploads[0][0][:, 0], [0.2, 3.1]) and np.array_equal(uploads[1][0][:, 0], [6.0, 6.6])
    assert np.array_equal(uploads[0][1], rows[:2, 4:7]) and np.array_equal(uploads[1][2], [1.0, 1.0])


def test_further_properties_become_user_defined_storage():
    """Properties beyond the MD set get rows in the user-property block (csrc/props.cu), numbered in declaration order; the MD
    slots go to the canonical names when they are declared, else to the first property of the matching type / volatility."""
    import lj_script
    from pairs_b200 import backend, kernelgen

    def charged(i, j):
        apply(force, delta(i, j) * charge[i] * charge[j])
        apply(field, delta(i, j) * charge[j])

    psim = lj_script.build("gpu", 8, 10, 20, 1)
    psim.add_property("charge", pairs.real(), 0.5)
    psim.add_property("field", pairs.vector(), volatile=True)
    psim.add_property("dipole", pairs.vector(), (0.0, 0.0, 1.0))
    st = psim._device_storage()
    assert st == {"position": "pos", "mass": "mass", "linear_velocity": "vel", "force": "force", "charge": ("x", 0, 1),
                  "field": ("x", 1, 3), "dipole": ("x", 4, 3), "uid": "uid", "shape": "shape", "flags": "flags", "type": "type"}
    assert psim._user_props() == [("charge", 1, False, [0.5]), ("field", 3, True, [0.0, 0.0, 0.0]), ("dipole", 3, False, [0.0, 0.0, 1.0])]
    _, _, src = kernelgen.translate(charged, st, {}, 1, {}, backend.jit_prelude())
    assert "a.xdata[0 * (size_t) a.cap + j]" in src and "acc_x1_2 = acc_x1_2 +" in src and "a.xdata[3 * (size_t) a.cap + i] =" in src
    assert backend.jit_check(src) > 1000
    # other names: the first non-volatile vector is the velocity, the first volatile one the force, the first real the mass
    q = pairs.simulation("x", [pairs.point_mass()], timesteps=1, double_prec=True)
    q.add_position("r")
    q.add_property("m", pairs.real(), 1.0)
    q.add_property("m2", pairs.real(), 2.0)
    q.add_property("v", pairs.vector())
    q.add_property("f", pairs.vector(), volatile=True)
    q.add_property("g", pairs.vector(), volatile=True)
    assert q._device_storage() == {"r": "pos", "m": "mass", "m2": ("x", 0, 1), "v": "vel", "f": "force", "g": ("x", 1, 3),
                                   "uid": "uid", "shape": "shape", "flags": "flags"}
    # setup(): any per-particle function (generic path, no FIXED filter); pair functions are rejected
    def init(i):
        m2[i] = 2.0 * m[i]
    q.setup(init)
    assert q.setup_functions[0]["family"] == "generic_setup"
    with pytest.raises(dsl.DslError, match="one particle argument"):
        q.setup(charged)
    _, _, src = kernelgen.translate(init, q._device_storage(), {}, 1, {}, backend.jit_prelude(), skip_fixed=False)
    assert "PB_FLAG_FIXED) != 0" not in src.split('extern "C"')[1]


def test_dem_script_with_a_generated_contact_model_plans_and_compiles():
    """A DEM script whose pair kernel is not recognised (here: forced) gets its contact model generated: the procedure list stays
    gravity / contact model / euler, the model is translated from the script's declared properties and the contact kernel built
    around it compiles for sm_100a (no GPU needed)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    from pairs_b200 import backend
    dsl.FORCE_GENERIC_CONTACT_MODEL = True
    try:
        psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10)
    finally:
        dsl.FORCE_GENERIC_CONTACT_MODEL = False
    assert [e["family"] for e in psim.functions] == ["gravity", "generic_pair", "euler"]
    name, src, nk = psim._translate_dem_model(psim.functions[1])
    assert name == "user_model_linear_spring_dashpot" and nk == 1 and "fp_friction_dynamic[1] = {0.5}" in src
    assert backend.jit_check_dem_model(src, name) > 10000
    # the stock script is still recognised
    assert [e["family"] for e in dem_script.build("gpu", (0.1, 0.015, 0.04), 10).functions] == ["gravity", "linear_spring_dashpot", "euler"]


def test_dem_script_with_a_user_property_plans_and_compiles():
    """DEM scripts may declare further properties and per-particle kernels: storage rows, procedure list and NVRTC compilation."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    from pairs_b200 import backend, kernelgen

    def odometer(i):
        travelled[i] += dt * length(linear_velocity[i])
        hits[i] = hits[i] + select(linear_velocity[i][2] > 0.0, 1, 0)

    psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10)
    psim.add_property('travelled', pairs.real(), 0.0)
    psim.add_property('hits', pairs.int32(), 0)
    psim.compute(odometer, symbols={'dt': 5e-5})
    assert [e["family"] for e in psim.functions] == ["gravity", "linear_spring_dashpot", "euler", "generic_particle"]
    st = psim._dem_storage()
    assert st["travelled"] == ("x", 0, 1) and st["hits"] == ("x", 1, 1, "i") and st["rotation_quat"] == "quat" and "normal" not in st
    assert psim._dem_user_props() == [("travelled", 1, False, [0.0]), ("hits", 1, False, [0.0])]
    _, name, src = kernelgen.translate(odometer, st, {}, 1, {"dt": 5e-5}, backend.jit_prelude())
    assert "a.xdata[0 * (size_t) a.cap + i] =" in src and "a.xdata[1 * (size_t) a.cap + i] = (double)" in src
    assert backend.jit_check(src) > 1000


def test_md_script_with_cell_lists_only_plans_generated_pair_kernels():
    """build_cell_lists() without build_neighbor_lists(): the recognised lennard_jones is generated with the cell-list traversal."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import lj_script
    psim = lj_script.build("gpu", 8, 10, 20, 1, cells_only=True)
    assert psim.neighbor_cutoff is None and psim.cell_spacing() == 2.5 + 0.3

    class Ctx:                                   # records what would be compiled / launched
        def __init__(self):
            self.compiled, self.launched = [], []

        def jit_compile(self, src, name):
            self.compiled.append((name, src))
            return len(self.compiled) - 1

        def jit_launch(self, handle, kind, cutoff=0.0):
            self.launched.append((handle, kind, cutoff))

        def initial_integrate(self, dt):
            pass

        def final_integrate(self, dt):
            pass

    ctx = Ctx()
    plan = [psim._bind(ctx, e) for e in psim.pre_step + psim.functions]
    assert [p["family"] for p in plan] == ["initial_integrate", "generic_pair", "final_integrate"]
    assert psim._native_md_params(plan[:1], plan[1:]) is None
    name, src = ctx.compiled[0]
    assert name == "user_lennard_jones" and "a.cell_start[c_lo]" in src
    plan[1]["call"]()
    assert ctx.launched == [(0, 2, 2.5)]
    # compute_half() with a generated pair kernel: half-list variant (launch kind 3), needs neighbour lists
    dsl.FORCE_GENERIC_NAMES = {"lennard_jones"}
    try:
        half = lj_script.build("gpu", 8, 10, 20, 1)
    finally:
        dsl.FORCE_GENERIC_NAMES = set()
    half.compute_half()
    ctx2 = Ctx()
    half._bind(ctx2, half.functions[0])["call"]()
    assert ctx2.launched == [(0, 3, 2.5)] and "atomicAdd(&a.force[j]" in ctx2.compiled[0][1]
    psim.compute_half()
    with pytest.raises(dsl.DslError, match="needs neighbour lists"):
        psim._bind(Ctx(), psim.functions[0])


def test_dem_script_with_a_reneighbouring_interval_is_planned_module_by_module():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10, reneighbor=3)
    assert psim.reneighbor_frequency == 3 and [e["family"] for e in psim.functions] == ["gravity", "linear_spring_dashpot", "euler"]


def test_mutated_example_kernels_are_not_mistaken_for_the_hand_written_families(tmp_path):
    """The hand-written CUDA kernels are bound by RECOGNISING the kernel text.  Every single-point mutation of the seven example
    kernels -- a literal changed, an operator flipped, a comparison reversed, two different operands swapped -- must fall out of its
    family (and go to the generic path) instead of silently running the unmodified hand-written kernel."""
    import copy
    import importlib.util
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    import lj_script
    flip = {ast.Add: ast.Sub, ast.Sub: ast.Add, ast.Mult: ast.Div, ast.Div: ast.Mult, ast.Lt: ast.Gt, ast.Gt: ast.Lt, ast.LtE: ast.GtE,
            ast.GtE: ast.LtE, ast.Eq: ast.NotEq, ast.And: ast.Or, ast.Or: ast.And}
    total = by_roles = by_symbols = 0
    for mod, names in ((lj_script, ("lennard_jones", "initial_integrate", "final_integrate")),
                       (dem_script, ("update_mass_and_inertia", "linear_spring_dashpot", "euler", "gravity"))):
        src = open(mod.__file__).read()
        tree = ast.parse(src)
        for fdef in [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]:
            family, orig_roles = dsl.recognise(getattr(mod, fdef.name))
            sites = [k for k, n in enumerate(ast.walk(fdef)) if isinstance(n, (ast.Constant, ast.BinOp, ast.Compare, ast.BoolOp, ast.AugAssign))]
            mutants = []
            for site in sites:
                for kind in ("value", "swap"):
                    m = copy.deepcopy(fdef)
                    node = list(ast.walk(m))[site]
                    if kind == "value":
                        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
                            node.value = node.value + 1
                        elif isinstance(node, (ast.BinOp, ast.AugAssign)) and type(node.op) in flip:
                            node.op = flip[type(node.op)]()
                        elif isinstance(node, ast.BoolOp):
                            node.op = flip[type(node.op)]()
                        elif isinstance(node, ast.Compare) and type(node.ops[0]) in flip:
                            node.ops = [flip[type(node.ops[0])]()]
                        else:
                            continue
                    else:
                        if not isinstance(node, ast.BinOp) or ast.dump(node.left) == ast.dump(node.right):
                            continue
                        node.left, node.right = node.right, node.left
                    mutants.append(m)
            assert len(mutants) >= 3, fdef.name
            text = "\n\n\n".join(ast.unparse(ast.fix_missing_locations(m)).replace(f"def {fdef.name}(", f"def mutant_{k}(") for k, m in enumerate(mutants))
            path = tmp_path / f"mutants_{fdef.name}.py"
            path.write_text(text + "\n")
            spec = importlib.util.spec_from_file_location(path.stem, path)
            mm = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mm)
            psim = mod.build("gpu") if mod is lj_script else mod.build("gpu", (0.1, 0.015, 0.04), 10)
            for k in range(len(mutants)):
                try:
                    got, roles = dsl.recognise(getattr(mm, f"mutant_{k}"))
                except dsl.DslError:
                    got, roles = None, {}
                # still the family's text up to names: then the names must be what gives it away (roles bound to the wrong arrays)
                if got == family and psim._roles_fit_the_storage({"family": got, "roles": roles}):
                    # ... or two SYMBOLS changed places (e.g. densityFluid_SI - densityParticle_SI): the hand-written kernel then
                    # receives their values in the exchanged roles, which is what the mutant computes
                    changed = [r for r in roles if roles[r] != orig_roles[r]]
                    assert changed and all(roles[r] not in psim.props and roles[r] not in psim.feature_props and
                                           roles[r] not in psim.contact_props for r in changed), (fdef.name, ast.unparse(mutants[k]))
                    by_symbols += 1
                elif got == family:
                    by_roles += 1
                total += 1
    assert total > 150 and by_roles >= 1


def test_recognised_kernels_keep_their_family_only_on_the_arrays_they_were_written_for():
    """Stock scripts: every kernel stays a hand-written family.  The text of initial_integrate applied to ANOTHER vector property, or
    dem.py's euler with torque and inv_inertia exchanged, is recognised structurally but re-classified as a generic kernel."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    import lj_legacy_script
    import lj_script
    for psim, want in ((lj_script.build("gpu"), ["initial_integrate", "lennard_jones", "final_integrate"]),
                       (dem_script.build("gpu", (0.1, 0.015, 0.04), 10), ["gravity", "linear_spring_dashpot", "euler"])):
        psim._reclassify()
        assert [e["family"] for e in psim.pre_step + psim.functions] == want
        assert all(not e["family"].startswith("generic") for e in psim.setup_functions)
    legacy = lj_legacy_script.build("gpu", 4, 3)
    legacy._reclassify()
    assert [e["family"] for e in legacy.functions] == ["lj_legacy", "euler_legacy"]

    def drift(i):
        path[i] += (dt * 0.5) * force[i] / mass[i]
        position[i] += dt * path[i]

    psim = lj_script.build("gpu")
    psim.add_property("path", pairs.vector())
    psim.compute(drift, symbols={"dt": 0.005})
    assert psim.functions[-1]["family"] == "initial_integrate" and psim.functions[-1]["roles"]["velocity"] == "path"
    psim._reclassify()
    assert [e["family"] for e in psim.functions] == ["lennard_jones", "final_integrate", "generic_particle"]


def test_checkpoint_files_round_trip_every_bit_and_split_over_ranks(tmp_path):
    """checkpoint_output() / read_checkpoint() host side with a stand-in context: two ranks write their locals (plus DEM contact rows),
    one rank -- and, separately, two other sub-boxes -- read them back: every double bit for bit, every row exactly once, contact rows
    with their owner."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
    import dem_script
    psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10)
    psim._reclassify()
    psim.checkpoint_output(str(tmp_path / "ck"), 5)
    assert psim._vtk_due(5) and psim._vtk_due(10) and not psim._vtk_due(7)
    rng = np.random.default_rng(3)
    C = 20

    class FakeCtx:
        contact_capacity = C

        def __init__(self, n, lo, hi, seed):
            r = np.random.default_rng(seed)
            self.n, self.lo, self.hi = n, lo, hi
            self.pos = np.column_stack([lo + (hi - lo) * r.random(n), 0.015 * r.random(n), 0.04 * r.random(n)])
            self.arr = {k: r.standard_normal((n, w)).squeeze() * 1e-3 for k, w in
                        (("linear_velocity", 3), ("mass", 1), ("angular_velocity", 3), ("radius", 1), ("inv_inertia", 9),
                         ("rotation_matrix", 9), ("rotation_quat", 4), ("normal", 3))}
            self.i = {"uid": np.arange(n, dtype=np.int32) + 1000 * seed, "type": np.zeros(n, np.int32), "flags": np.zeros(n, np.int32),
                      "shape": np.zeros(n, np.int32)}
            self.num = r.integers(0, 4, n).astype(np.int32)

        def counts(self):
            return self.n, 0

        def ints(self, name):
            return self.i[name]

        def real(self, name):
            return self.pos if name == "position" else self.arr[name]

        def dem_download(self, name, n):
            return self.arr[name]

        def dem_download_contacts(self, n):
            r = np.random.default_rng(7)
            return {"num_contacts": self.num, "contact_lists": r.integers(1, 999, (n, C)).astype(np.int32),
                    "is_sticking": r.integers(0, 2, (n, C)).astype(np.int32), "tangential_spring_displacement": r.standard_normal((n, C, 3)),
                    "impact_velocity_magnitude": r.standard_normal((n, C))}

        def dem_download_contact_extras(self, n):          # two lanes of a further contact property
            return np.random.default_rng(8).standard_normal((n, C, 2))

        def decomposition(self):
            return {"nranks": np.array([2 if self.hi - self.lo < 0.09 else 1, 1, 1]), "subdom": np.array([self.lo, self.hi, 0.0, 0.015, 0.0, 0.04])}

    a, b = FakeCtx(7, 0.0, 0.05, 1), FakeCtx(5, 0.05, 0.1, 2)
    psim._checkpoint_write(a, 5, 0, 2)
    psim._checkpoint_write(b, 5, 1, 2)
    whole = psim._checkpoint_read(FakeCtx(0, 0.0, 0.1, 9), str(tmp_path / "ck"), 5)
    assert np.array_equal(whole["uid"], np.concatenate([a.i["uid"], b.i["uid"]]))
    assert np.array_equal(whole["position"], np.concatenate([a.pos, b.pos]))                 # bit for bit through "%.17g"
    for k in ("linear_velocity", "angular_velocity", "inv_inertia", "rotation_quat", "radius", "mass"):
        assert np.array_equal(whole[k], np.concatenate([a.arr[k], b.arr[k]])), k
    assert len(whole["contacts"]) == int(a.num.sum() + b.num.sum())
    assert set(whole["contacts"][:, 0].astype(int)) <= set(whole["uid"].tolist())
    first = np.nonzero(a.num)[0][0]                       # columns: the seven of dem.py's table, then the further lanes
    assert whole["contacts"].shape[1] == 9 and np.array_equal(whole["contacts"][0, 7:], a.dem_download_contact_extras(7)[first, 0])
    # re-split on other sub-boxes: every row lands on exactly one rank
    left = psim._checkpoint_read(FakeCtx(0, 0.0, 0.03, 9), str(tmp_path / "ck"), 5)
    right = psim._checkpoint_read(FakeCtx(0, 0.03, 0.1, 9), str(tmp_path / "ck"), 5)
    assert len(left["uid"]) + len(right["uid"]) >= len(whole["uid"]) - 1            # (a row within 1e-5 of the cut belongs to nobody,
    assert not set(left["uid"].tolist()) & set(right["uid"].tolist())               #  as in runtime/read_from_file.hpp:106)
    assert rng is not None


def test_contact_layout_puts_further_contact_properties_into_extra_lanes():
    """The first integer / vector / real contact property take examples/dem.py's three columns; every further one gets lanes of the
    contact's extra doubles (mapping/funcs.py:230-263 gives each add_contact_property() its own array), defaults included."""
    import dem_script
    psim = dem_script.build("gpu", (0.1, 0.015, 0.04), 10, more_contact_props=True)
    contact, defaults, extra = psim._contact_layout()
    assert contact == {"is_sticking": "c_stick", "tangential_spring_displacement": "c_tsd", "impact_velocity_magnitude": "c_ivm",
                       "tsd_seen": "cx:vec:0", "contact_age": "cx:real:3", "hits": "cx:int:4"}
    assert extra == [0.0, 0.0, 0.0, -1.0, 3.0] and defaults["c_ivm"] == 0.0
    assert [e["family"] for e in psim.functions] == ["gravity", "generic_pair", "euler"]
    name, src, nk = psim._translate_dem_model(psim.functions[1])
    assert src.startswith("#define PB_DEM_NX 5\n") and "cx[3] = " in src and nk == 1
    plain = dem_script.build("gpu", (0.1, 0.015, 0.04), 10)
    assert plain._contact_layout()[2] == []
