"""User-defined particle properties (csrc/props.cu, kernelgen.py) on the GPU: tests/scripts/props_script.py -- examples/md.py plus
six properties beyond the MD set, used by a setup() function, the pair kernel and both integrators -- against the run of the
REFERENCE's code generator on the same text (oracle/build_ref.py variant md_props_t1 -> tests/golden/md_props_t1.npz), and the
structural operations (sort, wrap, growth, ghosts, upload / download) through the C-ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import props_common as pc
from tests.util import by_id, rel_err_force

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "md_props_t1.npz"))


def test_user_property_script_matches_the_reference_generator_golden(capsys):
    import props_script
    z = _golden()
    # the first two iterations: the reference has not permuted anything yet, lattice order of the golden = upload order = tag
    psim1 = props_script.build("gpu", 8, 1, 20, 1)
    ctx1 = psim1.generate()
    capsys.readouterr()
    tag = ctx1.ints("tag")
    assert np.array_equal(by_id(tag, ctx1.download_property("scale")), z["scale_1"])          # the setup() function, bit for bit
    # the force of iteration 0 is lattice round-off (|f| ~ 1e-13, different digits in every summation order): the velocity it
    # kicks, and dt * v, agree to the last bits only
    assert np.abs(by_id(tag, ctx1.download_property("path")) - z["path_1"]).max() <= 1e-15
    for name, arr in (("force", ctx1.real("force")), ("pull", ctx1.download_property("pull")), ("work", ctx1.download_property("work"))):
        assert rel_err_force(by_id(tag, arr), z[f"{name}_1"]) <= 1e-12, name
    assert np.abs(by_id(tag, ctx1.download_property("heat")) - z["heat_1"]).max() <= 1e-13     # dt * f.v with f = that round-off
    assert np.array_equal(by_id(tag, ctx1.download_property("ups")), z["ups_1"].astype(np.float64))
    # the whole run: 100 steps, 6 reneighbourings (sort + wrap each time)
    psim = props_script.build("gpu", 8, 100, 20, 1)
    ctx = psim.generate()
    capsys.readouterr()
    assert len(psim.thermo_log) == 101
    for (ts, t, p), t_ref in zip(psim.thermo_log, z["temperature"]):
        assert abs(t - t_ref) <= 1e-9 * t_ref, ts
    assert ctx.counts() == (int(z["nlocal"][100]), int(z["nghost"][100]))
    # scale is different for every lattice site and never changes: it is the identity both runs are ordered by
    scale = ctx.download_property("scale")
    og, orf = np.argsort(scale), np.argsort(z["scale_100"])
    assert len(np.unique(scale)) == len(scale) and np.array_equal(scale[og], z["scale_100"][orf])
    for name, arr in (("position", ctx.real("position")), ("linear_velocity", ctx.real("linear_velocity")),
                      ("heat", ctx.download_property("heat")), ("work", ctx.download_property("work")),
                      ("path", ctx.download_property("path")), ("pull", ctx.download_property("pull")),
                      ("ups", ctx.download_property("ups"))):
        ref = z[f"{name}_100"][orf]
        assert np.abs(ref).max() > 0.0 and np.abs(arr[og] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), name
    box = 8 * pow(4.0 / 0.8442, 1.0 / 3.0)
    pc.check_identity(ctx.real("position"), ctx.download_property("path"), scale, box, props_script.XLEN)


def test_vocabulary_script_matches_the_reference_generator_golden(capsys):
    """tests/scripts/vocab_script.py (skip_when, cross, is_point_mass, integer properties and operators, n-ary min / max,
    normalized, ...) against the reference generator's run of the same text (variant md_vocab_t1): the module-level bit-for-bit
    pin is tests/test_kernelgen.py; here the NVRTC-compiled kernels run the program's first iterations on the device."""
    import vocab_script
    z = np.load(os.path.join(ROOT, "tests", "golden", "md_vocab_t1.npz"))
    psim = vocab_script.build("gpu", 8, 10, 20, 1)
    ctx = psim.generate()
    capsys.readouterr()
    assert len(psim.thermo_log) == 11 and ctx.counts() == (int(z["nlocal"][10]), int(z["nghost"][10]))
    for (ts, t, p), t_ref in zip(psim.thermo_log, z["temperature"]):
        assert abs(t - t_ref) <= 1e-9 * t_ref, ts
    psim1 = vocab_script.build("gpu", 8, 1, 20, 1)
    ctx1 = psim1.generate()
    capsys.readouterr()
    tag = ctx1.ints("tag")
    assert rel_err_force(by_id(tag, ctx1.real("force")), z["force_1"]) <= 1e-12
    assert np.abs(by_id(tag, ctx1.real("linear_velocity")) - z["linear_velocity_1"]).max() <= 1e-12


def test_generated_dem_contact_model_reproduces_the_built_in_run_bit_for_bit(capsys):
    """examples/dem.py with its contact model GENERATED (kernelgen.translate_dem_model -> NVRTC build of the contact kernel around
    it) instead of recognised: the model's arithmetic is the hand-written one's operation for operation
    (tests/test_kernelgen.py), so 300 iterations of the 420-sphere case -- falling, first impacts, sticking contacts -- end in
    identical bits: positions, velocities, angular velocities and the whole contact table."""
    import dem_script
    from pairs_b200 import dsl
    from tests import dem_common as dc
    ref_ctx = dem_script.build("gpu", dc.DOMAIN, 300).generate()
    dsl.FORCE_GENERIC_CONTACT_MODEL = True
    try:
        psim = dem_script.build("gpu", dc.DOMAIN, 300)
    finally:
        dsl.FORCE_GENERIC_CONTACT_MODEL = False
    assert [e["family"] for e in psim.functions] == ["gravity", "generic_pair", "euler"]
    ctx = psim.generate()
    capsys.readouterr()
    n = ctx.counts()[0]
    assert n == ref_ctx.counts()[0] == 422
    for name in ("position", "linear_velocity"):
        assert np.array_equal(ctx.real(name), ref_ctx.real(name)), name
    assert np.array_equal(ctx.dem_download("angular_velocity", n), ref_ctx.dem_download("angular_velocity", n))
    a, b = ctx.dem_download_contacts(n), ref_ctx.dem_download_contacts(n)
    assert a["num_contacts"].sum() > 100
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_contact_model_with_further_contact_properties_keeps_them_with_the_contact(capsys):
    """a contact table with MORE than examples/dem.py's three properties.  dem_script.build(more_contact_props=True)
    declares a second vector, a second real (default -1) and a second integer (default 3) contact property and appends
    `tsd_seen = tsd; age += 1.0; hits += 2` to dem.py's model.  The extra state changes no force, so after 300 iterations the run
    equals the built-in one bit for bit -- particles and the three standard columns -- while every live contact's further
    properties hold what those statements say: the displacement copy is the displacement, hits = 3 + 2 (age + 1), and the ages are
    the numbers of iterations the contacts have existed (their maximum reaches back to the first impacts).  Rows are compacted
    (clear_unused_contact_history) and spatially re-sorted during the run, so lanes that did not follow their slot would show."""
    import dem_script
    from tests import dem_common as dc
    ref_ctx = dem_script.build("gpu", dc.DOMAIN, 300).generate()
    psim = dem_script.build("gpu", dc.DOMAIN, 300, more_contact_props=True)
    assert [e["family"] for e in psim.functions] == ["gravity", "generic_pair", "euler"]
    ctx = psim.generate()
    capsys.readouterr()
    n = ctx.counts()[0]
    assert n == ref_ctx.counts()[0] == 422 and ctx.contact_extra_lanes == 5 and ref_ctx.contact_extra_lanes == 0
    for name in ("position", "linear_velocity"):
        assert np.array_equal(ctx.real(name), ref_ctx.real(name)), name
    assert np.array_equal(ctx.dem_download("angular_velocity", n), ref_ctx.dem_download("angular_velocity", n))
    a, b = ctx.dem_download_contacts(n), ref_ctx.dem_download_contacts(n)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    cx = ctx.dem_download_contact_extras(n)
    live = np.arange(cx.shape[1])[None, :] < a["num_contacts"][:, None]
    assert live.sum() > 100
    assert np.array_equal(cx[live][:, :3], a["tangential_spring_displacement"][live])
    age, hits = cx[live][:, 3], cx[live][:, 4]
    assert np.array_equal(hits, 3.0 + 2.0 * (age + 1.0)) and age.min() >= 0.0 and np.array_equal(age, np.round(age))
    assert 20 <= age.max() < 300 and len(np.unique(age)) > 10
    # the host <-> device path of the lanes: what is uploaded comes back
    back = np.random.default_rng(3).standard_normal(cx.shape)
    ctx.dem_upload_contact_extras(back)
    assert np.array_equal(ctx.dem_download_contact_extras(n), back)


def test_further_contact_properties_match_the_reference_generated_code(capsys):
    """The same script through the REFERENCE: oracle/build_ref.py variant dem_more_t1 is examples/dem.py with the three further
    contact properties and the three statements, generated and compiled by the reference itself; tests/golden/dem_more_t1.npz holds
    what its contact tables contain at the end of iteration 300.  Ours: same particles (positions to 1e-12), the same contacts
    (partner uids per particle), and in every one of them the same age and hit count exactly and the displacement copy to 1e-12."""
    import dem_script
    from tests import dem_common as dc
    z = np.load(os.path.join(ROOT, "tests", "golden", "dem_more_t1.npz"))
    ctx = dem_script.build("gpu", dc.DOMAIN, 300, more_contact_props=True).generate()
    capsys.readouterr()
    n = ctx.counts()[0]
    assert n == int(z["nlocal"][0]) == 422
    o, r = np.argsort(ctx.ints("uid")), np.argsort(z["uid"])
    pref = z["position"][r]
    assert np.abs(ctx.real("position")[o] - pref).max() <= 1e-12 * np.abs(pref[:n - 2]).max()
    c, cx = ctx.dem_download_contacts(n), ctx.dem_download_contact_extras(n)
    assert np.array_equal(c["num_contacts"][o], z["num_contacts"][r]) and z["num_contacts"].sum() > 100
    scale = np.abs(z["tsd_seen"]).max()
    checked = 0
    for a, b in zip(o, r):
        m = int(z["num_contacts"][b])
        ours = {int(c["contact_lists"][a, k]): cx[a, k] for k in range(m)}
        ref = {int(z["contact_lists"][b, k]): (z["tsd_seen"][b, k], z["contact_age"][b, k], z["hits"][b, k]) for k in range(m)}
        assert set(ours) == set(ref)
        for partner, (seen, age, hits) in ref.items():
            lanes = ours[partner]
            assert lanes[3] == age and lanes[4] == float(hits), (partner, lanes, age, hits)
            assert np.abs(lanes[:3] - seen).max() <= 1e-12 * scale
            checked += 1
    assert checked == int(z["num_contacts"].sum())


def test_dem_script_with_a_generated_per_particle_kernel_reproduces_the_built_in_run(capsys):
    """A DEM procedure list that is not exactly gravity / model / euler runs module by module with the user bodies generated: here
    gravity, euler and the set-up function update_mass_and_inertia are sent through the generic path (matrix / quaternion algebra
    included; the built-in and the generated kernel call the same device sin / cos), 150 iterations (before the native loop's first spatial re-sort) must end in
    the bits of the native run."""
    import dem_script
    from pairs_b200 import dsl
    from tests import dem_common as dc
    ref_ctx = dem_script.build("gpu", dc.DOMAIN, 150).generate()
    dsl.FORCE_GENERIC_NAMES = {"gravity", "euler", "update_mass_and_inertia"}
    try:
        psim = dem_script.build("gpu", dc.DOMAIN, 150)
    finally:
        dsl.FORCE_GENERIC_NAMES = set()
    assert [e["family"] for e in psim.functions] == ["generic_particle", "linear_spring_dashpot", "generic_particle"]
    assert [e["family"] for e in psim.setup_functions] == ["generic_setup"]
    ctx = psim.generate()
    capsys.readouterr()
    n = ctx.counts()[0]
    for name in ("position", "linear_velocity", "force"):
        assert np.array_equal(ctx.real(name), ref_ctx.real(name)), name
    for name in ("angular_velocity", "rotation_quat", "rotation_matrix", "inv_inertia"):
        assert np.array_equal(ctx.dem_download(name, n), ref_ctx.dem_download(name, n)), name
    a, b = ctx.dem_download_contacts(n), ref_ctx.dem_download_contacts(n)
    assert np.array_equal(a["num_contacts"], b["num_contacts"]) and a["num_contacts"].sum() > 0       # first impacts only at 150 iterations
    live = np.arange(a["contact_lists"].shape[1])[None, :] < a["num_contacts"][:, None]      # dead slots keep what the last tenant left
    for k in a:
        if k != "num_contacts":
            assert np.array_equal(a[k][live], b[k][live]), k


def test_dem_script_with_a_user_property_and_an_extra_kernel(capsys):
    """A DEM script that declares a further property (the distance every sphere has travelled) and a further per-particle kernel
    that integrates it: the trajectory is that of the plain script (the extra kernel touches nothing else; 150 iterations, staged
    loop), and the travelled distance equals what numpy integrates from the velocities of consecutive iterations."""
    import dem_script
    import pairs
    from tests import dem_common as dc

    def odometer(i):
        travelled[i] += dt * length(linear_velocity[i])

    ref_ctx = dem_script.build("gpu", dc.DOMAIN, 150).generate()
    psim = dem_script.build("gpu", dc.DOMAIN, 150)
    psim.add_property('travelled', pairs.real(), 0.0)
    psim.compute(odometer, symbols={'dt': 5e-5})
    assert [e["family"] for e in psim.functions] == ["gravity", "linear_spring_dashpot", "euler", "generic_particle"]
    ctx = psim.generate()
    capsys.readouterr()
    assert np.array_equal(ctx.real("position"), ref_ctx.real("position")) and np.array_equal(ctx.real("linear_velocity"), ref_ctx.real("linear_velocity"))
    trav = ctx.download_property("travelled")
    n = len(trav)
    # spheres start at ~1 m/s (150 integrated iterations of 5e-5 s; the lowest ones land after ~1 mm and creep from then on), most of
    # them are still falling at the end; the two FIXED half-spaces do not move
    moving = (ctx.ints("flags") & 4) == 0
    assert moving.sum() == n - 2 and np.all(trav[~moving] == 0.0)
    assert np.all(trav[moving] > 2e-4) and np.all(trav[moving] < 3.0 * 150 * 5e-5)
    assert 0.8 * 150 * 5e-5 < np.median(trav[moving]) < 1.3 * 150 * 5e-5


def test_md_script_with_cell_lists_only_matches_the_reference_run_without_neighbour_lists(capsys):
    """examples/md.py with build_cell_lists() instead of build_neighbor_lists(): the pair kernel is generated to walk cell 0 and the
    27 stencil cells.  That is NOT the neighbour-list run: a pair that was farther apart than cutoff + skin when the lists were
    built never enters them, but the cell walk evaluates it as soon as it is inside the cutoff (in the reference's own two runs the
    temperatures part at iteration 15, by 3.5e-7).  Golden = the reference generator's run of the same script
    (oracle/build_ref.py variant md_cells_t1 -> tests/golden/md_cells_t1.npz): thermo of 61 iterations to 1e-9, forces of iterations
    1 and 20 to 1e-12, and the run must differ from the neighbour-list golden exactly where the reference's does."""
    import lj_script
    z = np.load(os.path.join(ROOT, "tests", "golden", "md_cells_t1.npz"))
    zl = np.load(os.path.join(ROOT, "tests", "golden", "md_t1.npz"))
    psim = lj_script.build("gpu", 8, 60, 20, 1, cells_only=True)
    ctx = psim.generate()
    assert psim.functions[0]["family"] in ("generic_pair", "lennard_jones") and len(psim.thermo_log) == 61
    for (ts, t, p), t_ref in zip(psim.thermo_log, z["temperature"]):
        assert abs(t - t_ref) <= 1e-9 * t_ref, (ts, t, t_ref)
    assert abs(psim.thermo_log[15][1] - zl["temperature"][15]) > 1e-8 * zl["temperature"][15]
    assert ctx.counts() == (int(z["nlocal"][-1]), int(z["nghost"][-1]))
    for steps in (1, 20):
        ps = lj_script.build("gpu", 8, steps, 20, 1, cells_only=True)
        c = ps.generate()
        f = by_id(c.ints("tag"), c.real("force"))
        fr = z[f"force_{steps}"]
        if steps == 20:      # the reference re-numbers its particles when they wrap; identify them through the exact lattice velocities? no:
            o, r = np.lexsort(c.real("position")[np.argsort(c.ints("tag"))].T[::-1]), np.lexsort(z["position_20"].T[::-1])
            assert np.abs(c.real("position")[np.argsort(c.ints("tag"))][o] - z["position_20"][r]).max() <= 1e-9
            f, fr = f[o], fr[r]
        assert rel_err_force(f, fr) <= 1e-12, steps
    capsys.readouterr()


def test_generated_pair_kernel_with_compute_half_matches_the_built_in_half_kernel(capsys):
    """compute_half() with the pair kernel generated (every pair once, atomic update of the partner): thermo of 40 iterations to 1e-9
    and the forces of iteration 1 to 1e-12 against the hand-written half-list kernel, which is pinned to the reference run with
    compute_half() enabled (tests/golden/md_half_t1.npz)."""
    import lj_script
    from pairs_b200 import dsl
    runs = {}
    for generic in (False, True):
        out = []
        for steps in (40, 1):
            dsl.FORCE_GENERIC_NAMES = {"lennard_jones"} if generic else set()
            try:
                psim = lj_script.build("gpu", 8, steps, 20, 1)
            finally:
                dsl.FORCE_GENERIC_NAMES = set()
            psim.compute_half()
            assert psim.functions[0]["family"] == ("generic_pair" if generic else "lennard_jones")
            out.append((psim, psim.generate()))
        runs[generic] = (out[0][0].thermo_log, by_id(out[1][1].ints("tag"), out[1][1].real("force")), out[0][1].ints("numneighs").mean())
    capsys.readouterr()
    (th_a, f_a, nn_a), (th_b, f_b, nn_b) = runs[False], runs[True]
    assert len(th_a) == len(th_b) == 41 and nn_a == nn_b and nn_a < 60
    for (ts, t, p), (_, t2, p2) in zip(th_a, th_b):
        assert abs(t - t2) <= 1e-9 * t and abs(p - p2) <= 1e-9 * abs(p), ts
    assert rel_err_force(f_b, f_a) <= 1e-12


def test_dem_script_with_a_reneighbouring_interval_matches_the_reference(capsys):
    """examples/dem.py with psim.reneighbor_every(3): exchange / borders / cell lists every third iteration, the ghosts' positions,
    linear AND angular velocities refreshed by synchronize in between (module-by-module loop).  State after iteration 300 against the
    reference's generated C++ for the same script (oracle variant dem_rn3_t1 -> tests/golden/dem_rn3_t1.npz), matched through uid."""
    import dem_script
    from tests import dem_common as dc
    z = np.load(os.path.join(ROOT, "tests", "golden", "dem_rn3_t1.npz"))
    psim = dem_script.build("gpu", dc.DOMAIN, 300, reneighbor=3)
    ctx = psim.generate()
    capsys.readouterr()
    n = 422
    assert ctx.counts()[0] == n
    o, r = np.argsort(ctx.ints("uid")), np.argsort(z["end_300_uid"])
    pref = z["end_300_position"][r]
    assert np.abs(ctx.real("position")[o] - pref).max() <= 1e-12 * np.abs(pref[:n - 2]).max()
    vref = z["end_300_linear_velocity"][r]
    assert np.abs(ctx.real("linear_velocity")[o] - vref).max() <= 1e-9 * np.abs(vref).max()
    c = ctx.dem_download_contacts(n)
    assert np.array_equal(c["num_contacts"][o], z["end_300_num_contacts"][r]) and c["num_contacts"].sum() > 50


def test_property_store_through_the_c_abi(capsys):
    """add / upload / download, defaults, capacity growth, ghosts carrying their source's values, volatile reset, the cell-order
    sort -- without any generated kernel."""
    from pairs_b200 import backend
    rng = np.random.default_rng(5)
    L = 12.0
    ctx = backend.Context(0)
    ctx.init_domain([0.0, L, 0.0, L, 0.0, L])
    ctx.add_property("q", 1, False, [1.5])
    n = 3000
    pos = rng.random((n, 3)) * L
    ctx.upload(pos, None, None, None, None, None, None)
    ctx.add_property("w", 3, False, [0.25, 0.5, 0.75])          # declared after the particles exist: rows are re-allocated
    ctx.add_property("acc", 3, True, None)
    assert np.all(ctx.download_property("q") == 1.5) and np.all(ctx.download_property("w") == [0.25, 0.5, 0.75])
    q = rng.random(n)
    w = rng.random((n, 3))
    ctx.upload_property("q", q)
    ctx.upload_property("w", w)
    ctx.upload_property("acc", np.ones((n, 3)))
    assert np.array_equal(ctx.download_property("q"), q) and np.array_equal(ctx.download_property("w"), w)
    ctx.setup_cells(2.8)
    ctx.exchange()                                   # wrap + cell-order sort
    ctx.borders()                                    # ghosts; grows the particle capacity (n + n/4 + 1024 slots reserved at upload)
    tag = ctx.ints("tag")
    assert not np.array_equal(tag, np.arange(n)) and np.array_equal(by_id(tag, ctx.download_property("q")), q)
    assert np.array_equal(by_id(tag, ctx.download_property("w")), w)
    nl, ng = ctx.counts()
    assert nl == n and ng > n // 2
    # every ghost carries the non-volatile values of the particle it is an image of
    tags_all = ctx.ints("tag", with_ghosts=True)
    assert np.array_equal(ctx.download_property("q", with_ghosts=True), q[tags_all])
    assert np.array_equal(ctx.download_property("w", with_ghosts=True), w[tags_all])
    ctx.reset_volatile()
    assert not ctx.download_property("acc").any() and np.array_equal(by_id(tag, ctx.download_property("w")), w)
    # a second, larger upload: capacity growth re-strides the rows, new particles start from the declared defaults
    n2 = 9000
    ctx.upload(rng.random((n2, 3)) * L, None, None, None, None, None, None)
    assert np.all(ctx.download_property("q") == 1.5) and np.all(ctx.download_property("w") == [0.25, 0.5, 0.75])
    with pytest.raises(backend.BackendError, match="already defined"):
        ctx.add_property("q", 1)


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2])
def test_further_contact_properties_migrate_with_their_contact(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29850 + world), os.path.join(ROOT, "tests", "scripts", "mgpu_contact_props_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu_contact_props_check ok" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_user_properties_follow_their_particle_between_ranks(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29800 + world), os.path.join(ROOT, "tests", "scripts", "mgpu_props_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu_props_check ok" in r.stdout, r.stdout[-4000:]
