"""CPU unit test of the PRODUCT's DEM arithmetic (pairs_b200/csrc/dem_math.h compiled for the host) against golden module-boundary
states produced by the reference's generated C++ (tests/golden/make_golden_dem.py): contact history bit-exact per partner uid,
euler / inertia bit-exact, forces and torques to summation order."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from tests import dem_common as dc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = ctypes.POINTER(ctypes.c_double)
I = ctypes.POINTER(ctypes.c_int)


def dp(a):
    return a.ctypes.data_as(D)


def ip(a):
    return a.ctypes.data_as(I)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("demhost") / "libdem_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "pairs_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host", "dem_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    P = (ctypes.c_double * 16)()
    lib.host_dem_params(P, *[ctypes.c_double(x) for x in (dc.DT, math.pi, dc.KAPPA, dc.LN_DRY, dc.COLLISION_TIME, dc.RHO_P, dc.RHO_F, dc.G)])
    return lib, P


@pytest.mark.parametrize("ts", [150, 300, 400])
def test_contact_kernel_and_euler_match_reference(host, ts):
    lib, P = host
    z = dc.gold()
    nl = int(z["nlocal"][ts])
    g = lambda tag, k: np.ascontiguousarray(z[f"{tag}_{ts}_{k}"])      # noqa: E731
    ntot = len(g("pre", "mass"))
    force, torque = g("pre", "force").copy(), g("pre", "torque").copy()
    num, cu = g("pre", "num_contacts").copy(), g("pre", "contact_lists").copy()
    used = np.zeros_like(cu)
    st, tsd, ivm = g("pre", "is_sticking").copy(), g("pre", "tangential_spring_displacement").copy(), g("pre", "impact_velocity_magnitude").copy()
    fs, fd = np.array(dc.FS), np.array(dc.FD)
    lib.host_dem_contacts(P, nl, ntot, dc.C, dc.NTYPES, dp(g("pre", "position")), dp(g("pre", "linear_velocity")), dp(g("pre", "angular_velocity")),
                          dp(g("pre", "mass")), dp(g("pre", "radius")), dp(g("pre", "normal")), ip(g("pre", "flags")), ip(g("pre", "shape")),
                          ip(g("pre", "uid")), ip(g("pre", "type")), dp(fs), dp(fd), ip(num), ip(cu), ip(used), ip(st), dp(tsd), dp(ivm),
                          dp(force), dp(torque))
    rf, rt = g("post", "force"), g("post", "torque")
    assert np.abs(force[:nl] - rf[:nl]).max() <= 1e-12 * np.abs(rf[:nl]).max()
    assert np.abs(torque[:nl] - rt[:nl]).max() <= 1e-12 * max(np.abs(rt[:nl]).max(), 1e-300)
    assert np.array_equal(num[:nl], g("post", "num_contacts")[:nl]) and num[:nl].sum() > 0
    ours = dc.contact_sets(num, cu, st, tsd, ivm, nl)
    ref = dc.contact_sets(g("post", "num_contacts"), g("post", "contact_lists"), g("post", "is_sticking"),
                          g("post", "tangential_spring_displacement"), g("post", "impact_velocity_magnitude"), nl)
    assert ours == ref                                     # DEM contact-history bookkeeping: bit-exact
    # euler from the reference's post-contact state
    pos, vel, w = g("pre", "position").copy(), g("pre", "linear_velocity").copy(), g("pre", "angular_velocity").copy()
    q, R = g("pre", "rotation_quat").copy(), g("pre", "rotation_matrix").copy()
    lib.host_dem_euler(P, nl, ip(g("pre", "flags")), dp(g("pre", "mass")), dp(rf), dp(rt), dp(g("pre", "inv_inertia")), dp(pos), dp(vel), dp(w),
                       dp(q), dp(R))
    for a, k in ((pos, "position"), (vel, "linear_velocity"), (w, "angular_velocity"), (q, "rotation_quat"), (R, "rotation_matrix")):
        assert np.array_equal(a[:nl], g("eul", k)[:nl]), k


def test_inverse_inertia_of_spheres_is_bit_exact(host):
    lib, _ = host
    z = dc.gold()
    m, r, ref = z["pre_150_mass"], z["pre_150_radius"], z["pre_150_inv_inertia"]
    for i in range(0, 400, 37):
        out = np.zeros(9)
        lib.host_dem_sphere_inv_inertia(ctypes.c_double(m[i]), ctypes.c_double(r[i]), dp(out))
        assert np.array_equal(out, ref[i])


def test_reference_golden_of_further_contact_properties_is_what_the_statements_say():
    """tests/golden/dem_more_t1.npz (the reference's generated C++ for examples/dem.py plus three further contact properties,
    tests/golden/make_golden_dem.py): in every live contact the displacement copy equals the displacement bit for bit, the age
    counts iterations from the default -1, hits = 3 + 2 (age + 1); the particles are those of the plain run (dem_t1): further
    contact state changes no force.  This is the behaviour the extra lanes of the contact rows reproduce (tests/test_gpu_props.py)."""
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, t1 = np.load(os.path.join(gold, "dem_more_t1.npz")), np.load(os.path.join(gold, "dem_t1.npz"))
    live = np.arange(20)[None, :] < z["num_contacts"][:, None]
    assert live.sum() > 100
    assert np.array_equal(z["tsd_seen"][live], z["tangential_spring_displacement"][live])
    age, hits = z["contact_age"][live], z["hits"][live]
    assert age.min() >= 0.0 and age.max() >= 20.0 and np.array_equal(age, np.round(age))
    assert np.array_equal(hits, (3 + 2 * (age + 1)).astype(np.int32))
    assert np.array_equal(z["position"], t1["end_300_position"]) and np.array_equal(z["contact_lists"][live], t1["end_300_contact_lists"][live])
