"""checkpoint_output() / read_checkpoint() (checkpoint dump of per-uid state for restartable runs): a run that
is interrupted, written to disk and continued by a NEW simulation ends where the uninterrupted run ends."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))


def _by(ids, a):
    return np.asarray(a)[np.argsort(ids)]


def test_md_run_continues_from_a_checkpoint(tmp_path, capsys):
    """The continuation from the files equals, bit for bit, the continuation from the same state handed over through the C-ABI
    (nothing is lost or rounded on disk), and stays close to the uninterrupted run.  "Close", not 1e-9: between reneighbourings
    the reference refreshes forwarded (edge / corner) ghosts one step late (sim/comm.py:45-54), a restart
    begins with fresh ghosts -- the reference's own restart would differ from its uninterrupted run in the same way."""
    import lj_script
    from pairs_b200.backend import Context
    from tests.test_gpu_md import CUT, DT, SKIN, box
    prefix = str(tmp_path / "md")
    nx = 6
    # uninterrupted: iterations 0..60; the checkpoint holds the state after iteration 40
    whole = lj_script.build("gpu", nx, 60, 20, 0, checkpoint=(prefix, 40)).generate()
    assert os.path.exists(prefix + "_40.csv") and os.path.exists(prefix + "_40.json") and os.path.exists(prefix + "_0.csv")
    # continued: a new simulation reads it and runs the remaining 20 iterations (its own counter starts at 0: iteration 0 has no
    # integration step, so 20 further integrations = timesteps 20)
    cont = lj_script.build("gpu", nx, 20, 20, 0, restart=(prefix, 40)).generate()
    # the same hand-over without files: the state after iteration 40 of an identical run, uploaded as arrays
    first = lj_script.build("gpu", nx, 40, 20, 0).generate()
    capsys.readouterr()
    direct = Context(0)
    direct.init_domain(box(nx))
    direct.setup_cells(CUT + SKIN)
    direct.set_lj_params(4, [1.0] * 16, [1.0] * 16)
    direct.upload(first.real("position"), first.real("linear_velocity"), first.real("mass"), first.ints("type"), first.ints("flags"),
                  first.ints("uid"), first.ints("shape"))
    direct.md_run(0, 21, DT, CUT, CUT + SKIN, CUT + SKIN, 20, 0)
    assert whole.counts()[0] == cont.counts()[0] == direct.counts()[0] == 4 * nx ** 3
    for name in ("position", "linear_velocity", "force"):
        assert np.array_equal(_by(cont.ints("tag"), cont.real(name)), _by(direct.ints("tag"), direct.real(name))), name
    # against the uninterrupted run (rows of the checkpoint = device order of `first` = tags of `cont`)
    xa = np.empty((4 * nx ** 3, 3))
    xa[:] = np.nan
    ta = first.ints("tag")                       # tag of the uninterrupted run's particle in row k of the checkpoint
    xw = _by(whole.ints("tag"), whole.real("position"))
    xc = _by(cont.ints("tag"), cont.real("position"))
    d = xc - xw[ta]
    L = box(nx)[1]
    d -= L * np.round(d / L)
    assert np.abs(d).max() <= 5e-3 and np.abs(d).mean() <= 2e-4


def test_dem_run_continues_from_a_checkpoint_with_its_contact_history(tmp_path, capsys):
    import dem_script
    from tests import dem_common as dc
    prefix = str(tmp_path / "dem")
    whole = dem_script.build("gpu", dc.DOMAIN, 330, checkpoint=(prefix, 300)).generate()      # contacts exist from ~150 on
    assert os.path.exists(prefix + "_300.contacts.csv")
    rows = np.loadtxt(prefix + "_300.contacts.csv", delimiter=",", ndmin=2)
    assert len(rows) > 100 and rows[:, 2].max() == 1            # live contacts, some of them sticking
    # (dem.py integrates in iteration 0 too: iterations 301..330 of the first run are iterations 0..29 of the continuation)
    cont = dem_script.build("gpu", dc.DOMAIN, 29, restart=(prefix, 300)).generate()
    capsys.readouterr()
    n = whole.counts()[0]
    assert cont.counts()[0] == n == 422
    ua, ub = whole.ints("uid"), cont.ints("uid")
    assert np.array_equal(np.sort(ua), np.sort(ub))
    scale = np.abs(whole.real("position")[:n - 2]).max()
    assert np.abs(_by(ua, whole.real("position")) - _by(ub, cont.real("position"))).max() <= 1e-10 * scale
    for name in ("angular_velocity", "rotation_quat"):
        a, b = _by(ua, whole.dem_download(name, n)), _by(ub, cont.dem_download(name, n))
        assert np.abs(a - b).max() <= 1e-8 * max(np.abs(a).max(), 1.0), name
    ca, cb = whole.dem_download_contacts(n), cont.dem_download_contacts(n)
    assert np.array_equal(_by(ua, ca["num_contacts"]), _by(ub, cb["num_contacts"]))
    sa = dc.contact_sets(_by(ua, ca["num_contacts"]), _by(ua, ca["contact_lists"]), _by(ua, ca["is_sticking"]),
                         _by(ua, ca["tangential_spring_displacement"]), _by(ua, ca["impact_velocity_magnitude"]), n)
    sb = dc.contact_sets(_by(ub, cb["num_contacts"]), _by(ub, cb["contact_lists"]), _by(ub, cb["is_sticking"]),
                         _by(ub, cb["tangential_spring_displacement"]), _by(ub, cb["impact_velocity_magnitude"]), n)
    assert [set(x) for x in sa] == [set(x) for x in sb]                    # the same partners per particle ...
    assert sum(x[k][0] for x in sa for k in x) == sum(x[k][0] for x in sb for k in x)      # ... with the same sticking flags
