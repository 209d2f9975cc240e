"""A user script against the `pairs` DSL (same calls and kernel bodies a P4IRS user writes, cf. the reference's
examples/md.py), parameterised for the tests:  python lj_script.py gpu [nx] [timesteps] [reneighbor] [thermo]"""
import sys

import pairs


def lennard_jones(i, j):
    sr2 = 1.0 / squared_distance(i, j)
    sr6 = sr2 * sr2 * sr2 * sigma6[i, j]
    apply(force, delta(i, j) * (48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j]))


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    position[i] += dt * linear_velocity[i]


def final_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]


def build(target="gpu", nx=8, timesteps=100, reneighbor=20, thermo=10, ntypes=4, cells_only=False, checkpoint=None, restart=None):
    dt = 0.005
    cutoff_radius = 2.5
    skin = 0.3
    psim = pairs.simulation("md", [pairs.point_mass()], timesteps=timesteps, double_prec=True)
    psim.target(pairs.target_gpu() if target == "gpu" else pairs.target_cpu())
    psim.add_position('position')
    psim.add_property('mass', pairs.real(), 1.0)
    psim.add_property('linear_velocity', pairs.vector())
    psim.add_property('force', pairs.vector(), volatile=True)
    psim.add_feature('type', ntypes)
    psim.add_feature_property('type', 'epsilon', pairs.real(), [1.0 for _ in range(ntypes * ntypes)])
    psim.add_feature_property('type', 'sigma6', pairs.real(), [1.0 for _ in range(ntypes * ntypes)])
    if restart is not None:          # continue from a checkpoint written by checkpoint_output()
        a = pow(4.0 / 0.8442, 1.0 / 3.0)
        psim.set_domain([0.0, 0.0, 0.0, nx * a, nx * a, nx * a])
        psim.read_checkpoint(*restart)
    else:
        psim.copper_fcc_lattice(nx, nx, nx, 0.8442, 1.44, ntypes)
    if checkpoint is not None:
        psim.checkpoint_output(checkpoint[0], checkpoint[1])
    psim.set_domain_partitioner(pairs.regular_domain_partitioner())
    psim.compute_thermo(thermo)
    psim.reneighbor_every(reneighbor)
    if cells_only:          # no Verlet lists: pair kernels walk the cell lists (sim/interaction.py:92-118)
        psim.build_cell_lists(cutoff_radius + skin)
    else:
        psim.build_neighbor_lists(cutoff_radius + skin)
    psim.compute(initial_integrate, symbols={'dt': dt}, pre_step=True, skip_first=True)
    psim.compute(lennard_jones, cutoff_radius)
    psim.compute(final_integrate, symbols={'dt': dt}, skip_first=True)
    return psim


if __name__ == "__main__":
    a = sys.argv[1:]
    psim = build(a[0] if a else "gpu", *[int(x) for x in a[1:5]])
    psim.generate()
