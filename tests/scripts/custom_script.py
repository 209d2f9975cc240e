"""examples/md.py with OTHER kernel bodies -- the text of oracle/build_ref.py's variant md_custom_t1: a softened Lennard-Jones
pair kernel (sqrt, select, two extra symbols, a non-uniform epsilon table) and integrators with a drag term.  None of them is a
hand-written kernel family: they run through the generic path (pairs_b200/kernelgen.py -> NVRTC)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pairs  # noqa: E402


def lennard_jones(i, j):
    rsq = squared_distance(i, j)
    r = sqrt(rsq)
    sr2 = 1.0 / rsq
    sr6 = sr2 * sr2 * sr2 * sigma6[i, j]
    f = 48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j] + select(r < rsoft, kspring * (rsoft - r) / r, 0.0)
    apply(force, delta(i, j) * f)


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * (force[i] - gamma * linear_velocity[i]) / mass[i]
    position[i] += dt * linear_velocity[i]


def final_integrate(i):
    linear_velocity[i] += (dt * 0.5) * (force[i] - gamma * linear_velocity[i]) / mass[i]


def build(target="gpu", nx=8, timesteps=100, reneigh=20, thermo=1):
    dt = 0.005
    cutoff_radius = 2.5
    skin = 0.3
    ntypes = 4
    sigma = 1.0
    epsilon = 1.0
    sigma6 = sigma ** 6
    rho = 0.8442
    temp = 1.44
    psim = pairs.simulation("md", [pairs.point_mass()], timesteps=timesteps, double_prec=True)
    psim.target(pairs.target_gpu() if target == "gpu" else pairs.target_cpu())
    psim.add_position('position')
    psim.add_property('mass', pairs.real(), 1.0)
    psim.add_property('linear_velocity', pairs.vector())
    psim.add_property('force', pairs.vector(), volatile=True)
    psim.add_feature('type', ntypes)
    psim.add_feature_property('type', 'epsilon', pairs.real(), [1.0 + 0.05 * ((i % ntypes) + (i // ntypes)) for i in range(ntypes * ntypes)])
    psim.add_feature_property('type', 'sigma6', pairs.real(), [epsilon for i in range(ntypes * ntypes)])
    psim.copper_fcc_lattice(nx, nx, nx, rho, temp, ntypes)
    psim.set_domain_partitioner(pairs.regular_domain_partitioner())
    psim.compute_thermo(thermo)
    psim.reneighbor_every(reneigh)
    psim.build_neighbor_lists(cutoff_radius + skin)
    psim.compute(initial_integrate, symbols={'dt': dt, 'gamma': 0.05}, pre_step=True, skip_first=True)
    psim.compute(lennard_jones, cutoff_radius, symbols={'kspring': 3.5, 'rsoft': 1.05})
    psim.compute(final_integrate, symbols={'dt': dt, 'gamma': 0.05}, skip_first=True)
    return psim


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else "gpu").generate()
