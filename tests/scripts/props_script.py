"""examples/md.py with USER-DEFINED PROPERTIES -- the text of oracle/build_ref.py's variant md_props_t1: beyond position / mass /
velocity / force the script declares a per-particle real that enters the pair force (scale, written once by a setup() function; its value is different for every
lattice site, so it doubles as the particle identity when the final states are compared),
a second volatile vector accumulated by the pair kernel (pull), two reals and a vector integrated by the per-particle kernels
(heat, work, path) and an integer counter (ups).  The non-volatile ones have to follow their particle through the cell-order sort, the periodic wrap and the
migration between ranks; none of the kernels is a hand-written family: they run through the generic path
(pairs_b200/kernelgen.py -> NVRTC) on the user-property rows of csrc/props.cu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pairs  # noqa: E402


def init_scale(i):
    scale[i] = 1.0 + 0.25 * (position[i][0] + 0.3719 * position[i][1] + 0.1137 * position[i][2]) / xlen


def lennard_jones(i, j):
    sr2 = 1.0 / squared_distance(i, j)
    sr6 = sr2 * sr2 * sr2 * sigma6[i, j]
    f = 48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon[i, j] * scale[i]
    apply(force, delta(i, j) * f)
    apply(pull, delta(i, j) * (sr6 * scale[i]))


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    position[i] += dt * linear_velocity[i]
    path[i] += dt * linear_velocity[i]
    heat[i] += dt * dot(force[i], linear_velocity[i])


def final_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    work[i] = work[i] + dot(pull[i], linear_velocity[i])
    if dot(pull[i], linear_velocity[i]) > 0.0:
        ups[i] += 1 + (uid[i] & 1)


XLEN = 13.0


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
    psim.add_property('scale', pairs.real(), 1.0)
    psim.add_property('heat', pairs.real(), 0.0)
    psim.add_property('work', pairs.real(), 0.0)
    psim.add_property('path', pairs.vector())
    psim.add_property('pull', pairs.vector(), volatile=True)
    psim.add_property('ups', pairs.int32(), 0)
    psim.add_feature('type', ntypes)
    psim.add_feature_property('type', 'epsilon', pairs.real(), [sigma for i in range(ntypes * ntypes)])
    psim.add_feature_property('type', 'sigma6', pairs.real(), [epsilon for i in range(ntypes * ntypes)])
    psim.copper_fcc_lattice(nx, nx, nx, rho, temp, ntypes)
    psim.set_domain_partitioner(pairs.regular_domain_partitioner())
    psim.compute_thermo(thermo)
    psim.setup(init_scale, symbols={'xlen': XLEN})
    psim.reneighbor_every(reneigh)
    psim.build_neighbor_lists(cutoff_radius + skin)
    psim.compute(initial_integrate, symbols={'dt': dt}, pre_step=True, skip_first=True)
    psim.compute(lennard_jones, cutoff_radius)
    psim.compute(final_integrate, symbols={'dt': dt}, skip_first=True)
    return psim


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else "gpu").generate()
