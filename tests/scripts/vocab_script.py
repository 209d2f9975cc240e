"""Kernels that use the part of the generic vocabulary the other scripts do not (the text of oracle/build_ref.py's variant
md_vocab_t1): skip_when, cross, is_point_mass, the integer properties uid / shape / type with % & | ^ ~, and / or / not, n-ary
min / max, normalized, length, squared_length, dot, zero_vector.  Only the translation of these functions is under test
(tests/test_kernelgen.py runs kernelgen's output against the modules the reference generator printed for the same text)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pairs  # noqa: E402


def lennard_jones(i, j):
    skip_when(uid[j] % 7 == 3)
    d = delta(i, j)
    dv = linear_velocity[j] - linear_velocity[i]
    w = cross(linear_velocity[i], dv)
    s = select(is_point_mass(j), 1.0, 0.5)
    m = (uid[i] & 3) + (uid[j] | 1) + (type[i] ^ type[j]) + (~uid[j] & 1)
    sr2 = 1.0 / squared_distance(i, j)
    u = normalized(dv)
    c = min(dot(u, linear_velocity[i]), length(dv), 1.5) + max(squared_length(dv), 0.25)
    apply(force, w * s + d * (sr2 * m) + u * c + zero_vector())


def initial_integrate(i):
    linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]
    position[i] += dt * linear_velocity[i]


def final_integrate(i):
    skip_when(not is_point_mass(i))
    if uid[i] % 2 == 0 and mass[i] > 0.5 or shape[i] == 1:
        linear_velocity[i] += (dt * 0.5) * force[i] / mass[i]


def build(target="gpu", nx=8, timesteps=10, reneigh=20, thermo=1):
    dt = 0.005
    cutoff_radius = 2.5
    skin = 0.3
    ntypes = 4
    psim = pairs.simulation("md", [pairs.point_mass()], timesteps=timesteps, double_prec=True)
    psim.target(pairs.target_gpu() if target == "gpu" else pairs.target_cpu())
    psim.add_position('position')
    psim.add_property('mass', pairs.real(), 1.0)
    psim.add_property('linear_velocity', pairs.vector())
    psim.add_property('force', pairs.vector(), volatile=True)
    psim.add_feature('type', ntypes)
    psim.add_feature_property('type', 'epsilon', pairs.real(), [1.0 for i in range(ntypes * ntypes)])
    psim.add_feature_property('type', 'sigma6', pairs.real(), [1.0 for i in range(ntypes * ntypes)])
    psim.copper_fcc_lattice(nx, nx, nx, 0.8442, 1.44, ntypes)
    psim.set_domain_partitioner(pairs.regular_domain_partitioner())
    psim.compute_thermo(thermo)
    psim.reneighbor_every(reneigh)
    psim.build_neighbor_lists(cutoff_radius + skin)
    psim.compute(initial_integrate, symbols={'dt': dt}, pre_step=True, skip_first=True)
    psim.compute(lennard_jones, cutoff_radius)
    psim.compute(final_integrate, symbols={'dt': dt}, skip_first=True)
    return psim


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else "gpu").generate()
