"""A user script written against the OLDER pairs API that the reference's examples/lj_onetype.py still uses (no shapes argument,
add_real_property / add_vector_property(vol=), from_file, bare rsq / delta symbols, scalar sigma6 / epsilon, explicit Euler,
target() called last).  python lj_legacy_script.py gpu [n] [timesteps]"""
import sys

import pairs


def lj(i, j):
    sr2 = 1.0 / rsq
    sr6 = sr2 * sr2 * sr2 * sigma6
    force[i] += delta * 48.0 * sr6 * (sr6 - 0.5) * sr2 * epsilon


def euler(i):
    velocity[i] += dt * force[i] / mass[i]
    position[i] += dt * velocity[i]


def build(target="gpu", n=6, timesteps=20):
    dt = 0.005
    cutoff_radius = 2.5
    skin = 0.3
    sigma = 1.0
    epsilon = 1.0
    sigma6 = sigma ** 6
    psim = pairs.simulation("lj", debug=True, timesteps=timesteps)
    psim.add_real_property('mass', 1.0)
    psim.add_position('position')
    psim.add_vector_property('velocity')
    psim.add_vector_property('force', vol=True)
    psim.from_file(f"data/minimd_setup_{n}x{n}x{n}.input", ['mass', 'position', 'velocity'])
    psim.build_neighbor_lists(cutoff_radius + skin)
    psim.vtk_output(f"output/test_{target}")
    psim.compute(lj, cutoff_radius, {'sigma6': sigma6, 'epsilon': epsilon})
    psim.compute(euler, symbols={'dt': dt})
    psim.target(pairs.target_gpu() if target == 'gpu' else pairs.target_cpu())
    return psim


if __name__ == "__main__":
    a = sys.argv[1:]
    build(a[0] if a else "gpu", *[int(x) for x in a[1:3]]).generate()
