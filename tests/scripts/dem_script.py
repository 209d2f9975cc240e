"""A user script against the `pairs` DSL for the DEM case (same calls and kernel bodies a P4IRS user writes, cf. the reference's
examples/dem.py), parameterised for the tests:  python dem_script.py gpu [xsize] [ysize] [zsize] [timesteps]"""
import math
import os
import sys

import pairs


def update_mass_and_inertia(i):
    rotation_matrix[i] = diagonal_matrix(1.0)
    rotation_quat[i] = default_quaternion()

    if is_sphere(i):
        inv_inertia[i] = inversed(diagonal_matrix(0.4 * mass[i] * radius[i] * radius[i]))

    else:
        mass[i] = infinity
        inv_inertia[i] = 0.0


def linear_spring_dashpot(i, j):
    delta_ij = -penetration_depth(i, j)
    skip_when(delta_ij < 0.0)

    meff = 1.0 / ((1.0 / mass[i]) + (1.0 / mass[j]))
    stiffness_norm = meff * (pi * pi + lnDryResCoeff * lnDryResCoeff) / \
                     (collisionTime_SI * collisionTime_SI)
    stiffness_tan = kappa * stiffness_norm
    damping_norm = -2.0 * meff * lnDryResCoeff / collisionTime_SI
    damping_tan = sqrt(kappa) * damping_norm

    velocity_wf_i = linear_velocity[i] + cross(angular_velocity[i], contact_point(i, j) - position[i])
    velocity_wf_j = linear_velocity[j] + cross(angular_velocity[j], contact_point(i, j) - position[j])

    rel_vel = -(velocity_wf_i - velocity_wf_j)
    rel_vel_n = dot(rel_vel, contact_normal(i, j)) * contact_normal(i, j)
    rel_vel_t = rel_vel - rel_vel_n
    fN = stiffness_norm * delta_ij * contact_normal(i, j) + damping_norm * rel_vel_n;

    tan_spring_disp = tangential_spring_displacement[i, j]
    impact_vel_magnitude = impact_velocity_magnitude[i, j]
    impact_magnitude = select(impact_vel_magnitude > 0.0, impact_vel_magnitude, length(rel_vel))
    sticking = is_sticking[i, j]

    rot_tan_disp = tan_spring_disp - contact_normal(i, j) * dot(tan_spring_disp, contact_normal(i, j))
    rot_tan_disp_len2 = squared_length(rot_tan_disp)
    new_tan_spring_disp = dt * rel_vel_t + \
                          select(rot_tan_disp_len2 <= 0.0,
                                 zero_vector(),
                                 rot_tan_disp * sqrt(squared_length(tan_spring_disp) / rot_tan_disp_len2))

    fTLS = stiffness_tan * new_tan_spring_disp + damping_tan * rel_vel_t
    fTLS_len = length(fTLS)
    t = normalized(fTLS)

    f_friction_abs_static = friction_static[i, j] * length(fN)
    f_friction_abs_dynamic = friction_dynamic[i, j] * length(fN)
    tan_vel_threshold = 1e-8

    cond1 = sticking == 1 and length(rel_vel_t) < tan_vel_threshold and fTLS_len < f_friction_abs_static
    cond2 = sticking == 1 and fTLS_len < f_friction_abs_dynamic
    f_friction_abs = select(cond1, f_friction_abs_static, f_friction_abs_dynamic)
    n_sticking = select(cond1 or cond2 or fTLS_len < f_friction_abs_dynamic, 1, 0)
    tangential_spring_displacement[i, j] = \
        select(not cond1 and not cond2 and stiffness_tan > 0.0,
               (f_friction_abs * t - damping_tan * rel_vel_t) / stiffness_tan,
               new_tan_spring_disp)

    impact_velocity_magnitude[i, j] = impact_magnitude
    is_sticking[i, j] = n_sticking

    fTabs = min(fTLS_len, f_friction_abs)
    fT = fTabs * t
    partial_force = fN + fT

    apply(force, partial_force)
    apply(torque, cross(contact_point(i, j) - position, partial_force))


def contact_model_with_more_properties():
    """linear_spring_dashpot followed by three statements on FURTHER contact properties (a second vector, a second real, a second
    integer; declared in build(more_contact_props=True)) -- the case SURVEY.md 8f names: the reference gives every
    add_contact_property() its own array, so a model may keep any per-contact state.  The text is composed from the function above
    and written to a file (kernels are translated from their source)."""
    import importlib.util
    import inspect
    import tempfile
    src = inspect.getsource(linear_spring_dashpot).replace("def linear_spring_dashpot(", "def spring_dashpot_more(")
    src += ("\n    tsd_seen[i, j] = tangential_spring_displacement[i, j]\n"
            "    contact_age[i, j] = contact_age[i, j] + 1.0\n"
            "    hits[i, j] = hits[i, j] + 2\n")
    fd, path = tempfile.mkstemp(prefix="dem_model_", suffix=".py")
    with os.fdopen(fd, "w") as f:
        f.write(src)
    spec = importlib.util.spec_from_file_location(os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.spring_dashpot_more


def euler(i):
    inv_mass = 1.0 / mass[i]
    position[i] += 0.5 * inv_mass * force[i] * dt * dt + linear_velocity[i] * dt
    linear_velocity[i] += inv_mass * force[i] * dt
    wdot = rotation_matrix[i] * (inv_inertia[i] * torque[i]) * transposed(rotation_matrix[i])
    phi = angular_velocity[i] * dt + 0.5 * wdot * dt * dt
    rotation_quat[i] = quaternion(phi, length(phi)) * rotation_quat[i]
    rotation_matrix[i] = quaternion_to_rotation_matrix(rotation_quat[i])
    angular_velocity[i] += wdot * dt


def gravity(i):
    volume = (4.0 / 3.0) * pi * radius[i] * radius[i] * radius[i]
    force[i][2] = force[i][2] - (densityParticle_SI - densityFluid_SI) * volume * gravity_SI


def build(target="gpu", domain=(0.1, 0.015, 0.04), timesteps=300, planes_file=None, per_cell=False, vtk=None, reneighbor=None,
          checkpoint=None, restart=None, contact_capacity=20, more_contact_props=False):
    diameter_SI, gravity_SI, densityFluid_SI, densityParticle_SI = 0.0029, 9.81, 1000, 2550
    generationSpacing_SI, initialVelocity_SI, dt_SI = 0.005, 1, 5e-5
    frictionCoefficient, restitutionCoefficient, collisionTime_SI, poissonsRatio = 0.5, 0.1, 5e-4, 0.22
    kappa = 2.0 * (1.0 - poissonsRatio) / (2.0 - poissonsRatio)
    minDiameter_SI, maxDiameter_SI = diameter_SI * 0.9, diameter_SI * 1.1
    linkedCellWidth = 1.01 * maxDiameter_SI
    ntypes = 1
    lnDryResCoeff = math.log(restitutionCoefficient)

    psim = pairs.simulation("dem", [pairs.sphere(), pairs.halfspace()], timesteps=timesteps, double_prec=True,
                            use_contact_history=True, particle_capacity=1000000, neighbor_capacity=contact_capacity)
    psim.target(pairs.target_gpu() if target == "gpu" else pairs.target_cpu())
    psim.add_position('position')
    psim.add_property('mass', pairs.real(), 1.0)
    psim.add_property('linear_velocity', pairs.vector())
    psim.add_property('angular_velocity', pairs.vector())
    psim.add_property('force', pairs.vector(), volatile=True)
    psim.add_property('torque', pairs.vector(), volatile=True)
    psim.add_property('radius', pairs.real(), 1.0)
    psim.add_property('normal', pairs.vector())
    psim.add_property('inv_inertia', pairs.matrix())
    psim.add_property('rotation_matrix', pairs.matrix())
    psim.add_property('rotation_quat', pairs.quaternion())
    psim.add_feature('type', ntypes)
    psim.add_feature_property('type', 'friction_static', pairs.real(), [0.0 for _ in range(ntypes * ntypes)])
    psim.add_feature_property('type', 'friction_dynamic', pairs.real(), [frictionCoefficient for _ in range(ntypes * ntypes)])
    psim.add_contact_property('is_sticking', pairs.int32(), 0)
    psim.add_contact_property('tangential_spring_displacement', pairs.vector(), [0.0, 0.0, 0.0])
    psim.add_contact_property('impact_velocity_magnitude', pairs.real(), 0.0)
    if more_contact_props:
        psim.add_contact_property('tsd_seen', pairs.vector(), [0.0, 0.0, 0.0])
        psim.add_contact_property('contact_age', pairs.real(), -1.0)          # -1: the first evaluation of a contact leaves age 0
        psim.add_contact_property('hits', pairs.int32(), 3)
    psim.set_domain([0.0, 0.0, 0.0, domain[0], domain[1], domain[2]])
    psim.set_domain_partitioner(pairs.regular_domain_partitioner_xy())
    psim.pbc([True, True, False])
    if restart is not None:          # continue from a checkpoint: no generator, no plane file, no update_mass_and_inertia (it would
        psim.read_checkpoint(*restart)          # reset the orientations)
    else:
        psim.dem_sc_grid(domain[0], domain[1], domain[2], generationSpacing_SI, diameter_SI, minDiameter_SI, maxDiameter_SI,
                         initialVelocity_SI, densityParticle_SI, ntypes)
    if restart is None and planes_file is None:
        # the two half-spaces examples/dem.py reads from data/planes.input (uid, type, mass, position, normal, flags): floor at the
        # origin, ceiling at the corner of the stock 0.8 x 0.015 x 0.2 box; infinite | fixed | global
        import tempfile
        rows = [(100000, 0, 1, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 13), (100001, 0, 1, (0.8, 0.015, 0.2), (0.0, 0.0, -1.0), 13)]
        fd, planes_file = tempfile.mkstemp(prefix="planes_", suffix=".input")
        with os.fdopen(fd, "w") as f:
            for uid, typ, m, x, nrm, fl in rows:
                f.write(",".join(str(v) for v in (uid, typ, m, *x, *nrm, fl)) + "\n")
    if restart is None:
        psim.read_particle_data(planes_file,
                                ['uid', 'type', 'mass', 'position', 'normal', 'flags'], pairs.halfspace())
        psim.setup(update_mass_and_inertia, {'densityParticle_SI': densityParticle_SI, 'pi': math.pi, 'infinity': math.inf})
    if checkpoint is not None:
        psim.checkpoint_output(checkpoint[0], checkpoint[1])
    if per_cell:
        psim.build_cell_lists(linkedCellWidth, store_neighbors_per_cell=True)
    else:
        psim.build_cell_lists(linkedCellWidth)
    if vtk is not None:
        psim.vtk_output(vtk[0], frequency=vtk[1])          # examples/dem.py:194
    psim.compute(gravity, symbols={'densityParticle_SI': densityParticle_SI, 'densityFluid_SI': densityFluid_SI,
                                   'gravity_SI': gravity_SI, 'pi': math.pi})
    psim.compute(contact_model_with_more_properties() if more_contact_props else linear_spring_dashpot, linkedCellWidth, symbols={'dt': dt_SI, 'pi': math.pi, 'kappa': kappa,
                                                                   'lnDryResCoeff': lnDryResCoeff, 'collisionTime_SI': collisionTime_SI})
    psim.compute(euler, symbols={'dt': dt_SI})
    if reneighbor is not None:          # oracle variant dem_rn3_t1: cell lists / ghosts every `reneighbor` iterations
        psim.reneighbor_every(reneighbor)
    return psim


if __name__ == "__main__":
    a = sys.argv[1:]
    dom = tuple(float(x) for x in a[1:4]) if len(a) >= 4 else (0.1, 0.015, 0.04)
    psim = build(a[0] if a else "gpu", dom, int(a[4]) if len(a) > 4 else 300)
    psim.generate()
