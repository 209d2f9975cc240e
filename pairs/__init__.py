"""`import pairs` -> the B200 backend's implementation of the pairs DSL surface (pairs_b200.dsl), so that scripts
written against rafaelravedutti/pairs (examples/md.py) run unchanged with this repository on PYTHONPATH."""
from pairs_b200.dsl import (DomainPartitioners, DslError, Shapes, Simulation, Target, Types, double, float, halfspace,  # noqa: F401,A004
                            int32, matrix, point_mass, quaternion, real, regular_domain_partitioner,
                            regular_domain_partitioner_xy, simulation, sphere, target_cpu, target_gpu, vector)
