"""Generates tests/golden/dem_t1.npz from the REFERENCE ITSELF: the reference generator's serial C++ for examples/dem.py on a
0.1 x 0.015 x 0.04 box (variant dem_t1 of oracle/build_ref.py: 420 spheres + 2 half-spaces, 700 steps), run in a fresh process.
Kept: state at module boundaries of three iterations (before / after linear_spring_dashpot, after euler; locals + ghosts) and the
end-of-iteration state of a few iterations.  Needs /root/reference (this container only).

    python tests/golden/make_golden_dem.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_worker  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MODULE_STEPS = [150, 300, 400]
END_STEPS = [0, 100, 200, 300, 399]
INPUT = ["position", "linear_velocity", "angular_velocity", "force", "torque", "mass", "radius", "normal", "inv_inertia", "rotation_matrix",
         "rotation_quat", "uid", "type", "flags", "shape", "num_contacts", "contact_lists", "is_sticking", "tangential_spring_displacement",
         "impact_velocity_magnitude", "particle_cell"]
POST = ["force", "torque", "num_contacts", "contact_lists", "contact_used", "is_sticking", "tangential_spring_displacement",
        "impact_velocity_magnitude"]
EUL = ["position", "linear_velocity", "angular_velocity", "rotation_quat", "rotation_matrix"]
END = ["position", "linear_velocity", "angular_velocity", "rotation_quat", "uid", "num_contacts", "contact_lists", "is_sticking",
       "tangential_spring_displacement", "impact_velocity_magnitude", "mass", "radius", "inv_inertia", "flags", "shape", "type", "normal"]

z = ref_worker.dump_dem("dem_t1", "/tmp/dem_t1_golden_raw.npz", 401, sorted(set(MODULE_STEPS + END_STEPS)))
out = {"nlocal": z["nlocal"], "nghost": z["nghost"], "module_steps": np.array(MODULE_STEPS), "end_steps": np.array(END_STEPS)}
for ts in MODULE_STEPS:
    for tag, names in (("pre", INPUT), ("post", POST), ("eul", EUL)):
        for k in names:
            out[f"{tag}_{ts}_{k}"] = z[f"{tag}_{ts}_{k}"]
for ts in END_STEPS:
    for k in END:
        out[f"end_{ts}_{k}"] = z[f"end_{ts}_{k}"]
path = os.path.join(HERE, "dem_t1.npz")
np.savez_compressed(path, **out)
print("dem_t1:", os.path.getsize(path) // 1024, "KiB,", int(z["nlocal"][0]), "locals,", int(z["nghost"][0]), "ghosts")


# variant dem_rn3_t1 (reneighbor_every(3)): end-of-iteration states only
z3 = ref_worker.dump_dem("dem_rn3_t1", "/tmp/dem_rn3_t1_golden_raw.npz", 301, [0, 150, 300])
out3 = {"nlocal": z3["nlocal"], "nghost": z3["nghost"], "end_steps": np.array([0, 150, 300])}
for ts in (0, 150, 300):
    for k in END:
        out3[f"end_{ts}_{k}"] = z3[f"end_{ts}_{k}"]
path3 = os.path.join(HERE, "dem_rn3_t1.npz")
np.savez_compressed(path3, **out3)
print("dem_rn3_t1:", os.path.getsize(path3) // 1024, "KiB")


# variant dem_fix_t1: the same run from the CORRECTED copy of the reference generator (oracle/build_ref.py corrected_generator:
# contact history packed before the leaver's slot is overwritten, into its own buffer, uid per record, matching offsets) --
# end-of-iteration states past the first periodic wrap of a particle with live contacts (between iterations 350 and 400)
FIX_STEPS = [300, 350, 400, 500, 600, 699]
zf = ref_worker.dump_dem("dem_fix_t1", "/tmp/dem_fix_t1_golden_raw.npz", 700, FIX_STEPS)
outf = {"nlocal": zf["nlocal"], "nghost": zf["nghost"], "end_steps": np.array(FIX_STEPS)}
for ts in FIX_STEPS:
    for k in ("position", "linear_velocity", "angular_velocity", "uid", "num_contacts", "contact_lists", "is_sticking",
              "tangential_spring_displacement", "impact_velocity_magnitude"):
        outf[f"end_{ts}_{k}"] = zf[f"end_{ts}_{k}"]
pathf = os.path.join(HERE, "dem_fix_t1.npz")
np.savez_compressed(pathf, **outf)
print("dem_fix_t1:", os.path.getsize(pathf) // 1024, "KiB")


# variant dem_more_t1: examples/dem.py with three FURTHER contact properties (second vector / real / integer) kept by its contact
# model -- what the reference's generated code holds in them at the end of iteration 300
MORE = ["uid", "position", "num_contacts", "contact_lists", "is_sticking", "cp:tangential_spring_displacement:3:real",
        "cp:tsd_seen:3:real", "cp:contact_age:1:real", "cp:hits:1:int"]
zm = ref_worker.dump_dem_end("dem_more_t1", "/tmp/dem_more_t1_golden_raw.npz", 300, MORE)
pathm = os.path.join(HERE, "dem_more_t1.npz")
np.savez_compressed(pathm, **{k: zm[k] for k in zm.files})
live = np.arange(20)[None, :] < zm["num_contacts"][:, None]
print("dem_more_t1:", os.path.getsize(pathm) // 1024, "KiB,", int(live.sum()), "live contacts, oldest", float(zm["contact_age"][live].max()))
