"""Generates tests/golden/md_t1.npz, md_t2.npz, md_cells_t1.npz, md_half_t1.npz, md_custom_t1.npz, md_props_t1.npz and md_vocab_t1.npz from the REFERENCE ITSELF: the reference generator's own serial C++
for examples/md.py (variants md_t1 / md_t2 of oracle/build_ref.py, i.e. nx = 8 resp. 12, thermo every step), run in a
fresh process.  Needs /root/reference (this container only); the fixtures are committed so that the oracle restatement
and the CUDA path can be pinned where the reference is absent.

    python tests/golden/make_golden_md.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_worker  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = {"md_t1": [0, 1, 19, 20, 100], "md_t2": [5, 60], "md_half_t1": [0, 1, 20, 100], "md_custom_t1": [0, 1, 20, 100], "md_props_t1": [0, 1, 20, 100], "md_vocab_t1": [0, 1, 10], "md_cells_t1": [0, 1, 19, 20, 60]}
NAMES = {"md_t1": ("position", "linear_velocity", "force"), "md_t2": ("position", "force"),
         "md_half_t1": ("position", "linear_velocity", "force"), "md_custom_t1": ("position", "linear_velocity", "force"),
         "md_props_t1": ("position", "linear_velocity", "force", "scale", "heat", "work", "path", "pull", "ups"),
         "md_vocab_t1": ("position", "linear_velocity", "force"), "md_cells_t1": ("position", "linear_velocity", "force")}
# user-defined properties of a variant (tests/scripts/props_script.py), recorded next to the MD set
EXTRA = {"md_props_t1": (("scale", 1), ("heat", 1), ("work", 1), ("path", 3), ("pull", 3))}
EXTRA_INT = {"md_props_t1": ("ups",)}
ONLY = sys.argv[1:]


def serial_thermo(s):
    m, v = s["mass"], s["linear_velocity"]
    t = 0.0
    for i in range(s["nlocal"]):          # left-to-right, as runtime/thermo.hpp:32-36
        t += m[i] * (v[i, 0] * v[i, 0] + v[i, 1] * v[i, 1] + v[i, 2] * v[i, 2])
    return t * (1.0 / (s["nlocal"] * 3 - 3))


for variant, keep in KEEP.items():
    if ONLY and variant not in ONLY:
        continue
    snaps = ref_worker.dump(variant, f"/tmp/{variant}_golden_raw.npz", extra=EXTRA.get(variant, ()), extra_int=EXTRA_INT.get(variant, ()))
    out = {"temperature": np.array([serial_thermo(s) for s in snaps]),
           "nlocal": np.array([s["nlocal"] for s in snaps]), "nghost": np.array([s["nghost"] for s in snaps]),
           "steps_kept": np.array(keep)}
    for k in keep:
        for name in NAMES[variant]:
            out[f"{name}_{k}"] = snaps[k][name]
    out["type"] = snaps[0]["type"]
    if variant == "md_half_t1":      # compute_half(): the half lists of the first iteration, rows trimmed to the longest list
        nn = snaps[0]["numneighs"]
        out["numneighs_0"] = nn
        out["neighborlists_0"] = snaps[0]["neighborlists"][:, :int(nn.max())].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f"{variant}.npz"), **out)
    print(variant, len(snaps), "steps;", os.path.getsize(os.path.join(HERE, f"{variant}.npz")) // 1024, "KiB")
