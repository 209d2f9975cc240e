"""Multi-GPU check of contact properties beyond examples/dem.py's three, one process per GPU (torchrun --nproc-per-node N): the
script of tests/scripts/dem_script.py with more_contact_props=True runs through the DSL on N ranks and writes a checkpoint every 10
iterations (the contact rows with their further lanes, per rank).  A migrating particle takes its whole contact row with it
(sim/comm.py PackContactHistoryData), and the further lanes are part of that row.  What is checked, from the files:
  * in every checkpoint, on every rank, every live contact's further properties are consistent: tsd_seen is the tangential
    displacement bit for bit, hits = 3 + 2 (age + 1), age a non-negative whole number;
  * age counts the iterations a contact has existed -- a lane lost in transit would restart from the default -- so for a contact
    found on rank r whose particle was owned by ANOTHER rank ten iterations earlier, and that existed then, the age must be exactly
    the age in the other rank's file plus 10; the script demands that such contacts exist."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))

DOMAIN = (0.1, 0.03, 0.04)
STEPS, EVERY = 800, 10


def read_checkpoint(prefix, ts, rank, world):
    man = json.load(open(f"{prefix}_{ts}.json"))
    col, k = {}, 0
    for name, width, _ in man["columns"]:
        col[name] = k
        k += width
    infix = f"_r{rank}" if world > 1 else ""
    rows = np.loadtxt(f"{prefix}_{ts}{infix}.csv", delimiter=",", ndmin=2)
    uids = set(rows[:, col["uid"]].astype(int).tolist()) if rows.size else set()
    fn = f"{prefix}_{ts}{infix}.contacts.csv"
    cont = np.loadtxt(fn, delimiter=",", ndmin=2) if os.path.getsize(fn) else np.zeros((0, 12))
    assert cont.shape[1] == 12                            # uid_i, uid_j, sticking, tsd x 3, ivm + the five further lanes
    assert np.array_equal(cont[:, 7:10], cont[:, 3:6])                                   # tsd_seen == tsd ("%.17g" keeps every bit)
    assert np.array_equal(cont[:, 11], 3.0 + 2.0 * (cont[:, 10] + 1.0)) and (cont[:, 10] >= 0.0).all()
    assert np.array_equal(cont[:, 10], np.round(cont[:, 10]))
    return uids, {(int(r[0]), int(r[1])): float(r[10]) for r in cont}


def main():
    import datetime
    import torch.distributed as dist
    import dem_script
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=180))
    box = [tempfile.mkdtemp(prefix="ckpt_contact_props_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    prefix = os.path.join(box[0], "ck")
    ctx = dem_script.build("gpu", DOMAIN, STEPS, more_contact_props=True, checkpoint=(prefix, EVERY)).generate()
    assert ctx.contact_extra_lanes == 5
    dist.barrier()                                        # every rank's files are complete
    ok = 1
    if rank == 0:
        try:
            crossed, rows, oldest = 0, 0, 0.0
            prev = None
            for ts in range(0, STEPS, EVERY):             # a checkpoint follows every iteration whose number is a multiple of EVERY
                cur = [read_checkpoint(prefix, ts, r, world) for r in range(world)]
                rows += sum(len(c[1]) for c in cur)
                oldest = max([oldest] + [g for c in cur for g in c[1].values()])
                if prev is not None:
                    for r in range(world):
                        for (a, b), age in cur[r][1].items():
                            if age < EVERY or a in prev[r][0]:
                                continue                  # younger than the interval, or its particle did not change owner
                            before = [p[1][(a, b)] for q, p in enumerate(prev) if q != r and a in p[0] and (a, b) in p[1]]
                            assert before == [age - EVERY], (ts, a, b, age, before)
                            crossed += 1
                prev = cur
            print(f"{rows} contact rows in {STEPS // EVERY} checkpoints per rank, oldest contact {oldest:.0f} iterations, {crossed} contacts "
                  "followed across a rank boundary")
            assert crossed > 0, "no particle carried a live contact across a rank boundary: the test did not exercise the migration"
            print(f"mgpu_contact_props_check ok: world {world}, {crossed} contacts crossed a rank boundary with their age intact")
        except Exception:
            import traceback
            traceback.print_exc()
            ok = 0
    dist.barrier()
    dist.destroy_process_group()
    del ctx
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
