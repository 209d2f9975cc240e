"""Multi-GPU check of the user-defined properties, one process per GPU (launch with torchrun --nproc-per-node N): the script of
tests/scripts/props_script.py runs through the DSL on N ranks (NCCL migration carries the non-volatile user rows inside the
exchange record, ghosts get them inside the border record).  A different rank grid gives a slightly different trajectory (the
reference's one-step lag of forwarded ghosts), so the comparison is made through invariants that do not
depend on the decomposition:
  * every lattice site's `scale` value exists on exactly one rank at the end (the multiset over all ranks is bit-identical to the
    one the setup() function produced): nothing lost, duplicated or altered in transit;
  * position - path, wrapped into the box, is the lattice site whose scale the particle carries (tests/props_common.py): `scale`
    (written once) and `path` (integrated every step) stayed attached to the SAME particle across ownership changes;
  * particles did change owner during the run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "scripts"))


def main():
    import torch.distributed as dist
    import props_script
    from tests import props_common as pc
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, steps = 12, 120
    box = nx * pow(4.0 / 0.8442, 1.0 / 3.0)
    # state right after set-up (0 loop iterations beyond ts = 0): who owns what at the start
    ctx0 = props_script.build("gpu", nx, 0, 20, 0).generate()
    start = {"scale": ctx0.download_property("scale")}
    ctx = props_script.build("gpu", nx, steps, 20, 0).generate()
    end = {"scale": ctx.download_property("scale"), "path": ctx.download_property("path"), "pos": ctx.real("position"),
           "heat": ctx.download_property("heat")}
    pc.check_identity(end["pos"], end["path"], end["scale"], box, props_script.XLEN)
    gathered = [None] * world
    dist.all_gather_object(gathered, {"start": start["scale"], "end": end["scale"], "moved": float(np.abs(end["path"]).max())})
    if rank == 0:
        s0 = np.concatenate([g["start"] for g in gathered])
        s1 = np.concatenate([g["end"] for g in gathered])
        assert len(s0) == 4 * nx ** 3 and len(np.unique(s0)) == len(s0)
        assert np.array_equal(np.sort(s0), np.sort(s1))
        changed = sum(len(np.setdiff1d(g["end"], g["start"])) for g in gathered)
        assert changed > 0, "no particle changed owner: the test did not exercise the migration"
        assert all(g["moved"] > 0.1 for g in gathered)
        print(f"mgpu_props_check ok: world {world}, {len(s0)} particles, {changed} ownership changes, {steps} steps")
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
