"""Multi-GPU parity check, one process per GPU (launch with torchrun --nproc-per-node N): every rank runs the MD loop on its
sub-box through the C-ABI with NCCL halo exchange / migration; the comparison (tests/mgpu_parity.py) is against the SINGLE-RANK
oracle of the same global system -- thermo, every particle's position, forces and the global neighbour-pair set -- and, for the
steps between reneighbourings (where the reference's forwarded ghosts are decomposition-dependent), against the restatement
holding the same N-rank decomposition.  bench.py runs the same check before its timed region when N > 1."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from pairs_b200 import backend
    from tests import mgpu_parity
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 45
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        report = mgpu_parity.check(backend, dist, rank, world, local, nx=nx, steps_c=steps)
        if rank == 0:
            print("mgpu_check ok:", json.dumps(report))
    finally:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
