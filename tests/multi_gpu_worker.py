"""torchrun worker for the multi-GPU parity test: one process per GPU, slab plans attached over CUDA IPC, the
step compared against the CPU oracle on each rank's own rows.  Exit code 0 = parity within 1e-10."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cfd_julia_b200 as vm
    from cfd_julia_b200.common import Plan
    from helpers import grid, noise_field, rel_l2, stable_dt, vm_field
    from oracle import oracle_c as oc

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = vm.default_library()

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    worst = 0.0
    for n, nt in ((256, 6), (1024, 3), (4096, 1)):
        dx, dy, _, _ = grid(n)
        dt = stable_dt(n, 1000.)
        w0 = vm_field(n) + 0.05 * noise_field(n, 7)
        p = Plan(lib, n, n, rank, world)
        p.attach_peers(gather)
        dist.barrier()
        wn = w0.copy(order="F")
        p.upload(wn)
        p.step(dx, dy, dt, 1000., nt)
        psi = np.zeros_like(w0)
        p.download(wn, psi)
        ref = w0.copy(order="F")
        _, s = oc.numerical(n, n, nt, dx, dy, dt, 1000., ref)
        nj = n // world
        rows = slice(rank * nj, (rank + 1) * nj + 2)  # this rank's ghosted rows j0 .. j0+NJ+1 (incl. neighbour halos)
        e1, e2 = rel_l2(wn[:, rows], ref[:, rows]), rel_l2(psi[:, rows], s[:, rows])
        worst = max(worst, e1, e2)
        print(f"rank {rank}/{world} n={n} steps={nt}: rel-L2 w {e1:.2e} psi {e2:.2e}", flush=True)
        dist.barrier()
        p.close()
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.destroy_process_group()
    return 0 if float(t.item()) < 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main())
