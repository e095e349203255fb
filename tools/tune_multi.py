"""torchrun worker: per-kernel timings of the slab-decomposed step for a list of option settings.
usage: python -m torch.distributed.run --nproc-per-node P tools/tune_multi.py N opt=val[,opt=val] ..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import cfd_julia_b200 as vm
    from cfd_julia_b200.common import Plan
    from bench import DT, RE, vm_initial_condition
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1])
    dx, w0 = vm_initial_condition(n, rank * (n // world), n // world)
    if n > 8192:
        DT = DT * (8192. / n)**2
    p = Plan(vm.default_library(), n, n, rank, world)

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    p.attach_peers(gather)
    dist.barrier()
    p.upload(w0)

    def measure(tag):
        p.step(dx, dx, DT, RE, 3)
        p.sync()
        dist.barrier()
        prof = p.profile_steps(dx, dx, DT, RE, 3)
        dist.barrier()
        p.step(dx, dx, DT, RE, 10)
        p.sync()
        ms = torch.tensor([p.step_elapsed_ms() / 10], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"P={world} {tag:28s} " + " ".join(f"{k}={v['ms'] / max(v['launches'], 1):.4f}" for k, v in prof.items())
                  + f"  step={ms.item():.3f} ms", flush=True)

    measure("defaults")
    for arg in sys.argv[2:]:
        for kv in arg.split(","):
            k, v = kv.split("=")
            p.set_option(k, int(v))
        dist.barrier()
        measure(arg)
    dist.barrier()
    p.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
