"""World-size-2 (or more) CPU run of the slab-decomposed step: one PROCESS per rank over torch.distributed/gloo,
the kernel bodies executed by the host emulator, the ranks' buffers shared through POSIX shared memory (the emulator's
stand-in for CUDA IPC) and dist.barrier() as the cross-rank barrier hook.  Same host code path as the multi-GPU run
(Plan.attach_peers, vmk_peer_export/import, barrier placement); compared against the oracle on each rank's rows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    from cfd_julia_b200._lib import BARRIER_FN, VmkLibrary
    from cfd_julia_b200.common import Plan
    from helpers import grid, noise_field, rel_l2, vm_field
    from oracle import oracle_c as oc
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    lib = VmkLibrary(os.path.join(ROOT, "tests", "emul", "libvmk_emul.so"), "vmke_")

    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    hook = BARRIER_FN(lambda _u: dist.barrier())
    worst = 0.0
    for n, nt in ((64, 4), (128, 2)):
        dx, dy, _, _ = grid(n)
        w0 = vm_field(n) + 0.1 * noise_field(n, 5)
        p = Plan(lib, n, n, rank, world)
        p.attach_peers(gather)
        lib.check(lib.barrier_hook(p.handle, hook, None))
        dist.barrier()
        wn = w0.copy(order="F")
        psi = np.zeros_like(w0)
        p.upload(wn)
        p.step(dx, dy, 1e-3, 1000., nt)
        p.download(wn, psi)
        ref = w0.copy(order="F")
        _, s = oc.numerical(n, n, nt, dx, dy, 1e-3, 1000., ref)
        nj = n // world
        rows = slice(rank * nj, (rank + 1) * nj + 2)
        worst = max(worst, rel_l2(wn[:, rows], ref[:, rows]), rel_l2(psi[:, rows], s[:, rows]))
        # the reference-signature entry point on a slab plan: every rank solves for its own columns
        f = np.asfortranarray(np.random.default_rng(2).uniform(-1, 1, (n, n)))
        sg = np.zeros_like(w0)
        sr = np.zeros_like(w0)
        lib.check(lib.fps(p.handle, dx, dy, f.ctypes.data, sg.ctypes.data, 1e-6))
        oc.fps(n, n, dx, dy, f, sr)
        own = slice(rank * nj + 1, (rank + 1) * nj + 1)
        worst = max(worst, rel_l2(sg[1:n + 1, own], sr[1:n + 1, own]))
        dist.barrier()
        p.close()
    print(f"rank {rank}/{world}: worst rel-L2 {worst:.2e}", flush=True)
    dist.destroy_process_group()
    return 0 if worst < 1e-12 else 1


if __name__ == "__main__":
    sys.exit(main())
