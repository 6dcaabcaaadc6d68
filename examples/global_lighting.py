"""Opt-in global lighting: every rank fits the 16 SH
lighting coefficients of its own view; with --global the 16x16+16 normal
equations are summed over the ranks by ONE ncclAllReduce of 272 doubles before
the pseudo inverse. The reference fits per view and never shares, so the
global mode deliberately differs from it.

torchrun --nproc-per-node 2 examples/global_lighting.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smvs_b200 import api, nccl_util, workload  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    wl = workload.build_workload(640, 480, 3, scale=2, shading=True, seed_index=rank)
    ctx = api.Context(local)
    wl.push_views_u8(ctx)
    wl.push_surface(ctx)
    own = ctx.fit_lighting()
    comm = nccl_util.create_comm()
    shared = ctx.fit_lighting(nccl_comm=comm)
    gathered = [torch.zeros(16, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(gathered, torch.as_tensor(shared, device="cuda"))
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        print("per-view light[0:4] ", np.round(own[:4], 5))
        print("global   light[0:4] ", np.round(shared[:4], 5))
        print("identical on all ranks:", same, "| differs from per-view:",
              not np.allclose(own, shared))
    nccl_util.destroy_comm(comm)
    ctx.close()
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
