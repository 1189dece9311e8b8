"""torch.distributed plumbing for K-sharded controllers (one process per GPU).

Only carries bytes between ranks: the 128-byte ncclUniqueId (MPPI_COMM_NCCL) or the 64-byte
CUDA-IPC handles of the peer mailboxes (MPPI_COMM_P2P).  The exchanges of the control step
itself never go through torch.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .controller import PointMassModel, comm_unique_id


def sharded_controller(nb_sim, steps, dt, state_dim, act_dim, *, comm="p2p", device=None, **kw):
    """Create this rank's shard of a controller over torch.distributed's default group."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.cuda.current_device() if device is None else device
    if comm == "nccl":
        idt = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return PointMassModel(nb_sim, steps, dt, state_dim, act_dim, device=device, rank=rank,
                              world_size=world, comm=capi.COMM_NCCL,
                              comm_id=bytes(idt.cpu().numpy().tobytes()), **kw)
    if comm != "p2p":
        raise ValueError(comm)
    ctl = PointMassModel(nb_sim, steps, dt, state_dim, act_dim, device=device, rank=rank,
                         world_size=world, comm=capi.COMM_P2P, **kw)
    mine = torch.frombuffer(bytearray(ctl.p2p_handle()), dtype=torch.uint8).cuda()
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctl.p2p_connect(b"".join(bytes(np.asarray(t.cpu()).tobytes()) for t in allh))
    dist.barrier()
    return ctl
