"""Multi-GPU: envs shard by contiguous index blocks (rank r owns global envs [r*N, (r+1)*N), seeds by
global index), the step path has no collective; one NCCL all-gather moves the end-of-rollout
observation batch.  The all-gather runs inside libdtsim.so (dts_allgather_obs) on a
communicator created from a unique id that is broadcast with torch.distributed."""
from __future__ import annotations

import ctypes as C
import glob
import os

import torch
import torch.distributed as dist


def find_libnccl() -> str:
    """The NCCL that torch itself loaded (nvidia-nccl wheel), else the system one."""
    cands = []
    try:
        import nvidia.nccl  # type: ignore
        for p in nvidia.nccl.__path__:
            cands += glob.glob(os.path.join(p, "lib", "libnccl.so*"))
    except Exception:
        pass
    cands += glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so*"))
    cands += glob.glob("/usr/lib/x86_64-linux-gnu/libnccl.so*")
    if not cands:
        raise RuntimeError("libnccl.so not found")
    return cands[0]


class ObsAllGather:
    def __init__(self, env, rank: int, world: int):
        self.env, self.rank, self.world = env, rank, world
        sim = env.sim
        sim._check(sim.lib.dts_comm_load(sim.h, find_libnccl().encode()), "dts_comm_load")
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            sim._check(sim.lib.dts_comm_unique_id(sim.h, buf), "dts_comm_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(env.device)
        dist.broadcast(uid, src=0)
        host = uid.cpu().numpy()
        sim._check(sim.lib.dts_comm_init(sim.h, host.ctypes.data_as(C.c_void_p), rank, world), "dts_comm_init")

    def all_gather(self, out: torch.Tensor, src: torch.Tensor = None) -> torch.Tensor:
        """out: u8[world, N, H, W, 3] on this device; src defaults to the env's obs batch."""
        src = self.env.obs if src is None else src
        assert out.is_contiguous() and src.is_contiguous() and out.numel() == self.world * src.numel()
        sim = self.env.sim
        sim._check(sim.lib.dts_allgather_obs(sim.h, src.data_ptr(), out.data_ptr(), src.numel() * src.element_size(),
                                             torch.cuda.current_stream(self.env.device).cuda_stream),
                   "dts_allgather_obs")
        return out
