#!/usr/bin/env python3
"""Multi-GPU check of the fused observation gather (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_fused_gather.py
Every rank steps its shard, arms the fused gather for the last step, and compares its gather buffer with an NCCL
all-gather (dts_allgather_obs) of the same observations.  Prints one line per rank; exit code 1 on mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_duckietown_b200.batched_env import BatchedDuckietownEnv  # noqa: E402
from gym_duckietown_b200.dist import FusedObsGather, ObsAllGather  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = 512
env = BatchedDuckietownEnv(N, "loop_obstacles", device=lr, camera_width=160, camera_height=120, domain_rand=False, seed=9,
                           auto_reset=True, device_reset=True, env_id_offset=rank * N)
env.reset()
fg = FusedObsGather(env, rank, world)
ag = ObsAllGather(env, rank, world)
ref = torch.empty((world,) + tuple(env.obs.shape), dtype=env.obs.dtype, device=env.device)
acts = torch.rand((8, N, 2), device=env.device) * 2 - 1
ok = True
for rollout in range(2):
    for t in range(8):
        if t == 7:
            fg.arm()
        env.step(acts[t])
    got = fg.finish()
    ag.all_gather(ref)
    torch.cuda.synchronize()
    same = bool(torch.equal(got, ref))
    distinct = bool(world == 1 or not torch.equal(ref[0], ref[-1]))
    ok &= same and distinct
    print(f"rank {rank}/{world} rollout {rollout}: fused == nccl: {same}; shards differ: {distinct}; mean {float(ref.float().mean()):.3f}", flush=True)
    dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
