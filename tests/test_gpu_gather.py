"""The fused end-of-rollout gather (dts_gather_*): the last step's rasteriser stores the frames into the gather buffer
besides the caller's tensor.  One-GPU form here (world = 1: the peer table holds this rank's own buffer); the
two-GPU form — cudaIpc peer mappings, NVLink stores, equality with the NCCL all-gather — is tools/check_fused_gather.py,
run under torchrun on a multi-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [("hwc", "uint8"), ("chw", "float32")])
def test_fused_gather_world1_equals_obs(fmt):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    from gym_duckietown_b200.dist import FusedObsGather

    env = BatchedDuckietownEnv(40, "loop_obstacles", camera_width=160, camera_height=120, domain_rand=False, seed=5,
                               auto_reset=True, device_reset=True)
    env.set_output_format(obs_layout=fmt[0], obs_dtype=fmt[1])
    env.reset()
    g = FusedObsGather(env, 0, 1)
    acts = torch.rand((6, 40, 2), device=env.device) * 2 - 1
    for t in range(5):
        env.step(acts[t])
    assert float(g.gathered.float().abs().sum()) == 0.0          # nothing written before arm()
    g.arm()
    obs, *_ = env.step(acts[5])
    out = g.finish()
    assert out.shape == (1,) + tuple(obs.shape) and torch.equal(out[0], obs) and float(obs.float().std()) > (1 if fmt[1] == 'uint8' else 0.02)
    before = out.clone()
    env.step(acts[0])                                               # not armed: the gather buffer keeps the rollout's frames
    torch.cuda.synchronize()
    assert torch.equal(g.gathered, before)
    env.check()
    env.close()
