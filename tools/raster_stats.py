"""Print DTS_STATS counters of the raster kernel for the bench workload (needs a -DDTS_STATS build)."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
m = sys.argv[1] if len(sys.argv) > 1 else "small_loop"
env = BatchedDuckietownEnv(4096, m, camera_width=160, camera_height=120, domain_rand=False, seed=1000, auto_reset=True, device_reset=True)
env.reset(render=False)
a = torch.rand((30, 4096, 2), device=env.device) * 2 - 1
for t in range(30):
    env.step(a[t], render=False)
torch.cuda.synchronize()
c0 = env.sim.debug_counters().astype(np.int64)
env.render_obs(); torch.cuda.synchronize()
c = env.sim.debug_counters().astype(np.int64) - c0
names = {8: "coarse bins", 9: "empty coarse", 10: "sum list len", 11: "live prims (per fine bin chunks)", 12: "ground live", 13: "simple fine bins", 14: "coarse bins with > 32 prims", 15: "…their list lengths", 16: "general prim iterations", 17: "…of which fully inside"}
for k, v in names.items():
    print(f"{v:40s} {c[k]:12d}  per env {c[k]/4096:10.1f}")
