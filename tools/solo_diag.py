import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
tag = sys.argv[1]
out = {}
for name, kw in (("udem1_dr", dict(domain_rand=True)), ("loop_obstacles", dict(domain_rand=False)), ("udem1_fish", dict(domain_rand=True, distortion=True, w=640, h=480, n=4))):
    w, h, n = kw.pop("w", 160), kw.pop("h", 120), kw.pop("n", 64)
    env = BatchedDuckietownEnv(n, name.split("_")[0] if name != "loop_obstacles" else name, camera_width=w, camera_height=h, seed=9, **kw)
    env.reset()
    a = torch.zeros((n, 2), device=env.device)
    for _ in range(2): env.step(a)
    out[name] = env.render_obs().cpu().numpy().copy()
    out[name + "_seg"] = env.render_obs(segment=True).cpu().numpy().copy()
    env.close()
np.savez(f"/root/repo/gpurun_out/solo_{tag}.npz", **out)
print(tag, {k: v.shape for k, v in out.items()})
