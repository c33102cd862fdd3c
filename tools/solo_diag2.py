import sys, numpy as np
sys.path.insert(0, "/root/repo")
from gym_duckietown_b200 import gymshim
tag = sys.argv[1]
env = gymshim.make("Duckietown-udem1-v0", camera_width=160, camera_height=120)
env.reset()
for _ in range(10):
    obs, _, _, _ = env.step(np.array([0.1, 0.1]))
first_obs = env.reset()
first_render = env.render("rgb_array")
second_obs, rew, done, info = env.step([0.0, 0.0])
top = env.render("top_down")
seg = env.render_obs(segment=True)
seg2 = env.render_obs(segment=True)
b = env.unwrapped._b if hasattr(env, "unwrapped") else env._b
print(tag, "seg[0,0]", seg[0, 0], "seg2[0,0]", seg2[0, 0], "obs[0,0]", second_obs[0, 0], "status", b.sim.status(), "counters", b.sim.debug_counters()[:4])
np.savez(f"/root/repo/gpurun_out/solo2_{tag}.npz", seg=seg, seg2=seg2, obs=second_obs, top=top, first=first_obs)
