"""Dynamic obstacles on the GPU: DuckieObj pedestrians / DuckiebotObj followers stepped inside
dts_step, their share of collision / proximity / spawn checks, and their per-env pose in the render.
Checked against the reference's own objects (tests/golden/dynamic_*.npz) and the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DYN_MAPS = ["loop_pedestrians", "loop_dyn_duckiebots"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def load_md(name, golden_dir):
    from gym_duckietown_b200 import maps
    md = maps.load_map(name)
    g = np.load(os.path.join(golden_dir, f"dynamic_{name}.npz"))
    for d, w in zip(md.dyn_objects, g["wiggle"]):   # the reference draws the wiggle from the unseeded global RNG
        d.wiggle = float(w)
    return md, g


def make_env(md, n, **kw):
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    args = dict(camera_width=160, camera_height=120, domain_rand=False, seed=1000)
    args.update(kw)
    return BatchedDuckietownEnv(n, md, **args)


def dyn_host(env, torch, map_id=0):
    from gym_duckietown_b200 import lib as L
    arr, nd = env.sim.dyn_state(map_id)
    t = torch.as_tensor(arr, device=env.device).view(L.DYN_FIELDS, nd, env.num_envs)
    return t.cpu().numpy()


@pytest.mark.parametrize("name", DYN_MAPS)
def test_obstacles_and_probes_vs_reference_golden(name, golden_dir, torch_cuda):
    """Parked agents; every obstacle of every env must follow the reference's DuckieObj / DuckiebotObj step for
    step, and dts_query_poses(dyn_env=...) must reproduce the reference's _collision / proximity_penalty2 probes."""
    torch = torch_cuda
    from gym_duckietown_b200 import lib as L
    md, g = load_md(name, golden_dir)
    env = make_env(md, 3)
    env.reset(render=False)
    zero = torch.zeros(3, 2, device=env.device)
    T = 640 if name == "loop_pedestrians" else 400
    qi = 0
    for t in range(T):
        env.step(zero, render=False)
        st = dyn_host(env, torch)
        for e in (0, 2):
            assert np.abs(st[L.DYN_PX, :, e] - g["pos"][t, :, 0]).max() <= 1e-9, t
            assert np.abs(st[L.DYN_PZ, :, e] - g["pos"][t, :, 2]).max() <= 1e-9, t
            assert np.abs(st[L.DYN_ANGLE, :, e] - g["angle"][t]).max() <= 1e-9, t
            assert np.abs(st[L.DYN_YROT, :, e] - g["y_rot"][t]).max() <= 1e-7, t
            cor = st[L.DYN_CORNERS:L.DYN_CORNERS + 8, :, e].T.reshape(-1, 4, 2)
            assert np.abs(cor - g["corners"][t]).max() <= 1e-9, t
            if name == "loop_pedestrians":
                assert np.array_equal(st[L.DYN_ACTIVE, :, e] != 0, g["active"][t]), t
        while qi < len(g["q_step"]) and g["q_step"][qi] == t:
            x, z, a = g["q_pose"][qi]
            outd, outi = env.sim.query_poses(0, np.array([x]), np.array([z]), np.array([a]), 1.0, dyn_env=1)
            assert bool(outi[0, 1]) == bool(g["q_coll"][qi]), (t, qi)
            assert abs(outd[0, 3] - g["q_prox"][qi]) <= 1e-9, (t, qi)
            qi += 1
    assert qi > 20
    env.close()


@pytest.mark.parametrize("name", DYN_MAPS)
def test_trajectory_with_obstacles_vs_oracle(name, golden_dir, torch_cuda):
    """Driving agents among moving obstacles: pose, reward, done, collision flags against the oracle env by env.
    Some agents are parked in an obstacle's path so that the collision comes from the obstacle's motion."""
    torch = torch_cuda
    import oracle as orc
    md, g = load_md(name, golden_dir)
    N, T = 48, 420
    env = make_env(md, N, max_steps=400, frame_skip=1)
    env.reset(render=False)
    torch.cuda.synchronize()
    st0 = {k: v.cpu().numpy().copy() for k, v in env.state.items()}
    px, pz, ang = st0["pos_x"].copy(), st0["pos_z"].copy(), st0["angle"].copy()
    # park the first agents where an obstacle will be later (the reference trace tells where)
    nd = len(md.dyn_objects)
    for k in range(12):
        s, when = k % nd, 250 + 6 * k
        px[k], pz[k] = g["pos"][when, s, 0], g["pos"][when, s, 2]
        ang[k] = 0.3 * k
    env.sim.reset(None, dict(pos_x=px, pos_z=pz, angle=ang))
    om = orc.OracleMap(md)
    cpu = [orc.OracleEnv(om, px[k], pz[k], ang[k], max_steps=400, dynamics=orc.OracleDynamics(om, wiggle=g["wiggle"]))
           for k in range(N)]
    acts = np.random.default_rng(7).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    acts[:, :12] = 0.0
    acts[:, 12:30, 0] = 0.25
    acts[:, 12:30, 1] *= 0.4
    flips = 0
    prev_col = np.zeros(N, bool)
    for t in range(T):
        _, rew, done, info = env.step(torch.from_numpy(acts[t]).to(env.device), render=False)
        s = {k: v.cpu().numpy() for k, v in info.items()}
        d = done.cpu().numpy()
        for k in range(N):
            o = cpu[k].step(acts[t, k])
            assert bool(s["collided"][k]) == bool(o.collided), (t, k)
            assert bool(d[k]) == bool(o.done) and s["done_code"][k] == o.done_code, (t, k)
            assert (s["tile_i"][k], s["tile_j"][k]) == (o.tile_i, o.tile_j), (t, k)
            assert abs(s["pos_x"][k] - o.pos_x) <= 1e-5 and abs(s["pos_z"][k] - o.pos_z) <= 1e-5, (t, k)
            assert abs(s["prox_penalty"][k] - o.prox) <= 1e-9, (t, k)
            assert abs(s["reward"][k] - o.reward) <= 1e-5 * max(1.0, abs(o.reward)), (t, k)
        col = s["collided"].astype(bool)
        flips += int(np.sum(col[:12] != prev_col[:12]))
        prev_col = col
    assert flips >= 4          # parked agents were run into (and, for pedestrians, left again)
    env.close()


def test_frame_skip_steps_obstacles_per_physics_update(golden_dir, torch_cuda):
    torch = torch_cuda
    from gym_duckietown_b200 import lib as L
    md, g = load_md("loop_pedestrians", golden_dir)
    env = make_env(md, 2, frame_skip=3)
    env.reset(render=False)
    zero = torch.zeros(2, 2, device=env.device)
    for t in range(90):
        env.step(zero, render=False)
    st = dyn_host(env, torch)
    assert np.abs(st[L.DYN_PX, :, 1] - g["pos"][269, :, 0]).max() <= 1e-9
    assert np.abs(st[L.DYN_PZ, :, 0] - g["pos"][269, :, 2]).max() <= 1e-9
    env.close()


@pytest.mark.parametrize("name", DYN_MAPS)
def test_render_moving_obstacles_vs_oracle(name, golden_dir, torch_cuda):
    """Each env draws its own copy of the obstacles where its state has them (O:123-148 with the stepped pos/y_rot)."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import lib as L
    md, g = load_md(name, golden_dir)
    N = 6
    env = make_env(md, N)
    env.reset(render=False)
    nd = len(md.dyn_objects)
    px, pz, ang = np.zeros(N), np.zeros(N), np.zeros(N)
    when = 262
    for k in range(N):      # look at obstacle k % nd from 0.45 m away
        s = k % nd
        ox, oz = g["pos"][when, s, 0], g["pos"][when, s, 2]
        a = 0.9 * k
        px[k], pz[k], ang[k] = ox - 0.45 * np.cos(a), oz + 0.45 * np.sin(a), a
    env.sim.reset(None, dict(pos_x=px, pos_z=pz, angle=ang))
    zero = torch.zeros(N, 2, device=env.device)
    obs = None
    for t in range(when + 1):
        obs, *_ = env.step(zero, render=(t == when))
    torch.cuda.synchronize()
    st = dyn_host(env, torch)
    state = {k: v.cpu().numpy() for k, v in env.state.items()}
    scene = orc.OracleScene(md)
    got = obs.cpu().numpy()
    for k in range(N):
        for s, d in enumerate(md.dyn_objects):
            scene.set_object_pose(d.object_index, (st[L.DYN_PX, s, k], d.pos[1], st[L.DYN_PZ, s, k]), st[L.DYN_YROT, s, k])
        want = scene.render(state["pos_x"][k], state["pos_z"][k], state["angle"][k], W=160, H=120)
        diff = np.abs(got[k].astype(int) - want.astype(int))
        assert diff.max() <= 1, (k, diff.max(), int((diff > 0).sum()))
    # the obstacle really is in view and really moved: the frame differs from the load-time scene
    scene0 = orc.OracleScene(md)
    moved = sum(int(np.any(scene0.render(state["pos_x"][k], state["pos_z"][k], state["angle"][k]) != got[k])) for k in range(N))
    assert moved >= N // 2
    env.close()


def test_obstacle_state_survives_resets_and_is_per_env(golden_dir, torch_cuda):
    """The reference keeps its object list across reset() (S:528-763 never rebuilds it); with auto-reset and
    device-side resets the spawn check sees the obstacles where they are (x.pos, S:1466)."""
    torch = torch_cuda
    from gym_duckietown_b200 import lib as L
    md, g = load_md("loop_pedestrians", golden_dir)
    N = 64
    env = make_env(md, N, auto_reset=True, device_reset=True, max_steps=40)
    env.reset(render=False)
    acts = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (300, N, 2)).astype(np.float32)).to(env.device)
    for t in range(300):
        env.step(acts[t], render=False)
    st = dyn_host(env, torch)
    ep = env.state["episode"].cpu().numpy()
    assert ep.min() >= 7                      # every env went through several resets
    for e in range(N):                        # and the obstacles never noticed
        assert np.abs(st[L.DYN_PX, :, e] - g["pos"][299, :, 0]).max() <= 1e-9
        assert np.array_equal(st[L.DYN_ACTIVE, :, e] != 0, g["active"][299])
    # re-uploading the map puts the obstacles back
    env.sim.upload_map(0, md)
    st = dyn_host(env, torch)
    assert np.abs(st[L.DYN_PX, :, 5] - np.array([d.pos[0] for d in md.dyn_objects])).max() == 0
    env.close()


def test_domain_rand_finish_walk_draws(golden_dir, torch_cuda):
    """finish_walk under domain_rand (O:424-429): |vel| ~ |N(0.02, 0.005)| with the sign flipped, wait in [3, 20)."""
    torch = torch_cuda
    from gym_duckietown_b200 import lib as L
    md, g = load_md("loop_pedestrians", golden_dir)
    N = 256
    env = make_env(md, N, domain_rand=True, device_reset=True)
    env.reset(render=False)
    zero = torch.zeros(N, 2, device=env.device)
    for t in range(275):                      # 240 waiting + 30 walking + margin: every duckie finished one walk
        env.step(zero, render=False)
    st = dyn_host(env, torch)
    vel, wait, act = st[L.DYN_VEL], st[L.DYN_WAIT], st[L.DYN_ACTIVE]
    assert np.all(act == 0) and np.all(vel < 0)
    assert 0.018 < np.abs(vel).mean() < 0.022 and 0.004 < np.abs(vel).std() < 0.006
    assert wait.max() < 19.0 + 1e-9 and wait.min() > 3.0 - 0.5
    assert len(np.unique(np.round(wait * 30))) > 10
    assert len(np.unique(vel)) > N            # per env, per duckie draws
    env.close()


@pytest.mark.parametrize("tag", ["plain", "dr"])
def test_traffic_lights_vs_reference_golden(tag, golden_dir, torch_cuda):
    """TrafficLightObj (O:434-462) on the device: per-light pattern and the card of the shared mesh, step for step
    against the reference's own objects; frames with either card against the raster oracle."""
    torch = torch_cuda
    import copy
    import oracle as orc
    from gym_duckietown_b200 import lib as L, maps
    g = np.load(os.path.join(golden_dir, "trafficlight_loop_trafficlights.npz"))
    md = copy.deepcopy(maps.load_map("loop_trafficlights"))
    tl = [i for i, d in enumerate(md.dyn_objects) if d.kind == maps.DYN_TRAFFICLIGHT]
    for k, i in enumerate(tl):
        md.dyn_objects[i].freq, md.dyn_objects[i].pattern = float(g[f"{tag}_freq"][k]), int(g[f"{tag}_pattern0"][k])
    N = 4
    env = make_env(md, N)
    env.reset(render=False)
    # look at the first light from 0.6 m
    o = md.objects[md.dyn_objects[tl[0]].object_index]
    a = np.array([0.2, 1.7, 3.3, 4.9])
    px, pz = o.pos[0] - 0.6 * np.cos(a), o.pos[2] + 0.6 * np.sin(a)
    env.sim.reset(None, dict(pos_x=px, pos_z=pz, angle=a))
    st = dyn_host(env, torch)
    assert np.all(st[L.DYN_SHOWN, tl[0], :] == g[f"{tag}_shown0"])
    zero = torch.zeros(N, 2, device=env.device)
    scene = orc.OracleScene(md)
    frames = {}
    T = len(g[f"{tag}_shown"])
    want_frames = {100, int(np.flatnonzero(np.diff(g[f"{tag}_shown"]))[0]) + 5}
    for t in range(T):
        obs, *_ = env.step(zero, render=t in want_frames)
        if t in want_frames:
            frames[t] = (obs.cpu().numpy().copy(), dyn_host(env, torch))
        if t % 7 == 0 or t in want_frames or t > T - 5:
            st = dyn_host(env, torch)
            for e in (0, N - 1):
                assert np.array_equal(st[L.DYN_PATTERN, tl, e], g[f"{tag}_pattern"][t]), t
                assert st[L.DYN_SHOWN, tl[0], e] == g[f"{tag}_shown"][t], t
    state = {k: v.cpu().numpy() for k, v in env.state.items()}
    cards = set()
    for t, (got, dst) in frames.items():
        card = int(g[f"{tag}_shown"][t])
        cards.add(card)
        scene.set_trafficlight_card(card)
        for e in range(N):
            for s, d in enumerate(md.dyn_objects):      # the walking duckie of this map: where env e has it
                if d.kind != maps.DYN_TRAFFICLIGHT:
                    scene.set_object_pose(d.object_index, (dst[L.DYN_PX, s, e], d.pos[1], dst[L.DYN_PZ, s, e]), dst[L.DYN_YROT, s, e])
            want = scene.render(state["pos_x"][e], state["pos_z"][e], state["angle"][e])
            diff = np.abs(got[e].astype(int) - want.astype(int))
            assert diff.max() <= 1, (t, e, int((diff > 0).sum()))
        scene.set_trafficlight_card(1 - card)
        other = scene.render(state["pos_x"][0], state["pos_z"][0], state["angle"][0])
        assert np.any(other != got[0])            # the card is in view: the other pattern would look different
    assert cards == {0, 1}
    env.close()


def test_randomize_maps_on_reset_device_equals_host_and_recreates_obstacles(golden_dir, torch_cuda):
    """S:541-544: every reset draws the map from the env's own stream and reloads it (obstacles back to load state)."""
    torch = torch_cuda
    from gym_duckietown_b200 import lib as L
    names = ["small_loop", "loop_pedestrians", "udem1", "loop_dyn_duckiebots"]
    N = 24
    dev = make_env(names, N, device_reset=True, randomize_maps_on_reset=True, seed=300)
    host = make_env(names, N, device_reset=False, randomize_maps_on_reset=True, seed=300)
    zero = torch.zeros(N, 2, device=dev.device)
    seen = set()
    for ep in range(3):
        dev.reset(render=False); host.reset(render=False)
        torch.cuda.synchronize()
        a = {k: v.cpu().numpy() for k, v in dev.state.items()}
        b = {k: v.cpu().numpy() for k, v in host.state.items()}
        assert np.array_equal(a["map_id"], b["map_id"]) and np.array_equal(a["map_id"], host.map_ids)
        for k in ("pos_x", "pos_z", "angle"):
            assert np.array_equal(a[k], b[k]), (ep, k)
        seen |= set(a["map_id"].tolist())
        for t in range(250):                       # past the pedestrians' 240-step wait: they are walking now
            dev.step(zero, render=False)
    assert len(seen) >= 3
    st_before = dyn_host(dev, torch, 1)
    e = int(np.flatnonzero(dev.state["map_id"].cpu().numpy() == 1)[0]) if (dev.state["map_id"] == 1).any() else None
    if e is not None:
        assert st_before[L.DYN_TIME, 0, e] > 8.0
    dev.reset(render=False)                         # reload: envs that land on map 1 have fresh obstacles
    torch.cuda.synchronize()
    mid = dev.state["map_id"].cpu().numpy()
    st = dyn_host(dev, torch, 1)
    md = dev.maps[1]
    for e in np.flatnonzero(mid == 1):
        assert np.all(st[L.DYN_TIME, :, e] == 0.0) and np.all(st[L.DYN_ACTIVE, :, e] == 0.0)
        assert np.array_equal(st[L.DYN_PX, :, e], np.array([d.pos[0] for d in md.dyn_objects]))
    dev.close(); host.close()
