"""GPU parity of the step-logic kernels (through the C ABI) against (a) golden vectors produced by
the reference's own code and (b) the CPU oracle on seeded trajectories.
Bar (BASELINE.json north_star): tile index / collision / done flags bit-exact; pose <= 1e-5; reward <= 1e-5 rel."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAPS = ["small_loop", "loop_obstacles", "udem1"]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def make_env(name, n, **kw):
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    args = dict(camera_width=160, camera_height=120, domain_rand=False, seed=1000)
    args.update(kw)
    return BatchedDuckietownEnv(n, name, **args)


@pytest.mark.parametrize("name", MAPS)
def test_pose_predicates_vs_reference_golden(name, golden_dir, torch_cuda):
    g = np.load(os.path.join(golden_dir, f"logic_{name}.npz"))
    env = make_env(name, 4)
    x, z, a = g["poses"].T
    outd, outi = env.sim.query_poses(0, x, z, a, 1.0)
    _, outi13 = env.sim.query_poses(0, x, z, a, 1.3)
    assert np.array_equal(outi[:, 5], g["ti"]) and np.array_equal(outi[:, 6], g["tj"])
    assert np.array_equal(outi[:, 7].astype(bool), g["drv"])
    assert np.array_equal(outi[:, 0].astype(bool), g["valid10"])
    assert np.array_equal(outi13[:, 0].astype(bool), g["valid13"])
    assert np.array_equal(outi[:, 1].astype(bool), g["coll1"])
    assert np.array_equal(outi[:, 2].astype(bool), g["coll2"])
    assert np.array_equal(outi[:, 3].astype(bool), g["inlane"])
    il = g["inlane"]
    assert np.abs(outd[il, 0] - g["dist"][il]).max() <= 1e-12
    assert np.abs(outd[il, 1] - g["dot"][il]).max() <= 1e-12
    assert np.abs(outd[il, 2] - g["ang"][il]).max() <= 1e-7
    assert np.all(np.isnan(outd[~il, 0]))
    assert np.abs(outd[:, 3] - g["prox"]).max() <= 1e-12
    env.close()


@pytest.mark.parametrize("name", MAPS)
def test_done_reward_vs_reference_golden(name, golden_dir, torch_cuda):
    """Place env k at golden pose k (dts_reset with host params), read back reward / done_code of
    _compute_done_reward at step_count=0... the max-steps branch is covered by the trajectory test."""
    torch = torch_cuda
    g = np.load(os.path.join(golden_dir, f"logic_{name}.npz"))
    sel = np.flatnonzero(g["step_count"] < g["max_steps"])
    env = make_env(name, len(sel))
    env.sim.reset(None, dict(pos_x=g["poses"][sel, 0], pos_z=g["poses"][sel, 1], angle=g["poses"][sel, 2]))
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert np.array_equal(st["done_code"], g["code"][sel])
    rel = np.abs(st["reward"] - g["reward"][sel]) / np.maximum(1.0, np.abs(g["reward"][sel]))
    assert rel.max() <= 1e-9
    assert np.array_equal(st["tile_i"], g["ti"][sel]) and np.array_equal(st["tile_j"], g["tj"][sel])
    assert np.array_equal(st["collided"].astype(bool), g["coll2"][sel])
    env.close()


@pytest.mark.parametrize("name,mode", [("small_loop", "vel_steer"), ("loop_obstacles", "vel_steer"),
                                       ("udem1", "pwm")])
def test_trajectory_vs_oracle(name, mode, torch_cuda):
    """N=64 envs, T=300 steps, auto-reset off (SURVEY 8d parity run)."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps

    N, T = 64, 300
    env = make_env(name, N, action_mode=mode, max_steps=250)
    env.reset(render=False)
    torch.cuda.synchronize()
    st0 = {k: v.cpu().numpy().copy() for k, v in env.state.items()}
    om = orc.OracleMap(maps.load_map(name))
    cpu = [orc.OracleEnv(om, st0["pos_x"][k], st0["pos_z"][k], st0["angle"][k], wheel_dist=st0["wheel_dist"][k],
                         action_mode=1 if mode == "vel_steer" else 0, max_steps=250) for k in range(N)]
    acts = np.random.default_rng(1234).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    # keep some envs alive long enough to hit max_steps: gentle forward actions
    acts[:, : N // 4, 0] = 0.12
    acts[:, : N // 4, 1] *= 0.3
    if mode == "pwm":
        acts[:, : N // 4, 1] = acts[:, : N // 4, 0] * (1 + 0.05 * acts[:, : N // 4, 1])
    acts[:, : N // 8, :] = 0.0  # parked robots reach max_steps for sure
    seen_codes = set()
    for t in range(T):
        _, rew, done, info = env.step(torch.from_numpy(acts[t]).to(env.device), render=False)
        torch.cuda.synchronize()
        s = {k: v.cpu().numpy() for k, v in info.items()}
        r32, d = rew.cpu().numpy(), done.cpu().numpy()
        for k in range(N):
            o = cpu[k].step(acts[t, k])
            assert (s["tile_i"][k], s["tile_j"][k]) == (o.tile_i, o.tile_j), (t, k)
            assert bool(d[k]) == bool(o.done) and s["done_code"][k] == o.done_code, (t, k)
            assert bool(s["collided"][k]) == bool(o.collided) and bool(s["in_lane"][k]) == bool(o.in_lane), (t, k)
            assert s["step_count"][k] == o.step_count
            assert abs(s["pos_x"][k] - o.pos_x) <= 1e-5 and abs(s["pos_z"][k] - o.pos_z) <= 1e-5, (t, k)
            dang = abs(s["angle"][k] - o.angle)
            assert min(dang, abs(dang - 2 * np.pi)) <= 1e-5, (t, k)
            assert abs(s["reward"][k] - o.reward) <= 1e-5 * max(1.0, abs(o.reward)), (t, k)
            assert abs(r32[k] - np.float32(o.reward)) <= 1e-3 * max(1.0, abs(o.reward))
            assert abs(s["speed"][k] - o.speed) <= 1e-5
            seen_codes.add(int(o.done_code))
    assert seen_codes == {0, 1, 2}, seen_codes  # every branch of _compute_done_reward was exercised
    env.close()


@pytest.mark.parametrize("name", MAPS)
def test_host_reset_matches_reference_through_gpu_predicates(name, golden_dir, torch_cuda):
    from test_reset_sampler import check_against_golden
    env = make_env(name, 2)

    def make_query(md):
        return lambda k, x, z, a, sf, hid: env.sim.query_poses(0, x, z, a, sf, hid)

    check_against_golden(name, golden_dir, make_query)
    env.close()


def test_device_reset_spawns_valid_poses(torch_cuda):
    """dts_reset_random: every spawned pose satisfies the reference's acceptance test (S:692-731)."""
    torch = torch_cuda
    env = make_env("loop_obstacles", 2048, device_reset=True, auto_reset=True, domain_rand=True)
    env.reset(render=False)
    torch.cuda.synchronize()
    s = {k: v.cpu().numpy() for k, v in env.state.items()}
    outd, outi = env.sim.query_poses(0, s["pos_x"], s["pos_z"], s["angle"], 1.3)
    assert outi[:, 0].all() and outi[:, 3].all() and not outi[:, 4].any()
    assert (np.abs(np.rad2deg(outd[:, 2])) < 60).all()
    assert len(np.unique(np.round(s["pos_x"], 6))) > 2000  # streams are distinct per env
    assert (np.abs(s["wheel_dist"] - 0.102) <= 0.0102 + 1e-12).all() and s["wheel_dist"].std() > 1e-3
    # auto-reset: after enough random steps every env has been through several episodes, all still valid
    acts = torch.rand((200, 2048, 2), device=env.device) * 2 - 1
    for t in range(200):
        env.step(acts[t], render=False)
    torch.cuda.synchronize()
    s = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert (s["episode"] >= 2).mean() > 0.2 and (s["step_count"] < 1500).all()  # many envs re-spawned
    outd, outi = env.sim.query_poses(0, s["pos_x"], s["pos_z"], s["angle"], 1.0)
    assert outi[:, 0].all()  # nobody is left sitting in an invalid pose
    env.close()


@pytest.mark.parametrize("name", MAPS)
@pytest.mark.parametrize("dr", [False, True])
def test_device_reset_reproduces_reference_reset_draw_for_draw(name, dr, golden_dir, torch_cuda):
    """dts_reset_random runs Simulator.reset() ON THE DEVICE from the env's numpy-compatible PCG64 stream:
    pose and every DR value equal what the REFERENCE's reset() produced for the same seeds (golden vectors from
    oracle/make_golden.py), first and second episode."""
    torch = torch_cuda
    g = np.load(os.path.join(golden_dir, f"reset_{name}.npz"))
    tag = "dr" if dr else "nodr"
    n = len(g["seeds"])
    env = make_env(name, n, domain_rand=dr, seed=int(g["seeds"][0]), device_reset=True)   # env k gets seed0 + k
    assert list(g["seeds"]) == list(range(int(g["seeds"][0]), int(g["seeds"][0]) + n))
    for ep in range(2):
        env.reset(render=False)
        torch.cuda.synchronize()
        st = {k: v.cpu().numpy() for k, v in env.state.items()}
        rows = np.arange(n) * 2 + ep
        assert np.array_equal(st["pos_x"], g[f"{tag}_cur_pos"][rows, 0])
        assert np.array_equal(st["pos_z"], g[f"{tag}_cur_pos"][rows, 2])
        assert np.array_equal(st["angle"], g[f"{tag}_cur_angle"][rows])
        assert np.array_equal(st["wheel_dist"], g[f"{tag}_wheel_dist"][rows])
        for k in range(n):
            r = env.sim.debug_episode(k)
            row = rows[k]
            assert r["cam_height"] == np.float32(g[f"{tag}_cam_height"][row])
            assert r["cam_angle_deg"] == np.float32(g[f"{tag}_cam_angle"][row])
            assert r["cam_fov_y_deg"] == np.float32(g[f"{tag}_cam_fov_y"][row])
            assert np.array_equal(r["horizon"], g[f"{tag}_horizon_color"][row].astype(np.float32))
            assert np.array_equal(r["ground"], g[f"{tag}_ground_color"][row].astype(np.float32))
            assert np.array_equal(r["ambient"], g[f"{tag}_ambient"][row, :3].astype(np.float32))
            assert np.array_equal(r["diffuse"], g[f"{tag}_diffuse"][row, :3].astype(np.float32))
            if dr:
                assert np.array_equal(r["cam_noise"], g[f"{tag}_camera_noise"][row].astype(np.float32))
            if ep == 0:  # first reset: GL_POSITION captured under the identity modelview
                assert np.array_equal(r["light_eye"], g[f"{tag}_light_pos"][row].astype(np.float32))
            vis = g[f"{tag}_obj_visible"][row]
            for oi in range(len(vis)):
                assert bool(r["hidden"][oi >> 5] >> (oi & 31) & 1) == (not vis[oi])
    env.close()


def test_device_reset_with_custom_randomizer_table_and_sim_colours(golden_dir, torch_cuda):
    """dts_config.dr_ops / num_tris_distractors / color_sky / color_ground: the DEVICE reset replays the reference's
    reset() under a custom Randomizer table draw for draw (golden produced by the reference's own Randomizer)."""
    import json
    torch = torch_cuda
    g = np.load(os.path.join(golden_dir, "reset_customdr_loop_obstacles.npz"))
    cfg, simkw = json.loads(str(g["config_json"])), json.loads(str(g["sim_json"]))
    n = len(g["seeds"])
    env = make_env("loop_obstacles", n, domain_rand=True, dynamics_rand=True, seed=int(g["seeds"][0]), device_reset=True,
                   randomization_config=cfg, **simkw)
    for ep in range(2):
        env.reset(render=False)
        torch.cuda.synchronize()
        st = {k: v.cpu().numpy() for k, v in env.state.items()}
        rows = np.arange(n) * 2 + ep
        assert np.array_equal(st["pos_x"], g["dr_cur_pos"][rows, 0]) and np.array_equal(st["angle"], g["dr_cur_angle"][rows])
        assert np.array_equal(st["wheel_dist"], g["dr_wheel_dist"][rows])
        for k in range(n):
            r, row = env.sim.debug_episode(k), rows[k]
            assert r["cam_height"] == np.float32(g["dr_cam_height"][row]) and r["cam_angle_deg"] == np.float32(g["dr_cam_angle"][row])
            assert r["cam_fov_y_deg"] == np.float32(g["dr_cam_fov_y"][row])
            assert np.array_equal(r["horizon"], g["dr_horizon_color"][row].astype(np.float32))
            assert np.array_equal(r["ground"], g["dr_ground_color"][row].astype(np.float32))
            assert np.array_equal(r["cam_noise"], g["dr_camera_noise"][row].astype(np.float32))
            assert np.array_equal(r["ambient"], g["dr_ambient"][row, :3].astype(np.float32))
            if ep == 0:
                assert np.array_equal(r["light_eye"], g["dr_light_pos"][row].astype(np.float32))
    env.close()


def test_randomize_maps_host_reset_keeps_the_stale_light_capture(torch_cuda):
    """the host path of randomize_maps_on_reset must not disturb pose / camera before the reset proper, or the
    stale model-view under which GL_LIGHT0 is captured (S:581) is wrong from the second episode on.  Host-drawn and
    device-drawn resets must agree on light_eye."""
    torch = torch_cuda
    names = ["small_loop", "loop_obstacles", "udem1"]
    envs = [make_env(names, 12, domain_rand=True, seed=300, randomize_maps_on_reset=True, device_reset=dev) for dev in (False, True)]
    acts = torch.full((12, 2), 0.4, device=envs[0].device)
    for ep in range(3):
        for e in envs:
            e.reset(render=False)
            for _ in range(4):
                e.step(acts, render=False)
        torch.cuda.synchronize()
        for k in range(12):
            a, b = envs[0].sim.debug_episode(k), envs[1].sim.debug_episode(k)
            assert np.array_equal(a["light_eye"], b["light_eye"]), (ep, k, a["light_eye"], b["light_eye"])
            assert np.array_equal(a["horizon"], b["horizon"])
    for e in envs:
        e.close()


def test_auto_reset_rollout_equals_reference_style_loop(torch_cuda):
    """Device auto-reset (step + respawn in one kernel, numpy-compatible streams) against the reference-style
    host loop `obs, r, done, _ = env.step(a); if done: env.reset()` built from the CPU oracle (step) and the host
    reset sampler (reference draw order, numpy Generator): same seeds, same actions -> same episodes."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.episode import EpisodeSampler
    from test_reset_sampler import oracle_query

    name, N, T = "loop_obstacles", 32, 400
    md = maps.load_map(name)
    env = make_env(name, N, seed=500, device_reset=True, auto_reset=True, max_steps=60)
    env.reset(render=False)
    torch.cuda.synchronize()
    host = EpisodeSampler(N, domain_rand=False)
    host.seed([500 + k for k in range(N)])
    q = oracle_query(md)
    first = host.sample(list(range(N)), [md] * N, q)
    om = orc.OracleMap(md)
    cpu = [orc.OracleEnv(om, first["pos_x"][k], first["pos_z"][k], first["angle"][k], wheel_dist=first["wheel_dist"][k],
                         max_steps=60) for k in range(N)]
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert np.array_equal(st["pos_x"], first["pos_x"]) and np.array_equal(st["angle"], first["angle"])
    acts = np.random.default_rng(3).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    acts[:, :, 0] = np.abs(acts[:, :, 0])          # drive forward: episodes end quickly (off-road, duckies, max_steps)
    resets = 0
    for t in range(T):
        _, rew, done, info = env.step(torch.from_numpy(acts[t]).to(env.device), render=False)
        torch.cuda.synchronize()
        s = {k: v.cpu().numpy() for k, v in info.items()}
        d, r64 = done.cpu().numpy(), s["reward"]
        for k in range(N):
            o = cpu[k].step(acts[t, k])
            assert bool(d[k]) == bool(o.done) and s["done_code"][k] == o.done_code, (t, k)
            assert abs(r64[k] - o.reward) <= 1e-5 * max(1.0, abs(o.reward)), (t, k)
            if o.done:   # reference-style: the caller resets; the device already did
                nxt = host.sample([k], [md], q)
                cpu[k] = orc.OracleEnv(om, nxt["pos_x"][0], nxt["pos_z"][0], nxt["angle"][0],
                                       wheel_dist=nxt["wheel_dist"][0], max_steps=60)
                assert s["pos_x"][k] == nxt["pos_x"][0] and s["pos_z"][k] == nxt["pos_z"][0] and s["angle"][k] == nxt["angle"][0], (t, k)
                assert s["step_count"][k] == 0
                resets += 1
            else:
                assert abs(s["pos_x"][k] - o.pos_x) <= 1e-5 and abs(s["pos_z"][k] - o.pos_z) <= 1e-5, (t, k)
                assert s["step_count"][k] == o.step_count
    assert resets > 3 * N      # every env went through several episodes
    env.close()


def test_frame_skip_and_dynamics_rand_vs_oracle(torch_cuda):
    """frame_skip=3 (S:1673: three physics updates per action, step_count counts frames) and dynamics_rand
    (per-env trim on the motor gains, S:746-748) against the oracle."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    N, T = 16, 40
    env = make_env("small_loop", N, frame_skip=3, dynamics_rand=True, seed=321, max_steps=90)
    captured = {}
    orig = env.sim.reset
    env.sim.reset = lambda mask, params, stream=0: (captured.update(params), orig(mask, params, stream))[1]
    env.reset(render=False)
    torch.cuda.synchronize()
    assert np.abs(captured["trim"]).max() > 1e-4           # the Randomizer's normal(0, 0.02) draw is in use
    st0 = {k: v.cpu().numpy().copy() for k, v in env.state.items()}
    om = orc.OracleMap(maps.load_map("small_loop"))
    cpu = [orc.OracleEnv(om, st0["pos_x"][k], st0["pos_z"][k], st0["angle"][k], wheel_dist=st0["wheel_dist"][k],
                         frame_skip=3, max_steps=90, trim=float(captured["trim"][k])) for k in range(N)]
    acts = np.random.default_rng(8).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    acts[:, :, 0] = 0.25
    acts[:, :, 1] *= 0.2
    for t in range(T):
        _, rew, done, info = env.step(torch.from_numpy(acts[t]).to(env.device), render=False)
        torch.cuda.synchronize()
        s = {k: v.cpu().numpy() for k, v in info.items()}
        for k in range(N):
            o = cpu[k].step(acts[t, k])
            assert s["step_count"][k] == o.step_count == 3 * (t + 1)
            assert abs(s["pos_x"][k] - o.pos_x) <= 1e-5 and abs(s["pos_z"][k] - o.pos_z) <= 1e-5, (t, k)
            assert s["done_code"][k] == o.done_code and abs(s["speed"][k] - o.speed) <= 1e-5, (t, k)
    env.close()


def test_device_reset_honours_user_tile_start_and_start_pose(torch_cuda):
    """S:659-686: `user_tile_start` beats the map's start_tile and consumes no draw; a map `start_pose` fixes the
    pose inside the start tile.  The device reset must land where the host replay of the reference's reset() does."""
    torch = torch_cuda
    import copy
    from gym_duckietown_b200 import maps
    N = 16
    dev = make_env("udem1", N, device_reset=True, user_tile_start=(1, 1), seed=77)
    host = make_env("udem1", N, device_reset=False, user_tile_start=(1, 1), seed=77)
    dev.reset(render=False); host.reset(render=False)
    torch.cuda.synchronize()
    a = {k: v.cpu().numpy() for k, v in dev.state.items()}
    b = {k: v.cpu().numpy() for k, v in host.state.items()}
    assert np.all(a["tile_i"] == 1) and np.all(a["tile_j"] == 1)
    for k in ("pos_x", "pos_z", "angle"):
        assert np.array_equal(a[k], b[k]), k
    dev.close(); host.close()
    md = copy.deepcopy(maps.load_map("small_loop"))
    md.start_tile, md.start_pose = (1, 1), [[0.31, 0.0, 0.22], 1.25]
    env = make_env(md, 4, device_reset=True, seed=5)
    env.reset(render=False)
    s = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert np.all(s["pos_x"] == 1 * md.tile_size + 0.31) and np.all(s["pos_z"] == 1 * md.tile_size + 0.22)
    assert np.all(s["angle"] == 1.25)
    env.close()


def test_device_reset_fixed_starts_vs_reference_golden(golden_dir, torch_cuda):
    """Device-side reset with user_tile_start / start_tile / start_pose against the reference's own reset() (2 episodes)."""
    torch = torch_cuda
    import copy
    from gym_duckietown_b200 import maps
    g = np.load(os.path.join(golden_dir, "reset_start_udem1.npz"))
    n = len(g["seeds"])
    assert list(g["seeds"]) == list(range(n))
    base = maps.load_map("udem1")
    md_tile = copy.deepcopy(base); md_tile.start_tile = tuple(int(v) for v in g["start_tile"])
    md_pose = copy.deepcopy(md_tile)
    sp = g["start_pose"]
    md_pose.start_pose = [[float(sp[0]), float(sp[1]), float(sp[2])], float(sp[3])]
    cases = {"user": (base, dict(user_tile_start=tuple(int(v) for v in g["user_tile_start"]))),
             "tile": (md_tile, {}), "pose": (md_pose, {})}
    for tag, (md, kw) in cases.items():
        env = make_env(md, n, device_reset=True, seed=0, **kw)
        for ep in range(2):
            env.reset(render=False)
            torch.cuda.synchronize()
            s = {k: v.cpu().numpy() for k, v in env.state.items()}
            rows = np.arange(n) * 2 + ep
            assert np.array_equal(s["pos_x"], g[f"{tag}_cur_pos"][rows, 0]), (tag, ep)
            assert np.array_equal(s["pos_z"], g[f"{tag}_cur_pos"][rows, 2]), (tag, ep)
            assert np.array_equal(s["angle"], g[f"{tag}_cur_angle"][rows]), (tag, ep)
        env.close()
