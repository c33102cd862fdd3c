"""GPU parity of the CUDA rasteriser (dts_render / dts_step through the C ABI) against the CPU raster
oracle (oracle/dt_oracle_raster.c), which defines pixel truth for this project (pixels are "parity
unpinned" w.r.t. the reference's OpenGL driver — see the oracle header).
Bar: <= 1 LSB per channel on every pixel (north_star); the spec'd arithmetic makes 0 the expectation."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def oracle_episode(orc, params, k, light_eye=None):
    ep = orc.default_episode()
    if params is None:
        return ep
    ep.cam_height = float(params["cam_height"][k]); ep.cam_angle_deg = float(params["cam_angle_deg"][k])
    ep.cam_fov_y_deg = float(params["cam_fov_y_deg"][k])
    for name, field in (("cam_noise", "cam_noise"), ("horizon_color", "horizon"), ("light_ambient", "ambient"),
                        ("light_diffuse", "diffuse"), ("ground_color", "ground")):
        for i in range(3):
            getattr(ep, field)[i] = float(np.float32(params[name][k][i]))
    for i in range(4):
        ep.light_eye[i] = float(np.float32(params["light_pos"][k][i]))  # first episode: identity modelview
    for i in range(8):
        ep.hidden[i] = int(params["obj_hidden"][k][i])
    return ep


def compare(gpu, cpu, tag):
    diff = np.abs(gpu.astype(np.int16) - cpu.astype(np.int16))
    bad = int((diff > 1).sum())
    out = os.environ.get("DTS_DUMP_DIR")
    if out and (bad or os.environ.get("DTS_DUMP_ALL")):
        from PIL import Image
        os.makedirs(out, exist_ok=True)
        k = int(np.argmax(diff.reshape(len(gpu), -1).max(1)))
        Image.fromarray(np.concatenate([gpu[k], cpu[k], np.minimum(diff[k] * 32, 255).astype(np.uint8)], 1)).save(
            os.path.join(out, f"{tag}.png"))
    assert bad == 0, f"{tag}: {bad} channel values differ by more than 1 LSB (max {diff.max()})"
    return int(diff.max()), float((diff > 0).mean())


@pytest.fixture(autouse=True)
def _default_tile_mode():
    import oracle as orc
    orc.lib().orr_set_tile_mode(1)
    yield
    orc.lib().orr_set_tile_mode(1)


@pytest.mark.parametrize("name,W,H,dr,tess", [
    ("small_loop", 160, 120, False, False), ("loop_obstacles", 160, 120, False, False), ("udem1", 160, 120, False, False),
    ("udem1", 160, 120, True, False), ("small_loop", 84, 84, False, False), ("loop_obstacles", 320, 240, True, False),
    ("small_loop", 160, 120, False, True), ("udem1", 160, 120, True, True), ("loop_obstacles", 90, 70, False, False),
])
def test_first_frame_vs_oracle(name, W, H, dr, tess, torch_cuda):
    """reset() with host-drawn (reference-order) episode parameters, compare the first observation."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    N = 48
    orc.lib().orr_set_tile_mode(0 if tess else 1)   # literal 98-triangle tiles vs one quad + analytic lattice lighting
    env = BatchedDuckietownEnv(N, name, camera_width=W, camera_height=H, domain_rand=dr, seed=1000, tessellate_tiles=tess)
    captured = {}
    orig = env.sim.reset
    env.sim.reset = lambda mask, params, stream=0: (captured.update(params), orig(mask, params, stream))[1]
    obs = env.reset()
    torch.cuda.synchronize()
    gpu = obs.cpu().numpy()
    sc = orc.OracleScene(maps.load_map(name))
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    cpu = np.stack([sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], oracle_episode(orc, captured, k), W, H, dr)
                    for k in range(N)])
    mx, frac = compare(gpu, cpu, f"first_{name}_{W}x{H}_{'dr' if dr else 'nodr'}{'_tess' if tess else ''}")
    assert gpu.std() > 10  # not a blank image
    env.close()


@pytest.mark.parametrize("name", ["loop_obstacles", "udem1"])
def test_rollout_frames_vs_oracle(name, torch_cuda):
    """Step with random actions (no auto-reset): poses wander over tile borders, off the road and
    into obstacles; every 10th frame is compared."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    N, T, W, H = 32, 60, 160, 120
    env = BatchedDuckietownEnv(N, name, camera_width=W, camera_height=H, domain_rand=False, seed=2000)
    env.reset()
    sc = orc.OracleScene(maps.load_map(name))
    acts = np.random.default_rng(99).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    acts[:, :, 0] = np.abs(acts[:, :, 0]) * 0.8 + 0.2  # drive forward so the view changes
    for t in range(T):
        obs, _, _, info = env.step(torch.from_numpy(acts[t]).to(env.device))
        if t % 10 == 9:
            torch.cuda.synchronize()
            st = {k: v.cpu().numpy() for k, v in info.items()}
            cpu = np.stack([sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], None, W, H, False) for k in range(N)])
            compare(obs.cpu().numpy(), cpu, f"roll_{name}_t{t}")
    env.close()


def test_adversarial_poses_vs_oracle(torch_cuda):
    """Cameras placed by hand: on tile corners, outside the map looking in, nose against a duckie,
    exactly axis-aligned headings — the clipper and the fill rule get exercised."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    md = maps.load_map("loop_obstacles")
    ts = md.tile_size
    poses = []
    for ang in (0.0, np.pi / 2, np.pi, -np.pi / 2, 0.3, 2.0):
        poses += [(1.0 * ts, 1.0 * ts, ang), (2.0 * ts, 1.5 * ts, ang), (-0.5 * ts, 3.0 * ts, ang),
                  (4.75 * ts - 0.09, 1.25 * ts, ang), (3.0 * ts, 3.0 * ts, ang), (8.5 * ts, 7.5 * ts, ang)]
    poses = np.array(poses)
    N, W, H = len(poses), 160, 120
    env = BatchedDuckietownEnv(N, "loop_obstacles", camera_width=W, camera_height=H, domain_rand=False, seed=1)
    env.sim.reset(None, dict(pos_x=poses[:, 0], pos_z=poses[:, 1], angle=poses[:, 2]))
    obs = env.render_obs()
    torch.cuda.synchronize()
    sc = orc.OracleScene(md)
    cpu = np.stack([sc.render(x, z, a, None, W, H, False) for x, z, a in poses])
    compare(obs.cpu().numpy(), cpu, "adversarial")
    env.close()


def test_many_envs_per_cta_vs_oracle(torch_cuda):
    """More envs than resident CTAs: every persistent CTA renders several frames back to back and reuses
    its slab / lattice table / staging buffers; the LAST envs of the batch are compared."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    N, W, H = 3000, 84, 84
    env = BatchedDuckietownEnv(N, "loop_obstacles", camera_width=W, camera_height=H, domain_rand=True, seed=77,
                               device_reset=True, auto_reset=True)
    obs = env.reset()
    acts = torch.rand((N, 2), device=env.device) * 2 - 1
    for _ in range(3):
        obs, _, _, info = env.step(acts)
    torch.cuda.synchronize()
    # device-side DR: read the episode records back through a second render of identical state is not
    # possible from Python, so use domain_rand params the oracle can see: re-reset with host params
    env2 = BatchedDuckietownEnv(N, "loop_obstacles", camera_width=W, camera_height=H, domain_rand=False, seed=77)
    st = {k: v.cpu().numpy() for k, v in info.items()}
    env2.sim.reset(None, dict(pos_x=st["pos_x"], pos_z=st["pos_z"], angle=st["angle"]))
    obs2 = env2.render_obs()
    torch.cuda.synchronize()
    sc = orc.OracleScene(maps.load_map("loop_obstacles"))
    sel = list(range(N - 40, N)) + list(range(1500, 1520))
    cpu = np.stack([sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], None, W, H, False) for k in sel])
    compare(obs2.cpu().numpy()[sel], cpu, "many_envs")
    env.close(); env2.close()


def test_second_episode_stale_light_vs_oracle(torch_cuda):
    """S:581: from the second reset on GL_LIGHT0's position is captured under the previous frame's camera
    matrix.  Device auto-reset does that on its own; the oracle is handed the eye-space light the product reports."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    N, W, H = 24, 160, 120
    env = BatchedDuckietownEnv(N, "udem1", camera_width=W, camera_height=H, domain_rand=True, seed=40,
                               device_reset=True, auto_reset=True)
    env.reset()
    obs = env.reset()          # second episode: stale modelview
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    assert (st["episode"] == 2).all()
    sc = orc.OracleScene(maps.load_map("udem1"))
    gpu = obs.cpu().numpy()
    cpu = []
    lights = []
    for k in range(N):
        r = env.sim.debug_episode(k)
        ep = orc.default_episode(cam_height=float(r["cam_height"]), cam_angle_deg=float(r["cam_angle_deg"]),
                                 cam_fov_y_deg=float(r["cam_fov_y_deg"]), cam_noise=r["cam_noise"], horizon=r["horizon"],
                                 ambient=r["ambient"], diffuse=r["diffuse"], light_eye=r["light_eye"], ground=r["ground"],
                                 hidden=[int(v) for v in r["hidden"]])
        lights.append(r["light_eye"].copy())
        cpu.append(sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], ep, W, H, True))
    lights = np.array(lights)
    assert (lights[:, 3] == 0).all() and np.abs(lights[:, :3]).max() > 100   # directional DR light, rotated into eye space
    assert np.abs(lights[:, 1] - np.clip(lights[:, 1], 170, 220)).max() > 1   # ... i.e. no longer the raw world-space draw
    compare(gpu, np.stack(cpu), "stale_light")
    env.close()


def test_full_resolution_frame_vs_oracle(torch_cuda):
    """640x480 (the reference's default camera), no distortion: many coarse bins, long far-field lists."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    N, W, H = 6, 640, 480
    env = BatchedDuckietownEnv(N, "udem1", camera_width=W, camera_height=H, domain_rand=False, seed=9)
    obs = env.reset()
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    sc = orc.OracleScene(maps.load_map("udem1"))
    cpu = np.stack([sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], None, W, H, False) for k in range(N)])
    compare(obs.cpu().numpy(), cpu, "full_res")
    env.close()


@pytest.mark.parametrize("name,segment,top_down,W,H", [
    ("loop_obstacles", True, False, 160, 120), ("udem1", True, False, 160, 120), ("small_loop", False, True, 160, 120),
    ("udem1", False, True, 200, 150), ("loop_obstacles", True, True, 160, 120)])
def test_segment_and_top_down_views_vs_oracle(name, segment, top_down, W, H, torch_cuda):
    """render_obs(segment=True) / _render_img(top_down=True) (dts_set_render_mode): lighting off, magenta clear and ground,
    segmentation textures and flat class colours; camera above the map with the agent's own mesh — bit-exact vs the oracle."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    N = 24
    env = BatchedDuckietownEnv(N, name, camera_width=W, camera_height=H, domain_rand=False, seed=77)
    env.reset(render=False)
    plain = env.render_obs().cpu().numpy().copy()
    got = env.render_obs(segment=segment, top_down=top_down).cpu().numpy().copy()
    again = env.render_obs().cpu().numpy()
    assert np.array_equal(plain, again)                       # the mode does not stick
    torch.cuda.synchronize()
    st = {k: v.cpu().numpy() for k, v in env.state.items()}
    sc = orc.OracleScene(maps.load_map(name))
    ref = np.stack([sc.render(st["pos_x"][k], st["pos_z"][k], st["angle"][k], None, W, H, segment=segment, top_down=top_down)
                    for k in range(N)])
    compare(got, ref, f"modes_{name}_{int(segment)}{int(top_down)}")
    assert (got != plain).mean() > 0.2
    if segment and not top_down:   # the sky is the magenta clear colour (an object may cover the corner pixel of some frames)
        assert (got[:, 0, 0] == np.array([255, 0, 255])).all(axis=1).mean() > 0.6
    env.check()
    env.close()


@pytest.mark.parametrize("name,W,H", [("loop_obstacles", 160, 120), ("loop_obstacles", 84, 84), ("udem1", 160, 120), ("udem1", 320, 240)])
def test_small_triangle_path_vs_oracle(name, W, H, torch_cuda):
    """Agents parked at 0.15 .. 2.5 m from the map's props, facing them: the props' triangles range from a few dozen pixels
    down to sub-pixel size, so both raster paths (one warp per prim; one LANE per tiny triangle with the shared
    depth / winner buffer) and the merge between them are exercised on the same frames — bit-exact vs the oracle."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    md = maps.load_map(name)
    rng = np.random.default_rng(5)
    poses = []
    for o in md.objects:
        for d in (0.15, 0.3, 0.5, 0.8, 1.2, 1.8, 2.5):
            a = rng.uniform(-np.pi, np.pi)
            x, z = o.pos[0] - d * np.cos(a), o.pos[2] + d * np.sin(a)      # get_dir_vec(a) = (cos a, 0, -sin a) points at the prop
            poses.append((x, z, a + rng.uniform(-0.25, 0.25)))
    poses = poses[:96]
    N = len(poses)
    env = BatchedDuckietownEnv(N, name, camera_width=W, camera_height=H, domain_rand=False, seed=3)
    P = np.array(poses)
    env.sim.reset(None, dict(pos_x=P[:, 0].copy(), pos_z=P[:, 1].copy(), angle=P[:, 2].copy(), map_id=np.zeros(N, np.int32)),
                  env._stream())
    got = env.render_obs().cpu().numpy()
    sc = orc.OracleScene(md)
    ref = np.stack([sc.render(x, z, a, None, W, H) for x, z, a in poses])
    compare(got, ref, f"tiny_{name}_{W}")
    env.check()
    env.close()


@pytest.mark.parametrize("name", ["small_loop", "loop_obstacles"])
def test_large_batch_exact_and_order_independent(name, torch_cuda):
    """2048 random cameras of one map: every frame EQUAL to the oracle's (the spec'd arithmetic makes 0 LSB the
    expectation, and rare events — a sample covered by two neighbouring tiles whose shared border snapped differently,
    depth ties, bins with > 32 records — only show up in large batches), and two renders of the same state equal each
    other (the bin lists are built with atomics: nothing may depend on their order)."""
    torch = torch_cuda
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    md = maps.load_map(name)
    ts = md.tile_size
    rng = np.random.default_rng(2024)
    N, W, H = 2048, 160, 120
    cells = np.array(md.drivable_tiles)[rng.integers(len(md.drivable_tiles), size=N)]
    px = (cells[:, 0] + rng.uniform(size=N)) * ts
    pz = (cells[:, 1] + rng.uniform(size=N)) * ts
    ang = rng.uniform(-np.pi, np.pi, size=N)
    env = BatchedDuckietownEnv(N, name, camera_width=W, camera_height=H, domain_rand=False, seed=5)
    env.sim.reset(None, dict(pos_x=px, pos_z=pz, angle=ang))
    a = env.render_obs().clone()
    b = env.render_obs(out=torch.empty_like(a))
    torch.cuda.synchronize()
    assert torch.equal(a, b), "two renders of the same state differ"
    sc = orc.OracleScene(md)
    cpu = sc.render_batch(px, pz, ang, [orc.default_episode() for _ in range(N)], W, H, False, threads=os.cpu_count() or 1)
    mx, frac = compare(a.cpu().numpy(), cpu, f"large_batch_{name}")
    assert mx == 0, f"{name}: max diff {mx} LSB on {frac:.2e} of the channel values"
    env.close()
