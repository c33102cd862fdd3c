"""GPU side of the GL call-trace pin (tests/test_gltrace.py has the CPU side): what k_frame_setup / k_geometry put into
frame memory — camera model-view, projection, the lit 8x8 lattice of every visible road tile — against (a) the
reference's own GL calls recorded in tests/golden/gltrace_<map>.npz (float32 tolerance) and (b) the raster oracle
(exact), and the device's stale-model-view light capture (S:581) against the eye-space GL_POSITION of the trace."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
F32 = 2.0 ** -23


def _close(a, b, scale=1.0, ulps=4):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) <= ulps * F32 * np.maximum(scale, np.maximum(np.abs(a), np.abs(b)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _params(g, idx, light_key="f_light_eye", stale=0):
    n = len(idx)
    vis = g["f_visible"][idx]
    hidden = np.zeros((n, 8), np.uint32)
    for r in range(n):
        for o, v in enumerate(vis[r]):
            if not v:
                hidden[r, o >> 5] |= np.uint32(1 << (o & 31))
    dr = g["f_dr"][idx][:, None]
    return dict(pos_x=g["f_pos"][idx][:, 0].copy(), pos_z=g["f_pos"][idx][:, 2].copy(), angle=g["f_angle"][idx].copy(),
                cam_height=g["f_cam_height"][idx], cam_angle_deg=g["f_cam_angle"][idx], cam_fov_y_deg=g["f_cam_fov_y"][idx],
                cam_noise=np.where(dr, g["f_camera_noise"][idx], 0.0), horizon_color=g["f_horizon"][idx],
                light_ambient=g["f_light_ambient"][idx][:, :3], light_diffuse=g["f_light_diffuse"][idx][:, :3],
                light_pos=g[light_key][idx], light_stale=np.full(n, stale, np.int32), ground_color=g["f_ground"][idx],
                obj_hidden=hidden, map_id=np.zeros(n, np.int32))


@pytest.mark.parametrize("name", ["small_loop", "loop_obstacles", "udem1"])
def test_frame_setup_and_lattice_vs_gl_trace_and_oracle(name, torch_cuda):
    torch = torch_cuda
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    from gym_duckietown_b200 import maps
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv
    from test_gltrace import _episode

    md = maps.load_map(name)
    g = np.load(os.path.join(GOLD, f"gltrace_{name}.npz"))
    sc = orc.OracleScene(md)
    W, H = int(g["width"]), int(g["height"])
    cells = md.grid_w * md.grid_h
    for dr in (False, True):
        idx = np.flatnonzero((g["f_dr"] == dr) & (g["f_mode"] == 0))
        env = BatchedDuckietownEnv(len(idx), name, camera_width=W, camera_height=H, domain_rand=dr, seed=1)
        env.sim.reset(None, _params(g, idx), env._stream())
        env.render_obs()
        torch.cuda.synchronize()
        worst_lat = 0.0
        n_seen = 0
        for r, f in enumerate(idx):
            d = env.sim.debug_frame(r, cells)
            view = g["f_view"][f].reshape(4, 4)
            assert _close(d["V"].reshape(3, 4), view[:3]).all(), f"frame {f}: camera model-view vs gluLookAt trace"
            proj = g["f_proj"][f].reshape(4, 4)
            assert _close(d["P"], [proj[0, 0], proj[1, 1], proj[2, 2], proj[2, 3]]).all()
            o = sc.debug_frame(g["f_pos"][f][0], g["f_pos"][f][2], g["f_angle"][f], _episode(orc, g, f), W, H, dr)
            assert np.allclose(d["V"], o["V"], rtol=0, atol=1e-13) and np.array_equal(d["P"], o["P"])
            seen = ~np.isnan(d["lattice"][:, 0, 0])   # tiles this frame emitted (a camera at the map's edge may see none)
            assert d["overflow"] == 0
            n_seen += int(seen.sum())
            assert np.array_equal(d["lattice"][seen], o["lattice"][seen]), f"frame {f}: lit lattice != raster oracle"
            if seen.any():
                worst_lat = max(worst_lat, float(np.abs(d["lattice"][seen] - o["lattice"][seen]).max()))
        assert n_seen > 4 * len(idx)
        env.close()


@pytest.mark.parametrize("name", ["small_loop", "udem1"])
def test_device_stale_light_capture_vs_gl_trace(name, torch_cuda):
    """Second-episode resets hand the RAW light position with light_stale=1: the device multiplies it by the camera
    matrix of the env's previous pose (k_reset_params / respawn), as GL did with the model-view left on the stack."""
    torch = torch_cuda
    from gym_duckietown_b200.batched_env import BatchedDuckietownEnv

    g = np.load(os.path.join(GOLD, f"gltrace_{name}.npz"))
    W, H = int(g["width"]), int(g["height"])
    second = np.flatnonzero((g["f_k"] == 0) & (g["f_episode"] == 1) & (g["f_mode"] == 0))
    for dr in (False, True):
        idx = second[g["f_dr"][second] == dr]
        prev = idx - 1            # last frame of episode 0 of the same simulator
        env = BatchedDuckietownEnv(len(idx), name, camera_width=W, camera_height=H, domain_rand=dr, seed=1)
        env.sim.reset(None, _params(g, prev), env._stream())                                  # put every env where episode 0 ended
        env.sim.reset(None, _params(g, idx, light_key="f_light_raw", stale=1), env._stream())   # reset #2: stale capture
        torch.cuda.synchronize()
        for r, f in enumerate(idx):
            got = env.sim.debug_episode(r)["light_eye"]
            want = g["f_light_eye"][f]
            assert _close(got, want, float(np.abs(g["f_light_raw"][f]).max()), ulps=6).all(), (f, got, want)
        env.close()
