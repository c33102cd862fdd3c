"""Host-side reset sampling (episode.py) against Simulator.reset() of the REFERENCE (golden vectors
from oracle/make_golden.py).  The spawn predicates come from the C oracle here (CPU suite); the GPU
suite repeats it through dts_query_poses."""
import os

import numpy as np
import pytest

import oracle as orc
from gym_duckietown_b200 import maps
from gym_duckietown_b200.episode import EpisodeSampler

MAPS = ["small_loop", "loop_obstacles", "udem1"]


def oracle_query(md):
    om = orc.OracleMap(md)

    def query(k, x, z, a, safety, hidden):
        n = len(x)
        outd, outi = np.full((n, 4), np.nan), np.zeros((n, 8), np.int32)
        for q in range(n):
            o = om.done_reward(x[q], z[q], a[q], 0)
            outd[q] = (o.lane_dist, o.lane_dot, o.lane_angle, o.prox)
            outi[q, 0] = om.valid_pose(x[q], z[q], a[q], safety)
            outi[q, 3] = o.in_lane
            bad = False
            for oi, ob in enumerate(md.objects):  # _inconvenient_spawn S:1461-1471
                if hidden[q, oi >> 5] >> (oi & 31) & 1:
                    continue
                d = np.linalg.norm(ob.pos - np.array([x[q], 0, z[q]]))
                bad |= bool(d < max(ob.max_coords) * 0.5 * ob.scale + 0.25)
            outi[q, 4] = bad
        return outd, outi
    return query


def check_against_golden(name, golden_dir, make_query):
    g = np.load(os.path.join(golden_dir, f"reset_{name}.npz"))
    md = maps.load_map(name)
    for tag, dr in (("nodr", False), ("dr", True)):
        s = EpisodeSampler(len(g["seeds"]), domain_rand=dr)
        s.seed([int(v) for v in g["seeds"]])
        envs = list(range(len(g["seeds"])))
        for ep in range(2):
            out = s.sample(envs, [md] * len(envs), make_query(md))
            rows = np.arange(len(envs)) * 2 + ep
            assert np.array_equal(out["pos_x"], g[f"{tag}_cur_pos"][rows, 0])
            assert np.array_equal(out["pos_z"], g[f"{tag}_cur_pos"][rows, 2])
            assert np.array_equal(out["angle"], g[f"{tag}_cur_angle"][rows])
            assert np.array_equal(out["wheel_dist"], g[f"{tag}_wheel_dist"][rows])
            assert np.array_equal(out["cam_height"], g[f"{tag}_cam_height"][rows])
            assert np.array_equal(out["cam_angle_deg"], g[f"{tag}_cam_angle"][rows])
            assert np.array_equal(out["cam_fov_y_deg"], g[f"{tag}_cam_fov_y"][rows])
            assert np.array_equal(out["horizon_color"], g[f"{tag}_horizon_color"][rows])
            assert np.array_equal(out["ground_color"], g[f"{tag}_ground_color"][rows])
            # the reference hands these to GL as float32 (ctypes GLfloat arrays, S:581-583)
            assert np.array_equal(out["light_pos"].astype(np.float32), g[f"{tag}_light_pos"][rows].astype(np.float32))
            assert np.array_equal(out["light_ambient"].astype(np.float32), g[f"{tag}_ambient"][rows, :3].astype(np.float32))
            assert np.array_equal(out["light_diffuse"].astype(np.float32), g[f"{tag}_diffuse"][rows, :3].astype(np.float32))
            if dr:
                assert np.array_equal(out["cam_noise"], g[f"{tag}_camera_noise"][rows])
            vis = g[f"{tag}_obj_visible"][rows]
            for q in range(len(envs)):
                for oi in range(vis.shape[1]):
                    assert bool(out["obj_hidden"][q, oi >> 5] >> (oi & 31) & 1) == (not vis[q, oi])


@pytest.mark.parametrize("name", MAPS)
def test_reset_draws_match_reference(name, golden_dir):
    check_against_golden(name, golden_dir, oracle_query)


def test_fixed_starts_match_reference(golden_dir):
    """user_tile_start / map start_tile / map start_pose (S:659-686) as executed by the reference's reset()."""
    import copy
    g = np.load(os.path.join(golden_dir, "reset_start_udem1.npz"))
    base = maps.load_map("udem1")
    seeds = [int(v) for v in g["seeds"]]
    envs = list(range(len(seeds)))
    md_tile = copy.deepcopy(base); md_tile.start_tile = tuple(int(v) for v in g["start_tile"])
    md_pose = copy.deepcopy(md_tile)
    sp = g["start_pose"]
    md_pose.start_pose = [[float(sp[0]), float(sp[1]), float(sp[2])], float(sp[3])]
    cases = {"user": (base, dict(user_tile_start=tuple(int(v) for v in g["user_tile_start"]))),
             "tile": (md_tile, {}), "pose": (md_pose, {})}
    for tag, (md, kw) in cases.items():
        s = EpisodeSampler(len(seeds), domain_rand=False, **kw)
        s.seed(seeds)
        for ep in range(2):
            out = s.sample(envs, [md] * len(envs), oracle_query(md))
            rows = np.arange(len(envs)) * 2 + ep
            assert np.array_equal(out["pos_x"], g[f"{tag}_cur_pos"][rows, 0]), tag
            assert np.array_equal(out["pos_z"], g[f"{tag}_cur_pos"][rows, 2]), tag
            assert np.array_equal(out["angle"], g[f"{tag}_cur_angle"][rows]), tag


def test_map_choice_draw_is_a_bounded_integer_draw():
    """randomize_maps_on_reset picks the map with `np_random.choice(self.map_names)` (S:541-542): for a Generator that
    is one `integers(0, n)` draw — which is what the host path and the device's Lemire sampler replay."""
    names = [f"map{k}" for k in range(7)]
    for seed in range(50):
        a, b = np.random.default_rng(seed), np.random.default_rng(seed)
        a.uniform(); b.uniform()                       # mid-stream, with numpy's cached 32-bit half possibly pending
        for _ in range(5):
            assert a.choice(names) == names[int(b.integers(0, len(names)))]
        assert a.bit_generator.state == b.bit_generator.state


def test_custom_randomizer_table_and_sim_colours_match_reference(golden_dir):
    """Randomizer(randomization_config_fp=custom) (randomizer.py:19-33: other ranges, extra keys drawn in sorted
    position) + non-default num_tris_distractors / color_sky / color_ground: the host sampler replays the
    reference's reset() draw for draw (golden from the reference's own Randomizer.randomize)."""
    import json
    g = np.load(os.path.join(golden_dir, "reset_customdr_loop_obstacles.npz"))
    cfg, simkw = json.loads(str(g["config_json"])), json.loads(str(g["sim_json"]))
    md = maps.load_map("loop_obstacles")
    seeds = [int(v) for v in g["seeds"]]
    s = EpisodeSampler(len(seeds), domain_rand=True, dynamics_rand=True, randomization_config=cfg, **simkw)
    s.seed(seeds)
    envs = list(range(len(seeds)))
    for ep in range(2):
        out = s.sample(envs, [md] * len(envs), oracle_query(md))
        rows = np.arange(len(envs)) * 2 + ep
        for ours, theirs in (("pos_x", None), ("angle", "cur_angle"), ("wheel_dist", "wheel_dist"), ("cam_height", "cam_height"),
                             ("cam_angle_deg", "cam_angle"), ("cam_fov_y_deg", "cam_fov_y"), ("horizon_color", "horizon_color"),
                             ("ground_color", "ground_color"), ("cam_noise", "camera_noise"), ("trim", "trim")):
            want = g["dr_cur_pos"][rows, 0] if theirs is None else g[f"dr_{theirs}"][rows]
            assert np.array_equal(out[ours], want), ours
        assert np.array_equal(out["light_pos"].astype(np.float32), g["dr_light_pos"][rows].astype(np.float32))
        assert np.array_equal(out["light_ambient"].astype(np.float32), g["dr_ambient"][rows, :3].astype(np.float32))


def test_dr_ops_from_config_is_sorted_and_typed():
    import json
    from gym_duckietown_b200 import lib as L
    g_cfg = {"trim": {"type": "normal", "loc": 0, "scale": 0.02}, "horz_mode": {"type": "int", "low": 0, "high": 4},
             "zz": {"type": "uniform", "low": 0, "high": 1, "size": 5},
             "light_pos": {"type": "uniform", "low": [-1, 2, -3], "high": [1, 3, 3], "size": 3}}
    ops = L.dr_ops_from_config(g_cfg)
    assert [o[2] for o in ops] == [L.DR_TARGETS["horz_mode"], L.DR_TARGETS["light_pos"], L.DR_TARGETS["trim"], 0]
    assert [o[0] for o in ops] == [L.DR_INT, L.DR_UNIFORM, L.DR_NORMAL, L.DR_UNIFORM] and ops[3][1] == 5
    assert list(ops[1][3]) == [-1, 2, -3] and list(ops[1][4]) == [1, 3, 3]
