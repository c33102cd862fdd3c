"""The C oracle (oracle/dt_oracle_logic.c) against golden vectors produced by the REFERENCE's own
code (oracle/make_golden.py -> tests/golden/logic_*.npz, action_map.npz)."""
import os

import numpy as np
import pytest

import oracle as orc
from gym_duckietown_b200 import maps

MAPS = ["small_loop", "loop_obstacles", "udem1"]


@pytest.mark.parametrize("name", MAPS)
def test_map_loader_matches_reference_load(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"logic_{name}.npz"))
    md = maps.load_map(name)
    assert np.array_equal(g["ref_curves"], md.curves)  # bit-exact
    assert [tuple(x) for x in g["ref_drivable_ij"]] == md.drivable_tiles
    if md.n_coll:
        assert np.array_equal(g["ref_coll_corners"], md.coll_corners)
        assert np.array_equal(g["ref_coll_norms"], md.coll_norms)
        assert np.array_equal(g["ref_coll_centers"], md.coll_centers)
        assert np.array_equal(g["ref_coll_radii"], md.coll_radii)


@pytest.mark.parametrize("name", MAPS)
def test_logic_oracle_vs_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"logic_{name}.npz"))
    om = orc.OracleMap(maps.load_map(name))
    n = len(g["poses"])
    for q in range(n):
        x, z, a = g["poses"][q]
        o = om.done_reward(x, z, a, int(g["step_count"][q]), int(g["max_steps"]))
        # integer / flag outputs: bit-exact
        assert (o.tile_i, o.tile_j) == (g["ti"][q], g["tj"][q]), q
        assert bool(o.done) == bool(g["done"][q]), q
        assert o.done_code == g["code"][q], q
        assert bool(o.collided) == bool(g["coll2"][q]), q
        assert bool(o.in_lane) == bool(g["inlane"][q]), q
        assert om.valid_pose(x, z, a, 1.3) == bool(g["valid13"][q]), q
        assert om.valid_pose(x, z, a, 1.0) == bool(g["valid10"][q]), q
        assert om.collision(x, z, a) == bool(g["coll1"][q]), q
        # floating point: 1e-9 abs (reference is numpy f64; differences are summation order)
        assert abs(o.prox - g["prox"][q]) <= 1e-12, q
        if g["inlane"][q]:
            assert abs(o.lane_dist - g["dist"][q]) <= 1e-12, q
            assert abs(o.lane_dot - g["dot"][q]) <= 1e-12, q
            assert abs(o.lane_angle - g["ang"][q]) <= 1e-7, q  # acos near |dot|=1 amplifies 1 ulp
        assert abs(o.reward - g["reward"][q]) <= 1e-9 * max(1.0, abs(g["reward"][q])), q


def test_action_map_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "action_map.npz"))
    for c, cfg in enumerate(g["cfgs"]):
        for q in range(len(g["actions"])):
            l, r = orc.action_map(g["actions"][q, 0], g["actions"][q, 1], g["wheel_dist"][q], *cfg)
            # reference ran under numpy>=2 scalar rules (float32 inputs stay float32 in part of the
            # expression); the port computes in float64 like numpy<=1.20, which the reference pins.
            assert abs(l - g["vels"][c, q, 0]) <= 2e-6 and abs(r - g["vels"][c, q, 1]) <= 2e-6


def test_dynamics_sanity_anchors():
    """duckietown_world is absent (parity unpinned): check the restated model's steady states,
    0.6 m/s at duty (1,1) and 7.5 rad/s at (-1,+1)."""
    om = orc.OracleMap(maps.load_map("small_loop"))
    env = orc.OracleEnv(om, 1.0, 1.0, 0.0, action_mode=0)
    for _ in range(200):
        env.step([1.0, 1.0])
    assert abs(env.s.u - 0.6) < 1e-6 and abs(env.s.w) < 1e-12
    env = orc.OracleEnv(om, 1.0, 1.0, 0.0, action_mode=0)
    for _ in range(200):
        env.step([-1.0, 1.0])
    assert abs(env.s.w - 7.5) < 1e-6 and abs(env.s.u) < 1e-12
    # command delay: nothing moves for the first 5 steps (0.15 s at 30 Hz)
    env = orc.OracleEnv(om, 1.0, 1.0, 0.3, action_mode=0)
    for k in range(5):
        o = env.step([1.0, 1.0])
        assert o.pos_x == 1.0 and o.speed == 0.0
    o = env.step([1.0, 1.0])
    assert o.speed > 0.0
