"""CPU checks of the raster oracle itself (no GPU): the two road-tile interpretations agree, the image has
the structure the reference scene must have, fisheye = LUT gather, determinism.  Pixels are "parity
unpinned" against the reference's OpenGL driver (cannot run here); these tests pin the oracle's internal
consistency and the reference-derived facts that do not need GL."""
import numpy as np
import pytest

import oracle as orc
from gym_duckietown_b200 import maps


@pytest.fixture(autouse=True)
def _restore_mode():
    yield
    orc.lib().orr_set_tile_mode(1)


def _poses(md, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        i, j = md.drivable_tiles[rng.integers(len(md.drivable_tiles))]
        out.append(((i + rng.uniform()) * md.tile_size, (j + rng.uniform()) * md.tile_size, rng.uniform(-np.pi, np.pi)))
    return out


@pytest.mark.parametrize("name", ["small_loop", "loop_obstacles", "udem1"])
def test_analytic_tiles_equal_literal_tessellation_up_to_edge_placement(name):
    """Tile mode 1 (one quad + analytic Gouraud lattice, the parity spec) vs mode 0 (the literal 98
    triangles of simulator.py:386-507): same image except rounding (+-1) and sub-1/64-px placement of
    tile outlines (mode 0 snaps the 8 lattice points of every tile border separately, so its outline is a 7-segment
    polyline within 1/128 px of mode 1's straight edge; a sample inside that sliver belongs to the other tile).
    There is NO 1-LSB bound between the two modes: a flipped sample changes its pixel by up to a quarter of the local
    contrast, and two can flip in one pixel.  Hard bounds asserted here (measured: 0.5-0.8 % / 0.05-0.1 % / 0.01-0.04 %):
    >1 LSB on < 1 % of the channel values, > 8 LSB on < 0.2 %, > 32 LSB on < 0.06 %, mean abs diff < 0.15 LSB."""
    md = maps.load_map(name)
    sc = orc.OracleScene(md)
    tot = big = big8 = big32 = 0
    absdiff = 0.0
    for x, z, a in _poses(md, 16, 3):
        orc.lib().orr_set_tile_mode(0)
        lit = sc.render(x, z, a).astype(int)
        orc.lib().orr_set_tile_mode(1)
        ana = sc.render(x, z, a).astype(int)
        d = np.abs(lit - ana)
        tot += d.size; big += int((d > 1).sum()); big8 += int((d > 8).sum()); big32 += int((d > 32).sum()); absdiff += d.sum()
    assert big / tot < 0.01, big / tot
    assert big8 / tot < 0.002 and big32 / tot < 0.0006, (big8 / tot, big32 / tot)
    assert absdiff / tot < 0.15


def test_scene_structure_small_loop():
    """Sky above the horizon is the clear colour (S:1753), the bottom rows are textured road or grass,
    an 84x84 and a 160x120 render of the same pose agree in mean colour (run_tests.py:17-22 style)."""
    md = maps.load_map("small_loop")
    sc = orc.OracleScene(md)
    ts = md.tile_size
    img = sc.render(1.5 * ts, 1.3 * ts, 0.0, W=160, H=120)
    sky = np.array([round(0.45 * 255), round(0.82 * 255), 255])
    assert (img[:20] == sky).all()
    assert img[100:].std() > 3 and (img[100:] != sky).any(axis=2).all()
    small = sc.render(1.5 * ts, 1.3 * ts, 0.0, W=84, H=84)
    assert abs(float(small.mean()) - float(img.mean())) < 12
    # determinism
    assert np.array_equal(img, sc.render(1.5 * ts, 1.3 * ts, 0.0, W=160, H=120))


def test_hidden_object_and_domain_rand_inputs_change_the_image():
    md = maps.load_map("loop_obstacles")
    sc = orc.OracleScene(md)
    ob = md.objects[0]
    x, z = ob.pos[0] - 0.35, ob.pos[2]
    base = sc.render(x, z, 0.0)
    ep = orc.default_episode()
    ep.hidden[0] = 1
    hidden = sc.render(x, z, 0.0, ep)
    assert (base != hidden).any()          # the duckie in front of the camera disappeared
    ep2 = orc.default_episode(horizon=(0.1, 0.2, 0.3), ground=(0.3, 0.1, 0.1), cam_fov_y_deg=60.0)
    assert (sc.render(x, z, 0.0, ep2)[0, 0] == np.array([26, 51, 76])).all()   # rint(255*c) of the new clear colour
    # camera noise only applies under domain_rand (S:1768)
    ep3 = orc.default_episode(cam_noise=(0.004, 0.004, -0.004))
    assert np.array_equal(sc.render(x, z, 0.0, ep3, domain_rand=False), base)
    assert (sc.render(x, z, 0.0, ep3, domain_rand=True) != base).any()


def test_fisheye_is_lut_gather():
    from gym_duckietown_b200.distortion import Distortion
    md = maps.load_map("small_loop")
    sc = orc.OracleScene(md)
    d = Distortion(160, 120)
    ts = md.tile_size
    plain = sc.render(1.5 * ts, 1.3 * ts, 0.1, W=160, H=120)
    fused = sc.render(1.5 * ts, 1.3 * ts, 0.1, W=160, H=120, lut=(d.rmapx, d.rmapy))
    assert np.array_equal(fused, d.distort(plain))
