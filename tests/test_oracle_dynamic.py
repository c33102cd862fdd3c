"""Dynamic obstacles: the C oracle (oracle/dt_oracle_dynamic.c) against DuckieObj / DuckiebotObj of the
REFERENCE stepped by its own code (tests/golden/dynamic_*.npz from oracle/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle as orc
from gym_duckietown_b200 import maps


@pytest.mark.parametrize("name", ["loop_pedestrians", "loop_dyn_duckiebots"])
def test_dynamic_objects_vs_reference(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"dynamic_{name}.npz"))
    md = maps.load_map(name)
    om = orc.OracleMap(md)
    dyn = orc.OracleDynamics(om, wiggle=g["wiggle"])
    T = len(g["pos"])
    qi = 0
    for t in range(T):
        dyn.step()
        for i in range(dyn.n):
            o = dyn.objs[i]
            assert np.abs(np.array(o.pos[:]) - g["pos"][t, i]).max() <= 1e-12, (t, i)
            assert abs(o.angle - g["angle"][t, i]) <= 1e-12 and abs(o.y_rot - g["y_rot"][t, i]) <= 1e-9, (t, i)
            assert np.abs(np.array([list(c) for c in o.corners]) - g["corners"][t, i]).max() <= 1e-12, (t, i)
            assert bool(o.active) == bool(g["active"][t, i]), (t, i)
        while qi < len(g["q_step"]) and g["q_step"][qi] == t:
            x, z, a = g["q_pose"][qi]
            hit = om.collision(x, z, a) or dyn.collision(x, z, a)          # S:1481-1489: static, then dynamic
            assert hit == bool(g["q_coll"][qi]), (t, qi)
            prox = om.proximity(x, z, a) + dyn.proximity(x, z, a)          # S:1454-1457
            assert abs(prox - g["q_prox"][qi]) <= 1e-12, (t, qi)
            qi += 1
    assert qi == len(g["q_step"]) and g["q_coll"].any() and not g["q_coll"].all()


@pytest.mark.parametrize("tag", ["plain", "dr"])
def test_traffic_lights_vs_reference(tag, golden_dir):
    """TrafficLightObj.step (O:455-462): per-light pattern and the card of the mesh all lights share."""
    g = np.load(os.path.join(golden_dir, "trafficlight_loop_trafficlights.npz"))
    md = maps.load_map("loop_trafficlights")
    om = orc.OracleMap(md)
    tl = [i for i, d in enumerate(md.dyn_objects) if d.kind == maps.DYN_TRAFFICLIGHT]
    assert [md.dyn_objects[i].object_index for i in tl] == list(g["tl_index"])
    freq, pat = [5.0] * len(md.dyn_objects), [0] * len(md.dyn_objects)
    for k, i in enumerate(tl):
        freq[i], pat[i] = float(g[f"{tag}_freq"][k]), int(g[f"{tag}_pattern0"][k])
    dyn = orc.OracleDynamics(om, freq=freq, pattern=pat)
    assert dyn.shown_card == int(g[f"{tag}_shown0"])
    for t in range(len(g[f"{tag}_shown"])):
        dyn.step()
        assert [dyn.objs[i].active for i in tl] == list(g[f"{tag}_pattern"][t]), t
        assert dyn.shown_card == g[f"{tag}_shown"][t], t
