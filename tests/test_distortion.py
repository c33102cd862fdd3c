"""Fisheye LUT (gym-duckietown_b200/distortion.py) against the reference's Distortion class
(tests/golden/fisheye.npz, produced by oracle/make_golden.py running distortion.py itself)."""
import hashlib
import os

import numpy as np


def test_fisheye_lut_bit_identical_to_reference(golden_dir):
    from gym_duckietown_b200.distortion import Distortion
    g = np.load(os.path.join(golden_dir, "fisheye.npz"))
    d = Distortion(640, 480)
    assert np.array_equal(d.new_camera_matrix, g["new_camera_matrix"])
    assert np.array_equal(d.rmapx[::8, ::8], g["rmapx_sub"]) and np.array_equal(d.rmapy[::8, ::8], g["rmapy_sub"])
    assert hashlib.sha256(d.rmapx.astype(np.float32).tobytes()).hexdigest() == str(g["sha_rmapx"])
    assert hashlib.sha256(d.rmapy.astype(np.float32).tobytes()).hexdigest() == str(g["sha_rmapy"])
    # the gather itself: cv2.remap(INTER_NEAREST) of the reference == rint() indexing
    img = np.random.default_rng(int(g["img_seed"])).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    out = d.distort(img)
    assert np.array_equal(out[::8, ::8], g["out_sub"])
    assert hashlib.sha256(out.tobytes()).hexdigest() == str(g["out_sha"])
