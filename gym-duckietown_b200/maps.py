"""Host-side map loader: MapFormat1 YAML -> flat arrays (`MapData`) ready for `dts_upload_map`.

Mirrors what the reference computes once at load time and then only reads on the hot path:
  * tile grid (kind, angle, drivable)                  simulator.py:788-860  (`_interpret_map`)
  * per-tile lane-centre Bezier control points         simulator.py:1151-1335 (`_get_curve`)
  * static-object OBB corners / SAT axes / safety radii simulator.py:933-1038, objects.py:33-63,
                                                       collision.py:64-106, 214-220
Nothing here runs per step; the CUDA kernels consume the arrays this module produces.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import yaml

from . import assets


class InvalidMapException(Exception):
    """Same role as gym_duckietown.exceptions.InvalidMapException (exceptions.py:10)."""


# kind ids shared with the kernels (textures are indexed by the same id)
TILE_KINDS = ["straight", "curve_left", "curve_right", "3way_left", "3way_right", "4way",
              "asphalt", "grass", "floor"]
KIND_ID = {k: i for i, k in enumerate(TILE_KINDS)}
DRIVABLE = {"straight", "curve_left", "curve_right", "3way_left", "3way_right", "4way"}  # S:840-848
_ORIENT = ["S", "E", "N", "W"]  # S:823 ; default "E" S:824

# Lane-centre cubic Bezier templates in tile units, (x, z) pairs, y == 0 (values of S:1164-1299).
_L, _H, _Q = 0.20, 0.50, 0.25
_STRAIGHT_A = [(-_L, -_H), (-_L, -_Q), (-_L, _Q), (-_L, _H)]
_STRAIGHT_B = [(_L, _H), (_L, _Q), (_L, -_Q), (_L, -_H)]
_LEFT_WIDE = [(-_L, -_H), (-_L, 0.0), (0.0, _L), (_H, _L)]
_RIGHT_TIGHT = [(-_L, -_H), (-_L, -_L), (-0.30, -_L), (-_H, -_L)]
_TEMPLATES = {
    "straight": [_STRAIGHT_A, _STRAIGHT_B],
    "curve_left": [_LEFT_WIDE, [(_H, -_L), (0.30, -_L), (_L, -0.30), (_L, -_H)]],
    "curve_right": [_RIGHT_TIGHT, [(-_H, _L), (-0.30, _L), (0.30, 0.0), (_L, -_H)]],
    "3way": [
        _STRAIGHT_A,
        _LEFT_WIDE,
        _STRAIGHT_B,
        [(_H, -_L), (0.30, -_L), (_L, -_L), (_L, -_H)],
        [(_L, _H), (_L, _L), (0.30, _L), (_H, _L)],
        [(_H, -_L), (0.30, -_L), (-_L, 0.0), (-_L, _H)],
    ],
    "4way": [_LEFT_WIDE, _STRAIGHT_A, _RIGHT_TIGHT],
}


def _rot_y(angle: float) -> np.ndarray:
    """graphics.gen_rot_matrix((0,1,0), angle) in closed form (G:268-283): quaternion
    (a, 0, -sin(angle/2), 0) -> rows [a2-c2, 0, 2ac], [0, a2+c2, 0], [-2ac, 0, a2-c2]."""
    a = math.cos(angle / 2.0)
    c = -math.sin(angle / 2.0)
    return np.array([
        [a * a - c * c, 0.0, 2 * a * c],
        [0.0, a * a + c * c, 0.0],
        [-2 * a * c, 0.0, a * a - c * c],
    ])


def tile_curves(kind: str, angle: int, i: int, j: int, ts: float) -> np.ndarray:
    """Control points [C,4,3] of the tile's lane-centre curves in world frame (S:1151-1335)."""
    key = "3way" if kind.startswith("3way") else ("4way" if kind.startswith("4way") else kind)
    if kind.startswith("straight"):
        key = "straight"
    if key not in _TEMPLATES:
        raise InvalidMapException(f"Cannot get bezier for kind {kind!r}")
    t = np.array(_TEMPLATES[key], dtype=np.float64)
    pts = np.zeros(t.shape[:2] + (3,))
    pts[..., 0] = t[..., 0]
    pts[..., 2] = t[..., 1]
    pts = pts * ts
    shift = np.array([(i + 0.5) * ts, 0, (j + 0.5) * ts])
    if key == "4way":  # three templates x four rotations (S:1305-1316)
        out = []
        for rot in np.arange(0, 4):
            p = np.matmul(pts, _rot_y(rot * math.pi / 2))
            p += shift
            out.append(p)
        return np.reshape(np.array(out), (12, 4, 3))
    p = np.matmul(pts, _rot_y(angle * math.pi / 2))
    p += shift
    return p


def _rotate_point(px, py, cx, cy, theta):
    """graphics.rotate_point (G:254-265)."""
    dx, dy = px - cx, py - cy
    ndx = dx * math.cos(theta) + dy * math.sin(theta)
    ndy = dy * math.cos(theta) - dx * math.sin(theta)
    return cx + ndx, cy + ndy


def obb_corners(pos, min_c, max_c, theta, scale) -> np.ndarray:
    """collision.generate_corners (C:64-79): 4 footprint corners [4,2] rotated about (px,pz)."""
    px, pz = pos[0], pos[-1]
    x0, x1 = min_c[0] * scale + px, max_c[0] * scale + px
    z0, z1 = min_c[-1] * scale + pz, max_c[-1] * scale + pz
    return np.array([_rotate_point(x, z, px, pz, theta) for x, z in ((x0, z0), (x1, z0), (x1, z1), (x0, z1))])


def obb_axes(corners: np.ndarray) -> np.ndarray:
    """collision.generate_norm (C:99-106): eigenvectors of the corner covariance, as rows.

    Deliberately numpy/LAPACK on the host: for square footprints LAPACK's
    choice of axes is not the box axes, and the SAT flags depend on it.
    """
    ca = np.cov(corners, y=None, rowvar=False, bias=True)
    _, vect = np.linalg.eig(ca)
    return vect.T


@dataclass
class MapObject:
    kind: str
    mesh_id: int
    pos: np.ndarray  # [3] world
    angle: float  # rad
    scale: float
    optional: bool
    static: bool
    collidable: bool
    corners: np.ndarray  # [4,2]
    axes: np.ndarray  # [2,2]
    safety_radius: float
    max_coords: np.ndarray  # mesh extents (for _inconvenient_spawn S:1467)


@dataclass
class DynObject:
    """A `static: false` prop: DuckieObj pedestrian (objects.py:339-432) or DuckiebotObj lane follower
    (objects.py:180-336).  Parameters are the reference's non-randomized defaults; under domain_rand the
    reference draws them from the GLOBAL numpy RNG (unseeded), so callers may override."""
    kind: int                 # 1 duckie, 2 duckiebot, 3 traffic light (static, only its card texture changes)
    object_index: int         # entry of MapData.objects (mesh, scale, initial pose)
    pos: np.ndarray
    angle: float
    corners: np.ndarray       # [4,2] generate_corners of the mesh footprint (C:64-79)
    axes: np.ndarray          # [2,2] generate_norm at load; the duckiebot never refreshes it (O:60-63, O:306-336)
    safety_radius: float
    # DuckieObj (O:339-366)
    walk_distance: float = 0.0
    vel: float = 0.02
    wait_time: float = 8.0
    wiggle: float = math.pi / 15
    # DuckiebotObj (O:183-227)
    follow_dist: float = 0.3
    velocity: float = 0.1
    gain: float = 2.0
    trim: float = 0.0
    radius: float = 0.0318
    k: float = 27.0
    limit: float = 1.0
    wheel_dist: float = 0.102
    robot_width: float = 0.13 + 0.02
    robot_length: float = 0.18
    # TrafficLightObj (O:434-462): seconds between card flips and the starting pattern
    freq: float = 5.0
    pattern: int = 0


DYN_DUCKIE, DYN_DUCKIEBOT, DYN_TRAFFICLIGHT = 1, 2, 3


@dataclass
class MapData:
    name: str
    tile_size: float
    grid_w: int
    grid_h: int
    tile_kind: np.ndarray  # int8 [H*W], -1 = empty
    tile_angle: np.ndarray  # int8
    tile_drivable: np.ndarray  # uint8
    tile_curve_off: np.ndarray  # int32
    tile_curve_cnt: np.ndarray  # int32
    curves: np.ndarray  # f64 [NC,4,3]
    drivable_tiles: List[tuple]  # (i, j) in reference order (row-major scan, S:810-860)
    objects: List[MapObject] = field(default_factory=list)
    coll_corners: np.ndarray = None  # f64 [K,2,4]  (S:1035 stores corners.T)
    coll_norms: np.ndarray = None  # f64 [K,2,2]
    coll_centers: np.ndarray = None  # f64 [K,3]
    coll_radii: np.ndarray = None  # f64 [K]
    start_tile: Optional[tuple] = None
    start_pose: Optional[list] = None
    meshes: List["assets.Mesh"] = field(default_factory=list)
    dyn_objects: List[DynObject] = field(default_factory=list)

    @property
    def n_coll(self) -> int:
        return 0 if self.coll_radii is None else len(self.coll_radii)


def parse_tile(tile: str):
    """'kind/orient' -> (kind, angle) per S:818-838."""
    tile = tile.strip()
    if tile == "empty":
        return None
    if "/" in tile:
        kind, orient = tile.split("/")
        kind, orient = kind.strip(" "), orient.strip(" ")
        angle = _ORIENT.index(orient)
    elif "4" in tile:
        kind, angle = "4way", _ORIENT.index("E")
    else:
        kind, angle = tile, _ORIENT.index("E")
    return kind, angle


def interpret_map(map_data: dict, name: str = "map") -> MapData:
    """Build MapData from parsed MapFormat1 YAML.  Raises InvalidMapException like S:877-879."""
    try:
        if "tile_size" not in map_data:
            raise InvalidMapException("Must now include explicit tile_size in the map data.")
        ts = float(map_data["tile_size"])
        rows = map_data["tiles"]
        gh, gw = len(rows), len(rows[0])
        assert gh > 0 and gw > 0
        kind_arr = np.full(gh * gw, -1, np.int8)
        ang_arr = np.zeros(gh * gw, np.int8)
        drv_arr = np.zeros(gh * gw, np.uint8)
        coff = np.zeros(gh * gw, np.int32)
        ccnt = np.zeros(gh * gw, np.int32)
        curves: List[np.ndarray] = []
        drivable_tiles = []
        nc = 0
        for j, row in enumerate(rows):
            if len(row) != gw:
                raise InvalidMapException("each row of tiles must have the same length")
            for i, t in enumerate(row):
                parsed = parse_tile(t)
                if parsed is None:
                    continue
                kind, angle = parsed
                idx = j * gw + i
                if kind not in KIND_ID:
                    raise InvalidMapException(f"unknown tile kind {kind!r}")
                kind_arr[idx] = KIND_ID[kind]
                ang_arr[idx] = angle
                if kind in DRIVABLE:
                    drv_arr[idx] = 1
                    c = tile_curves(kind, angle, i, j, ts)
                    coff[idx], ccnt[idx] = nc, len(c)
                    nc += len(c)
                    curves.append(c)
                    drivable_tiles.append((i, j))
        md = MapData(
            name=name, tile_size=ts, grid_w=gw, grid_h=gh, tile_kind=kind_arr, tile_angle=ang_arr,
            tile_drivable=drv_arr, tile_curve_off=coff, tile_curve_cnt=ccnt,
            curves=np.concatenate(curves, 0) if curves else np.zeros((0, 4, 3)),
            drivable_tiles=drivable_tiles)
        _load_objects(md, map_data)
        if "start_tile" in map_data:
            md.start_tile = tuple(int(c) for c in map_data["start_tile"])
        if "start_pose" in map_data:
            md.start_pose = map_data["start_pose"]
        return md
    except InvalidMapException:
        raise
    except Exception as e:  # the reference wraps everything (S:877-879)
        raise InvalidMapException(f"Cannot load map data: {e}") from e


def _load_objects(md: MapData, map_data: dict) -> None:
    objs = map_data.get("objects") or []
    if isinstance(objs, dict):
        descs = list(objs.values())
    elif isinstance(objs, list):
        descs = objs
    else:
        raise ValueError(objs)
    mesh_ids: Dict[str, int] = {}
    corners_l, norms_l, centers_l, radii_l = [], [], [], []
    for desc in descs:
        kind = desc["kind"]
        if kind == "floor_tag":  # S:971-972
            continue
        mesh_key = "sign_generic:" + kind if kind.startswith("sign") else kind
        if mesh_key not in mesh_ids:
            mesh_ids[mesh_key] = len(md.meshes)
            md.meshes.append(assets.get_mesh(kind))
        mesh = md.meshes[mesh_ids[mesh_key]]
        # placement: README semantics, see DESIGN.md "object placement" (get_transform is absent)
        px, pz = float(desc["pos"][0]), float(desc["pos"][1])
        # through the cartesian frame and back, as interpret_object does (S:938-943, S:1651)
        cart_y = (md.grid_h - pz) * md.tile_size
        pos = np.array([px * md.tile_size, 0.0, md.grid_h * md.tile_size - cart_y])
        rot = math.radians(float(desc.get("rotate", 0.0)))
        angle = math.atan2(math.sin(rot), math.cos(rot))  # weird_from_cartesian re-wraps (S:1642)
        assert not ("height" in desc and "scale" in desc), "cannot specify both height and scale"
        if "height" in desc:
            scale = desc["height"] / mesh.max_coords[1]  # S:976-977
        else:
            scale = desc.get("scale", 1.0)
        static = desc.get("static", True)
        if not static and kind not in ("duckie", "duckiebot"):
            raise InvalidMapException(f"Object kind unknown: dynamic {kind!r}")  # S:1013-1015
        corners = obb_corners(pos, mesh.min_coords, mesh.max_coords, angle, scale)
        axes = obb_axes(corners)
        ext = np.max([abs(mesh.min_coords), abs(mesh.max_coords)], axis=0)  # C:218
        radius = 1.8 * (np.linalg.norm([ext[0], ext[2]]) * scale)  # S:150, O:53
        collidable = static and kind != "trafficlight"  # S:1027-1030
        md.objects.append(MapObject(
            kind=kind, mesh_id=mesh_ids[mesh_key], pos=pos, angle=angle, scale=float(scale),
            optional=bool(desc.get("optional", False)), static=bool(static), collidable=collidable,
            corners=corners, axes=axes, safety_radius=radius, max_coords=mesh.max_coords))
        if not static:
            md.dyn_objects.append(DynObject(
                kind=DYN_DUCKIE if kind == "duckie" else DYN_DUCKIEBOT, object_index=len(md.objects) - 1, pos=pos.copy(),
                angle=angle, corners=corners.copy(), axes=axes.copy(), safety_radius=float(radius),
                walk_distance=md.tile_size))   # DuckieObj(..., self.road_tile_size) S:1010
        elif kind == "trafficlight":   # TrafficLightObj S:999-1000: stepped like the others, never collides
            md.dyn_objects.append(DynObject(
                kind=DYN_TRAFFICLIGHT, object_index=len(md.objects) - 1, pos=pos.copy(), angle=angle,
                corners=corners.copy(), axes=axes.copy(), safety_radius=float(radius)))
        if collidable:
            corners_l.append(corners.T)
            norms_l.append(axes)
            centers_l.append(pos)
            radii_l.append(radius)
    k = len(radii_l)
    md.coll_corners = np.stack(corners_l, 0) if k else np.zeros((0, 2, 4))
    md.coll_norms = np.stack(norms_l, 0) if k else np.zeros((0, 2, 2))
    md.coll_centers = np.array(centers_l, dtype=np.float64).reshape(k, 3)
    md.coll_radii = np.array(radii_l, dtype=np.float64)


_MAP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")
_CACHE: Dict[str, MapData] = {}


def list_maps() -> List[str]:
    return sorted(f[:-5] for f in os.listdir(_MAP_DIR) if f.endswith(".yaml"))


def load_map(map_name: str) -> MapData:
    """Resolve like S:765-786: a path to a .yaml, or a bare map name looked up in the map dir."""
    if os.path.isfile(map_name):
        path, name = map_name, os.path.basename(map_name)[:-5]
    else:
        name = map_name
        path = os.path.join(_MAP_DIR, f"{map_name}.yaml")
    if not os.path.isfile(path):
        raise InvalidMapException(f"map file not found: {path}")
    key = os.path.abspath(path)
    if key not in _CACHE:
        with open(path, "r") as f:
            data = yaml.safe_load(f)
        _CACHE[key] = interpret_map(data, name)
    return _CACHE[key]
