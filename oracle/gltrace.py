"""Recording OpenGL mock: run the REFERENCE's own render code and capture what it asks GL to do.

TEST INFRASTRUCTURE ONLY (needs /root/reference; used by oracle/make_golden.py to write
tests/golden/gltrace_*.npz).  Nothing in the product imports this file.

The reference's pixels come out of an OpenGL driver that cannot run here, but everything UP TO the driver is
plain Python: `Simulator._render_img` (simulator.py:1707-1951), `_init_vlists` (:386-527), the reset-time
lighting (:564-586), `WorldObj.render` (objects.py:123-148) and `Texture.bind` (graphics.py:52-56) are a
sequence of fixed-function GL calls.  `GLRecorder` stands in for `pyglet.gl` / `pyglet.graphics` under the
stub-import harness (oracle/refstub.py), keeps the state a GL context would keep (matrix stacks, current
colour, bound texture, enables, GL_LIGHT0) with the argument types of the real entry points (GLfloat
arguments are rounded to float32, GLdouble ones are not) and logs one record per draw call.  The record is
the INPUT of rasterisation as the reference produces it: model-view and projection matrices, light
parameters (GL_POSITION already multiplied by the model-view current at glLightfv time — the "stale
model-view" quirk of simulator.py:581), current colour, vertex arrays, texture, draw order.

Matrix composition follows the OpenGL 2.1 specification formulas (glRotate, glTranslate, glScale,
gluLookAt, gluPerspective) in float64; a driver would do the same products in float32.
"""
from __future__ import annotations

import ctypes
import math
from typing import List

import numpy as np


def _rot(angle_deg, x, y, z):
    n = math.sqrt(x * x + y * y + z * z)
    if n == 0:
        return np.eye(4)
    x, y, z = x / n, y / n, z / n
    c, s = math.cos(math.radians(angle_deg)), math.sin(math.radians(angle_deg))
    return np.array([
        [x * x * (1 - c) + c, x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0.0],
        [y * x * (1 - c) + z * s, y * y * (1 - c) + c, y * z * (1 - c) - x * s, 0.0],
        [x * z * (1 - c) - y * s, y * z * (1 - c) + x * s, z * z * (1 - c) + c, 0.0],
        [0.0, 0.0, 0.0, 1.0]])


def _f64(v):
    """ctypes c_double conversion: size-1 numpy arrays (DR multipliers, randomizer.py sizes) convert like scalars."""
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


def _f32(v):
    return float(np.float32(_f64(v)))


class GLenum(str):
    """A GL enum that is its own name; bit-ors of enums (glClear masks) concatenate."""

    def __or__(self, other):
        return GLenum(f"{self}|{other}")

    __ror__ = __or__


class VList:
    """pyglet.graphics.vertex_list stand-in: keeps the attribute arrays, logs draws."""

    def __init__(self, rec, vid, count, attrs):
        self.rec, self.vid, self.count = rec, vid, count
        self.attrs = {}
        for fmt, data in attrs:
            # pyglet formats: v3f t2f n3f c3f c4B ...: f -> float32 array, B -> ubyte
            dt = np.float32 if fmt[2] == "f" else np.uint8
            self.attrs[fmt[0]] = np.asarray(list(data), dtype=np.float64).astype(dt).reshape(count, int(fmt[1]))

    def draw(self, mode):
        self.rec._draw("vlist", self.vid, mode)

    def delete(self):
        pass


class GLRecorder:
    def __init__(self):
        self.mode = "GL_MODELVIEW"
        self.stacks = {"GL_MODELVIEW": [np.eye(4)], "GL_PROJECTION": [np.eye(4)], "GL_TEXTURE": [np.eye(4)]}
        self.color = np.array([1.0, 1.0, 1.0, 1.0])
        self.enabled = set()
        self.texture = None
        self.light = {}            # (light, pname) -> values (GL_POSITION: eye space)
        self.light_model_ambient = np.array([0.2, 0.2, 0.2, 1.0])
        self.clear_color = np.zeros(4)
        self.perspective = None
        self.events: List[dict] = []
        self.vlists: List[VList] = []
        self.tex_names = {}        # texture id -> name
        self.label = None          # set by the harness: which draw group comes next (mesh kind ...)

    # ---------------------------------------------------------------- module surface
    GLfloat = ctypes.c_float
    GLdouble = ctypes.c_double
    GLubyte = ctypes.c_ubyte
    GLuint = ctypes.c_uint
    GLint = ctypes.c_int

    def __getattr__(self, name):
        if name.startswith("GL_") or name.startswith("GLU_"):
            return GLenum(name)               # enums are their own names
        if name.startswith("gl") or name.startswith("glu"):
            return lambda *a, **k: None       # everything not modelled is a no-op
        raise AttributeError(name)

    @property
    def top(self):
        return self.stacks[self.mode][-1]

    def _mul(self, m):
        self.stacks[self.mode][-1] = self.top @ m

    def glMatrixMode(self, mode):
        self.mode = mode

    def glLoadIdentity(self):
        self.stacks[self.mode][-1] = np.eye(4)

    def glPushMatrix(self):
        self.stacks[self.mode].append(self.top.copy())

    def glPopMatrix(self):
        self.stacks[self.mode].pop()

    def glTranslatef(self, x, y, z):
        t = np.eye(4)
        t[:3, 3] = [_f32(x), _f32(y), _f32(z)]
        self._mul(t)

    def glScalef(self, x, y, z):
        self._mul(np.diag([_f32(x), _f32(y), _f32(z), 1.0]))

    def glRotatef(self, a, x, y, z):
        self._mul(_rot(_f32(a), _f32(x), _f32(y), _f32(z)))

    def gluPerspective(self, fovy, aspect, zn, zf):
        fovy, aspect, zn, zf = _f64(fovy), _f64(aspect), _f64(zn), _f64(zf)
        f = 1.0 / math.tan(math.radians(fovy) / 2.0)
        self.perspective = (fovy, aspect, zn, zf)
        self._mul(np.array([[f / aspect, 0, 0, 0], [0, f, 0, 0], [0, 0, (zf + zn) / (zn - zf), 2 * zf * zn / (zn - zf)],
                            [0, 0, -1, 0]], dtype=float))

    def gluLookAt(self, ex, ey, ez, cx, cy, cz, ux, uy, uz):
        e = np.array([_f64(ex), _f64(ey), _f64(ez)])
        c = np.array([_f64(cx), _f64(cy), _f64(cz)])
        up = np.array([_f64(ux), _f64(uy), _f64(uz)])
        f = c - e
        f = f / np.linalg.norm(f)
        s = np.cross(f, up)
        s = s / np.linalg.norm(s)
        u = np.cross(s, f)
        m = np.eye(4)
        m[0, :3], m[1, :3], m[2, :3] = s, u, -f
        t = np.eye(4)
        t[:3, 3] = -e
        self._mul(m @ t)
        self.events.append(dict(kind="lookat", eye=e, center=c, up=up, modelview=self.stacks["GL_MODELVIEW"][-1].copy()))

    def glColor3f(self, r, g, b):
        self.color = np.array([_f32(r), _f32(g), _f32(b), 1.0])

    def glColor4f(self, r, g, b, a):
        self.color = np.array([_f32(r), _f32(g), _f32(b), _f32(a)])

    def glEnable(self, cap):
        self.enabled.add(cap)

    def glDisable(self, cap):
        self.enabled.discard(cap)

    def glBindTexture(self, target, tid):
        self.texture = tid

    def glClearColor(self, r, g, b, a):
        self.clear_color = np.array([_f32(r), _f32(g), _f32(b), _f32(a)])

    def glClear(self, mask):
        self.events.append(dict(kind="clear", color=self.clear_color.copy()))

    def glLightModelfv(self, pname, params):
        if pname == "GL_LIGHT_MODEL_AMBIENT":
            self.light_model_ambient = np.array(list(params), float)

    def glLightfv(self, light, pname, params):
        v = np.array(list(params), float)      # a (GLfloat * n) array: already float32 values
        rec = dict(kind="light", light=light, pname=pname, values=v.copy(), modelview=self.stacks["GL_MODELVIEW"][-1].copy())
        if pname == "GL_POSITION":
            v = self.stacks["GL_MODELVIEW"][-1] @ v    # stored in eye coordinates (GL 2.1 section 2.14.1)
            rec["eye"] = v.copy()
        self.light[(light, pname)] = v
        self.events.append(rec)

    # ---------------------------------------------------------------- draws
    def vertex_list(self, count, *attrs):
        v = VList(self, len(self.vlists), count, attrs)
        self.vlists.append(v)
        return v

    def _draw(self, what, ident, mode, extra=None):
        ev = dict(kind="draw", what=what, id=ident, mode=mode, label=self.label,
                  modelview=self.stacks["GL_MODELVIEW"][-1].copy(), projection=self.stacks["GL_PROJECTION"][-1].copy(),
                  color=self.color.copy(), texture=self.texture if "GL_TEXTURE_2D" in self.enabled else None,
                  lighting="GL_LIGHTING" in self.enabled, light0="GL_LIGHT0" in self.enabled,
                  color_material="GL_COLOR_MATERIAL" in self.enabled, normalize="GL_NORMALIZE" in self.enabled,
                  rescale_normal="GL_RESCALE_NORMAL" in self.enabled,
                  light_pos_eye=self.light.get(("GL_LIGHT0", "GL_POSITION"), np.array([0, 0, 1, 0.0])).copy(),
                  light_ambient=self.light.get(("GL_LIGHT0", "GL_AMBIENT"), np.array([0, 0, 0, 1.0])).copy(),
                  light_diffuse=self.light.get(("GL_LIGHT0", "GL_DIFFUSE"), np.array([1, 1, 1, 1.0])).copy(),
                  light_model_ambient=self.light_model_ambient.copy())
        if extra:
            ev.update(extra)
        self.events.append(ev)

    def mesh_draw(self, kind, segment):
        """Hook for the FakeMesh stand-in of ObjMesh.render (objmesh.py:360-375)."""
        self._draw("mesh", kind, "GL_TRIANGLES", dict(segment=bool(segment)))

    def take(self) -> List[dict]:
        ev, self.events = self.events, []
        return ev


class _FakeTex:
    def __init__(self, tid):
        self.target, self.id = "GL_TEXTURE_2D", tid


def attach(rec: GLRecorder):
    """Route the stubbed pyglet.gl / pyglet.graphics of oracle/refstub.py into `rec`."""
    import sys

    import refstub
    refstub.install()
    gl = sys.modules["pyglet.gl"]
    for name in ("glMatrixMode", "glLoadIdentity", "glPushMatrix", "glPopMatrix", "glTranslatef", "glScalef",
                 "glRotatef", "gluPerspective", "gluLookAt", "glColor3f", "glColor4f", "glEnable", "glDisable",
                 "glBindTexture", "glClearColor", "glClear", "glLightModelfv", "glLightfv"):
        setattr(gl, name, getattr(rec, name))
    for name in ("GLfloat", "GLdouble", "GLubyte", "GLuint", "GLint"):
        setattr(gl, name, getattr(GLRecorder, name))

    class _Enums:
        def __init__(self, mod):
            self.mod = mod

    # enums: any GL_* attribute resolves to its own name (set lazily through the module's __getattr__)
    orig_getattr = type(gl).__getattr__

    def patched(self, name):
        if self is gl and (name.startswith("GL_") or name.startswith("GLU_")):
            setattr(self, name, GLenum(name))
            return GLenum(name)
        return orig_getattr(self, name)

    type(gl).__getattr__ = patched
    sys.modules["pyglet.graphics"].vertex_list = rec.vertex_list
    sys.modules["pyglet"].graphics.vertex_list = rec.vertex_list
    # meshes: the stand-in ObjMesh logs its draw with the state current at that point
    refstub.FakeMesh.render = lambda self, segment=False: rec.mesh_draw(getattr(self, "kind", "?"), segment)
    S, C, G, O = refstub.modules()
    counter = {"n": 0}

    def fake_load_texture(path, segment=False, segment_into_color=None):
        counter["n"] += 1
        t = _FakeTex(counter["n"])
        rec.tex_names[t.id] = (str(path), bool(segment))
        return t

    G.Texture.tex_cache.clear()   # class-level cache of Texture objects: ids of an earlier recorder must not survive
    rec._fake_load_texture = fake_load_texture
    S.load_texture = fake_load_texture
    G.load_texture = fake_load_texture
    sys.modules["duckietown_world"].get_texture_file = lambda name: [name]
    S.get_texture_file = lambda name: [name]
    G.get_texture_file = lambda name: [name]
    return S, C, G, O


def make_sim(rec: GLRecorder, raw_map: dict, mesh_extents: dict, *, domain_rand: bool, seed: int, width=160, height=120):
    """A reference Simulator (no GL context) whose render path runs for real against the recorder."""
    from unittest import mock

    import refstub
    # a new Simulator = a fresh GL context: identity matrices, default state
    rec.mode = "GL_MODELVIEW"
    rec.stacks = {"GL_MODELVIEW": [np.eye(4)], "GL_PROJECTION": [np.eye(4)], "GL_TEXTURE": [np.eye(4)]}
    rec.enabled, rec.light, rec.texture = set(), {}, None
    sim = refstub.build_reference_sim(raw_map, mesh_extents, domain_rand=domain_rand, seed=seed)
    S, C, G, O = refstub.modules()
    S.load_texture = rec._fake_load_texture    # build_reference_sim installs a plain no-op loader: ours carries ids
    if "render_obs" in sim.__dict__:
        del sim.__dict__["render_obs"]         # the harness's no-op: we want the class's real render_obs
    sim.camera_width, sim.camera_height = width, height
    sim.shadow_window = mock.MagicMock()
    sim.multi_fbo, sim.final_fbo = 1, 2
    sim.multi_fbo_human, sim.final_fbo_human = 3, 4
    sim.img_array = np.zeros((height, width, 3), np.uint8)
    sim.img_array_human = np.zeros((600, 800, 3), np.uint8)
    sim.mesh = refstub.FakeMesh([0, 0, 0], [1, 1, 1])
    sim.mesh.kind = "agent_duckiebot"
    sim.cam_offset = np.array([0, 0, 0])
    sim._init_vlists()                         # road_vlist / ground_vlist through the recording vertex_list
    return sim
