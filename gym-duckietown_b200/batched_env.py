"""BatchedDuckietownEnv — N independent Duckietown agents per GPU behind the reference's reset/step API.

`step(actions)` is `Simulator.step()` (simulator.py:1669-1683) for every env at once: the action and
observation buffers are caller-visible torch CUDA tensors, the work runs on
`torch.cuda.current_stream()` inside libdtsim.so, nothing synchronises.  Constructor keywords are the
reference's (simulator.py:207-232, envs/duckietown_env.py:15) plus `num_envs`, `device`,
`auto_reset`, `device_reset`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import lib as L
from .episode import EpisodeSampler
from .maps import MapData, load_map


class BatchedDuckietownEnv:
    def __init__(self, num_envs: int, map_name: Union[str, Sequence[str]] = "udem1", *, device: int = 0,
                 max_steps: int = 1500, domain_rand: bool = True, frame_rate: float = 30, frame_skip: int = 1,
                 camera_width: int = 640, camera_height: int = 480, robot_speed: float = 1.2,
                 accept_start_angle_deg: float = 60, user_tile_start=None, seed: Optional[int] = None,
                 distortion: bool = False, dynamics_rand: bool = False, camera_rand: bool = False,
                 color_ground=(0.15, 0.15, 0.15), color_sky=(0.45, 0.82, 1), num_tris_distractors: int = 12,
                 gain=1.0, trim=0.0, radius=0.0318, k=27.0, limit=1.0,
                 action_mode: str = "vel_steer", auto_reset: bool = False, device_reset: bool = False,
                 cycle_maps: bool = False, env_id_offset: int = 0, tessellate_tiles: bool = False,
                 randomize_maps_on_reset: bool = False, randomization_config=None):
        if not torch.cuda.is_available():
            raise L.DtsError("BatchedDuckietownEnv needs a CUDA device; there is no CPU implementation")
        if camera_rand:
            raise NotImplementedError("camera_rand needs carnivalmirror (distortion.py:58-83); out of scope")
        names = [map_name] if isinstance(map_name, (str, MapData)) else list(map_name)
        self.maps: List[MapData] = [n if isinstance(n, MapData) else load_map(n) for n in names]   # parsed maps pass through
        self.num_envs, self.device_index = num_envs, device
        self.device = torch.device("cuda", device)
        self.camera_width, self.camera_height = camera_width, camera_height
        self.max_steps, self.domain_rand, self.distortion = max_steps, domain_rand, distortion
        self.frame_rate, self.delta_time, self.frame_skip = frame_rate, 1.0 / frame_rate, frame_skip
        self.robot_speed = robot_speed
        self.auto_reset, self.device_reset, self.cycle_maps = auto_reset, device_reset, cycle_maps
        self.randomize_maps_on_reset = randomize_maps_on_reset
        if cycle_maps and randomize_maps_on_reset:
            raise ValueError("cycle_maps (MultiMapEnv) and randomize_maps_on_reset are two different reset policies")
        if auto_reset and not device_reset:
            raise ValueError("auto_reset re-spawns on the device: pass device_reset=True")
        flags = (L.FLAG_AUTO_RESET if auto_reset else 0) | (L.FLAG_DOMAIN_RAND if domain_rand else 0) | \
                (L.FLAG_DISTORTION if distortion else 0) | (L.FLAG_DYNAMICS_RAND if dynamics_rand else 0) | \
                (L.FLAG_TESSELLATE if tessellate_tiles else 0)
        self.cfg = L.default_config(
            num_envs=num_envs, device=device, cam_width=camera_width, cam_height=camera_height, max_steps=max_steps,
            frame_skip=int(frame_skip), action_mode=L.ACTION_VEL_STEER if action_mode == "vel_steer" else L.ACTION_PWM,
            flags=flags, max_maps=len(self.maps), cycle_maps=len(self.maps) if cycle_maps else 0,
            random_maps=len(self.maps) if randomize_maps_on_reset else 0,
            frame_rate=float(frame_rate), robot_speed=robot_speed, accept_start_angle_deg=float(accept_start_angle_deg),
            gain=gain, trim=trim, radius=radius, k=k, limit=limit, seed=0 if seed is None else int(seed),
            env_id_offset=env_id_offset, num_tris_distractors=int(num_tris_distractors),
            color_sky=tuple(color_sky), color_ground=tuple(color_ground))
        if randomization_config is not None:   # Randomizer(randomization_config_fp=...) randomizer.py:19-33
            from .episode import load_dr_config
            L.set_dr_ops(self.cfg, L.dr_ops_from_config(load_dr_config(randomization_config)))
        self.sim = L.Sim(self.cfg)
        for i, md in enumerate(self.maps):
            self.sim.upload_map(i, md, tuple(user_tile_start) if user_tile_start else None)
        if distortion:
            from .distortion import Distortion
            self.camera_model = Distortion(camera_width, camera_height)
            self.sim.set_fisheye_lut(self.camera_model.rmapx, self.camera_model.rmapy)
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((num_envs, camera_height, camera_width, 3), dtype=torch.uint8, device=self.device)
            self.reward = torch.zeros(num_envs, dtype=torch.float32, device=self.device)
            self._done_u8 = torch.zeros(num_envs, dtype=torch.uint8, device=self.device)
            self.state: Dict[str, torch.Tensor] = {
                k_: torch.as_tensor(v, device=self.device) for k_, v in self.sim.state_arrays().items()}
        self.sampler = EpisodeSampler(
            num_envs, domain_rand=domain_rand, dynamics_rand=dynamics_rand, accept_start_angle_deg=accept_start_angle_deg,
            num_tris_distractors=num_tris_distractors, color_ground=color_ground, color_sky=color_sky,
            user_tile_start=user_tile_start, randomization_config=randomization_config)
        self.map_ids = np.zeros(num_envs, np.int32)
        self._first_reset = True
        self.resize = None
        self.output_format = dict(obs_layout="hwc", obs_dtype="uint8", reward="raw", discrete_actions=False,
                                  action_vel_scale=1.0)
        self.seed(seed)

    def set_output_format(self, obs_layout: Optional[str] = None, obs_dtype: Optional[str] = None,
                          reward: Optional[str] = None, discrete_actions: Optional[bool] = None,
                          action_vel_scale: Optional[float] = None):
        """Fuse the reference's wrapper stack into the step kernels (dts_set_output_format): obs_layout 'hwc' |
        'chw' (ImgWrapper) | 'cwh' (PyTorchObsWrapper), obs_dtype 'uint8' | 'float32' (NormalizeWrapper: /255),
        reward 'raw' | 'dt' (DtRewardWrapper), discrete_actions (DiscreteWrapper: ids in actions[:, 0]),
        action_vel_scale (ActionWrapper: 0.8).  Re-allocates `self.obs` in the new shape / dtype."""
        f = self.output_format
        for key, val in (("obs_layout", obs_layout), ("obs_dtype", obs_dtype), ("reward", reward),
                         ("discrete_actions", discrete_actions), ("action_vel_scale", action_vel_scale)):
            if val is not None:
                f[key] = val
        lay = {"hwc": L.OBS_HWC, "chw": L.OBS_CHW, "cwh": L.OBS_CWH}[f["obs_layout"]]
        dt = {"uint8": L.OBS_U8, "float32": L.OBS_F32_UNIT}[f["obs_dtype"]]
        rw = {"raw": L.REWARD_RAW, "dt": L.REWARD_DT}[f["reward"]]
        self.sim.set_output_format(lay, dt, rw, L.ACTIONS_DISCRETE3 if f["discrete_actions"] else L.ACTIONS_CONTINUOUS,
                                   f["action_vel_scale"])
        self._alloc_obs()
        return self

    @property
    def obs_size(self):
        """(height, width) of the observations the env emits: the camera's, or the fused ResizeWrapper's target."""
        return (self.resize[1], self.resize[0]) if self.resize else (self.camera_height, self.camera_width)

    def _alloc_obs(self):
        f = self.output_format
        H, W = self.obs_size
        shape = {"hwc": (H, W, 3), "chw": (3, H, W), "cwh": (3, W, H)}[f["obs_layout"]]
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((self.num_envs,) + shape, device=self.device,
                                   dtype=torch.uint8 if f["obs_dtype"] == "uint8" else torch.float32)

    def set_resize(self, resize_w: Optional[int], resize_h: Optional[int]):
        """ResizeWrapper (wrappers.py:111-141) on the device: render at the camera size, emit `resize_w` x `resize_h`
        observations (cv2.INTER_CUBIC's 8-bit fixed-point arithmetic) in the current layout / dtype.  None switches off."""
        self.resize = (int(resize_w), int(resize_h)) if resize_w else None
        self.sim.set_resize(*(self.resize or (0, 0)))
        self._alloc_obs()
        return self

    def sim_resize_only(self, frames: torch.Tensor) -> torch.Tensor:
        """The device ResizeWrapper pass on caller-supplied full-size frames u8[N, H, W, 3] -> self.obs."""
        if tuple(frames.shape) != (self.num_envs, self.camera_height, self.camera_width, 3) or frames.dtype != torch.uint8:
            raise ValueError("frames must be uint8 [num_envs, camera_height, camera_width, 3]")
        frames = frames.to(self.device).contiguous()
        self.sim.resize_frames(frames.data_ptr(), self.obs.data_ptr(), self._stream())
        return self.obs

    # ------------------------------------------------------------------ gym-like surface
    def seed(self, seed=None):
        """Env k gets seed+k (global index), like one reference env per seed (S:1043-1045)."""
        off = self.cfg.env_id_offset
        seeds = [None if seed is None else int(seed) + off + k for k in range(self.num_envs)]
        self.sampler.seed(seeds)
        if self.device_reset:   # the device continues the very same numpy streams (np_random.cuh)
            self.sim.seed_streams(self.sampler.rngs)
        return seeds

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask: Optional[torch.Tensor] = None, render: bool = True) -> torch.Tensor:
        """Simulator.reset() (S:528-763) for the masked envs (all if None); returns the obs batch."""
        mask_ptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mask_ptr = mask.data_ptr()
        if self.device_reset:
            self.sim.reset_random(mask_ptr, self._stream())
        else:
            envs = list(range(self.num_envs)) if mask is None else torch.nonzero(mask).flatten().tolist()
            if envs:
                if self.cycle_maps and not self._first_reset:  # MultiMapEnv.reset, envs/multimap_env.py:46
                    self.map_ids[envs] = (self.map_ids[envs] + 1) % len(self.maps)
                if self.randomize_maps_on_reset:   # np_random.choice(self.map_names) S:541-542: first draw of the reset
                    for e in envs:
                        self.map_ids[e] = int(self.sampler.rngs[e].integers(0, len(self.maps)))
                    # _load_map (S:544) comes before the spawn loop: re-create the drawn maps' obstacles first, so that
                    # the spawn predicates below see them at their load-time places.  Only the map id and the
                    # obstacles change here: pose / camera of the episode that just ended stay in place for the
                    # stale-modelview light capture of the reset proper (S:581).
                    self.sim.assign_maps(mask_ptr, self.map_ids.copy(), self._stream())
                dense = self.sampler.sample(envs, [self.maps[self.map_ids[e]] for e in envs], self._query_for(envs))
                params = {}
                for key, val in dense.items():
                    full = np.zeros((self.num_envs,) + val.shape[1:], val.dtype)
                    full[envs] = val
                    params[key] = full
                params["map_id"] = self.map_ids.copy()
                self.sim.reset(mask_ptr, params, self._stream())
        self._first_reset = False
        if render:
            self.sim.render(self.obs.data_ptr(), self._stream())
        return self.obs

    def _query_for(self, envs):
        def query(k, x, z, a, safety, hidden):
            return self.sim.query_poses(int(self.map_ids[envs[k]]), x, z, a, safety, hidden, dyn_env=int(envs[k]),
                                        stream=self._stream())
        return query

    def step(self, actions: torch.Tensor, render: bool = True, out=None):
        """actions f32[N,2] on this device: [vel, steering] (DuckietownEnv.step) or wheel duty
        (Simulator.step) depending on action_mode.  Returns (obs u8[N,H,W,3], reward f32[N], done bool[N], info).
        `out=(obs, reward, done_u8)` writes into caller-provided CUDA tensors instead of the env's own."""
        if actions.device != self.device or actions.dtype != torch.float32 or tuple(actions.shape) != (self.num_envs, 2):
            raise ValueError("actions must be a float32 CUDA tensor of shape [num_envs, 2] on the env's device")
        actions = actions.contiguous()
        obs, reward, done = (self.obs, self.reward, self._done_u8) if out is None else out
        self.sim.step(actions.data_ptr(), obs.data_ptr() if render else None, reward.data_ptr(), done.data_ptr(),
                      self._stream())
        return obs, reward, done.view(torch.bool), self.state

    def render_obs(self, segment: bool = False, top_down: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """render_obs(segment) (S:1953-1972) of the current state; `top_down` gives _render_img(top_down=True)'s camera."""
        tgt = self.obs if out is None else out
        if segment or top_down:
            self.sim.set_render_mode(segment, top_down)
            try:
                self.sim.render(tgt.data_ptr(), self._stream())
            finally:
                self.sim.set_render_mode(False, False)
        else:
            self.sim.render(tgt.data_ptr(), self._stream())
        return tgt

    # convenience views --------------------------------------------------------------------------
    @property
    def cur_pos(self) -> torch.Tensor:
        s = self.state
        return torch.stack([s["pos_x"], torch.zeros_like(s["pos_x"]), s["pos_z"]], dim=1)

    @property
    def cur_angle(self) -> torch.Tensor:
        return self.state["angle"]

    def launch_count(self) -> int:
        return self.sim.launch_count()

    def check(self):
        """Synchronise and raise if any render since creation ran out of its frame-memory capacity (prim slab or bin
        lists, sized from the scene's upper bounds at the first render) — such frames are left at the clear colour."""
        torch.cuda.synchronize(self.device)
        if int(self.sim.debug_counters()[0]) != 0 or (self.sim.status() & 1):
            raise L.DtsError("render frame memory overflowed: some frames were not drawn")

    def close(self):
        self.sim.close()


class HostPipeline:
    """Host-facing stepping with the copies taken off the critical path.

    `submit(actions_host)` enqueues: pinned-host -> device copy of the actions, `dts_step`, and a
    device -> pinned-host copy of obs / reward / done on a second stream; `result(ticket)` blocks until
    that step's host buffers are complete.  With `depth` slots in flight the PCIe transfer of step k
    overlaps the kernels of step k+1 (callers whose next action does not depend on the previous
    observation — replay, open-loop or random-action rollouts as in the reference's benchmark.py — get the
    full overlap; a closed-loop caller simply calls result() before the next submit())."""

    def __init__(self, env: BatchedDuckietownEnv, depth: int = 2):
        self.env, self.depth = env, depth
        dev, n = env.device, env.num_envs
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                act=torch.empty((n, 2), dtype=torch.float32, device=dev),
                obs=torch.empty_like(env.obs), rew=torch.empty_like(env.reward), done=torch.empty_like(env._done_u8),
                h_obs=torch.empty(tuple(env.obs.shape), dtype=env.obs.dtype).pin_memory(),
                h_rew=torch.empty(n, dtype=torch.float32).pin_memory(),
                h_done=torch.empty(n, dtype=torch.uint8).pin_memory(),
                computed=torch.cuda.Event(), copied=torch.cuda.Event(), busy=False))
        self.ticket = 0

    def submit(self, actions_host: torch.Tensor) -> int:
        """actions_host: float32 [N,2] CPU tensor (pinned for a truly asynchronous copy)."""
        s = self.slots[self.ticket % self.depth]
        cur = torch.cuda.current_stream(self.env.device)
        if s["busy"]:
            cur.wait_event(s["copied"])          # the slot's previous D2H must be done before we overwrite its buffers
        s["act"].copy_(actions_host, non_blocking=True)
        self.env.step(s["act"], out=(s["obs"], s["rew"], s["done"]))
        s["computed"].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["computed"])
            s["h_obs"].copy_(s["obs"], non_blocking=True)
            s["h_rew"].copy_(s["rew"], non_blocking=True)
            s["h_done"].copy_(s["done"], non_blocking=True)
            s["copied"].record(self.copy_stream)
        s["busy"] = True
        self.ticket += 1
        return self.ticket - 1

    def result(self, ticket: int):
        """(obs u8[N,H,W,3], reward f32[N], done bool[N]) pinned host tensors of that step; valid until the slot
        is reused `depth` submits later."""
        s = self.slots[ticket % self.depth]
        s["copied"].synchronize()
        return s["h_obs"], s["h_rew"], s["h_done"].view(torch.bool)
