// dts_common.cuh — device-side data layout shared by the dtsim kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/dtsim.h"

namespace dts {

// ------------------------------------------------------------------ map data resident in HBM
// One DMap per uploaded map; all pointers are device memory owned by the handle.  Read-only on
// the hot path and shared by every env, so it lives in L2 after the first touch.
struct DObject {
  float pos[3];
  float scale;
  float y_rot_deg;
  int32_t mesh_id;
  int32_t optional;
  float spawn_rad;   // max(max_coords)*0.5*scale + MIN_SPAWN_OBJ_DIST   (simulator.py:1467)
  int32_t tri_offset, tri_count;
  float bound_rad;   // object-space bounding radius (culling)
  float centre[3];   // object-space bounding-sphere centre
  int32_t dyn_slot;  // -1 static; else slot of the per-env dynamic state that supplies pos / y_rot / card
  int32_t alt_from, alt_to;   // traffic-light card swap (texture ids), -1 none
  int32_t seg_tex, pad_;      // flat class-colour texture of the mesh under segment=True (-1 none)
  double dpos[3];    // float64 position (x.pos in _inconvenient_spawn S:1466)
};

// A dynamic obstacle's constants (dts_dyn_object); its evolving state lives in DMap::dyn_state.
struct DDyn {
  int32_t kind, object_index;
  double pos_y;
  double norms[4];
  double safety_radius;
  double walk_distance, wiggle, angle0;   // duckie: heading = heading_vec(angle at load) never changes (O:357)
  double follow_dist, velocity, gain, trim, radius, k, limit, wheel_dist, robot_width, robot_length;
  double freq;        // traffic light
  int32_t tl_first;   // slot of the map's first traffic light (holds the shared card), -1 if none
  int32_t pad;
};

// Textures of a map live in ONE pool allocation; `info` is what a prim carries to find its texels without a dependent
// table lookup: (byte offset in the pool >> 8) | log2(w) << 24 | log2(h) << 28.
struct DTexture { const uint8_t* rgba; int32_t w, h; uint32_t info, pad; };

struct DMap {
  double tile_size;
  int32_t grid_w, grid_h;
  int32_t n_tiles;              // grid_w * grid_h
  const int8_t* tile_kind;
  const int8_t* tile_angle;
  const uint8_t* tile_drivable;
  const int16_t* tile_tex;
  const int32_t* tile_curve_off;
  const int32_t* tile_curve_cnt;
  const double* curves;         // [n][4][3]
  int32_t n_coll;
  const double* coll_corners;   // [K][2][4]
  const double* coll_norms;     // [K][2][2]
  const double* coll_centers;   // [K][3]
  const double* coll_radii;     // [K]
  int32_t start_i, start_j;     // user_tile_start / map `start_tile` (S:659-671) or -1
  int32_t has_start_pose;       // map `start_pose` S:679-686
  double start_pose[3];         // x, z offset inside the start tile, angle
  int32_t n_drivable;
  const int32_t* drivable_ij;   // [n_drivable][2] in reference scan order (S:806-860)
  int32_t n_objects;
  const DObject* objects;
  int32_t n_tris;
  const float* tri_pos;         // [T][3][3]
  const float* tri_nrm;
  const float* tri_uv;          // [T][3][2]
  const float* tri_col;         // [T][3][3]
  const int16_t* tri_tex;
  int32_t n_textures;
  const DTexture* textures;
  const uint8_t* tex_pool;      // all RGBA8 textures, each 256-byte aligned
  const int16_t* tex_segment;   // [n_textures] segment=True replacement of each texture (or the same index)
  DObject agent;                // top-down views: the agent's own mesh (tri_count 0 = none)
  int32_t n_dyn;
  const DDyn* dyn;              // [n_dyn]
  double* dyn_state;            // [DTS_DYN_FIELDS][n_dyn][num_envs]: mutable, per env, survives resets
  const double* dyn_init;       // [DTS_DYN_FIELDS][n_dyn] load-time values (a map reload re-creates the obstacles)
  int32_t valid;
};

// One env's view of its map's dynamic obstacles.
struct DynRef {
  const DDyn* par;
  double* st;
  int32_t n_dyn, n_envs, e;
  __device__ __forceinline__ double& f(int field, int slot) const {
    return st[((size_t)field * n_dyn + slot) * n_envs + e];
  }
};
__device__ __forceinline__ DynRef dyn_ref(const DMap& m, int n_envs, int e) {
  return DynRef{m.dyn, m.dyn_state, e >= 0 ? m.n_dyn : 0, n_envs, e};
}

// ------------------------------------------------------------------ per-env state, SoA in HBM
// One thread per env in the logic kernels: consecutive threads touch consecutive doubles.
struct DState {
  int32_t n;
  // dynamics (cartesian frame of duckietown_world: x right, y up; S:1629-1652)
  double *cx, *cy, *ctheta, *vu, *vw;
  double* fifo;          // [DTS_MAX_DELAY][2][n]  pending (left,right) duty, slot 0 = oldest
  // simulator-frame pose and per-step outputs
  double *pos_x, *pos_z, *angle, *speed, *reward, *lane_dist, *lane_dot, *lane_angle, *prox;
  double *wheel_dist, *trim;
  int32_t *step_count, *tile_i, *tile_j, *map_id, *episode;
  uint8_t *done_code, *in_lane, *collided;
  uint64_t* rng;         // [6][n] numpy PCG64 stream per env: state hi, lo, inc hi, lo, has_uint32, cached uint32
  struct RenderEp* rep;  // [n] per-episode render parameters (AoS: one CTA reads one record)
};

// Per-episode render inputs (simulator.py:546-614, 1768): 128 bytes, read by one CTA per frame.
struct __align__(16) RenderEp {
  float cam_height, cam_angle_deg, cam_fov_y_deg, pad0;
  float cam_noise[3]; float pad1;
  float horizon[3]; float pad2;
  float ambient[3]; float pad3;       // GL_LIGHT0 ambient
  float diffuse[3]; float pad4;       // GL_LIGHT0 diffuse
  float light_eye[4];                 // GL_POSITION as stored by GL: already in eye space
  float ground[3]; float pad5;
  uint32_t hidden[8];                 // bit o = object o invisible
};

struct DynParams { double u1, u2, u3, w1, w2, w3, uar, ual, war, wal; int32_t delay_steps; };

struct StepCfg {
  double dt, robot_speed, accept_angle_deg;
  double gain, trim, radius, k, limit;
  DynParams dyn;
  int32_t frame_skip, max_steps, action_mode, flags;
  int32_t reward_mode, action_map;      // dts_output_format
  int32_t random_maps, pad0;
  double action_vel_scale;
  uint64_t seed;
  int64_t env_id_offset;
  // Simulator.__init__ keywords read by reset() (S:226-230) and the Randomizer table (randomizer.py:19-89)
  int32_t num_tris_distractors, n_dr_ops;
  double color_sky[3], color_ground[3];
  dts_dr_op dr_ops[DTS_MAX_DR_OPS];
};

}  // namespace dts
