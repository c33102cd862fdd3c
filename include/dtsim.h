/* dtsim.h — C ABI of libdtsim.so, the B200-native batched Duckietown step() hot path.
 *
 * The reference (duckietown/gym-duckietown @5c2a586) has no FFI for this path: Simulator.step()
 * is Python all the way down to the OpenGL driver.  This header is the boundary we introduce
 * beneath its gym.Env surface.  Every entry point names the reference code it
 * replaces (paths relative to /root/reference/src/gym_duckietown).  Plain C, POD structs, raw
 * pointers and sizes; no torch types.  All device work is stream-ordered on the cudaStream_t the
 * caller passes (as a void*; NULL = legacy default stream); no entry point synchronises.
 *
 * Return value: 0 = ok, non-zero = error; dts_last_error() gives the message.
 */
#ifndef DTSIM_H
#define DTSIM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTS_ABI_VERSION 3
#define DTS_MAX_DELAY 16     /* command-delay line depth (steps): 0.15 s at up to 100 Hz (MotionBlurWrapper steps at 90 Hz) */
#define DTS_MAX_OBJECTS 256  /* per map; visibility bitmask is 8 x u32 */

typedef struct dts_sim dts_sim; /* opaque, one per GPU, not thread-safe */

/* done_code values (simulator.py:1685-1705 DoneRewardInfo.done_code) */
enum { DTS_IN_PROGRESS = 0, DTS_INVALID_POSE = 1, DTS_MAX_STEPS = 2 };
/* action_mode */
enum { DTS_ACTION_PWM = 0,      /* Simulator.step(action=[u_left,u_right])        simulator.py:1669 */
       DTS_ACTION_VEL_STEER = 1 /* DuckietownEnv.step(action=[vel, steering])  envs/duckietown_env.py:36-59 */ };
/* flags */
enum { DTS_FLAG_AUTO_RESET = 1,   /* done envs are re-spawned on device inside dts_step */
       DTS_FLAG_DOMAIN_RAND = 2,  /* simulator.py:213  (camera noise S:1768, DR sampling in device resets) */
       DTS_FLAG_DISTORTION = 4,   /* simulator.py:223  fisheye gather fused into the render (distortion.py:118) */
       DTS_FLAG_DYNAMICS_RAND = 8,/* simulator.py:224  per-env trim on the motor gains (S:746-748) */
       DTS_FLAG_TESSELLATE = 16   /* draw every road tile as the literal 7x7 quads of simulator.py:386-507
                                     instead of one quad + analytic lattice lighting (DESIGN.md render spec) */ };

/* One entry of the reference's domain-randomization table (randomization/randomizer.py:19-89): Randomizer.randomize
 * draws every key of the JSON config in SORTED key order from the env's np_random, whether or not domain_rand is on
 * (S:546).  `target` says which render / dynamics parameter the draw feeds; keys the simulator never reads are still
 * drawn (DTS_DR_NONE) so that the stream stays aligned with the reference. */
enum { DTS_DR_INT = 0,      /* rng.integers(low, high, size)   randomizer.py:55-58 */
       DTS_DR_UNIFORM = 1,  /* rng.uniform(low, high, size)    randomizer.py:69 */
       DTS_DR_NORMAL = 2    /* rng.normal(loc, scale, size)    randomizer.py:77 */ };
enum { DTS_DR_NONE = 0, DTS_DR_CAMERA_ANGLE, DTS_DR_CAMERA_FOV_Y, DTS_DR_CAMERA_HEIGHT, DTS_DR_CAMERA_NOISE,
       DTS_DR_HORZ_MODE, DTS_DR_LIGHT_POS, DTS_DR_TRIM };
#define DTS_MAX_DR_OPS 16
typedef struct {
  int32_t type, size, target, reserved;
  double a[3], b[3];        /* low / loc and high / scale; element k of a size-3 draw uses a[k], b[k]; size > 3 uses [0] */
} dts_dr_op;

typedef struct {
  int32_t abi_version;      /* DTS_ABI_VERSION */
  int32_t num_envs;         /* N independent agents on this GPU */
  int32_t device;           /* CUDA ordinal */
  int32_t cam_width;        /* simulator.py:216 (default 640) */
  int32_t cam_height;       /* simulator.py:217 (default 480) */
  int32_t max_steps;        /* simulator.py:210 (1500) */
  int32_t frame_skip;       /* simulator.py:215 (1) */
  int32_t action_mode;
  int32_t flags;
  int32_t max_maps;         /* slots for dts_upload_map */
  int32_t cycle_maps;       /* >0: every reset advances the env's map id modulo this count
                               (MultiMapEnv.reset round-robin, envs/multimap_env.py:44-49) */
  int32_t random_maps;      /* >0: every reset draws the env's map uniformly from the first `random_maps` slots and
                               re-creates its obstacles (randomize_maps_on_reset: np_random.choice + _load_map, S:541-544) */
  double frame_rate;        /* simulator.py:214 (30) */
  double robot_speed;       /* simulator.py:218 (1.2): the constant used in the reward, S:1702 */
  double accept_start_angle_deg; /* simulator.py:219 (60), device-side spawn only */
  /* DuckietownEnv constructor (envs/duckietown_env.py:15-33) */
  double gain, trim, radius, k, limit;
  /* duckietown_world DB18 PWM dynamics — source absent, restated (DESIGN.md "dynamics"):
   *   u' = u + dt(-u1 u - u2 w + u3 w^2 + uar*R + ual*L),  w' = w + dt(-w1 w - w2 u - w3 u w + war*R - wal*L)
   *   q' = q * exp(dt * [u', 0, w']),  commands delayed by `delay` seconds.            call sites S:746-755, S:2083-2086 */
  double dyn_u1, dyn_u2, dyn_u3, dyn_w1, dyn_w2, dyn_w3, dyn_uar, dyn_ual, dyn_war, dyn_wal;
  double dyn_delay;         /* 0.15 (S:748-750) */
  uint64_t seed;            /* device-side spawn/DR streams: stream k = hash(seed, global_env_id) */
  int64_t env_id_offset;    /* global index of local env 0 (multi-GPU shards keep seeds GPU-count independent) */
  /* Simulator.__init__ keywords that reset() reads (simulator.py:226-230): device resets honour them */
  int32_t num_tris_distractors; /* 12: 3*n vertices drawn per reset S:621-629 (never visible, draws consumed) */
  int32_t n_dr_ops;         /* 0 = the reference's default_dr.json; else the custom table below, in sorted key order */
  double color_sky[3];      /* BLUE_SKY S:108 (float64 like the Python floats _perturb multiplies) */
  double color_ground[3];   /* (0.15, 0.15, 0.15) S:228 */
  dts_dr_op dr_ops[DTS_MAX_DR_OPS];
} dts_config;

typedef struct { int32_t width, height; const uint8_t* rgba; /* [height][width][4], row 0 = t=0 */ } dts_texture;

typedef struct {            /* one placed prop: objects.py:33-66 (WorldObj), render O:123-148 */
  double pos[3];            /* float64 like WorldObj.pos (_inconvenient_spawn S:1461-1471); rendered as float32 */
  float scale;
  float y_rot_deg;
  int32_t mesh_id;
  int32_t optional;         /* hidden w.p. 1/2 at reset under domain_rand (S:653-654) */
  int32_t dyn_slot;         /* -1: static; else index into dts_map_blob.dyn — pose / card come from the per-env state */
  int32_t alt_tex_from;     /* traffic light: triangles textured `alt_tex_from` show `alt_tex_to` while the shared */
  int32_t alt_tex_to;       /*   card is on pattern 1 (mesh.textures[0] = texs[pattern] O:453,462); -1 = none */
  int32_t reserved;
} dts_object;

/* One `static: false` obstacle (S:973-1017): DuckieObj pedestrian O:339-432 or DuckiebotObj lane follower
 * O:180-336, with its state at map load.  Every env carries its own copy of the evolving state (per map), which
 * like the reference's object list survives resets.  Under domain_rand the reference draws vel / wait_time /
 * follow_dist ... from the GLOBAL numpy RNG at construction (O:348-350, O:197-205): the host supplies them. */
enum { DTS_DYN_DUCKIE = 1, DTS_DYN_DUCKIEBOT = 2,
       DTS_DYN_TRAFFICLIGHT = 3 /* TrafficLightObj O:434-462: static, never collides; flips its card every freq s */ };
#define DTS_MAX_DYN 32
typedef struct {
  int32_t kind;             /* DTS_DYN_* */
  int32_t object_index;     /* entry of objects[] carrying mesh / scale / height */
  double pos[3], angle;     /* WorldObj.pos, .angle O:54-57 */
  double corners[4][2];     /* obj_corners (x,z) O:60 */
  double norms[2][2];       /* obj_norm rows O:61 — the duckiebot never refreshes them (O:306-336) */
  double safety_radius;     /* O:66 */
  double walk_distance, vel, wait_time, wiggle;                       /* DuckieObj O:343-366 */
  double follow_dist, velocity, gain, trim, radius, k, limit, wheel_dist, robot_width, robot_length; /* O:196-227 */
  double freq;              /* TrafficLightObj O:445-450 (5; randint(4,7) under domain_rand) */
  int32_t pattern;          /* starting pattern (0; randint(0,2) under domain_rand) */
  int32_t reserved;
} dts_dyn_object;

/* Per-env state of one dynamic obstacle: DTS_DYN_FIELDS doubles, device layout [field][slot][env]. */
enum { DTS_DYN_PX = 0, DTS_DYN_PZ, DTS_DYN_ANGLE, DTS_DYN_YROT, DTS_DYN_CORNERS /* 8: x0 z0 .. x3 z3 */,
       DTS_DYN_START_X = 12, DTS_DYN_START_Z, DTS_DYN_WAIT, DTS_DYN_VEL, DTS_DYN_TIME, DTS_DYN_ACTIVE, DTS_DYN_FIELDS,
       /* traffic lights reuse two fields: their own pattern, and — in the FIRST light's slot — the card the mesh
        * they all share currently shows (every flip assigns it; the last writer wins, like the reference) */
       DTS_DYN_PATTERN = DTS_DYN_ACTIVE, DTS_DYN_SHOWN = DTS_DYN_WAIT };

typedef struct {
  int32_t tri_offset, tri_count;
  int32_t seg_flat_tex;     /* texture every triangle of this mesh shows under segment=True: the flat class colour
                               gen_segmentation_color(mesh_name) (objmesh.py:260-290); -1 = none */
  int32_t reserved;
} dts_mesh;

/* What the reference's wrapper stacks do to actions, observations and rewards, fused into the step kernels
 * (SURVEY 8f-3).  W = src/gym_duckietown/wrappers.py, LW = learning/utils/wrappers.py. */
enum { DTS_OBS_HWC = 0,     /* render_obs S:1953-1972: [H][W][3] */
       DTS_OBS_CHW = 1,     /* ImgWrapper LW:73-87 transpose(2,0,1): [3][H][W] */
       DTS_OBS_CWH = 2      /* PyTorchObsWrapper W:93-110 transpose(2,1,0): [3][W][H] */ };
enum { DTS_OBS_U8 = 0,
       DTS_OBS_F32_UNIT = 1 /* NormalizeWrapper LW:56-70: (obs - 0) / (255 - 0) as float32 */ };
enum { DTS_REWARD_RAW = 0,
       DTS_REWARD_DT = 1    /* DtRewardWrapper LW:90-102: -1000 -> -10, r > 0 -> r + 10, else r + 4 */ };
enum { DTS_ACTIONS_CONTINUOUS = 0,
       DTS_ACTIONS_DISCRETE3 = 1 /* DiscreteWrapper W:8-33: 0 left [0.6,+1], 1 right [0.6,-1], 2 forward [0.7,0];
                                    actions[e][0] carries the id, actions[e][1] is ignored */ };
typedef struct {
  int32_t obs_layout, obs_dtype, reward_mode, action_map;
  double action_vel_scale;  /* ActionWrapper LW:106-112: action[0] * 0.8 before the env sees it (1.0 = off) */
} dts_output_format;

/* Everything the hot path reads about one map, prepared on the host (maps.py):
 * tile grid S:788-860, curves S:1151-1335, static collidables S:1019-1038 / S:919-931,
 * meshes objmesh.py:65-293, textures graphics.py:70-169. All pointers are HOST memory, copied. */
typedef struct {
  double tile_size;
  int32_t grid_w, grid_h;
  const int8_t* tile_kind;        /* [grid_h*grid_w], -1 = no tile */
  const int8_t* tile_angle;       /* 0..3 (S,E,N,W) */
  const uint8_t* tile_drivable;
  const int16_t* tile_tex;        /* texture index per tile, -1 = none */
  const int32_t* tile_curve_off;  /* first curve of the tile */
  const int32_t* tile_curve_cnt;  /* 0, 2, 6 or 12 */
  int32_t n_curves;
  const double* curves;           /* [n_curves][4][3] */
  int32_t n_coll;
  const double* coll_corners;     /* [n_coll][2][4]  x row, z row */
  const double* coll_norms;       /* [n_coll][2][2]  two SAT axes as rows */
  const double* coll_centers;     /* [n_coll][3] */
  const double* coll_radii;       /* [n_coll] safety radii */
  int32_t n_objects;
  const dts_object* objects;
  int32_t n_meshes;
  const dts_mesh* meshes;
  int32_t n_tris;
  const float* tri_pos;           /* [n_tris][3][3] */
  const float* tri_nrm;           /* [n_tris][3][3] */
  const float* tri_uv;            /* [n_tris][3][2] */
  const float* tri_col;           /* [n_tris][3][3] */
  const int16_t* tri_tex;         /* [n_tris] texture index, -1 = untextured */
  int32_t n_textures;
  const dts_texture* textures;
  int32_t start_tile[2];          /* `user_tile_start` (S:659-662) else map `start_tile` (S:867-871) else {-1,-1}: device resets spawn there */
  int32_t n_dyn;                  /* <= DTS_MAX_DYN */
  int32_t has_start_pose;         /* map `start_pose` (S:874-876): device resets then place the agent at start_pose */
  const dts_dyn_object* dyn;      /* [n_dyn] in the order of the map's object list (update order S:1570-1584) */
  double start_pose[3];           /* x offset, z offset inside the start tile, angle (S:679-686) */
  const int16_t* tex_segment;     /* [n_textures] what Texture.bind(segment=True) / get_mesh(name, True) show instead of texture t
                                     (graphics.py:52-56, 70-130; objmesh.py:268-290); NULL or -1 = unchanged */
  int32_t agent_mesh;             /* mesh drawn at the agent's pose in top-down views (self.mesh, S:864, S:1923-1929); -1 = none */
  int32_t reserved2;
} dts_map_blob;

/* Per-episode inputs produced by Simulator.reset() (simulator.py:528-763), one
 * entry per env of the batch (SoA, HOST pointers, length num_envs; entries of unmasked envs are
 * ignored).  Any pointer may be NULL = keep the reference's non-randomized default. */
typedef struct {
  const int32_t* map_id;
  const double* pos_x;  const double* pos_z;  const double* angle;   /* cur_pos / cur_angle S:740-741 */
  const double* wheel_dist;      /* S:597 */
  const double* trim;            /* S:747 (dynamics_rand) */
  const float* cam_height;       /* S:602,612 */
  const float* cam_angle_deg;    /* S:605,613 */
  const float* cam_fov_y_deg;    /* S:608,614 */
  const float* cam_noise;        /* [N][3] S:1768-1769 */
  const float* horizon_color;    /* [N][3] S:551-562 */
  const float* light_ambient;    /* [N][3] S:573-574 */
  const float* light_diffuse;    /* [N][3] S:575-576 */
  const float* light_pos;        /* [N][4] S:565-570, S:581 */
  const int32_t* light_stale;    /* [N] 0: GL_POSITION captured under identity modelview (first reset);
                                        1: under the previous frame's camera matrix (S:581) */
  const float* ground_color;     /* [N][3] S:594 */
  const uint32_t* obj_hidden;    /* [N][8] bit o set = object o invisible this episode (S:653-656) */
} dts_episode_params;

/* Device pointers to the library-owned SoA state (all length num_envs unless noted). */
typedef struct {
  double* pos_x; double* pos_z; double* angle;     /* cur_pos[0], cur_pos[2], cur_angle */
  double* speed;                                   /* S:1568 */
  double* reward;                                  /* f64 shadow of the f32 reward output */
  double* lane_dist; double* lane_dot; double* lane_angle_rad;   /* LanePosition S:1409 (NaN if not in lane) */
  double* prox_penalty;                            /* S:1430-1459 */
  double* wheel_dist;
  int32_t* step_count;
  int32_t* tile_i; int32_t* tile_j;                /* get_grid_coords(cur_pos) S:1134 */
  int32_t* map_id;
  int32_t* episode;                                /* resets so far */
  uint8_t* done_code;
  uint8_t* in_lane;
  uint8_t* collided;                               /* _collision() of the current pose S:1473 */
} dts_state_view;

/* Simulator.__init__ (simulator.py:207-384) minus map loading. */
int dts_create(const dts_config* cfg, dts_sim** out);
/* Simulator._load_map/_interpret_map/_load_objects (simulator.py:765-931) : device copy of one map. */
int dts_upload_map(dts_sim* sim, int map_id, const dts_map_blob* blob);
/* Distortion.rmapx/rmapy (distortion.py:85-125): LUT of the fused fisheye gather, [H][W] each (HOST pointers).  The
 * rasteriser renders every output pixel AT the source position rint(rmap) names — obs[y,x] = undistorted[rint(rmapy),
 * rint(rmapx)], 0 when that falls outside (distortion.py:118, cv2.remap INTER_NEAREST / BORDER_CONSTANT) — so no
 * undistorted frame is materialised and no second pass runs.  Fails if the LUT scatters one 32x8 output bin over a
 * source region too wide for the rasteriser's int32 edge functions. */
int dts_set_fisheye_lut(dts_sim* sim, const float* rmapx, const float* rmapy, int width, int height);
/* Simulator.reset() (simulator.py:528-763) with host-drawn episode parameters.
 * mask_dev: device u8[num_envs] (NULL = all). */
int dts_reset(dts_sim* sim, const uint8_t* mask_dev, const dts_episode_params* params, void* stream);
/* Simulator.seed() (simulator.py:1043-1045): one numpy PCG64 stream per env for the DEVICE-side resets.
 * streams[N][6] (HOST) = state_hi, state_lo, inc_hi, inc_lo, has_uint32, uinteger of
 * numpy.random.Generator(PCG64(SeedSequence(seed))).bit_generator.state; mask_host u8[N] or NULL = all. */
int dts_seed_streams(dts_sim* sim, const uint8_t* mask_host, const uint64_t* streams);
/* Simulator.reset() with pose / DR drawn ON THE DEVICE from the env's numpy-compatible stream, draw for draw
 * in the reference's order (np_random.cuh): same seeds -> same episodes as the reference. */
int dts_reset_random(dts_sim* sim, const uint8_t* mask_dev, void* stream);
/* Simulator.step() (simulator.py:1669-1683) for all envs: actions f32[N][2] -> obs u8[N][H][W][3]
 * (NULL = skip rendering), reward f32[N], done u8[N]. All DEVICE pointers. */
int dts_step(dts_sim* sim, const float* actions_dev, void* obs_dev, float* reward_dev, uint8_t* done_dev,
             void* stream);
/* Simulator.render_obs() (simulator.py:1953-1972) of the current state. */
int dts_render(dts_sim* sim, void* obs_dev, void* stream);
/* Render variants of _render_img (S:1707-1951) for subsequent dts_render / dts_step calls (0 = the agent camera):
 *   DTS_RENDER_SEGMENT   segment=True: lighting off (S:1730-1733), magenta clear + ground (S:1752, 1808), no distractors
 *                        (S:1814), segmentation textures / flat per-class mesh colours
 *   DTS_RENDER_TOP_DOWN  top_down=True: camera above the map centre looking down (S:1786-1798), the agent's own mesh drawn
 *                        at cur_pos (S:1923-1929) */
enum { DTS_RENDER_SEGMENT = 1, DTS_RENDER_TOP_DOWN = 2 };
int dts_set_render_mode(dts_sim* sim, int mode);
/* Select the fused wrapper behaviour for subsequent dts_step / dts_render calls (default: all zero, scale 1).
 * obs_dev then holds num_envs * 3 * H * W elements of uint8 or float32 in the chosen layout. */
int dts_set_output_format(dts_sim* sim, const dts_output_format* fmt);
/* ResizeWrapper (wrappers.py:111-141: cv2.resize(obs, (resize_w, resize_h), interpolation=cv2.INTER_CUBIC)) on the device:
 * subsequent dts_step / dts_render calls render at the camera size into a library buffer and write obs_dev as
 * num_envs x 3 x out_h x out_w elements in the selected layout / dtype (OpenCV's 8-bit fixed-point bicubic; matches cv2
 * within 1 LSB).  out_w = out_h = 0 switches it off. */
int dts_set_resize(dts_sim* sim, int out_w, int out_h);
/* MotionBlurWrapper (learning/utils/wrappers.py:8-54): out f64[n] = np.average of four u8 frame batches with `weights`
 * (numpy's evaluation order), and the knobs that wrapper turns on the wrapped env: delta_time / frame_skip (it divides
 * env.delta_time by 3 and drives update_physics itself, LW:13-14) and the action convention (wheel commands). */
int dts_blend4(dts_sim* sim, const uint8_t* const frames_dev[4], const double weights[4], double* out_dev, uint64_t n, void* stream);
int dts_set_timing(dts_sim* sim, double delta_time, int frame_skip, int action_mode);
/* The resize pass alone, on caller-supplied frames: src u8[num_envs][cam_h][cam_w][3] -> dst in the selected layout / dtype. */
int dts_resize_frames(dts_sim* sim, const uint8_t* src_dev, void* dst_dev, void* stream);
int dts_get_state(dts_sim* sim, dts_state_view* out);
/* Batched pose predicates for host callers — the de-facto public helpers of Simulator:
 * _valid_pose (S:1494), _collision(get_agent_corners()) (S:1473, run_tests.py:50), get_lane_pos2 (S:1371),
 * proximity_penalty2 (S:1430), _inconvenient_spawn (S:1461), get_grid_coords (S:1134), _drivable_pos (S:1411).
 * HOST pointers, synchronous. query[n][4] = x, z, angle, safety_factor; hidden[n][8] object-visibility
 * bitmasks or NULL; out_f64[n][4] = lane dist, dot_dir, angle_rad (NaN when not in a lane), proximity;
 * out_i32[n][8] = valid, collision (offset once), collision (as _valid_pose sees it), in_lane,
 * inconvenient_spawn, tile_i, tile_j, drivable.  Dynamic obstacles are taken where env `dyn_env` currently has
 * them (check_collision O:265/368, proximity O:271/374, x.pos in S:1466); dyn_env < 0 ignores them. */
int dts_query_poses(dts_sim* sim, int map_id, int dyn_env, int n, const double* query, const uint32_t* hidden,
                    double* out_f64, int32_t* out_i32, void* stream);
/* randomize_maps_on_reset, host-drawn (S:541-544): give the masked envs the map ids in map_id_host[N] and re-create
 * those maps' obstacles for them (_load_map) WITHOUT touching pose, episode counter or render record — the reset that
 * follows still sees the previous episode's last camera (stale GL_LIGHT0 position, S:581). */
int dts_assign_maps(dts_sim* sim, const uint8_t* mask_dev, const int32_t* map_id_host, void* stream);
/* Device pointer to the dynamic-obstacle state of map `map_id`: f64[DTS_DYN_FIELDS][n_dyn][num_envs] (NULL, 0 for a
 * map without dynamic obstacles).  Re-uploading the map puts every env's obstacles back to their load-time state. */
int dts_get_dyn_state(dts_sim* sim, int map_id, double** state_dev, int32_t* n_dyn);
/* End-of-rollout observation all-gather across the GPU shards of one box; the step path
 * itself has no collective.  libnccl is dlopen'ed from `libnccl_path` (the torch-bundled copy); rank 0
 * creates a unique id, the caller broadcasts its 128 bytes (torch.distributed), every rank inits.
 * send: u8[bytes_per_rank] on this GPU, recv: u8[world*bytes_per_rank]. */
int dts_comm_load(dts_sim* sim, const char* libnccl_path);
int dts_comm_unique_id(dts_sim* sim, uint8_t out[128]);
int dts_comm_init(dts_sim* sim, const uint8_t id[128], int rank, int world);
int dts_allgather_obs(dts_sim* sim, const void* send_dev, void* recv_dev, uint64_t bytes_per_rank, void* stream);
/* Sticky status, readable WITHOUT synchronising (a mapped host word the kernels write): bit 0 = some frame since
 * creation overflowed its render frame memory (prim slab / bin lists) and was left incomplete.  dts_step and
 * dts_render return non-zero once it is set (the frame that overflowed may be one or two calls back). */
int dts_status(dts_sim* sim);
/* Per-kernel device timing of the render launches (bench.py's roofline): when enabled, every dts_render brackets its
 * kernels with CUDA events on the caller's stream.  dts_profile_read synchronises, returns the summed milliseconds of
 * [0] k_frame_setup, [1] k_geometry, [2] k_bin, [3] k_raster, [4] post passes (resize) and the number of frames
 * they cover, and clears the accumulators. */
int dts_profile_enable(dts_sim* sim, int level);   /* 0 off; 1 events around k_raster only; 2 around every render kernel */
int dts_profile_read(dts_sim* sim, double ms_out[8], int64_t* frames);
/* FUSED end-of-rollout gather (the path's one exchange step) — instead of running an all-gather after the last
 * step, the last step's rasteriser stores every frame straight into the gather buffers of all GPUs of the box:
 *   dts_gather_alloc   this rank's buffer u8[world][bytes_per_rank] (returned in *buf_dev) and its 64-byte cudaIpc handle;
 *   dts_gather_open    the other ranks' handles (the caller exchanges them, e.g. torch.distributed all_gather), mapped
 *                      as peer memory (NVLink / NVSwitch);
 *   dts_gather_next    the NEXT dts_step / dts_render also writes its observations, in the selected layout / dtype,
 *                      to slot `rank` of every rank's buffer while it rasterises (one extra store per peer and word).
 * A rank's buffer is complete once every rank's step has finished: the caller orders that (stream sync + barrier), exactly
 * as it orders the use of an all-gather's output.  dts_allgather_obs remains as the NCCL baseline of the same exchange. */
int dts_gather_alloc(dts_sim* sim, uint64_t bytes_per_rank, int rank, int world, uint8_t handle_out[64], void** buf_dev);
int dts_gather_open(dts_sim* sim, const uint8_t* handles /* [world][64] */);
int dts_gather_next(dts_sim* sim);
/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
uint64_t dts_launch_count(dts_sim* sim);
/* debug: the env's per-episode render record as 36 x 32-bit words (cam_height, cam_angle_deg, cam_fov_y_deg, -,
 * cam_noise[3], -, horizon[3], -, ambient[3], -, diffuse[3], -, light_eye[4], ground[3], -, hidden u32[8]). */
int dts_debug_episode(dts_sim* sim, int env, void* out144);
/* debug (tests/test_gpu_gltrace.py): what k_frame_setup / k_geometry produced for `env` in the last dts_render — camera
 * model-view V (row-major 3x4, float64), projection P00 P11 P22 P23, counts = {prims, lattices, overflow, (prim, bin) pairs of the whole batch}, and the lit
 * 8x8 lattice [n_cells][64][3] of every emitted road tile by grid cell i * grid_h + j (NaN where culled).  Synchronises. */
int dts_debug_frame(dts_sim* sim, int env, double V[12], float P[4], int32_t counts[4], float* lattice_by_cell, int n_cells);
/* 32 diagnostic counters: [0] != 0 -> a render scratch buffer overflowed (frame incomplete). */
int dts_debug_counters(dts_sim* sim, int32_t out[32]);
const char* dts_last_error(dts_sim* sim); /* sim may be NULL: error of the last failed dts_create */
void dts_destroy(dts_sim* sim);

#ifdef __cplusplus
}
#endif
#endif /* DTSIM_H */
