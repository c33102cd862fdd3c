/* dt_oracle_dynamic.c — CPU ORACLE (test infrastructure): dynamic obstacles of gym-duckietown, restated from
 * objects.py — DuckieObj pedestrians (O:339-432) and DuckiebotObj pure-pursuit followers (O:180-336) — with the
 * per-step update of simulator.py:1570-1584 and their share of _collision (S:1487-1489) and
 * proximity_penalty2 (S:1455-1457).  Pinned against the reference's own classes stepped through
 * oracle/refstub.py (tests/golden/dynamic_*.npz, tests/test_oracle_dynamic.py). */
#include <math.h>
#include <string.h>

#include "dt_oracle.h"

static const double PI = 3.14159265358979323846;

/* DuckieObj.step O:396-422 (+ finish_walk O:424-432, non-randomized branch) */
static void duckie_step(orc_dyn* o, double dt) {
  o->time += dt;
  if (!o->active) {
    o->wait_time -= dt;
    if (o->wait_time <= 0) o->active = 1;
    return;
  }
  double vx = o->heading[0] * o->vel, vy = o->heading[1] * o->vel, vz = o->heading[2] * o->vel;
  o->pos[0] += vx; o->pos[1] += vy; o->pos[2] += vz;   /* self.center (aliases self.pos afterwards) */
  for (int k = 0; k < 4; k++) { o->corners[k][0] += vx; o->corners[k][1] += vz; }
  double dx = o->pos[0] - o->start[0], dy = o->pos[1] - o->start[1], dz = o->pos[2] - o->start[2];
  double distance = sqrt(dx * dx + dy * dy + dz * dz);
  if (distance > o->walk_distance) {
    memcpy(o->start, o->pos, sizeof o->start);
    o->angle += PI;
    o->active = 0;
    o->vel *= -1;
    o->wait_time = 8;
  }
  double angle_delta = o->wiggle * sin(48 * o->time);
  o->y_rot = (o->angle + angle_delta) * (180 / PI);
}

/* DuckiebotObj._update_pos O:281-336 */
static void duckiebot_update_pos(orc_dyn* o, double vel, double steer, double dt) {
  double k_r_inv = (o->gain + o->trim) / o->k, k_l_inv = (o->gain - o->trim) / o->k;
  double omega_r = (vel + 0.5 * steer * o->wheel_dist) / o->radius;
  double omega_l = (vel - 0.5 * steer * o->wheel_dist) / o->radius;
  double u_r = omega_r * k_r_inv, u_l = omega_l * k_l_inv;
  double ur = fmax(fmin(u_r, o->limit), -o->limit), ul = fmax(fmin(u_l, o->limit), -o->limit);
  if (ul == ur) {  /* straight: no rotation, and the bounding box is NOT refreshed (O:305-307) */
    o->pos[0] = o->pos[0] + dt * ul * cos(o->angle);
    o->pos[1] = o->pos[1] + dt * ul * 0;
    o->pos[2] = o->pos[2] + dt * ul * -sin(o->angle);
    return;
  }
  double w = (ur - ul) / o->wheel_dist;
  double r = (o->wheel_dist * (ul + ur)) / (2 * (ul - ur));
  double rot = w * dt;
  double rvx = sin(o->angle), rvz = cos(o->angle);
  double px = o->pos[0], pz = o->pos[2];
  double cx = px + r * rvx, cz = pz + r * rvz;
  /* graphics.rotate_point G:254-265 */
  double ddx = px - cx, ddy = pz - cz;
  double ndx = ddx * cos(rot) + ddy * sin(rot), ndy = ddy * cos(rot) - ddx * sin(rot);
  o->pos[0] = cx + ndx; o->pos[2] = cz + ndy;
  o->angle += rot;
  o->y_rot += rot * 180 / PI;
  /* agent_boundbox C:9-34 with the duckiebot's own width / length */
  double fx = cos(o->angle), fz = -sin(o->angle), sx = sin(o->angle), sz = cos(o->angle);
  double hw = 0.5 * o->robot_width, hl = 0.5 * o->robot_length;
  o->corners[0][0] = o->pos[0] - hw * sx - hl * fx; o->corners[0][1] = o->pos[2] - hw * sz - hl * fz;
  o->corners[1][0] = o->pos[0] + hw * sx - hl * fx; o->corners[1][1] = o->pos[2] + hw * sz - hl * fz;
  o->corners[2][0] = o->pos[0] + hw * sx + hl * fx; o->corners[2][1] = o->pos[2] + hw * sz + hl * fz;
  o->corners[3][0] = o->pos[0] - hw * sx + hl * fx; o->corners[3][1] = o->pos[2] - hw * sz + hl * fz;
}

/* DuckiebotObj.step_duckiebot O:229-263 */
static void duckiebot_step(const orc_map* m, orc_dyn* o, double dt) {
  double cp[3], ct[3], cur[3], dummy[3];
  if (!orc_closest_curve_point(m, o->pos[0], o->pos[2], o->angle, cp, ct)) return; /* the reference raises here */
  double lookup = o->follow_dist;
  int found = 0;
  for (int it = 0; it < 1000 && !found; it++) {
    double fx = cp[0] + ct[0] * lookup, fz = cp[2] + ct[2] * lookup;
    found = orc_closest_curve_point(m, fx, fz, o->angle, cur, dummy);
    if (!found) lookup *= 0.5;
  }
  if (!found) return;
  double vx = cur[0] - o->pos[0], vy = cur[1] - o->pos[1], vz = cur[2] - o->pos[2];
  double n = sqrt(vx * vx + vy * vy + vz * vz);
  vx /= n; vy /= n; vz /= n;
  double dot = sin(o->angle) * vx + 0 * vy + cos(o->angle) * vz;   /* get_right_vec S:2066 */
  double steering = o->gain * -dot;
  duckiebot_update_pos(o, o->velocity, steering, dt);
}

/* TrafficLightObj.step O:455-462.  Python's round(time, 3) is restated as rint(time*1000)/1000: the same value
 * unless time*1000 sits within an ulp of a .5 tie, which multiples of 1/frame_rate never do. */
static void trafficlight_step(orc_dyn* objs, orc_dyn* o, double dt) {
  o->time += dt;
  double r = rint(o->time * 1000.0) / 1000.0;
  if (fmod(r, o->freq) == 0) {
    o->active ^= 1;
    objs[o->tl_first].shown = o->active;   /* self.mesh.textures[0] = ... on the mesh all lights share */
  }
}

/* the update loop of update_physics S:1570-1584 for the dynamic objects of one env */
void orc_dyn_step_all(const orc_map* m, orc_dyn* objs, int n, double dt) {
  for (int i = 0; i < n; i++) {
    if (objs[i].kind == 2) duckiebot_step(m, &objs[i], dt);
    else if (objs[i].kind == 3) trafficlight_step(objs, &objs[i], dt);
    else duckie_step(&objs[i], dt);
  }
}

static void mm4(const double ax[2], const double* cx, const double* cz, double* lo, double* hi) {
  double mn = INFINITY, mx = -INFINITY;
  for (int k = 0; k < 4; k++) { double v = ax[0] * cx[k] + ax[1] * cz[k]; if (v < mn) mn = v; if (v > mx) mx = v; }
  *lo = mn; *hi = mx;
}
static int ov(double a0, double a1, double b0, double b1) { return (a0 <= b0 && b0 <= a1) || (b0 <= a0 && a0 <= b1); }

/* check_collision O:265-269 / O:368-372 -> intersects_single_obj C:162-186, for the agent box centred like
 * orc_collision (centre = get_agent_corners(px,pz,angle)) */
int orc_dyn_collision(const orc_dyn* objs, int n, double px, double pz, double angle) {
  double cx[4], cz[4];
  orc_agent_corners(px, pz, angle, cx, cz);
  const double an[2][2] = {{sin(angle), cos(angle)}, {cos(angle), -sin(angle)}};
  for (int i = 0; i < n; i++) {
    if (objs[i].kind == 3) continue;   /* static WorldObj.check_collision -> False (O:150-158) */
    double ox[4], oz[4];
    for (int k = 0; k < 4; k++) { ox[k] = objs[i].corners[k][0]; oz[k] = objs[i].corners[k][1]; }
    int hit = 1;
    double a0, a1, b0, b1;
    for (int a = 0; a < 2 && hit; a++) { mm4(an[a], cx, cz, &a0, &a1); mm4(an[a], ox, oz, &b0, &b1); if (!ov(a0, a1, b0, b1)) hit = 0; }
    for (int a = 0; a < 2 && hit; a++) {
      const double* on = objs[i].norm + 2 * a;
      mm4(on, cx, cz, &a0, &a1); mm4(on, ox, oz, &b0, &b1);
      if (!ov(a0, a1, b0, b1)) hit = 0;
    }
    if (hit) return 1;
  }
  return 0;
}

/* obj.proximity O:271-279 / O:374-383 summed over the dynamic objects (S:1455-1457); (px,pz) = cur_pos */
double orc_dyn_proximity(const orc_dyn* objs, int n, double px, double pz, double angle) {
  double off = 0.066 - (0.18 / 2);
  double ax = px + off * cos(angle), az = pz + off * -sin(angle);
  double r1 = (fmax(0.18, 0.13 + 0.02) / 2) * 1.8, tot = 0.0;
  for (int i = 0; i < n; i++) {
    if (objs[i].kind == 3) continue;   /* static WorldObj.proximity -> 0 (O:160-168) */
    double dx = ax - objs[i].pos[0], dy = 0 - objs[i].pos[1], dz = az - objs[i].pos[2];
    double score = sqrt(dx * dx + dy * dy + dz * dz) - r1 - objs[i].safety_radius;
    tot += score < 0 ? score : 0;
  }
  return tot;
}

/* _valid_pose S:1494-1534 with the dynamic half of _collision: the agent box is built from the already
 * shifted centre (like orc_valid_pose). */
int orc_valid_pose_dyn(const orc_map* m, const orc_dyn* objs, int n, double px, double pz, double angle,
                       double safety_factor, uint8_t* collided) {
  uint8_t col = 0, drv = 0;
  orc_valid_pose(m, px, pz, angle, safety_factor, &col, &drv);
  if (!col) {
    double off = 0.066 - (0.18 / 2);
    col = (uint8_t)orc_dyn_collision(objs, n, px + off * cos(angle), pz + off * -sin(angle), angle);
  }
  if (collided) *collided = col;
  return !col && drv;
}

/* Simulator.step S:1669-1683 on a map with dynamic obstacles: per physics update the agent moves, then every
 * object steps (S:1551-1584); done/reward (S:1685-1705, S:1654-1667) see the moved objects. */
void orc_step_dynamic(const orc_map* m, const orc_dyn_params* dp, orc_dyn_state* s, int* step_count, double* last_px,
                      double* last_pz, const double action[2], int action_mode, double wheel_dist,
                      const double env5[5], int frame_skip, double dt, int max_steps, double robot_speed,
                      orc_dyn* objs, int n, orc_step_out* o) {
  double cmd[2] = {action[0], action[1]};
  if (action_mode == 1) orc_action_map(action[0], action[1], wheel_dist, env5[0], env5[1], env5[2], env5[3], env5[4], cmd);
  cmd[0] = fmax(-1.0, fmin(1.0, cmd[0]));
  cmd[1] = fmax(-1.0, fmin(1.0, cmd[1]));
  double px = *last_px, pz = *last_pz, ang = 0, speed = 0;
  for (int f = 0; f < frame_skip; f++) {
    double ppx = px, ppz = pz;
    orc_dyn_step(s, dp, cmd, dt);
    orc_weird_from_cartesian(m, s, &px, &pz, &ang);
    (*step_count)++;
    speed = sqrt((px - ppx) * (px - ppx) + (pz - ppz) * (pz - ppz)) / dt;
    orc_dyn_step_all(m, objs, n, dt);
  }
  *last_px = px; *last_pz = pz;
  orc_done_reward(m, px, pz, ang, *step_count, max_steps, robot_speed, o);   /* static part + lane pose */
  o->speed = speed;
  o->prox += orc_dyn_proximity(objs, n, px, pz, ang);
  int valid = orc_valid_pose_dyn(m, objs, n, px, pz, ang, 1.0, &o->collided);
  if (!valid) { o->done = 1; o->done_code = 1; o->reward = -1000.0; }
  else if (*step_count >= max_steps) { o->done = 1; o->done_code = 2; o->reward = 0.0; }
  else {
    o->done = 0; o->done_code = 0;
    if (o->in_lane) o->reward = +1.0 * robot_speed * o->lane_dot + -10 * fabs(o->lane_dist) + +40 * o->prox;
    else o->reward = 40 * o->prox;
  }
}
