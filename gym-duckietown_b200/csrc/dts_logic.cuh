// dts_logic.cuh — device functions for the non-rendering half of Simulator.step():
// action map, delayed PWM dynamics, tile lookup, valid-pose, OBB SAT, safety circles, lane pose,
// reward/done, and the device-side spawn.  float64 like the reference (numpy defaults).
// Citations: S = simulator.py, C = collision.py, G = graphics.py, E = envs/duckietown_env.py.
#pragma once
#include "dts_common.cuh"
#include "np_random.cuh"

namespace dts {

// robot constants S:118-177
constexpr double kRobotWidth = 0.13 + 0.02;
constexpr double kRobotLength = 0.18;
constexpr double kCamForward = 0.066;
constexpr double kCentreOff = kCamForward - (kRobotLength / 2);          // _actual_center S:2109
constexpr double kAgentSafetyRad = ((kRobotLength > kRobotWidth ? kRobotLength : kRobotWidth) / 2) * 1.8;  // S:153
constexpr double kRewardInvalidPose = -1000.0;                           // S:175
constexpr int kMaxSpawnAttempts = 5000;                                  // S:177

struct LanePose { double dist, dot_dir, angle_rad; bool in_lane; };

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmax(lo, fmin(hi, v)); }

// E:36-59 — [vel, steering] -> (u_left, u_right), clamped to +-limit
__device__ __forceinline__ void action_to_pwm(double vel, double steer, double baseline, const StepCfg& c,
                                              double& u_l, double& u_r) {
  const double k_r_inv = (c.gain + c.trim) / c.k;
  const double k_l_inv = (c.gain - c.trim) / c.k;
  const double omega_r = (vel + 0.5 * steer * baseline) / c.radius;
  const double omega_l = (vel - 0.5 * steer * baseline) / c.radius;
  u_r = fmax(fmin(omega_r * k_r_inv, c.limit), -c.limit);
  u_l = fmax(fmin(omega_l * k_l_inv, c.limit), -c.limit);
}

// get_grid_coords S:1134-1149 + _get_tile S:1053-1063. Returns flat tile index or -1.
__device__ __forceinline__ int tile_at(const DMap& m, double x, double z, int& i, int& j) {
  i = (int)floor(x / m.tile_size);
  j = (int)floor(z / m.tile_size);
  if (i < 0 || i >= m.grid_w || j < 0 || j >= m.grid_h) return -1;
  const int idx = j * m.grid_w + i;
  return m.tile_kind[idx] < 0 ? -1 : idx;
}
// _drivable_pos S:1411-1428
__device__ __forceinline__ bool drivable_at(const DMap& m, double x, double z) {
  int i, j;
  const int idx = tile_at(m, x, z, i, j);
  return idx >= 0 && m.tile_drivable[idx] != 0;
}

__device__ __forceinline__ void project4(double ax, double az, const double* xs, const double* zs, double& lo,
                                         double& hi) {
  double v = ax * xs[0] + az * zs[0];
  lo = v; hi = v;
#pragma unroll
  for (int k = 1; k < 4; k++) {
    v = ax * xs[k] + az * zs[k];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
}
// C:50-61 closed-interval overlap
__device__ __forceinline__ bool intervals_touch(double a0, double a1, double b0, double b1) {
  return (a0 <= b0 && b0 <= a1) || (b0 <= a0 && a0 <= b1);
}

// _collision S:1473-1492 / intersects C:129-159 for the agent box centred at (bx,bz).
// Agent SAT axes = right / forward unit vectors (eigenvectors of the 0.15x0.18 corner covariance,
// C:99-106, up to sign and order, to which the interval test is indifferent).
__device__ inline bool agent_hits_static(const DMap& m, double bx, double bz, double angle) {
  if (m.n_coll == 0) return false;
  double sn, cs;
  sincos(angle, &sn, &cs);
  const double fx = cs, fz = -sn, rx = sn, rz = cs;  // get_dir_vec / get_right_vec S:2056-2073
  const double hw = 0.5 * kRobotWidth, hl = 0.5 * kRobotLength;
  double ax[4], az[4];  // agent_boundbox C:9-34 corner order
  ax[0] = bx - hw * rx - hl * fx; az[0] = bz - hw * rz - hl * fz;
  ax[1] = bx + hw * rx - hl * fx; az[1] = bz + hw * rz - hl * fz;
  ax[2] = bx + hw * rx + hl * fx; az[2] = bz + hw * rz + hl * fz;
  ax[3] = bx - hw * rx + hl * fx; az[3] = bz - hw * rz + hl * fz;
  double aR0, aR1, aF0, aF1;
  project4(rx, rz, ax, az, aR0, aR1);
  project4(fx, fz, ax, az, aF0, aF1);
  for (int k = 0; k < m.n_coll; k++) {
    const double* ox = m.coll_corners + k * 8;
    const double* oz = ox + 4;
    const double* on = m.coll_norms + k * 4;
    double lo, hi, lo2, hi2;
    project4(rx, rz, ox, oz, lo, hi);
    if (!intervals_touch(aR0, aR1, lo, hi)) continue;
    project4(fx, fz, ox, oz, lo, hi);
    if (!intervals_touch(aF0, aF1, lo, hi)) continue;
    project4(on[0], on[1], ax, az, lo, hi);
    project4(on[0], on[1], ox, oz, lo2, hi2);
    if (!intervals_touch(lo, hi, lo2, hi2)) continue;
    project4(on[2], on[3], ax, az, lo, hi);
    project4(on[2], on[3], ox, oz, lo2, hi2);
    if (!intervals_touch(lo, hi, lo2, hi2)) continue;
    return true;
  }
  return false;
}

// _valid_pose S:1494-1534.  NB the collision box is built from the already-shifted centre and
// get_agent_corners shifts it again (S:1502 + S:1521 -> S:2114): offset applied twice.
__device__ inline bool valid_pose(const DMap& m, double px, double pz, double angle, double safety, bool* collided) {
  double sn, cs;
  sincos(angle, &sn, &cs);
  const double fx = cs, fz = -sn, rx = sn, rz = cs;
  const double qx = px + kCentreOff * fx, qz = pz + kCentreOff * fz;
  const double sw = safety * 0.5 * kRobotWidth, sl = safety * 0.5 * kRobotLength;
  const bool all_drivable = drivable_at(m, qx, qz) && drivable_at(m, qx - sw * rx, qz - sw * rz) &&
                            drivable_at(m, qx + sw * rx, qz + sw * rz) && drivable_at(m, qx + sl * fx, qz + sl * fz);
  const bool hit = agent_hits_static(m, qx + kCentreOff * fx, qz + kCentreOff * fz, angle);
  if (collided) *collided = hit;
  return !hit && all_drivable;
}

// proximity_penalty2 S:1430-1459 with C:189-211
__device__ inline double proximity_penalty(const DMap& m, double px, double pz, double angle) {
  if (m.n_coll == 0) return 0.0;
  double sn, cs;
  sincos(angle, &sn, &cs);
  const double qx = px + kCentreOff * cs, qz = pz + kCentreOff * -sn;
  bool touching = false;
  double acc = 0.0;
  for (int k = 0; k < m.n_coll; k++) {
    const double* c = m.coll_centers + 3 * k;
    const double dx = c[0] - qx, dy = c[1], dz = c[2] - qz;
    const double d = sqrt(dx * dx + dy * dy + dz * dz);
    const double r2 = m.coll_radii[k];
    const double dif = kAgentSafetyRad - r2, sum = kAgentSafetyRad + r2;
    if ((dif * dif <= d * d && d * d <= sum * sum) || d < fabs(dif)) touching = true;
    const double s = d - kAgentSafetyRad - r2;
    if (s < 0) acc += s;
  }
  return touching ? acc : 0.0;
}

__device__ __forceinline__ void bezier_at(const double* cp, double t, double& x, double& y, double& z) {  // G:286-297
  const double s = 1 - t;
  const double b0 = s * s * s, b1 = 3 * t * (s * s), b2 = 3 * (t * t) * s, b3 = t * t * t;
  x = b0 * cp[0]; x += b1 * cp[3]; x += b2 * cp[6]; x += b3 * cp[9];
  y = b0 * cp[1]; y += b1 * cp[4]; y += b2 * cp[7]; y += b3 * cp[10];
  z = b0 * cp[2]; z += b1 * cp[5]; z += b2 * cp[8]; z += b3 * cp[11];
}

// closest_curve_point S:1337-1369 + get_lane_pos2 S:1371-1409
__device__ inline LanePose lane_pose(const DMap& m, double px, double pz, double angle) {
  LanePose r;
  r.dist = r.dot_dir = r.angle_rad = __longlong_as_double(0x7ff8000000000000LL);
  r.in_lane = false;
  int ti, tj;
  const int idx = tile_at(m, px, pz, ti, tj);
  if (idx < 0 || !m.tile_drivable[idx]) return r;
  const double* cv = m.curves + (size_t)m.tile_curve_off[idx] * 12;
  const int nc = m.tile_curve_cnt[idx];
  double sn, cs;
  sincos(angle, &sn, &cs);
  const double dirx = cs, dirz = -sn;
  // argmax_c (P3-P0).dir — S:1355-1362 divide every chord by ONE Frobenius norm, kept for the rounding
  double fro = 0.0;
  for (int c = 0; c < nc; c++) {
    const double hx = cv[c * 12 + 9] - cv[c * 12], hy = cv[c * 12 + 10] - cv[c * 12 + 1], hz = cv[c * 12 + 11] - cv[c * 12 + 2];
    fro += hx * hx; fro += hy * hy; fro += hz * hz;
  }
  fro = sqrt(fro);
  int best = 0;
  double bestv = -1e300;
  for (int c = 0; c < nc; c++) {
    const double hx = (cv[c * 12 + 9] - cv[c * 12]) / fro, hz = (cv[c * 12 + 11] - cv[c * 12 + 2]) / fro;
    const double v = hx * dirx + hz * dirz;
    if (v > bestv) { bestv = v; best = c; }  // first maximum wins (np.argmax)
  }
  const double* cp = cv + best * 12;
  double lo = 0.0, hi = 1.0;  // bezier_closest G:316-333
#pragma unroll 1
  for (int lvl = 0; lvl < 8; lvl++) {
    const double mid = (lo + hi) * 0.5;
    double ax, ay, az, bx, by, bz;
    bezier_at(cp, lo, ax, ay, az);
    bezier_at(cp, hi, bx, by, bz);
    const double dlo = sqrt((ax - px) * (ax - px) + ay * ay + (az - pz) * (az - pz));
    const double dhi = sqrt((bx - px) * (bx - px) + by * by + (bz - pz) * (bz - pz));
    if (dlo < dhi) hi = mid; else lo = mid;
  }
  const double t = (lo + hi) * 0.5, s = 1 - t;
  double qx, qy, qz;
  bezier_at(cp, t, qx, qy, qz);
  double tx = 3 * (s * s) * (cp[3] - cp[0]); tx += 6 * s * t * (cp[6] - cp[3]); tx += 3 * (t * t) * (cp[9] - cp[6]);   // G:300-313
  double ty = 3 * (s * s) * (cp[4] - cp[1]); ty += 6 * s * t * (cp[7] - cp[4]); ty += 3 * (t * t) * (cp[10] - cp[7]);
  double tz = 3 * (s * s) * (cp[5] - cp[2]); tz += 6 * s * t * (cp[8] - cp[5]); tz += 3 * (t * t) * (cp[11] - cp[8]);
  const double nrm = sqrt(tx * tx + ty * ty + tz * tz);
  tx /= nrm; tz /= nrm;
  const double dd = clampd(dirx * tx + dirz * tz, -1.0, 1.0);
  const double rvx = -tz, rvz = tx;  // cross(tangent, +y)
  r.dist = (px - qx) * rvx + (pz - qz) * rvz;
  r.angle_rad = acos(dd);
  if (dirx * rvx + dirz * rvz < 0) r.angle_rad = -r.angle_rad;
  r.dot_dir = dd;
  r.in_lane = true;
  return r;
}

// One integration step of the restated duckietown_world model (DESIGN.md "dynamics"; call site S:2083-2086).
__device__ __forceinline__ void dynamics_step(double& x, double& y, double& th, double& u, double& w, double l,
                                              double r, const DynParams& p, double trim, double dt) {
  l = clampd(l, -1.0, 1.0);
  r = clampd(r, -1.0, 1.0);
  const double uar = p.uar * (1.0 + trim), ual = p.ual * (1.0 - trim);
  const double war = p.war * (1.0 + trim), wal = p.wal * (1.0 - trim);
  const double du = -p.u1 * u - p.u2 * w + p.u3 * w * w + (uar * r + ual * l);
  const double dw = -p.w1 * w - p.w2 * u - p.w3 * u * w + (war * r - wal * l);
  const double u1 = u + dt * du, w1 = w + dt * dw;
  const double a = dt * w1, vx = dt * u1;
  double sa, ca;  // sin(a)/a, (1-cos a)/a
  if (fabs(a) < 1e-9) { sa = 1.0 - a * a / 6.0; ca = a / 2.0; }
  else { double s_, c_; sincos(a, &s_, &c_); sa = s_ / a; ca = (1.0 - c_) / a; }
  const double tx = sa * vx, ty = ca * vx;
  double sn, cs;
  sincos(th, &sn, &cs);
  x += cs * tx - sn * ty;
  y += sn * tx + cs * ty;
  th += a;
  u = u1;
  w = w1;
}

}  // namespace dts
