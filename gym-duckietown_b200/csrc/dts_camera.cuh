// dts_camera.cuh — agent camera matrices of _render_img (simulator.py:1758-1803), float64,
// row-major 3x4 [R|t].  Shared by the render kernel and by the device reset (the reference captures
// GL_LIGHT0's position under whatever modelview the previous frame left, S:581).
#pragma once
#include "dts_common.cuh"

namespace dts {

constexpr double kDeg2Rad = 0.017453292519943295;

// modelview = Rx(cam_angle) * T(0,0,CAMERA_FORWARD_DIST) * gluLookAt(eye, eye+dir, +y)   S:1780-1803
__host__ __device__ inline void camera_view(double px, double pz, double angle, const RenderEp& ep, bool domain_rand,
                                            double V[12]) {
  double ex = px, ey = 0.0, ez = pz;
  if (domain_rand) { ex += (double)ep.cam_noise[0]; ey += (double)ep.cam_noise[1]; ez += (double)ep.cam_noise[2]; }  // S:1768-1769
  ey += (double)ep.cam_height;                                                                                       // S:1780
  double fx = cos(angle), fy = 0.0, fz = -sin(angle);  // get_dir_vec S:2056
  const double fn = sqrt(fx * fx + fy * fy + fz * fz);
  fx /= fn; fy /= fn; fz /= fn;
  // s = f x up, up = (0,1,0)
  double sx = fy * 0.0 - fz * 1.0, sy = fz * 0.0 - fx * 0.0, sz = fx * 1.0 - fy * 0.0;
  const double sn = sqrt(sx * sx + sy * sy + sz * sz);
  sx /= sn; sy /= sn; sz /= sn;
  const double ux = sy * fz - sz * fy, uy = sz * fx - sx * fz, uz = sx * fy - sy * fx;  // u = s x f
  double L[12] = {sx, sy, sz, -(sx * ex + sy * ey + sz * ez),
                  ux, uy, uz, -(ux * ex + uy * ey + uz * ez),
                  -fx, -fy, -fz, (fx * ex + fy * ey + fz * ez)};
  L[11] += (double)0.066f;  // glTranslatef(0, 0, CAMERA_FORWARD_DIST) S:1784: a GLfloat argument (S:131)
  const double th = (double)ep.cam_angle_deg * kDeg2Rad;
  const double c = cos(th), s = sin(th);
  for (int k = 0; k < 4; k++) {
    V[k] = L[k];
    V[4 + k] = c * L[4 + k] - s * L[8 + k];
    V[8 + k] = s * L[4 + k] + c * L[8 + k];
  }
}

// top_down=True (S:1786-1798): gluLookAt from (a, H, b) to (a, 0, b - 0.01), up +y, a / b = half the map extents,
// H = (max(a, b) + 0.1) / tan(fov_y / 2); no camera tilt / forward offset in this view.
__host__ __device__ inline void top_down_view(double grid_w, double grid_h, double tile_size, double fov_y_deg, double V[12]) {
  const double a = (grid_w * tile_size) / 2, b = (grid_h * tile_size) / 2;
  const double H = ((a > b ? a : b) + 0.1) / tan(fov_y_deg * kDeg2Rad / 2);
  const double ex = a, ey = H, ez = b;
  double fx = 0.0, fy = 0.0 - H, fz = (b - 0.01) - b;
  const double fn = sqrt(fx * fx + fy * fy + fz * fz);
  fx /= fn; fy /= fn; fz /= fn;
  double sx = fy * 0.0 - fz * 1.0, sy = fz * 0.0 - fx * 0.0, sz = fx * 1.0 - fy * 0.0;   // s = f x up
  const double sn = sqrt(sx * sx + sy * sy + sz * sz);
  sx /= sn; sy /= sn; sz /= sn;
  const double ux = sy * fz - sz * fy, uy = sz * fx - sx * fz, uz = sx * fy - sy * fx;   // u = s x f
  const double L[12] = {sx, sy, sz, -(sx * ex + sy * ey + sz * ez),
                        ux, uy, uz, -(ux * ex + uy * ey + uz * ez),
                        -fx, -fy, -fz, (fx * ex + fy * ey + fz * ez)};
  for (int k = 0; k < 12; k++) V[k] = L[k];
}

// Eye-space GL_POSITION for a light given under modelview V: positional (w=1) or direction (w=0).
__host__ __device__ inline void light_to_eye(const double V[12], const float lp[4], float out[4]) {
  const double x = lp[0], y = lp[1], z = lp[2], w = lp[3];
  out[0] = (float)(V[0] * x + V[1] * y + V[2] * z + V[3] * w);
  out[1] = (float)(V[4] * x + V[5] * y + V[6] * z + V[7] * w);
  out[2] = (float)(V[8] * x + V[9] * y + V[10] * z + V[11] * w);
  out[3] = (float)w;
}

}  // namespace dts
