// dts_render.cu — batched software rasteriser for the agent camera (Simulator._render_img,
// simulator.py:1707-1951) on sm_100a.  No tensor cores: there is no dense contraction here.
//
// One persistent CTA (256 threads = 8 warps) renders one env at a time:
//   G  geometry   warp-per-draw-item (ground / map tile / placed mesh): model-view in f64->f32,
//                 per-vertex fixed-function lighting (tile lattice 8x8 shared through smem), frustum
//                 cull, guard-band/near clip, snap to 1/64 px, triangle setup -> 128-byte PrimRec
//                 appended to the CTA's slab in HBM/L2; per-16x16-bin counters in smem
//   B  binning    exclusive scan of the bin counters, second pass scatters prim indices per bin
//   R  raster     bin by bin: 64 PrimRecs at a time staged HBM->smem (edge functions re-based to the
//                 bin so lanes work in int32), 8 warps each own an 8x4 pixel block, one pixel per
//                 lane with its 4 MSAA samples (depth, colour, prim id) in registers
//   O  output     box resolve -> u8, 16x16x3 tile staged in smem and stored as 16-byte rows
// Arithmetic follows the render spec in oracle/dt_oracle_raster.c / DESIGN.md bit for bit
// (compiled with -fmad=false; fmaf() is spelled out where the spec has one).
//
// HBM traffic per env-frame: obs store W*H*3 B (compulsory) + PrimRec slab write/read (~128 B x
// visible triangles, L2-resident) + texels (shared by all envs, L2-resident).
#include "dts_camera.cuh"
#include "dts_kernels.h"

namespace dts {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kBin = 16;          // bin edge in pixels
constexpr int kChunk = 64;        // prims staged per pass
constexpr float kGuard = 4.0f;
constexpr int kSub = 64;          // sub-pixel units per pixel
__constant__ int c_sx[4] = {24, 56, 8, 40};
__constant__ int c_sy[4] = {8, 24, 40, 56};

struct Vtx { float cx, cy, cz, cw, r, g, b, u, v; };

struct __align__(16) PrimRec {   // 128 B in the CTA's HBM slab
  int32_t X[3], Y[3];            // snapped vertices, orientation normalised (area > 0)
  float f0[7], fx[7], fy[7];     // planes anchored at vertex 0: z, q=1/w, u*q, v*q, r*q, g*q, b*q
  int32_t id_tex;                // draw id << 8 | (texture index + 1)
  uint32_t bbox;                 // bin bbox: bx0 | by0<<8 | bx1<<16 | by1<<24
  int32_t px0y0, px1y1;          // pixel bbox (x | y<<16)
  int32_t pad;
};
static_assert(sizeof(PrimRec) == 128, "PrimRec must be 128 bytes");

struct __align__(16) BinPrim {   // smem, per staged prim, values re-based to the current bin
  int32_t E0[3], A[3], B[3];     // E_k(x,y) = E0_k + A_k*x + B_k*y, x,y in 1/64 px from the bin corner
  int32_t x0, y0;                // anchor vertex relative to the bin corner (sub-pixels)
  float f0[7], fx[7], fy[7];
  int32_t id;                    // draw id
  int32_t tex_w, tex_h;          // 0 = untextured
  const uint8_t* tex;
  int32_t px0, py0, px1, py1;    // pixel bbox relative to the bin
  int32_t live;                  // 0 = trivially rejected for this bin
  int32_t pad;
};

struct Xform { float MV[12], N[9]; };

struct Shared {
  RenderEp ep;
  double V[12];
  float P00, P11, P22, P23;
  int n_prims, n_pairs, overflow;
  Vtx lattice[kWarps][64];
  BinPrim chunk[kChunk];
  uint8_t tile[kBin][kBin * 3];
};

// MV = V * T(t) * S(sc) * Ry(c,s), N = rot(V) * Ry / sc — float64 then rounded (spec)
__device__ __forceinline__ void model_view(const double* V, double tx, double ty, double tz, double sc, double c,
                                           double s, Xform& x) {
  const double R[9] = {c, 0.0, s, 0.0, 1.0, 0.0, -s, 0.0, c};
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double a = V[4 * r + 0] * R[0 + k] + V[4 * r + 1] * R[3 + k] + V[4 * r + 2] * R[6 + k];
      x.MV[4 * r + k] = (float)(a * sc);
      x.N[3 * r + k] = (float)(a / sc);
    }
    x.MV[4 * r + 3] = (float)(V[4 * r + 0] * tx + V[4 * r + 1] * ty + V[4 * r + 2] * tz + V[4 * r + 3]);
  }
}

// fixed-function transform & lighting of one vertex (float32, operation order = spec)
__device__ __forceinline__ Vtx shade_vertex(const Xform& x, const Shared& sh, float px, float py, float pz, float nx,
                                            float ny, float nz, float cr, float cg, float cb, float u, float v) {
  float e[3], ne[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    float t = x.MV[4 * r] * px;
    t = t + x.MV[4 * r + 1] * py;
    t = t + x.MV[4 * r + 2] * pz;
    e[r] = t + x.MV[4 * r + 3];
    float q = x.N[3 * r] * nx;
    q = q + x.N[3 * r + 1] * ny;
    ne[r] = q + x.N[3 * r + 2] * nz;
  }
  const float* lp = sh.ep.light_eye;
  float lx, ly, lz;
  if (lp[3] == 0.0f) { lx = lp[0]; ly = lp[1]; lz = lp[2]; }
  else { lx = lp[0] - e[0]; ly = lp[1] - e[1]; lz = lp[2] - e[2]; }
  float len = lx * lx;
  len = len + ly * ly;
  len = len + lz * lz;
  len = sqrtf(len);
  float ndl = 0.0f;
  if (len > 0.0f) {
    lx = lx / len; ly = ly / len; lz = lz / len;
    ndl = ne[0] * lx;
    ndl = ndl + ne[1] * ly;
    ndl = ndl + ne[2] * lz;
    if (!(ndl > 0.0f)) ndl = 0.0f;
  }
  Vtx o;
  const float col[3] = {cr, cg, cb};
  float lit[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float s = 0.3f + sh.ep.ambient[k];
    s = s + ndl * sh.ep.diffuse[k];
    const float c = col[k] * s;
    lit[k] = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
  }
  o.r = lit[0]; o.g = lit[1]; o.b = lit[2];
  o.u = u; o.v = v;
  o.cx = sh.P00 * e[0];
  o.cy = sh.P11 * e[1];
  o.cz = sh.P22 * e[2] + sh.P23;
  o.cw = -e[2];
  return o;
}

__device__ __forceinline__ float plane_dist(const Vtx& a, int pl) {
  switch (pl) {
    case 0: return a.cz + a.cw;
    case 1: return a.cw - a.cz;
    case 2: return a.cx + kGuard * a.cw;
    case 3: return kGuard * a.cw - a.cx;
    case 4: return a.cy + kGuard * a.cw;
    default: return kGuard * a.cw - a.cy;
  }
}

// 0 = visible without clipping, 1 = needs the clipper, 2 = invisible (outside one true-frustum plane)
__device__ __forceinline__ int classify(const Vtx& a, const Vtx& b, const Vtx& c) {
  const Vtx* v[3] = {&a, &b, &c};
  int out[6] = {0, 0, 0, 0, 0, 0}, need = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float x = v[k]->cx, y = v[k]->cy, z = v[k]->cz, w = v[k]->cw;
    out[0] += !(z + w >= 0.0f); out[1] += !(w - z >= 0.0f);
    out[2] += x < -w; out[3] += x > w; out[4] += y < -w; out[5] += y > w;
    need |= !(x + kGuard * w >= 0.0f) | !(kGuard * w - x >= 0.0f) | !(y + kGuard * w >= 0.0f) |
            !(kGuard * w - y >= 0.0f);
  }
  need |= out[0] | out[1];
#pragma unroll
  for (int p = 0; p < 6; p++) if (out[p] == 3) return 2;
  return need ? 1 : 0;
}

struct EmitCtx {
  Shared* sh;
  PrimRec* prims;
  int* bin_count;
  int max_prims, W, H, bins_x;
};

// screen mapping + triangle setup (spec steps 5-7) and append to the slab
__device__ __forceinline__ void setup_and_emit(const EmitCtx& ec, const Vtx& a, const Vtx& b, const Vtx& c, int id,
                                               int tex) {
  const Vtx* vs[3] = {&a, &b, &c};
  int X[3], Y[3];
  float zw[3], q[3];
  const float Wf = (float)ec.W, Hf = (float)ec.H;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float iw = 1.0f / vs[k]->cw;
    const float nx = vs[k]->cx * iw, ny = vs[k]->cy * iw, nz = vs[k]->cz * iw;
    const float sx = (nx * 0.5f + 0.5f) * Wf;
    const float sy = (0.5f - ny * 0.5f) * Hf;
    X[k] = (int)rintf(sx * 64.0f);
    Y[k] = (int)rintf(sy * 64.0f);
    zw[k] = nz * 0.5f + 0.5f;
    q[k] = iw;
  }
  const long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(X[2] - X[0]) * (Y[1] - Y[0]);
  if (area2 == 0) return;
  const int i1 = area2 < 0 ? 2 : 1, i2 = area2 < 0 ? 1 : 2;
  const int x0 = X[0], y0 = Y[0], x1 = X[i1], y1 = Y[i1], x2 = X[i2], y2 = Y[i2];
  int minx = min(x0, min(x1, x2)), maxx = max(x0, max(x1, x2));
  int miny = min(y0, min(y1, y2)), maxy = max(y0, max(y1, y2));
  int px0 = max(minx >> 6, 0), px1 = min(maxx >> 6, ec.W - 1);
  int py0 = max(miny >> 6, 0), py1 = min(maxy >> 6, ec.H - 1);
  if (px0 > px1 || py0 > py1) return;
  PrimRec r;
  r.X[0] = x0; r.X[1] = x1; r.X[2] = x2;
  r.Y[0] = y0; r.Y[1] = y1; r.Y[2] = y2;
  const float dx1 = (float)(x1 - x0) * 0.015625f, dy1 = (float)(y1 - y0) * 0.015625f;
  const float dx2 = (float)(x2 - x0) * 0.015625f, dy2 = (float)(y2 - y0) * 0.015625f;
  const float areaf = dx1 * dy2 - dx2 * dy1;
  const float ia = 1.0f / areaf;
  const Vtx* p0 = vs[0]; const Vtx* p1 = vs[i1]; const Vtx* p2 = vs[i2];
  const float q0 = q[0], q1 = q[i1], q2 = q[i2];
  const float v0[7] = {zw[0], q0, p0->u * q0, p0->v * q0, p0->r * q0, p0->g * q0, p0->b * q0};
  const float v1[7] = {zw[i1], q1, p1->u * q1, p1->v * q1, p1->r * q1, p1->g * q1, p1->b * q1};
  const float v2[7] = {zw[i2], q2, p2->u * q2, p2->v * q2, p2->r * q2, p2->g * q2, p2->b * q2};
#pragma unroll
  for (int at = 0; at < 7; at++) {
    const float d1 = v1[at] - v0[at], d2 = v2[at] - v0[at];
    r.f0[at] = v0[at];
    r.fx[at] = (d1 * dy2 - d2 * dy1) * ia;
    r.fy[at] = (d2 * dx1 - d1 * dx2) * ia;
  }
  r.id_tex = (id << 8) | (tex + 1);
  const int bx0 = px0 / kBin, bx1 = px1 / kBin, by0 = py0 / kBin, by1 = py1 / kBin;
  r.bbox = (uint32_t)bx0 | ((uint32_t)by0 << 8) | ((uint32_t)bx1 << 16) | ((uint32_t)by1 << 24);
  r.px0y0 = px0 | (py0 << 16);
  r.px1y1 = px1 | (py1 << 16);
  r.pad = 0;
  const int slot = atomicAdd(&ec.sh->n_prims, 1);
  if (slot >= ec.max_prims) { ec.sh->overflow = 1; return; }
  // 128-byte record as 8 x 16-byte stores
  const int4* src = reinterpret_cast<const int4*>(&r);
  int4* dst = reinterpret_cast<int4*>(ec.prims + slot);
#pragma unroll
  for (int k = 0; k < 8; k++) dst[k] = src[k];
  for (int by = by0; by <= by1; by++)
    for (int bx = bx0; bx <= bx1; bx++) atomicAdd(&ec.bin_count[by * ec.bins_x + bx], 1);
}

__device__ __forceinline__ Vtx clip_lerp(const Vtx& in, const Vtx& out, float din, float dout) {
  const float t = din / (din - dout);
  Vtx o;
  const float* a = reinterpret_cast<const float*>(&in);
  const float* b = reinterpret_cast<const float*>(&out);
  float* c = reinterpret_cast<float*>(&o);
#pragma unroll
  for (int k = 0; k < 9; k++) { const float d = b[k] - a[k]; c[k] = a[k] + t * d; }
  return o;
}

// rare path: Sutherland-Hodgman against near, far and the guard band (spec step 4), then fan
__device__ __noinline__ void clip_and_emit(const EmitCtx& ec, const Vtx& a, const Vtx& b, const Vtx& c, int id,
                                           int tex) {
  Vtx poly[12], tmp[12];
  float d[12];
  poly[0] = a; poly[1] = b; poly[2] = c;
  int n = 3;
  for (int pl = 0; pl < 6; pl++) {
    int any_out = 0, all_out = 1;
    for (int k = 0; k < n; k++) { d[k] = plane_dist(poly[k], pl); const int o = !(d[k] >= 0.0f); any_out |= o; all_out &= o; }
    if (pl == 0 && n == 3) {
      // the oracle's trivial reject looks at all six planes of the ORIGINAL triangle before clipping
      for (int p2 = 0; p2 < 6; p2++) {
        int cnt = 0;
        for (int k = 0; k < 3; k++) cnt += !(plane_dist(poly[k], p2) >= 0.0f);
        if (cnt == 3) return;
      }
    }
    if (!any_out) continue;
    if (all_out) return;
    int m = 0;
    for (int k = 0; k < n; k++) {
      const int k2 = (k + 1 == n) ? 0 : k + 1;
      const bool in1 = d[k] >= 0.0f, in2 = d[k2] >= 0.0f;
      if (in1) tmp[m++] = poly[k];
      if (in1 && !in2) tmp[m++] = clip_lerp(poly[k], poly[k2], d[k], d[k2]);
      else if (!in1 && in2) tmp[m++] = clip_lerp(poly[k2], poly[k], d[k2], d[k]);
    }
    n = m;
    for (int k = 0; k < n; k++) poly[k] = tmp[k];
    if (n < 3) return;
  }
  for (int k = 1; k + 1 < n; k++) setup_and_emit(ec, poly[0], poly[k], poly[k + 1], id, tex);
}

__device__ __forceinline__ void process_triangle(const EmitCtx& ec, const Vtx& a, const Vtx& b, const Vtx& c, int id,
                                                 int tex) {
  const int cls = classify(a, b, c);
  if (cls == 2) return;
  if (cls == 0) setup_and_emit(ec, a, b, c, id, tex);
  else clip_and_emit(ec, a, b, c, id, tex);
}

}  // namespace

size_t render_scratch_bytes(int n_ctas, int max_prims, int max_pairs, size_t undistorted_frame_bytes) {
  const size_t slab = (size_t)max_prims * sizeof(PrimRec) + (((size_t)max_pairs * sizeof(uint16_t) + 255) & ~size_t(255));
  return (size_t)n_ctas * (slab + undistorted_frame_bytes) + 256;
}

__global__ void __launch_bounds__(kThreads, 2)
k_render(const DState S, const DMap* __restrict__ maps, RenderCfg rc, uint8_t* __restrict__ obs,
         uint8_t* __restrict__ scratch, int max_prims, int max_pairs, uint8_t* __restrict__ undist,
         const float* __restrict__ lut_x, const float* __restrict__ lut_y, int32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const int W = rc.width, H = rc.height;
  const int bins_x = (W + kBin - 1) / kBin, bins_y = (H + kBin - 1) / kBin, n_bins = bins_x * bins_y;
  int* bin_count = reinterpret_cast<int*>(smem_raw + ((sizeof(Shared) + 15) & ~size_t(15)));
  int* bin_start = bin_count + n_bins;
  const size_t slab = (size_t)max_prims * sizeof(PrimRec) + (((size_t)max_pairs * sizeof(uint16_t) + 255) & ~size_t(255));
  PrimRec* prims = reinterpret_cast<PrimRec*>(scratch + (size_t)blockIdx.x * slab);
  uint16_t* pairs = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(prims) + (size_t)max_prims * sizeof(PrimRec));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool dr = (rc.flags & DTS_FLAG_DOMAIN_RAND) != 0;
  const bool fisheye = (rc.flags & DTS_FLAG_DISTORTION) != 0;
  const size_t frame_bytes = (size_t)W * H * 3;
  uint8_t* my_undist = fisheye ? undist + (size_t)blockIdx.x * frame_bytes : nullptr;

  for (int env = blockIdx.x; env < rc.n_envs; env += gridDim.x) {
    const DMap& m = maps[S.map_id[env]];
    uint8_t* out = fisheye ? my_undist : obs + (size_t)env * frame_bytes;
    // ---------------------------------------------------------------- per-frame setup
    if (tid < (int)(sizeof(RenderEp) / 4)) reinterpret_cast<uint32_t*>(&sh.ep)[tid] = reinterpret_cast<const uint32_t*>(&S.rep[env])[tid];
    for (int b = tid; b < n_bins; b += kThreads) bin_count[b] = 0;
    __syncthreads();
    if (tid == 0) {
      camera_view(S.pos_x[env], S.pos_z[env], S.angle[env], sh.ep, dr, sh.V);
      const double f = 1.0 / tan((double)sh.ep.cam_fov_y_deg * kDeg2Rad / 2.0), aspect = (double)W / (double)H;
      const double zn = 0.04, zf = 100.0;                                     // gluPerspective S:1761
      sh.P00 = (float)(f / aspect); sh.P11 = (float)f;
      sh.P22 = (float)((zf + zn) / (zn - zf)); sh.P23 = (float)(2.0 * zf * zn / (zn - zf));
      sh.n_prims = 0; sh.n_pairs = 0; sh.overflow = 0;
    }
    __syncthreads();
    // ---------------------------------------------------------------- G: geometry, warp per draw item
    EmitCtx ec{&sh, prims, bin_count, max_prims, W, H, bins_x};
    const int n_tiles = m.grid_w * m.grid_h;
    const int n_items = 1 + n_tiles + m.n_objects;
    for (int item = warp; item < n_items; item += kWarps) {
      Xform x;
      if (item == 0) {
        // ground quad S:1805-1812: glScalef(50,0.01,50) applied to (+-1,-0.8,+-1), world-space +y normal
        if (lane < 2) {
          model_view(sh.V, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, x);
          const float gy = (float)(-0.8 * 0.01);
          const float P[4][3] = {{-50.f, gy, 50.f}, {-50.f, gy, -50.f}, {50.f, gy, -50.f}, {50.f, gy, 50.f}};
          const int i1 = lane == 0 ? 1 : 2, i2 = lane == 0 ? 2 : 3;
          const float* g = sh.ep.ground;
          const Vtx a = shade_vertex(x, sh, P[0][0], P[0][1], P[0][2], 0.f, 1.f, 0.f, g[0], g[1], g[2], 0.f, 0.f);
          const Vtx b = shade_vertex(x, sh, P[i1][0], P[i1][1], P[i1][2], 0.f, 1.f, 0.f, g[0], g[1], g[2], 0.f, 0.f);
          const Vtx c = shade_vertex(x, sh, P[i2][0], P[i2][1], P[i2][2], 0.f, 1.f, 0.f, g[0], g[1], g[2], 0.f, 0.f);
          process_triangle(ec, a, b, c, lane, -1);
        }
      } else if (item <= n_tiles) {
        // road tile S:1852-1884: draw order i outer, j inner; the tile's 8x8 lattice is lit once (2 verts/lane)
        const int t = item - 1, ti = t / m.grid_h, tj = t - ti * m.grid_h;
        const int idx = tj * m.grid_w + ti;
        if (m.tile_kind[idx] < 0) continue;
        const int quarter = (m.tile_angle[idx] + 2) & 3;                     // glRotatef(angle*90+180) S:1873
        const double cs = quarter == 0 ? 1.0 : (quarter == 2 ? -1.0 : 0.0), sn = quarter == 1 ? 1.0 : (quarter == 3 ? -1.0 : 0.0);
        const double ts = m.tile_size;
        model_view(sh.V, (ti + 0.5) * ts, 0.0, (tj + 0.5) * ts, 1.0, cs, sn, x);
        int outside[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int vi = lane + 32 * h, a = vi >> 3, b = vi & 7;             // a: u index (x), b: v index (z)
          const float lx = (float)(-ts / 2 + ((double)a / 7.0) * ts), lz = (float)(-ts / 2 + ((double)b / 7.0) * ts);
          const Vtx v = shade_vertex(x, sh, lx, 0.0f, lz, 0.f, 1.f, 0.f, 1.f, 1.f, 1.f, (float)((double)a / 7.0),
                                     (float)(1.0 - (double)b / 7.0));
          sh.lattice[warp][vi] = v;
          outside[0] += !(v.cz + v.cw >= 0.0f); outside[1] += !(v.cw - v.cz >= 0.0f);
          outside[2] += v.cx < -v.cw; outside[3] += v.cx > v.cw; outside[4] += v.cy < -v.cw; outside[5] += v.cy > v.cw;
        }
        bool culled = false;
#pragma unroll
        for (int p = 0; p < 6; p++) culled |= __all_sync(0xffffffffu, outside[p] == 2);
        __syncwarp();
        if (culled) continue;
        const int tex = m.tile_tex[idx];
        const int base_id = 2 + 98 * t;
        for (int k = lane; k < 98; k += 32) {                                // S:407-433 quad order, (0,1,2)(0,2,3) split
          const int quad = k >> 1, half = k & 1, a = quad / 7, b = quad - 7 * a;
          const Vtx& v0 = sh.lattice[warp][a * 8 + b];
          const Vtx& v2 = sh.lattice[warp][(a + 1) * 8 + b + 1];
          const Vtx& v1 = half == 0 ? sh.lattice[warp][(a + 1) * 8 + b] : v2;
          const Vtx& v2b = half == 0 ? v2 : sh.lattice[warp][a * 8 + b + 1];
          process_triangle(ec, v0, v1, v2b, base_id + k, tex);
        }
        __syncwarp();
      } else {
        // placed mesh S:1905-1907, O:123-148: T(pos) S(scale) Ry(y_rot)
        const int o = item - 1 - n_tiles;
        if (sh.ep.hidden[o >> 5] >> (o & 31) & 1u) continue;
        const DObject& ob = m.objects[o];
        double sn, cs;
        sincos((double)ob.y_rot_deg * kDeg2Rad, &sn, &cs);
        model_view(sh.V, (double)ob.pos[0], (double)ob.pos[1], (double)ob.pos[2], (double)ob.scale, cs, sn, x);
        {  // conservative bounding-sphere cull in eye space against the four side planes and near
          const float cx_ = x.MV[0] * ob.centre[0] + x.MV[1] * ob.centre[1] + x.MV[2] * ob.centre[2] + x.MV[3];
          const float cy_ = x.MV[4] * ob.centre[0] + x.MV[5] * ob.centre[1] + x.MV[6] * ob.centre[2] + x.MV[7];
          const float cz_ = x.MV[8] * ob.centre[0] + x.MV[9] * ob.centre[1] + x.MV[10] * ob.centre[2] + x.MV[11];
          const float rad = ob.bound_rad * ob.scale * 1.001f + 1e-4f;
          const float hx = rsqrtf(sh.P00 * sh.P00 + 1.0f), hy = rsqrtf(sh.P11 * sh.P11 + 1.0f);
          bool outside_ = cz_ - rad > -0.04f;                                  // entirely behind the near plane
          outside_ |= (sh.P00 * cx_ + cz_) * hx > rad * 1.01f;                 // right plane: P00*x <= -z
          outside_ |= (-sh.P00 * cx_ + cz_) * hx > rad * 1.01f;
          outside_ |= (sh.P11 * cy_ + cz_) * hy > rad * 1.01f;
          outside_ |= (-sh.P11 * cy_ + cz_) * hy > rad * 1.01f;
          if (outside_) continue;
        }
        int base_id = 2 + 98 * n_tiles;
        for (int q = 0; q < o; q++) base_id += m.objects[q].tri_count;
        for (int k = lane; k < ob.tri_count; k += 32) {
          const size_t ti = (size_t)ob.tri_offset + k;
          const float* p = m.tri_pos + ti * 9;
          const float* n = m.tri_nrm + ti * 9;
          const float* uv = m.tri_uv + ti * 6;
          const float* c = m.tri_col + ti * 9;
          Vtx v[3];
#pragma unroll
          for (int j = 0; j < 3; j++)
            v[j] = shade_vertex(x, sh, p[3 * j], p[3 * j + 1], p[3 * j + 2], n[3 * j], n[3 * j + 1], n[3 * j + 2],
                                c[3 * j], c[3 * j + 1], c[3 * j + 2], uv[2 * j], uv[2 * j + 1]);
          process_triangle(ec, v[0], v[1], v[2], base_id + k, m.tri_tex[ti]);
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- B: scan + scatter
    const int n_prims = min(sh.n_prims, max_prims);
    if (warp == 0) {  // exclusive scan of bin_count by one warp
      int carry = 0;
      for (int base = 0; base < n_bins; base += 32) {
        const int b = base + lane;
        const int v = b < n_bins ? bin_count[b] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t_ = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t_; }
        if (b < n_bins) bin_start[b] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) { sh.n_pairs = carry; if (carry > max_pairs) sh.overflow = 1; }
    }
    __syncthreads();
    const bool pairs_ok = sh.n_pairs <= max_pairs;
    for (int b = tid; b < n_bins; b += kThreads) bin_count[b] = 0;   // reuse as fill cursors
    __syncthreads();
    if (pairs_ok) {
      for (int p = tid; p < n_prims; p += kThreads) {
        const uint32_t bb = prims[p].bbox;
        const int bx0 = bb & 255, by0 = (bb >> 8) & 255, bx1 = (bb >> 16) & 255, by1 = bb >> 24;
        for (int by = by0; by <= by1; by++)
          for (int bx = bx0; bx <= bx1; bx++) {
            const int b = by * bins_x + bx;
            pairs[bin_start[b] + atomicAdd(&bin_count[b], 1)] = (uint16_t)p;
          }
      }
    }
    __syncthreads();
    if (tid == 0 && sh.overflow) atomicOr(err, 1);
    // ---------------------------------------------------------------- R: raster, bin by bin
    const int sbx = (warp & 1) * 8, sby = (warp >> 1) * 4;   // this warp's 8x4 block inside the bin
    const int lx = sbx + (lane & 7), ly = sby + (lane >> 3); // this lane's pixel inside the bin
    const float clr[3] = {sh.ep.horizon[0], sh.ep.horizon[1], sh.ep.horizon[2]};
    for (int bin = 0; bin < n_bins; bin++) {
      const int bx = bin % bins_x, by = bin / bins_x;
      const int count = pairs_ok ? bin_count[bin] : 0;
      const int start = bin_start[bin];
      float z[4], cr[4], cg[4], cb[4];
      int wid[4];
#pragma unroll
      for (int s = 0; s < 4; s++) { z[s] = 1.0f; cr[s] = clr[0]; cg[s] = clr[1]; cb[s] = clr[2]; wid[s] = 0x7fffffff; }
      for (int c0 = 0; c0 < count; c0 += kChunk) {
        const int nch = min(kChunk, count - c0);
        __syncthreads();   // previous chunk fully consumed
        if (tid < nch) {
          const PrimRec& r = prims[pairs[start + c0 + tid]];
          BinPrim& bp = sh.chunk[tid];
          const int ox = bx * kBin * kSub, oy = by * kBin * kSub;
          const int X0 = r.X[0], X1 = r.X[1], X2 = r.X[2], Y0 = r.Y[0], Y1 = r.Y[1], Y2 = r.Y[2];
          const int ax[3] = {X1, X2, X0}, ay[3] = {Y1, Y2, Y0}, bxv[3] = {X2, X0, X1}, byv[3] = {Y2, Y0, Y1};
          int live = 1;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const int dx = bxv[k] - ax[k], dy = byv[k] - ay[k];
            const int bias = (dy > 0 || (dy == 0 && dx < 0)) ? 0 : 1;
            // E(x,y) = dx*(y-ay) - dy*(x-ax) - bias  at the bin corner, exact in 64 bits
            long long e0 = (long long)dx * (oy - ay[k]) - (long long)dy * (ox - ax[k]) - bias;
            // inside the bin |A*x+B*y| < 2^29: beyond +-2^30 the sign is decided for every sample
            if (e0 < -(1LL << 30)) live = 0;
            if (e0 > (1LL << 30)) e0 = (1LL << 30);
            bp.E0[k] = (int)e0; bp.A[k] = -dy; bp.B[k] = dx;
          }
          bp.x0 = X0 - ox; bp.y0 = Y0 - oy;
#pragma unroll
          for (int k = 0; k < 7; k++) { bp.f0[k] = r.f0[k]; bp.fx[k] = r.fx[k]; bp.fy[k] = r.fy[k]; }
          bp.id = r.id_tex >> 8;
          const int tex = (r.id_tex & 255) - 1;
          if (tex >= 0) { const DTexture t = m.textures[tex]; bp.tex = t.rgba; bp.tex_w = t.w; bp.tex_h = t.h; }
          else { bp.tex = nullptr; bp.tex_w = 0; bp.tex_h = 0; }
          bp.px0 = (r.px0y0 & 0xffff) - bx * kBin; bp.py0 = (r.px0y0 >> 16) - by * kBin;
          bp.px1 = (r.px1y1 & 0xffff) - bx * kBin; bp.py1 = (r.px1y1 >> 16) - by * kBin;
          bp.live = live;
        }
        __syncthreads();
        for (int k = 0; k < nch; k++) {
          const BinPrim& bp = sh.chunk[k];
          // warp-uniform reject: prim's pixel bbox vs this warp's 8x4 block
          if (!bp.live || bp.px1 < sbx || bp.px0 > sbx + 7 || bp.py1 < sby || bp.py0 > sby + 3) continue;
          const int pxs = lx * kSub, pys = ly * kSub;
          int mask = 0;
#pragma unroll
          for (int s = 0; s < 4; s++) {
            const int xs = pxs + c_sx[s], ys = pys + c_sy[s];
            const int e0 = bp.E0[0] + bp.A[0] * xs + bp.B[0] * ys;
            const int e1 = bp.E0[1] + bp.A[1] * xs + bp.B[1] * ys;
            const int e2 = bp.E0[2] + bp.A[2] * xs + bp.B[2] * ys;
            if ((e0 | e1 | e2) >= 0) mask |= 1 << s;
          }
          if (!mask) continue;
          const float cdx = (float)(pxs + 32 - bp.x0) * 0.015625f, cdy = (float)(pys + 32 - bp.y0) * 0.015625f;
          float qq = fmaf(bp.fy[1], cdy, fmaf(bp.fx[1], cdx, bp.f0[1]));
          if (!(qq > 1e-20f)) qq = 1e-20f;
          const float rq = 1.0f / qq;
          const float u = fmaf(bp.fy[2], cdy, fmaf(bp.fx[2], cdx, bp.f0[2])) * rq;
          const float v = fmaf(bp.fy[3], cdy, fmaf(bp.fx[3], cdx, bp.f0[3])) * rq;
          float c3[3];
          c3[0] = fmaf(bp.fy[4], cdy, fmaf(bp.fx[4], cdx, bp.f0[4])) * rq;
          c3[1] = fmaf(bp.fy[5], cdy, fmaf(bp.fx[5], cdx, bp.f0[5])) * rq;
          c3[2] = fmaf(bp.fy[6], cdy, fmaf(bp.fx[6], cdx, bp.f0[6])) * rq;
          if (bp.tex) {
            const float tx = u * (float)bp.tex_w - 0.5f, ty = v * (float)bp.tex_h - 0.5f;
            const float txf = floorf(tx), tyf = floorf(ty);
            const float ffx = tx - txf, ffy = ty - tyf;
            const int ti0 = ((int)txf) & (bp.tex_w - 1), ti1 = (ti0 + 1) & (bp.tex_w - 1);
            const int tj0 = ((int)tyf) & (bp.tex_h - 1), tj1 = (tj0 + 1) & (bp.tex_h - 1);
            const uchar4* tp = reinterpret_cast<const uchar4*>(bp.tex);
            const uchar4 t00 = __ldg(tp + tj0 * bp.tex_w + ti0), t10 = __ldg(tp + tj0 * bp.tex_w + ti1);
            const uchar4 t01 = __ldg(tp + tj1 * bp.tex_w + ti0), t11 = __ldg(tp + tj1 * bp.tex_w + ti1);
            const float a0[3] = {(float)t00.x, (float)t00.y, (float)t00.z}, a1[3] = {(float)t10.x, (float)t10.y, (float)t10.z};
            const float b0[3] = {(float)t01.x, (float)t01.y, (float)t01.z}, b1[3] = {(float)t11.x, (float)t11.y, (float)t11.z};
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
              const float ta = fmaf(ffx, a1[ch] - a0[ch], a0[ch]);
              const float tb = fmaf(ffx, b1[ch] - b0[ch], b0[ch]);
              const float tc = fmaf(ffy, tb - ta, ta);
              c3[ch] = tc * (c3[ch] * 0.00392156862745098f);
            }
          }
#pragma unroll
          for (int s = 0; s < 4; s++) {
            if (!(mask >> s & 1)) continue;
            const float sdx = (float)(pxs + c_sx[s] - bp.x0) * 0.015625f, sdy = (float)(pys + c_sy[s] - bp.y0) * 0.015625f;
            const float zs = fmaf(bp.fy[0], sdy, fmaf(bp.fx[0], sdx, bp.f0[0]));
            if (zs < z[s] || (zs == z[s] && bp.id < wid[s])) {   // GL_LESS in draw order
              z[s] = zs; wid[s] = bp.id; cr[s] = c3[0]; cg[s] = c3[1]; cb[s] = c3[2];
            }
          }
        }
      }
      // ------------------------------------------------------------ O: resolve + store
      {
        float c;
        c = ((cr[0] + cr[1]) + (cr[2] + cr[3])) * 0.25f; c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
        sh.tile[ly][lx * 3 + 0] = (uint8_t)rintf(c * 255.0f);
        c = ((cg[0] + cg[1]) + (cg[2] + cg[3])) * 0.25f; c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
        sh.tile[ly][lx * 3 + 1] = (uint8_t)rintf(c * 255.0f);
        c = ((cb[0] + cb[1]) + (cb[2] + cb[3])) * 0.25f; c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
        sh.tile[ly][lx * 3 + 2] = (uint8_t)rintf(c * 255.0f);
      }
      __syncthreads();
      {
        const int gx0 = bx * kBin, gy0 = by * kBin;
        const int cols = min(kBin, W - gx0), rows = min(kBin, H - gy0);
        if (((W * 3) & 15) == 0 && cols == kBin) {
          if (tid < rows * 3) {   // 3 x 16-byte stores per pixel row of the tile
            const int row = tid / 3, part = tid - row * 3;
            *reinterpret_cast<int4*>(out + ((size_t)(gy0 + row) * W + gx0) * 3 + part * 16) =
                *reinterpret_cast<const int4*>(&sh.tile[row][part * 16]);
          }
        } else {
          for (int t_ = tid; t_ < rows * cols * 3; t_ += kThreads) {
            const int row = t_ / (cols * 3), col = t_ - row * cols * 3;
            out[((size_t)(gy0 + row) * W + gx0) * 3 + col] = sh.tile[row][col];
          }
        }
      }
      __syncthreads();
    }
    // ---------------------------------------------------------------- fisheye gather (distortion.py:118)
    if (fisheye) {
      __threadfence_block();
      __syncthreads();
      uint8_t* dst = obs + (size_t)env * frame_bytes;
      for (int p = tid; p < W * H; p += kThreads) {
        const int sx = (int)rintf(__ldg(lut_x + p)), sy = (int)rintf(__ldg(lut_y + p));
        uint8_t r = 0, g = 0, b = 0;
        if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
          const uint8_t* s = my_undist + ((size_t)sy * W + sx) * 3;
          r = s[0]; g = s[1]; b = s[2];
        }
        dst[(size_t)p * 3] = r; dst[(size_t)p * 3 + 1] = g; dst[(size_t)p * 3 + 2] = b;
      }
      __syncthreads();
    }
  }
}

int launch_render(const DState& S, const DMap* maps, const RenderCfg& rc, uint8_t* obs, void* scratch, int n_ctas,
                  int max_prims, int max_pairs, const float* lut_x, const float* lut_y, int32_t* err_flag,
                  cudaStream_t st) {
  const int bins = ((rc.width + kBin - 1) / kBin) * ((rc.height + kBin - 1) / kBin);
  const size_t smem = ((sizeof(Shared) + 15) & ~size_t(15)) + (size_t)bins * 2 * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_render, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  const size_t slab = (size_t)max_prims * sizeof(PrimRec) + (((size_t)max_pairs * sizeof(uint16_t) + 255) & ~size_t(255));
  uint8_t* undist = reinterpret_cast<uint8_t*>(scratch) + (size_t)n_ctas * slab;
  k_render<<<n_ctas, kThreads, smem, st>>>(S, maps, rc, obs, reinterpret_cast<uint8_t*>(scratch), max_prims, max_pairs,
                                           undist, lut_x, lut_y, err_flag);
  return 1;
}

}  // namespace dts
