// placeholder: replaced by the real rasteriser
#include "dts_kernels.h"
namespace dts {
size_t render_scratch_bytes(int, int, int) { return 256; }
__global__ void k_clear(const DState S, int w, int h, uint8_t* obs) {
  const size_t n = (size_t)S.n * w * h;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const RenderEp& r = S.rep[p / ((size_t)w * h)];
    for (int k = 0; k < 3; k++) obs[3 * p + k] = (uint8_t)rintf(fminf(fmaxf(r.horizon[k], 0.f), 1.f) * 255.f);
  }
}
int launch_render(const DState& S, const DMap*, const RenderCfg& rc, uint8_t* obs, void*, int, int, int, const float*,
                  const float*, int32_t*, cudaStream_t st) {
  k_clear<<<1024, 256, 0, st>>>(S, rc.width, rc.height, obs);
  return 1;
}
}  // namespace dts
