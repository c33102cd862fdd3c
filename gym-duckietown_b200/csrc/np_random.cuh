// np_random.cuh — NumPy-compatible random streams on the device.
//
// The reference draws every reset quantity from `numpy.random.Generator(PCG64(SeedSequence(seed)))`
// (simulator.py:1043-1045 via gym.utils.seeding.np_random; draw sites simulator.py:546-736,
// randomization/randomizer.py:36-91).  To let DEVICE-side resets (dts_reset_random / auto-reset) stay on
// that stream draw for draw, this header restates the published algorithms behind the four numpy calls the
// reference makes:
//   PCG64 (pcg_setseq_128_xsl_rr_64): state = state * 0x2360ED051FC65DA44385DF649FCCF645 + inc (mod 2^128),
//          output rotr64(hi ^ lo, hi >> 58) of the NEW state; 32-bit draws return the low half and cache the high
//   Generator.uniform(lo, hi)   = lo + (hi - lo) * (next64 >> 11) * 2^-53
//   Generator.integers(lo, hi)  = Lemire's nearly-divisionless rejection on 32-bit draws (ranges < 2^32)
//   Generator.normal(loc, sc)   = loc + sc * ziggurat(256 layers), tables in np_ziggurat_tables.h
// The host seeds the streams (numpy itself computes SeedSequence -> initial state) and uploads
// (state, inc, has_uint32, uinteger) per env with dts_seed_streams.  tests/test_gpu_logic.py checks the
// device draws against numpy through the reference's own reset() golden vectors.
#pragma once
#include <cstdint>

#include "np_ziggurat_tables.h"

namespace dts {

struct NpStream {
  unsigned __int128 state, inc;
  uint32_t has32, cache32;

  __device__ __forceinline__ uint64_t next64() {
    const unsigned __int128 mult = ((unsigned __int128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;
    state = state * mult + inc;
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
  }
  __device__ __forceinline__ uint32_t next32() {
    if (has32) { has32 = 0; return cache32; }
    const uint64_t n = next64();
    has32 = 1;
    cache32 = (uint32_t)(n >> 32);
    return (uint32_t)n;
  }
  __device__ __forceinline__ double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
  __device__ __forceinline__ double uniform(double lo, double hi) { return lo + (hi - lo) * next_double(); }
  // Generator.integers(lo, hi): hi exclusive, hi - lo <= 2^32
  __device__ inline int integers(int lo, int hi) {
    const uint32_t rng = (uint32_t)(hi - 1 - lo);
    if (rng == 0) return lo;
    const uint32_t rng_excl = rng + 1u;
    uint64_t m = (uint64_t)next32() * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
      const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
      while (leftover < threshold) { m = (uint64_t)next32() * rng_excl; leftover = (uint32_t)m; }
    }
    return lo + (int)(m >> 32);
  }
  __device__ inline double standard_normal() {
    const double R = 3.6541528853610087963519472518, INV_R = 0.27366123732975827203338247596;
    for (;;) {
      uint64_t r = next64();
      const int idx = (int)(r & 0xff);
      r >>= 8;
      const int sign = (int)(r & 1);
      const uint64_t rabs = (r >> 1) & 0x000fffffffffffffULL;
      double x = (double)rabs * np_wi_double[idx];
      if (sign) x = -x;
      if (rabs < np_ki_double[idx]) return x;
      if (idx == 0) {
        for (;;) {
          const double xx = -INV_R * log1p(-next_double());
          const double yy = -log1p(-next_double());
          if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -(R + xx) : R + xx;
        }
      } else if (((np_fi_double[idx - 1] - np_fi_double[idx]) * next_double() + np_fi_double[idx]) < exp(-0.5 * x * x)) {
        return x;
      }
    }
  }
  __device__ __forceinline__ double normal(double loc, double scale) { return loc + scale * standard_normal(); }
};

}  // namespace dts
