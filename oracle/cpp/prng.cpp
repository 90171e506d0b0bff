// JAX threefry2x32 PRNG (legacy, non-partitionable layout) restated in C++ -- ORACLE / CPU BASELINE ONLY.
// TEST INFRASTRUCTURE: nothing under geomjax_b200/ may link or call this (see oracle/__init__.py).
//
// Restates oracle/prng.py (itself a restatement of jax/_src/prng.py `threefry_2x32`, `_threefry_split`,
// `_threefry_random_bits`; jax/_src/random.py `_uniform`, `_normal_real`; XLA `ErfInv32`), for the call sites
//   jax.random.split(rng_key, 2)       rmhmc/rmhmc.py:158, lmcmc/lmc.py:164, lmcmonge/lmc.py:196
//   jax.random.normal                  util.py:81-82
//   jax.random.bernoulli               mcmc/proposal.py:178
//   split(key, num_samples|num_chains) examples/funnel/main.py:18,22
// Compiled with -ffp-contract=off: every float operation of the normal transform is individually rounded, as in
// the NumPy oracle and the CUDA kernels, so draws are bit-identical across the three.
#include "geom_cpu.h"

#include <cmath>
#include <cstring>

namespace ocpu {

static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1) {
  static const int RA[4] = {13, 15, 26, 6}, RB[4] = {17, 29, 16, 24};
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int g = 1; g <= 5; ++g) {
    const int* R = (g & 1) ? RA : RB;
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl(x1, R[i]);
      x1 ^= x0;
    }
    x0 += ks[g % 3];
    x1 += ks[(g + 1) % 3] + (uint32_t)g;
  }
  *o0 = x0;
  *o1 = x1;
}

// element j of random_bits(key, (n,)) in legacy mode: counters are hashed pairwise (first half, second half),
// an odd n is padded with one zero counter
uint32_t bits_elem(Key k, uint32_t j, uint32_t n) {
  const uint32_t h = (n + 1) / 2;
  const bool lo = j < h;
  const uint32_t c0 = lo ? j : j - h;
  uint32_t c1 = c0 + h;
  if (c1 >= n) c1 = 0;  // the pad
  uint32_t o0, o1;
  threefry2x32(k.a, k.b, c0, c1, &o0, &o1);
  return lo ? o0 : o1;
}

Key split_index(Key k, uint32_t num, uint32_t idx) {
  Key r;
  r.a = bits_elem(k, 2 * idx, 2 * num);
  r.b = bits_elem(k, 2 * idx + 1, 2 * num);
  return r;
}

static inline float as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline uint32_t as_u32(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

static inline float unit_float(uint32_t bits) { return as_float((bits >> 9) | 0x3F800000u) - 1.0f; }

float uniform01(Key k) {  // uniform(key, ()) in [0, 1)
  const float f = unit_float(bits_elem(k, 0, 1));
  return std::fmax(0.0f, f * 1.0f + 0.0f);
}

// fdlibm / musl log1pf, every operation a rounded float32 operation (oracle/prng.py::log1p_f32)
static float log1p_f32(float x) {
  const float Lg1 = (float)(0xAAAAAA / 16777216.0), Lg2 = (float)(0xCCCE13 / 33554432.0),
              Lg3 = (float)(0x91E9EE / 33554432.0), Lg4 = (float)(0xF89E26 / 67108864.0);
  const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
  const uint32_t ix = as_u32(x);
  const bool small = (ix < 0x3ED413D0u) || ((ix >> 31) == 1);
  const bool tiny = small && ((ix << 1) < 0x67000000u);
  if (tiny) return x;
  const bool k0 = small && (ix <= 0xBE95F619u);
  int k;
  float c, f;
  if (k0) {
    k = 0;
    c = 0.0f;
    f = x;
  } else {
    const float u = 1.0f + x;
    const uint32_t iu = as_u32(u) + (0x3F800000u - 0x3F3504F3u);
    k = (int)(iu >> 23) - 0x7F;
    c = (k >= 2) ? 1.0f - (u - x) : x - (u - 1.0f);
    c = (k < 25) ? c / u : 0.0f;
    f = as_float((iu & 0x007FFFFFu) + 0x3F3504F3u) - 1.0f;
  }
  const float s = f / (2.0f + f);
  const float z = s * s;
  const float w = z * z;
  const float t1 = w * (Lg2 + w * Lg4);
  const float t2 = z * (Lg1 + w * Lg3);
  const float R = t2 + t1;
  const float hfsq = (0.5f * f) * f;
  const float dk = (float)k;
  float r = s * (hfsq + R);
  r = r + (dk * ln2_lo + c);
  r = r - hfsq;
  r = r + f;
  r = r + dk * ln2_hi;
  return r;
}

static float erfinv_f32(float x) {
  static const float LT[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                              -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float GE[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                              -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  if (std::fabs(x) == 1.0f) return x * 3.402823466e+38f;
  const float t = x * x;
  float w = -log1p_f32(-t);
  const bool lt = w < 5.0f;
  w = lt ? w - 2.5f : std::sqrt(w) - 3.0f;
  const float* cf = lt ? LT : GE;
  float p = cf[0];
  for (int i = 1; i < 9; ++i) p = cf[i] + p * w;
  return p * x;
}

float normal_from_bits(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float f = unit_float(bits);
  const float u = std::fmax(lo, f * (1.0f - lo) + lo);
  return 1.41421354f * erfinv_f32(u);
}

void normal(Key k, int n, float* out) {
  for (int j = 0; j < n; ++j) out[j] = normal_from_bits(bits_elem(k, (uint32_t)j, (uint32_t)n));
}

}  // namespace ocpu

extern "C" {
// test surface: the draws the transitions consume
void ocpu_normal(const uint32_t* key, int n, float* out) { ocpu::normal(ocpu::Key{key[0], key[1]}, n, out); }
float ocpu_uniform(const uint32_t* key) { return ocpu::uniform01(ocpu::Key{key[0], key[1]}); }
void ocpu_split_index(const uint32_t* key, uint32_t num, uint32_t idx, uint32_t* out) {
  const ocpu::Key r = ocpu::split_index(ocpu::Key{key[0], key[1]}, num, idx);
  out[0] = r.a;
  out[1] = r.b;
}
}
