// device_math.cuh -- fp32 helpers of the sm_100a path tracer.
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, so a
// fused multiply-add exists only where fmaf() is written, and +,-,*,/,sqrtf are
// IEEE-754 correctly rounded.  dot3/cross3 and the polynomial sin/cos/exp/atan2/
// acos below use the association fixed in DESIGN.md ("Arithmetic contract"), which
// is what lets tests/ compare the kernels bit-for-bit with the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace crt {

#define CRT_MAXFLOAT 1.0e15f
#define CRT_FLT_EPS 1.0e-5f
#define CRT_PI 3.14159265358979f
#define CRT_2PI 6.28318530717959f
#define CRT_INV_PI 0.318309886183791f
#define CRT_INV_2PI 0.159154943091895f
#define CRT_MIN_THROUGHPUT 1.0e-3f
#define CRT_MIN_CONTRIBUTION 1.0e-2f

struct v3 { float x, y, z; };

#define CRT_HD __host__ __device__ __forceinline__

CRT_HD v3 V(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
CRT_HD v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
CRT_HD v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
CRT_HD v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
CRT_HD v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
CRT_HD float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
CRT_HD v3 cross3(v3 a, v3 b)
{
  return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
CRT_HD v3 normalize3(v3 a) { return vscale(a, 1.0f / sqrtf(dot3(a, a))); }
// select-form min/max (NaN and signed-zero behaviour of the specification)
CRT_HD float minf(float a, float b) { return a < b ? a : b; }
CRT_HD float maxf(float a, float b) { return a > b ? a : b; }
CRT_HD bool any_gt(v3 a, float s) { return a.x > s || a.y > s || a.z > s; }
CRT_HD bool all_lt(v3 a, float s) { return a.x < s && a.y < s && a.z < s; }

// sin and cos of 2*pi*x, x in [0,1]: quadrant reduction + Horner polynomials.
__device__ __forceinline__ void sincos2pi(float x, float& s, float& c)
{
  float y = x * 4.0f;
  int q = (int)(y + 0.5f);
  float r = y - (float)q;
  float a = r * 1.57079632679490f;
  float a2 = a * a;
  float ps = fmaf(a2, 2.75573192e-6f, -1.98412698e-4f);
  ps = fmaf(a2, ps, 8.33333333e-3f);
  ps = fmaf(a2, ps, -1.66666667e-1f);
  ps = fmaf(a2 * a, ps, a);
  float pc = fmaf(a2, 2.48015873e-5f, -1.38888889e-3f);
  pc = fmaf(a2, pc, 4.16666667e-2f);
  pc = fmaf(a2, pc, -0.5f);
  pc = fmaf(a2, pc, 1.0f);
  switch (q & 3) {
    case 0: s = ps; c = pc; break;
    case 1: s = pc; c = -ps; break;
    case 2: s = -ps; c = -pc; break;
    default: s = -pc; c = ps; break;
  }
}

// e^x, clamped to [-87, 88].
__device__ __forceinline__ float exp_poly(float x)
{
  x = minf(maxf(x, -87.0f), 88.0f);
  float n = floorf(fmaf(x, 1.44269504088896f, 0.5f));
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.42860682030941723e-6f, r);
  float p = fmaf(r, 1.98412698e-4f, 1.38888889e-3f);
  p = fmaf(r, p, 8.33333333e-3f);
  p = fmaf(r, p, 4.16666667e-2f);
  p = fmaf(r, p, 1.66666667e-1f);
  p = fmaf(r, p, 0.5f);
  p = fmaf(r, p, 1.0f);
  p = fmaf(r, p, 1.0f);
  return p * __uint_as_float((uint32_t)((int)n + 127) << 23);
}

__device__ __forceinline__ float atan2_poly(float y, float x)
{
  float ax = fabsf(x), ay = fabsf(y);
  float mx = maxf(ax, ay), mn = minf(ax, ay);
  if (mx == 0.0f) return 0.0f;
  float a = mn / mx;
  float s = a * a;
  float p = fmaf(s, -0.01172120f, 0.05265332f);
  p = fmaf(s, p, -0.11643287f);
  p = fmaf(s, p, 0.19354346f);
  p = fmaf(s, p, -0.33262347f);
  p = fmaf(s, p, 0.99997726f);
  float r = p * a;
  if (ay > ax) r = 1.57079632679490f - r;
  if (x < 0.0f) r = CRT_PI - r;
  if (y < 0.0f) r = -r;
  return r;
}

__device__ __forceinline__ float acos_poly(float x)
{
  float ax = minf(fabsf(x), 1.0f);
  float p = fmaf(ax, -0.0187293f, 0.0742610f);
  p = fmaf(ax, p, -0.2121144f);
  p = fmaf(ax, p, 1.5707288f);
  float r = sqrtf(1.0f - ax) * p;
  return x < 0.0f ? CRT_PI - r : r;
}

// SeedRand / RandFloat, SURVEY A.8.
CRT_HD uint32_t seed_rand(uint32_t frame_seed, uint32_t px, uint32_t py, uint32_t size_x, uint32_t radius)
{
  uint32_t s = (py / radius) * size_x + px / radius + frame_seed;
  s = (s + 0x479ab41du) + (s << 8);
  s = (s ^ 0xe4aa10ceu) ^ (s >> 5);
  s = (s + 0x9942f0a6u) - (s << 14);
  s = (s ^ 0x5aedd67du) ^ (s >> 3);
  s = (s + 0x17bea992u) + (s << 7);
  return s;
}

CRT_HD float rand_float(uint32_t& state)
{
  uint32_t s = state;
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  state = s;
  return minf((float)s * 2.3283064365386963e-10f, 0.99999994f);
}

}  // namespace crt
