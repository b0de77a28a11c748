// kernels.cuh -- sm_100a kernels of the wavefront path tracer.
//
// Replaces the GLSL programs OCCT runs for CADRays' V3d_View::Redraw()
// (src/Launcher/AppViewer.cxx:1047): SceneNearestHit / SceneAnyHit (SURVEY 8(a)
// rows a5-a7), PathTrace + layered BSDF + light sampling (rows a1-a4), RNG (a9),
// ray generation / accumulation / display (a10).
//
// Design: wavefront.  Path state lives in SoA float4 arrays indexed by path slot;
// compacted queues of slot indices are produced with warp-aggregated atomics
// (__ballot_sync + __popc, one atomicAdd per warp).  Kernels read the queue length
// from device memory, so a whole wave is enqueued without a host round trip.
// Traversal kernels are persistent (grid = SM count x resident CTAs): each lane
// pulls its next ray from a warp-local pool when its ray has finished (in batches
// of finished lanes, trace_persistent); camera rays are walked in lockstep per
// 8x4 pixel tile; shading kernels grid-stride with a grid that is a multiple
// of the SM count.  One wave (normally two half-frames of it on two streams):
//   k_extend_primary_lockstep -> k_shade<FIRST>(0) -> [k_trace_dual(d) -> k_tail(d) -> k_shade(d)]... -> k_connect -> k_resolve
#pragma once
#include "device_math.cuh"
#include "host_scene.hpp"   // kTriStride, reference encodings

// threads per CTA of the persistent traversal kernels
#ifndef CRT_TRACE_BLOCK
#define CRT_TRACE_BLOCK 128
#endif

namespace crt {

constexpr int kStackSize = 72;            // 32 top + 32 bottom levels + sentinel + slack
constexpr int kStackSizeQuad = 104;       // 4-wide: up to 3 pushes per level, half as many levels
constexpr int32_t kSentinel = 0x7ffffffe; // "leave the instance" marker on the stack
constexpr int32_t kNoRef = 0x7fffffff;

struct DeviceScene {
  const float4* __restrict__ nodes;       // 4 x float4 per inner node (see host_scene.hpp)
  const float4* __restrict__ tri_verts;   // 3 x float4 per triangle
  const float4* __restrict__ tri_nrm;     // 3 x float4 per triangle
  const float4* __restrict__ inst;        // 4 x float4 per instance
  const float4* __restrict__ mats;        // 8 x float4 per material (crt_bsdf)
  const uint8_t* __restrict__ mat_class;  // shading class of each material (kClass*)
  uint32_t mats_in_smem;                  // k_shade stages the material table in shared memory (n_mats <= kSmemMats)
  const float4* __restrict__ lights;      // 2 x float4 per light (shader form)
  const float4* __restrict__ env;         // lat-long texels, rgb_
  const float2* __restrict__ tri_uv;      // 3 x float2 per triangle
  const uchar4* __restrict__ tex_data;    // RGBA8 texels of all base-colour textures
  const uint32_t* __restrict__ tex_table; // offset, width, height per texture
  uint32_t n_tex;
  int32_t top_root;
  uint32_t n_mats, n_lights, env_w, env_h;
  float scene_eps;
  // top of the top-level tree (first nodes in breadth-first order), 80-byte stride (64 B node + 16 B pad
  // so that scattered shared-memory reads spread over the banks); staged per CTA by one TMA bulk copy
  const float4* __restrict__ top_cache;
  uint32_t n_top_cache;
};

struct DeviceParams {
  int32_t max_depth;
  float max_radiance;
  int32_t two_sided;
  uint32_t rng_radius;          // 8 in coherent mode else 1
  float aperture_radius, focal_dist;
  int32_t env_as_background;    // already and-ed with "a map exists"
  int32_t russian_roulette;
  float background[3];
  // camera basis (host computed, same formulas as the oracle)
  float eye[3], cu[3], cv[3], cw[3];
  float hw, hh;
  int32_t is_ortho;
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;    // 8x4 pixel tiles per warp
  // the part of the frame this launch works on: 8x4 tiles [tile0, tile0 + n_tiles) in row-major tile order (the whole
  // frame unless a wave is split into parts that run on several streams); path slot = sample * n_tiles * 32 + tile * 32 + lane
  uint32_t tile0, n_tiles;
  // path slot layout of the wave: a warp holds G = sample_group consecutive samples of a block of 32 / G pixels
  // (G = 1: one sample of an 8x4 pixel tile, slot = k * per_sample + tile * 32 + lane; 4: a 4x2 block; 8: 2x2; 16: 2x1;
  // 32: one pixel).  Camera rays of one pixel differ by the sub-pixel jitter only, so a warp's rays -- and the bounce
  // and shadow rays made from them -- start closer together.  Which (pixel, sample) a slot holds never changes a path's
  // arithmetic; only slot_to_sample / sample_to_slot know the layout.
  uint32_t sample_group;
};

// block of pixels a warp covers for G samples per warp: width and height as shifts
__device__ __forceinline__ void group_block(uint32_t G, uint32_t& bw_s, uint32_t& bh_s)
{
  bw_s = G >= 32u ? 0u : (G >= 8u ? 1u : (G >= 4u ? 2u : 3u));     // 8, 4, 2, 2, 1 pixels wide for G = 1, 4, 8, 16, 32
  bh_s = G >= 16u ? 0u : (G >= 4u ? 1u : 2u);                      // 4, 2, 2, 1, 1 pixels high
}

__device__ __forceinline__ void slot_to_sample(const DeviceParams& P, uint32_t slot, uint32_t per_sample, uint32_t& px, uint32_t& py, uint32_t& k)
{
  const uint32_t G = P.sample_group;
  if (G > 1u) {
    uint32_t bw_s, bh_s;
    group_block(G, bw_s, bh_s);
    const uint32_t kq = slot / (G * per_sample), in = slot - kq * G * per_sample;
    const uint32_t tile = P.tile0 + in / (32u * G), w = (in >> 5) % G, l = in & 31u;
    const uint32_t ppb_s = bw_s + bh_s;                       // log2(pixels per block) = log2(32 / G)
    const uint32_t pb = l & ((1u << ppb_s) - 1u), ks = l >> ppb_s;
    const uint32_t cols_s = 3u - bw_s;                        // blocks per tile row = 8 / bw
    const uint32_t bx = w & ((1u << cols_s) - 1u), by = w >> cols_s;
    px = (tile % P.tiles_x) * 8u + (bx << bw_s) + (pb & ((1u << bw_s) - 1u));
    py = (tile / P.tiles_x) * 4u + (by << bh_s) + (pb >> bw_s);
    k = kq * G + ks;
  } else {
    k = slot / per_sample;
    const uint32_t in = slot - k * per_sample;
    const uint32_t tile = P.tile0 + (in >> 5), lane = in & 31u;
    px = (tile % P.tiles_x) * 8u + (lane & 7u);
    py = (tile / P.tiles_x) * 4u + (lane >> 3);
  }
}

// slot of sample k of the pixel with tile-order index `in` (tile * 32 + 8x4 lane) of this part of the frame
__device__ __forceinline__ size_t sample_to_slot(const DeviceParams& P, uint32_t in, uint32_t k, uint32_t per_sample)
{
  const uint32_t G = P.sample_group;
  if (G <= 1u) return (size_t)k * per_sample + in;
  uint32_t bw_s, bh_s;
  group_block(G, bw_s, bh_s);
  const uint32_t lane = in & 31u, lx8 = lane & 7u, ly4 = lane >> 3;
  const uint32_t bx = lx8 >> bw_s, lx = lx8 & ((1u << bw_s) - 1u), by = ly4 >> bh_s, ly = ly4 & ((1u << bh_s) - 1u);
  const uint32_t w = (by << (3u - bw_s)) + bx, pb = (ly << bw_s) + lx;
  const uint32_t l = ((k % G) << (bw_s + bh_s)) + pb;
  return (size_t)(k / G) * G * per_sample + (size_t)(in >> 5) * 32u * G + w * 32u + l;
}

struct Counters {   // mirrors crt_stats
  unsigned long long rays_nearest, rays_any, n_inner, n_leaf, n_tri, n_switch, shaded_hits, samples;
  unsigned long long n_inner_any, n_leaf_any, n_tri_any, n_switch_any;
  unsigned long long n_boxes, n_boxes_any;
};

struct PathState {
  float4* ray_o;      // o.xyz, w = implicit pdf of the direction (for MIS)
  float4* ray_d;      // d.xyz, w = bits: bit0 = inside medium
  float4* thr;        // throughput.xyz, w = bits(rng state)
  float4* rad;        // radiance.xyz
  float4* hit;        // t, u, v, bits(triangle slot, -1 = miss)
  int32_t* hit_inst;
  uint32_t* queue[2]; // compacted active path slots, ping-pong by depth parity
  float4* sh_o;       // shadow ray origin.xyz, w = tmax
  float4* sh_d;       // shadow ray dir.xyz, w = bits(path slot)
  float4* sh_c;       // throughput * contribution
  uint32_t* n_active; // [max_depth + 1]
  uint32_t* n_shadow; // [max_depth]
  uint32_t* work_extend;   // [max_depth] persistent-kernel work counters
  uint32_t* work_connect;  // [max_depth]
};

// Path state is streamed once per kernel (gigabytes per wave); the BVH and triangles are re-read by
// every ray.  State accesses use the streaming cache operators (ld.global.cs / st.global.cs: evict
// first) so they do not displace scene data from L1/L2.  CRT_STREAMING=0 builds the plain variant.
#ifndef CRT_STREAMING
#define CRT_STREAMING 1
#endif
template <typename T> __device__ __forceinline__ T ld_stream(const T* p) { return CRT_STREAMING ? __ldcs(p) : *p; }
template <typename T> __device__ __forceinline__ void st_stream(T* p, const T& v) { if (CRT_STREAMING) __stcs(p, v); else *p = v; }

// 64-byte records (BVH node, instance) are fetched with two 256-bit loads (LDG.E.256, new on
// sm_100): half the L1 tag lookups of four 128-bit loads for the same bytes.  CRT_LDG256=0 = 4 x 128.
#ifndef CRT_LDG256
#define CRT_LDG256 1
#endif
__device__ __forceinline__ void ld_record64(const float4* p, float4& a, float4& b, float4& c, float4& d)
{
#if CRT_LDG256
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w), "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + 2));
#else
  a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
#endif
}

__device__ __forceinline__ void ld_triangle(const float4* p, float4& a, float4& b, float4& c)
{
#if CRT_LDG256 && CRT_TRI_STRIDE == 4
  float4 pad;
  ld_record64(p, a, b, c, pad);
#else
  a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2);
#endif
}

// ------------------------------------------------------------------ traversal

struct Hit { float t, u, v; int32_t tri; int32_t inst; };

__device__ __forceinline__ float inv_dir(float d)
{
  float a = 1.0f / maxf(fabsf(d), 8.271806125530277e-25f);
  return d < 0.0f ? -a : a;
}

struct Ray {
  v3 o, d, inv, oinv;
  __device__ __forceinline__ void setup(v3 o_, v3 d_)
  {
    o = o_; d = d_;
    inv = V(inv_dir(d.x), inv_dir(d.y), inv_dir(d.z));
    oinv = V(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
  }
};

// IntersectTriangle, SURVEY A.4 (same operation order as the oracle's tri_test).
__device__ __forceinline__ bool tri_test(v3 o, v3 d, v3 p0, v3 p1, v3 p2, float& t, float& u, float& v)
{
  v3 e0 = vsub(p1, p0);
  v3 e1 = vsub(p0, p2);
  v3 nn = cross3(e1, e0);
  v3 to = vsub(p0, o);
  float rcp = 1.0f / dot3(nn, d);
  float tt = dot3(nn, to) * rcp;
  v3 k = cross3(d, to);
  float uu = dot3(k, e1) * rcp;
  float vv = dot3(k, e0) * rcp;
  if (tt >= 0.0f && uu >= 0.0f && vv >= 0.0f && uu + vv <= 1.0f) { t = tt; u = uu; v = vv; return true; }
  return false;
}

__device__ __forceinline__ v3 xf_point(float4 r0, float4 r1, float4 r2, v3 p)
{
  return V(fmaf(r0.z, p.z, fmaf(r0.y, p.y, r0.x * p.x)) + r0.w,
           fmaf(r1.z, p.z, fmaf(r1.y, p.y, r1.x * p.x)) + r1.w,
           fmaf(r2.z, p.z, fmaf(r2.y, p.y, r2.x * p.x)) + r2.w);
}
__device__ __forceinline__ v3 xf_vector(float4 r0, float4 r1, float4 r2, v3 p)
{
  return V(fmaf(r0.z, p.z, fmaf(r0.y, p.y, r0.x * p.x)),
           fmaf(r1.z, p.z, fmaf(r1.y, p.y, r1.x * p.x)),
           fmaf(r2.z, p.z, fmaf(r2.y, p.y, r2.x * p.x)));
}

// SceneNearestHit / SceneAnyHit over the 64-byte two-child nodes.
// One thread per ray, stack of child references in local memory.  References:
// >= 0 inner node, bit31 set = leaf (bit30 set = instance, else first triangle),
// kDone = traversal finished.
//
// Control flow is the "while-while" form: an inner loop that only walks inner nodes,
// then one leaf / instance step, with no `continue` across the two.  The loops are
// structured so the compiler can place a reconvergence point after the inner loop;
// a flat loop with `continue` in its branches leaves warps permanently fragmented
// (measured: 5 of 32 lanes active per instruction, profiles/r01_*).
constexpr int32_t kDone = -1;

// Traversal stack storage.  Local memory (one private array per thread) costs up to 32 L1 wavefronts per
// push/pop when the lanes of a warp sit at different depths; CRT_SMEM_STACK = D > 0 keeps D levels per
// thread in shared memory laid out [level][thread] (bank = lane: always one wavefront).
#ifndef CRT_SMEM_STACK
#define CRT_SMEM_STACK 0
#endif
struct LocalStack {
  int32_t* base;
  __device__ __forceinline__ int32_t& operator[](int i) const { return base[i]; }
};
struct SharedStack {
  int32_t* base;       // &smem[threadIdx.x]; levels 0 .. CRT_SMEM_STACK-1
  int32_t* overflow;   // deeper levels spill to a private local array (rare: 2-level trees of depth > 28)
  __device__ __forceinline__ int32_t& operator[](int i) const
  {
    return i < CRT_SMEM_STACK ? base[i * CRT_TRACE_BLOCK] : overflow[i - CRT_SMEM_STACK];
  }
};

// The pop of a lane that missed both children of a node is predicated into the node loop's common path instead of
// being its own divergent region (13 % of the warp instructions of the bounce launches ran there with 3.5 lanes):
// traversal 16.23 -> 15.92 ms per step.  Only leaving an instance stays a branch.  (The lockstep walk of the camera rays
// keeps the branch: its lanes mostly pop together, and the predicated form measured 16.05 ms.)
#ifndef CRT_FLAT_POP
#define CRT_FLAT_POP 1
#endif
#ifndef CRT_KEEP_WORLD_RAY
#define CRT_KEEP_WORLD_RAY 0
#endif
template <class Stack>
__device__ __forceinline__ int32_t stack_pop(const Stack& stack, int& sp, Ray& r, v3 org, v3 dir)
{
  if (sp == 0) return kDone;
  int32_t c = stack[--sp];
  if (c == kSentinel) {          // leaving an instance: back to the world-space ray
    r.setup(org, dir);           // (recomputed: cheaper than six more live registers, measured)
    if (sp == 0) return kDone;
    c = stack[--sp];
  }
  return c;
}

template <bool ANY, bool COUNT>
__device__ __forceinline__ bool traverse(const DeviceScene& S, v3 org, v3 dir, float tmax, Hit& hit, Counters& cnt)
{
  hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.tri = -1; hit.inst = -1;
  int32_t cur = S.top_root;
  if (cur == kNoRef) cur = kDone;
  if (!(dot3(dir, dir) > 0.0f) || !(dot3(org, org) >= 0.0f)) cur = kDone;
  int32_t stack[kStackSize];
  int sp = 0;
  int32_t inst = -1;
  Ray r;
  r.setup(org, dir);
  bool found = false;
  while (cur != kDone) {
    // ---- inner nodes: test both children, descend into the nearer one, push the farther
    while (cur >= 0) {
      if (COUNT) { if (ANY) { cnt.n_inner_any++; cnt.n_boxes_any += 2; } else { cnt.n_inner++; cnt.n_boxes += 2; } }
      const float4* nd = S.nodes + 4 * (size_t)cur;
      float4 n0, n1, n2, n3;
        ld_record64(nd, n0, n1, n2, n3);
      const float c0x0 = fmaf(n0.x, r.inv.x, r.oinv.x), c0x1 = fmaf(n0.y, r.inv.x, r.oinv.x);
      const float c0y0 = fmaf(n0.z, r.inv.y, r.oinv.y), c0y1 = fmaf(n0.w, r.inv.y, r.oinv.y);
      const float c0z0 = fmaf(n2.x, r.inv.z, r.oinv.z), c0z1 = fmaf(n2.y, r.inv.z, r.oinv.z);
      const float c1x0 = fmaf(n1.x, r.inv.x, r.oinv.x), c1x1 = fmaf(n1.y, r.inv.x, r.oinv.x);
      const float c1y0 = fmaf(n1.z, r.inv.y, r.oinv.y), c1y1 = fmaf(n1.w, r.inv.y, r.oinv.y);
      const float c1z0 = fmaf(n2.z, r.inv.z, r.oinv.z), c1z1 = fmaf(n2.w, r.inv.z, r.oinv.z);
      const float te0 = fmaxf(fmaxf(fminf(c0x0, c0x1), fminf(c0y0, c0y1)), fminf(c0z0, c0z1));
      const float tx0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fmaxf(c0z0, c0z1));
      const float te1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fminf(c1z0, c1z1));
      const float tx1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fmaxf(c1z0, c1z1));
      const bool h0 = fmaxf(te0, 0.0f) <= fminf(tx0, hit.t);
      const bool h1 = fmaxf(te1, 0.0f) <= fminf(tx1, hit.t);
      const int32_t r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
      if (h0 | h1) {
        const bool swap = h1 && (!h0 || te1 < te0);   // go to child 1 first
        cur = swap ? r1 : r0;
        if (h0 & h1) stack[sp++] = swap ? r0 : r1;
      } else {
        cur = stack_pop(stack, sp, r, org, dir);
      }
    }
    if (cur == kDone) break;
    if ((uint32_t)cur & 0x40000000u) {
      // ---- top-level leaf: enter the instance (ray to object space, not renormalised)
      if (COUNT) { if (ANY) cnt.n_switch_any++; else cnt.n_switch++; }
      inst = (int32_t)((uint32_t)cur & 0x3fffffffu);
      const float4* ir = S.inst + 4 * (size_t)inst;
      float4 m0, m1, m2, m3;
        ld_record64(ir, m0, m1, m2, m3);
      r.setup(xf_point(m0, m1, m2, org), xf_vector(m0, m1, m2, dir));
      stack[sp++] = kSentinel;
      cur = __float_as_int(m3.x);
    } else {
      // ---- bottom-level leaf: triangles until the "last" flag
      if (COUNT) { if (ANY) cnt.n_leaf_any++; else cnt.n_leaf++; }
      uint32_t tri = (uint32_t)cur & 0x3fffffffu;
      bool more = true;
      while (more) {
        if (COUNT) { if (ANY) cnt.n_tri_any++; else cnt.n_tri++; }
        const float4* tv = S.tri_verts + kTriStride * (size_t)tri;
        float4 a, b, c;
        ld_triangle(tv, a, b, c);
        float t, u, v;
        more = __float_as_int(b.w) == 0;
        if (tri_test(r.o, r.d, V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), t, u, v) && t < hit.t) {
          hit.t = t; hit.u = u; hit.v = v; hit.tri = (int32_t)tri; hit.inst = inst;
          found = true;
          if (ANY) more = false;
        }
        ++tri;
      }
      cur = (ANY && found) ? kDone : stack_pop(stack, sp, r, org, dir);
    }
  }
  return found;
}

// Persistent-thread form of the same traversal: every lane owns one ray at a time and
// takes the next one from a warp-local pool the moment its ray finishes, instead of
// idling until the slowest ray of the warp is done.  The pool is refilled kChunk rays
// at a time with one atomicAdd per warp on a per-launch work counter; lanes pick from
// it with a ballot prefix (no further atomics).  Per outer iteration a lane does:
// retire/refill -> walk inner nodes until a leaf reference -> one leaf or instance step.
// Identical arithmetic and visiting order per ray as traverse<>, so results and work
// counters are unchanged.  Policy supplies load(index) / store(token, hit, found).
#ifndef CRT_CHUNK
#define CRT_CHUNK 64
#endif
// Lanes leave the loop that walks inner nodes once fewer than CRT_INNER_EXIT lanes of the warp are still walking while
// another lane waits for its leaf / instance step (1 = walk until every lane is done).  With OCCT's leaves of 5
// triangles this lost (round 1: +-0 ... -10 %); with leaves of 2 the step the waiting lanes get to sooner is short, and
// it pays: lanes per instruction 10.1 -> 13.6, warp instructions -20 %, traversal 17.9 -> 16.5 ms per step together
// with 8 instead of 9 resident CTAs (N = 4 / 6 / 8 / 10 / 12 / 16: 17.19 / 17.00 / 16.95 / 17.00 / 17.19 / 17.81 ms at 9 CTAs).
#ifndef CRT_INNER_EXIT
#define CRT_INNER_EXIT 8
#endif
#ifndef CRT_PREFETCH
#define CRT_PREFETCH 0
#endif

// Finished lanes of the persistent driver retire and take their next ray together, once CRT_REFILL_MIN of them wait
// (or no lane of the warp has work left): 1 / 2 / 4 / 8 / 12 / 16 / 24 -> 16.55 / 16.37 / 16.34 / 16.23 / 16.30 / 16.54 /
// 17.90 ms of traversal per step (C2).
#ifndef CRT_REFILL_MIN
#define CRT_REFILL_MIN 8
#endif
#ifndef CRT_SMEM_TOP
#define CRT_SMEM_TOP 0
#endif
// Resident CTAs per SM the traversal kernels are compiled for.  Round 1 (ms of traversal per step, C2 / C5 flattened):
// 7 CTAs (72 registers) 20.33 / 29.64, 8 (64) 19.78 / 28.02, 9 (56) 19.64 / 27.20, 10 (48) 20.29 / 27.61, 12 (40)
// 21.28 / 27.82.  With the node-loop exit threshold the walk issues 20 % fewer instructions and leans on the L1 data
// pipe instead (84 % busy); the spills of the 56-register build then cost more than the ninth CTA hides:
// 7 / 8 / 9 / 10 CTAs: 17.14 / 16.47 / 16.95 / 17.41 ms (C2).
#ifndef CRT_TRACE_MIN_BLOCKS
#define CRT_TRACE_MIN_BLOCKS 8
#endif
constexpr uint32_t kChunk = CRT_CHUNK;

// MODE 0: closest hit for every ray, 1: any hit for every ray, 2: per ray (Policy::load says which;
// used by the fused "connect(d) + extend(d+1)" launch).
template <int MODE, bool COUNT, bool QUAD, class Policy>
__device__ __forceinline__ void trace_persistent(const DeviceScene& S, uint32_t n, uint32_t* work, Counters& cnt,
                                                 const Policy& pol)
{
  const unsigned FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t pool_next = 0, pool_end = 0;   // warp-uniform
  bool drained = false;                   // warp-uniform: the launch's queue is exhausted
  // rays a warp takes per atomic: kChunk while there is plenty of work; one ray per lane when the launch has fewer
  // rays than that for every warp (late bounces, small frames), so that the work spreads over all warps instead of
  // a few warps walking two rounds of long rays while the rest of the GPU idles
  const uint32_t chunk = (n >= gridDim.x * (blockDim.x >> 5) * kChunk) ? kChunk : 32u;
  int32_t cur = kDone;
  bool has_ray = false, found = false, any_ray = (MODE == 1);
  uint32_t token = 0;
  v3 org = V(0, 0, 0), dir = V(0, 0, 1);
  Ray r;
  r.setup(org, dir);
  Hit hit;
  hit.t = 0.0f; hit.u = 0.0f; hit.v = 0.0f; hit.tri = -1; hit.inst = -1;
#if CRT_SMEM_STACK > 0
  __shared__ int32_t s_stack[CRT_SMEM_STACK * CRT_TRACE_BLOCK];
  int32_t stack_overflow[kStackSize - CRT_SMEM_STACK];
  const SharedStack stack{ s_stack + threadIdx.x, stack_overflow };
#else
  int32_t stack_mem[QUAD ? kStackSizeQuad : kStackSize];
  const LocalStack stack{ stack_mem };
#endif
#if CRT_SMEM_TOP > 0
  // stage the top of the top-level tree in shared memory: one cp.async.bulk (TMA, UBLKCP) per CTA,
  // completion through an mbarrier transaction count
  __shared__ __align__(128) float4 s_top[CRT_SMEM_TOP * 5];
  __shared__ __align__(8) unsigned long long s_bar;
  const int32_t n_cache = (int32_t)min(S.n_top_cache, (uint32_t)CRT_SMEM_TOP);
  if (n_cache > 0) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_top);
    const uint32_t bytes = (uint32_t)n_cache * 80u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(S.top_cache), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t landed = 0;
    while (!landed) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(landed) : "r"(bar), "r"(0) : "memory");
    }
  }
#endif
  int sp = 0;
  int32_t inst = -1;
  for (;;) {
    // ---- retire finished rays, hand out new ones
    const bool need = cur == kDone;
    const unsigned m = __ballot_sync(FULL, need);
#if CRT_REFILL_MIN > 1
    // finished lanes wait until CRT_REFILL_MIN of them can retire and refill together (or no lane has work left):
    // fewer, fuller passes through the retire / refill code and its dependent state loads
    const bool do_refill = m != 0u && ((int)__popc(m) >= CRT_REFILL_MIN || m == FULL);
#else
    const bool do_refill = m != 0u;
#endif
    if (do_refill && cur == kDone && has_ray) { pol.store(token, hit, found, any_ray); has_ray = false; }
    if (do_refill) {
      if (pool_next == pool_end && !drained) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(work, chunk);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) drained = true;
        else { pool_next = base; pool_end = min(base + chunk, n); }
      }
      const uint32_t avail = pool_end - pool_next;
      const uint32_t rank = __popc(m & ((1u << lane) - 1u));
      if (need && rank < avail) {
        float tmax;
        bool a = (MODE == 1);
        token = pol.load(pool_next + rank, org, dir, tmax, a);
        if (MODE == 2) any_ray = a;
        hit.t = tmax; hit.u = 0.0f; hit.v = 0.0f; hit.tri = -1; hit.inst = -1;
        found = false; sp = 0; inst = -1;
        cur = S.top_root;
        if (cur == kNoRef) cur = kDone;
        if (!(dot3(dir, dir) > 0.0f) || !(dot3(org, org) >= 0.0f)) cur = kDone;
        r.setup(org, dir);
        has_ray = true;
        if (COUNT) { if (any_ray) cnt.rays_any++; else cnt.rays_nearest++; }
      }
      pool_next += min(avail, (uint32_t)__popc(m));
    }
    const unsigned live = __ballot_sync(FULL, has_ray);
    if (drained && live == 0) break;
#if CRT_REFILL_MIN > 1
    const int n_live = __popc(__ballot_sync(FULL, cur != kDone));    // lanes with work (finished lanes may be waiting to retire)
#else
    const int n_live = __popc(live);
#endif

    // ---- inner nodes.  Lanes leave the loop once fewer than CRT_INNER_EXIT lanes are still walking inner nodes
    // while another lane of the warp waits for its leaf / instance step (measurements at the macro's definition).
    while (cur >= 0) {
      if (!QUAD) {
        if (COUNT) { if (any_ray) { cnt.n_inner_any++; cnt.n_boxes_any += 2; } else { cnt.n_inner++; cnt.n_boxes += 2; } }
        const float4* nd = S.nodes + 4 * (size_t)cur;
        float4 n0, n1, n2, n3;
        #if CRT_SMEM_TOP > 0
        if (cur < n_cache) {
          const float4* sn = s_top + 5 * cur;
          n0 = sn[0]; n1 = sn[1]; n2 = sn[2]; n3 = sn[3];
        } else
#endif
        ld_record64(nd, n0, n1, n2, n3);
        const float c0x0 = fmaf(n0.x, r.inv.x, r.oinv.x), c0x1 = fmaf(n0.y, r.inv.x, r.oinv.x);
        const float c0y0 = fmaf(n0.z, r.inv.y, r.oinv.y), c0y1 = fmaf(n0.w, r.inv.y, r.oinv.y);
        const float c0z0 = fmaf(n2.x, r.inv.z, r.oinv.z), c0z1 = fmaf(n2.y, r.inv.z, r.oinv.z);
        const float c1x0 = fmaf(n1.x, r.inv.x, r.oinv.x), c1x1 = fmaf(n1.y, r.inv.x, r.oinv.x);
        const float c1y0 = fmaf(n1.z, r.inv.y, r.oinv.y), c1y1 = fmaf(n1.w, r.inv.y, r.oinv.y);
        const float c1z0 = fmaf(n2.z, r.inv.z, r.oinv.z), c1z1 = fmaf(n2.w, r.inv.z, r.oinv.z);
        const float te0 = fmaxf(fmaxf(fminf(c0x0, c0x1), fminf(c0y0, c0y1)), fminf(c0z0, c0z1));
        const float tx0 = fminf(fminf(fmaxf(c0x0, c0x1), fmaxf(c0y0, c0y1)), fmaxf(c0z0, c0z1));
        const float te1 = fmaxf(fmaxf(fminf(c1x0, c1x1), fminf(c1y0, c1y1)), fminf(c1z0, c1z1));
        const float tx1 = fminf(fminf(fmaxf(c1x0, c1x1), fmaxf(c1y0, c1y1)), fmaxf(c1z0, c1z1));
        const bool h0 = fmaxf(te0, 0.0f) <= fminf(tx0, hit.t);
        const bool h1 = fmaxf(te1, 0.0f) <= fminf(tx1, hit.t);
        const int32_t r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
#if CRT_PREFETCH
        // both children's records are requested while this node's slab tests run
        if (r0 >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.nodes + 4 * (size_t)r0));
        if (r1 >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.nodes + 4 * (size_t)r1));
#endif
#if CRT_FLAT_POP
        {   // the pop of a lane that missed both children is predicated into the common path (a short divergent region per
            // trip costs more issue slots than a few masked instructions); only leaving an instance stays a branch
          const bool swap = h1 && (!h0 || te1 < te0);
          const bool any = h0 | h1;
          if (h0 & h1) stack[sp++] = swap ? r0 : r1;
          int32_t nxt = swap ? r1 : r0;
          const bool do_pop = !any && sp > 0;
          if (!any) nxt = kDone;
          if (do_pop) nxt = stack[--sp];
          cur = nxt;
          if (cur == kSentinel) {
            r.setup(org, dir);
            cur = kDone;
            if (sp > 0) cur = stack[--sp];
          }
        }
#else
        if (h0 | h1) {
          const bool swap = h1 && (!h0 || te1 < te0);
          cur = swap ? r1 : r0;
          if (h0 & h1) stack[sp++] = swap ? r0 : r1;
        } else {
          cur = stack_pop(stack, sp, r, org, dir);
        }
#endif
      }
      else {
        // QUAD_BVH: 128-byte node = 4 child boxes + 4 references; all children tested, sorted by entry
        // distance with the 5-comparator network (0,1)(2,3)(0,2)(1,3)(1,2), pushed far-to-near
        const float4* nd = S.nodes + 8 * (size_t)cur;
        float4 a0, a1, a2, a3, a4, a5, rf, pd;
        ld_record64(nd, a0, a1, a2, a3);
        ld_record64(nd + 4, a4, a5, rf, pd);
        const float b[24] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w,
                              a3.x, a3.y, a3.z, a3.w, a4.x, a4.y, a4.z, a4.w, a5.x, a5.y, a5.z, a5.w };
        const int32_t ref[4] = { __float_as_int(rf.x), __float_as_int(rf.y), __float_as_int(rf.z), __float_as_int(rf.w) };
        float te[4];
        int32_t id[4];
        int nbox = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float x0 = fmaf(b[6 * c + 0], r.inv.x, r.oinv.x), x1 = fmaf(b[6 * c + 3], r.inv.x, r.oinv.x);
          const float y0 = fmaf(b[6 * c + 1], r.inv.y, r.oinv.y), y1 = fmaf(b[6 * c + 4], r.inv.y, r.oinv.y);
          const float z0 = fmaf(b[6 * c + 2], r.inv.z, r.oinv.z), z1 = fmaf(b[6 * c + 5], r.inv.z, r.oinv.z);
          const float e = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
          const float x = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
          const bool present = ref[c] != kNoRef;
          const bool h = present && fmaxf(e, 0.0f) <= fminf(x, hit.t);
          te[c] = h ? e : 3.0e38f;
          id[c] = h ? ref[c] : kNoRef;
          nbox += present ? 1 : 0;
        }
        if (COUNT) { if (any_ray) { cnt.n_inner_any++; cnt.n_boxes_any += nbox; } else { cnt.n_inner++; cnt.n_boxes += nbox; } }
#define CRT_CSWAP(i, j) { const bool sw = te[j] < te[i]; const float tf = sw ? te[j] : te[i]; te[j] = sw ? te[i] : te[j]; te[i] = tf; \
                          const int32_t ti = sw ? id[j] : id[i]; id[j] = sw ? id[i] : id[j]; id[i] = ti; }
        CRT_CSWAP(0, 1) CRT_CSWAP(2, 3) CRT_CSWAP(0, 2) CRT_CSWAP(1, 3) CRT_CSWAP(1, 2)
#undef CRT_CSWAP
        if (id[3] != kNoRef) stack[sp++] = id[3];
        if (id[2] != kNoRef) stack[sp++] = id[2];
        if (id[1] != kNoRef) stack[sp++] = id[1];
        cur = id[0] != kNoRef ? id[0] : stack_pop(stack, sp, r, org, dir);
      }
#if CRT_INNER_EXIT > 1
      {
        const int walking = __popc(__activemask());
        if (walking < CRT_INNER_EXIT && walking < n_live) break;
      }
#endif
    }
    // ---- one leaf / instance step
    if (cur < 0 && cur != kDone) {
      if ((uint32_t)cur & 0x40000000u) {
        if (COUNT) { if (any_ray) cnt.n_switch_any++; else cnt.n_switch++; }
        inst = (int32_t)((uint32_t)cur & 0x3fffffffu);
        const float4* ir = S.inst + 4 * (size_t)inst;
        float4 m0, m1, m2, m3;
        ld_record64(ir, m0, m1, m2, m3);
        r.setup(xf_point(m0, m1, m2, org), xf_vector(m0, m1, m2, dir));
        stack[sp++] = kSentinel;
        cur = __float_as_int(m3.x);
      } else {
        if (COUNT) { if (any_ray) cnt.n_leaf_any++; else cnt.n_leaf++; }
        uint32_t tri = (uint32_t)cur & 0x3fffffffu;
        bool more = true;
        while (more) {
          if (COUNT) { if (any_ray) cnt.n_tri_any++; else cnt.n_tri++; }
          const float4* tv = S.tri_verts + kTriStride * (size_t)tri;
          float4 a, b, c;
          ld_triangle(tv, a, b, c);
          float t, u, v;
          more = __float_as_int(b.w) == 0;
          if (tri_test(r.o, r.d, V(a.x, a.y, a.z), V(b.x, b.y, b.z), V(c.x, c.y, c.z), t, u, v) && t < hit.t) {
            hit.t = t; hit.u = u; hit.v = v; hit.tri = (int32_t)tri; hit.inst = inst;
            found = true;
            if (any_ray) more = false;
          }
          ++tri;
        }
        cur = (any_ray && found) ? kDone : stack_pop(stack, sp, r, org, dir);
      }
    }
  }
}

// ------------------------------------------------------------------ BSDF

struct Bsdf {
  v3 Kc; float Kc_w;
  v3 Kd;
  v3 Ks; float Ks_w;
  v3 Kt;
  v3 Fc, Fb;
};

__device__ __forceinline__ float fresnel_dielectric4(float cos_i, float cos_t, float eta_i, float eta_t)
{
  float parl = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t);
  float perp = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t);
  return (parl * parl + perp * perp) * 0.5f;
}

__device__ __forceinline__ float fresnel_dielectric(float cos_i, float index)
{
  float eta_i = cos_i > 0.0f ? 1.0f : index;
  float eta_t = cos_i > 0.0f ? index : 1.0f;
  float sin_t2 = (eta_i * eta_i) / (eta_t * eta_t) * (1.0f - cos_i * cos_i);
  if (sin_t2 < 1.0f) return fresnel_dielectric4(fabsf(cos_i), sqrtf(1.0f - sin_t2), eta_i, eta_t);
  return 1.0f;
}

__device__ __forceinline__ float fresnel_conductor(float cos_i, float eta, float k)
{
  float tmp = 2.0f * eta * cos_i;
  float tmp1 = eta * eta + k * k;
  float s_perp = (tmp1 - tmp + cos_i * cos_i) / (tmp1 + tmp + cos_i * cos_i);
  float tmp2 = tmp1 * cos_i * cos_i;
  float s_parl = (tmp2 - tmp + 1.0f) / (tmp2 + tmp + 1.0f);
  return (s_perp + s_parl) * 0.5f;
}

// Graphic3d_Fresnel::Serialize() encoding (MaterialEditor.cxx:209-255).
// conductor / dielectric interfaces out of line (OL: the instantiations of the shading kernels for scenes without coat
// and transmission -- fresnel_media is inlined at eight places, the two exact models are most of its code and such
// scenes rarely use them: k_shade -3.7 % on C2; scenes with glass keep them inline, where the call costs 2 %)
__device__ __noinline__ float fresnel_exact(float cos_i, float fx, float fy, float fz)
{
  if (fx > -2.5f) return fresnel_conductor(fabsf(cos_i), fy, fz);
  return fresnel_dielectric(cos_i, fy);
}

template <bool OL = false>
__device__ __forceinline__ v3 fresnel_media(float cos_i, v3 f)
{
  if (f.x > -0.5f) {
    float m = 1.0f - fabsf(cos_i);
    float m2 = m * m;
    float m5 = m2 * m2 * m;
    return V(f.x + (1.0f - f.x) * m5, f.y + (1.0f - f.y) * m5, f.z + (1.0f - f.z) * m5);
  }
  if (f.x > -1.5f) return V(f.z, f.z, f.z);
  if (OL) { const float c = fresnel_exact(cos_i, f.x, f.y, f.z); return V(c, c, c); }
  if (f.x > -2.5f) { float c = fresnel_conductor(fabsf(cos_i), f.y, f.z); return V(c, c, c); }
  float c = fresnel_dielectric(cos_i, f.y);
  return V(c, c, c);
}

__device__ __forceinline__ float ggx_d(float mz, float a)
{
  float a2 = a * a;
  float q = fmaf(mz * mz, a2 - 1.0f, 1.0f);
  return a2 / (CRT_PI * q * q);
}

__device__ __forceinline__ float smith_g1(v3 dir, v3 m, float a)
{
  if (dot3(dir, m) * dir.z <= 0.0f) return 0.0f;
  float c2 = dir.z * dir.z;
  float tan2 = (1.0f - c2) / c2;
  return 2.0f / (1.0f + sqrtf(fmaf(a * a, tan2, 1.0f)));
}

template <bool OL = false>
__device__ __forceinline__ v3 eval_ggx(v3 wi, v3 wo, v3 fresnel, float a)
{
  if (wi.z <= 0.0f || wo.z <= 0.0f) return V(0, 0, 0);
  v3 h = normalize3(vadd(wi, wo));
  float d = ggx_d(h.z, a);
  float g = smith_g1(wo, h, a) * smith_g1(wi, h, a);
  return vscale(fresnel_media<OL>(dot3(wo, h), fresnel), d * g / (4.0f * wo.z));
}

__device__ __forceinline__ float eval_lambert(v3 wi, v3 wo)
{
  return (wi.z <= 0.0f || wo.z <= 0.0f) ? 0.0f : wi.z * CRT_INV_PI;
}

template <bool OL = false>
__device__ __forceinline__ v3 eval_bsdf_layered(const Bsdf& b, v3 wi, v3 wo, bool two_sided)
{
  if (two_sided) { wi.z = fabsf(wi.z); wo.z = fabsf(wo.z); }
  v3 r = vscale(b.Kd, eval_lambert(wi, wo));
  if (b.Ks_w > CRT_FLT_EPS) r = vadd(r, vmul(b.Ks, eval_ggx<OL>(wi, wo, b.Fb, b.Ks_w)));
  v3 cf = fresnel_media<OL>(wo.z, b.Fc);
  r = vmul(r, V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z));
  if (b.Kc_w > CRT_FLT_EPS) r = vadd(r, vmul(b.Kc, eval_ggx<OL>(wi, wo, b.Fc, b.Kc_w)));
  return r;
}

__device__ __forceinline__ float ggx_pdf_term(float hz, float a, float wi_dot_h)
{
  return ggx_d(hz, a) * fabsf(hz) * 0.25f / wi_dot_h;
}

template <bool OL = false>
__device__ __forceinline__ float bsdf_pdf_layered(const Bsdf& b, v3 wo, v3 wi, v3 weight)
{
  v3 cf = fresnel_media<OL>(wo.z, b.Fc);
  v3 ct = V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z);
  float pc = dot3(vmul(b.Kc, cf), weight);
  float pd = dot3(vmul(b.Kd, ct), weight);
  float ps = dot3(vmul(b.Ks, ct), weight);
  float pt = dot3(vmul(b.Kt, ct), weight);
  float pdf = 0.0f;
  if (wi.z * wo.z > 0.0f) {
    v3 h = normalize3(vadd(wi, wo));
    float wh = dot3(wi, h);
    pdf = pd * fabsf(wi.z * CRT_INV_PI);
    if (b.Kc_w > CRT_FLT_EPS) pdf += pc * ggx_pdf_term(h.z, b.Kc_w, wh);
    if (b.Ks_w > CRT_FLT_EPS) pdf += ps * ggx_pdf_term(h.z, b.Ks_w, wh);
  }
  return pdf / ((pc + pd) + (ps + pt));
}

__device__ __forceinline__ v3 sample_lambert(v3 wo, v3& wi, float& pdf, uint32_t& rng, bool two_sided)
{
  float k1 = rand_float(rng);
  float k2 = rand_float(rng);
  float sn, cs;
  sincos2pi(k1, sn, cs);
  float r = sqrtf(k2);
  v3 w = V(cs * r, sn * r, sqrtf(1.0f - k2));
  if (two_sided && wo.z < 0.0f) w.z = -w.z;
  wi = w;
  pdf *= fabsf(w.z) * CRT_INV_PI;
  if (two_sided) return V(1, 1, 1);
  return wo.z >= 0.0f ? V(1, 1, 1) : V(0, 0, 0);
}

template <bool OL = false>
__device__ __forceinline__ v3 sample_ggx(v3 wo, v3& wi, v3 fresnel, float a, float& pdf, uint32_t& rng, bool two_sided)
{
  float k1 = rand_float(rng);
  float k2 = rand_float(rng);
  float tan2 = a * a * k1 / (1.0f - k1);
  float cos_m = 1.0f / sqrtf(1.0f + tan2);
  float sin_m = sqrtf(maxf(1.0f - cos_m * cos_m, 0.0f));
  float sn, cs;
  sincos2pi(k2, sn, cs);
  v3 m = V(cs * sin_m, sn * sin_m, cos_m);
  pdf *= ggx_d(cos_m, a) * cos_m;
  bool flip = two_sided && wo.z < 0.0f;
  if (flip) wo.z = -wo.z;
  float cos_d = dot3(wo, m);
  v3 w = V(fmaf(2.0f * cos_d, m.x, -wo.x), fmaf(2.0f * cos_d, m.y, -wo.y), fmaf(2.0f * cos_d, m.z, -wo.z));
  wi = w;
  if (w.z <= 0.0f || wo.z <= 0.0f) return V(0, 0, 0);
  pdf /= 4.0f * cos_d;
  float g = smith_g1(wo, m, a) * smith_g1(w, m, a);
  if (flip) wi.z = -w.z;
  return vscale(fresnel_media<OL>(cos_d, fresnel), (g * cos_d) / (wo.z * cos_m));
}

__device__ __forceinline__ v3 transmitted(float index, v3 wo)
{
  float eta = wo.z > 0.0f ? 1.0f / index : index;
  float sin_t2 = eta * eta * (1.0f - wo.z * wo.z);
  float cos_t = sqrtf(1.0f - minf(sin_t2, 1.0f));
  if (wo.z > 0.0f) cos_t = -cos_t;
  return normalize3(V(-eta * wo.x, -eta * wo.y, cos_t));
}

// LEAN: the scene's material table has no coat and no transmission (Kc = Kt = 0 for every material), so those
// lobes can never be chosen; the instantiation drops their code.  pc and pt are still computed (they are +0) so
// that every remaining value is the same bit pattern as in the full version.
template <bool LEAN>
__device__ __forceinline__ float sample_bsdf_layered(const Bsdf& b, v3 wo, v3& wi, v3& weight, bool& inside,
                                                     uint32_t& rng, bool two_sided)
{
  float pdf = 0.0f;
  v3 cf = fresnel_media<LEAN>(wo.z, b.Fc);
  v3 ct = V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z);
  float pc = dot3(vmul(b.Kc, cf), weight);
  float pd = dot3(vmul(b.Kd, ct), weight);
  float ps = dot3(vmul(b.Ks, ct), weight);
  float pt = dot3(vmul(b.Kt, ct), weight);
  float total = (pc + pd) + (ps + pt);
  float ksi = total * rand_float(rng);
  wi = V(0, 0, 1);
  if (!LEAN && ksi < pc) {
    pdf = pc / total;
    weight = vmul(weight, vscale(b.Kc, 1.0f / pdf));
    if (b.Kc_w < CRT_FLT_EPS) {
      weight = vmul(weight, cf);
      wi = V(-wo.x, -wo.y, wo.z);
      pdf = CRT_MAXFLOAT;
    } else {
      weight = vmul(weight, sample_ggx<LEAN>(wo, wi, b.Fc, b.Kc_w, pdf, rng, two_sided));
    }
  } else if (ksi < total) {
    weight = vmul(weight, ct);
    if (ksi < pc + pd) {
      pdf = pd / total;
      weight = vmul(weight, vscale(b.Kd, 1.0f / pdf));
      weight = vmul(weight, sample_lambert(wo, wi, pdf, rng, two_sided));
    } else if (LEAN || ksi < (pc + pd) + ps) {
      pdf = ps / total;
      weight = vmul(weight, vscale(b.Ks, 1.0f / pdf));
      if (b.Ks_w < CRT_FLT_EPS) {
        weight = vmul(weight, fresnel_media<LEAN>(wo.z, b.Fb));
        wi = V(-wo.x, -wo.y, wo.z);
        pdf = CRT_MAXFLOAT;
      } else {
        weight = vmul(weight, sample_ggx<LEAN>(wo, wi, b.Fb, b.Ks_w, pdf, rng, two_sided));
      }
    } else {
      pdf = pt / total;
      weight = vmul(weight, vscale(b.Kt, 1.0f / pdf));
      float index = b.Fc.x > -2.5f ? 1.0f : b.Fc.y;
      wi = transmitted(index, wo);
      inside = !inside;
      pdf = CRT_MAXFLOAT;
    }
  }
  if (!(total >= CRT_FLT_EPS)) weight = V(0, 0, 0);
  return pdf;
}

// ------------------------------------------------------------------ frame, lights, env

struct Frame { v3 x, y, z; };

__device__ __forceinline__ Frame build_frame(v3 n)
{
  v3 ax = V(n.z, 0.0f, -n.x);
  v3 ay = V(0.0f, -n.z, n.y);
  float lx = dot3(ax, ax), ly = dot3(ay, ay);
  Frame f;
  if (lx > ly) {
    ax = vscale(ax, 1.0f / sqrtf(lx));
    ay = cross3(ax, n);
  } else {
    ay = vscale(ay, 1.0f / sqrtf(ly));
    ax = cross3(ay, n);
  }
  f.x = ax; f.y = ay; f.z = n;
  return f;
}
__device__ __forceinline__ v3 to_local(v3 v, const Frame& f) { return V(dot3(v, f.x), dot3(v, f.y), dot3(v, f.z)); }
__device__ __forceinline__ v3 from_local(v3 v, const Frame& f)
{
  return vadd(vadd(vscale(f.x, v.x), vscale(f.y, v.y)), vscale(f.z, v.z));
}

__device__ __forceinline__ float cone_pdf(float cos_max) { return 1.0f / (CRT_2PI - cos_max * CRT_2PI); }

__device__ __forceinline__ v3 ld_rgb(const float4* p, int i) { float4 t = __ldg(p + i); return V(t.x, t.y, t.z); }

__device__ __forceinline__ v3 env_lookup(const DeviceScene& S, v3 d)
{
  if (S.env_w == 0) return V(0, 0, 0);
  float u = (atan2_poly(d.y, d.x) + CRT_PI) * CRT_INV_2PI;
  float v = acos_poly(d.z) * CRT_INV_PI;
  float fx = fmaf(u, (float)S.env_w, -0.5f);
  float fy = fmaf(v, (float)S.env_h, -0.5f);
  float flx = floorf(fx), fly = floorf(fy);
  float ax = fx - flx, ay = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int w = (int)S.env_w, h = (int)S.env_h;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % w) + w) % w; x1 = ((x1 % w) + w) % w;
  y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0);
  y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
  v3 c00 = ld_rgb(S.env, y0 * w + x0), c10 = ld_rgb(S.env, y0 * w + x1);
  v3 c01 = ld_rgb(S.env, y1 * w + x0), c11 = ld_rgb(S.env, y1 * w + x1);
  v3 top = vadd(vscale(c00, 1.0f - ax), vscale(c10, ax));
  v3 bot = vadd(vscale(c01, 1.0f - ax), vscale(c11, ax));
  return vadd(vscale(top, 1.0f - ay), vscale(bot, ay));
}

// textureLod(sampler, st, 0) with GL_LINEAR / GL_REPEAT, t = 0 at the bottom row of the image file.
__device__ __forceinline__ float4 tex_lookup(const DeviceScene& S, uint32_t tex, float u, float v)
{
  const uint32_t off = __ldg(S.tex_table + 3 * tex);
  const int w = (int)__ldg(S.tex_table + 3 * tex + 1), h = (int)__ldg(S.tex_table + 3 * tex + 2);
  const float fx = fmaf(u, (float)w, -0.5f);
  const float fy = fmaf(1.0f - v, (float)h, -0.5f);
  const float flx = floorf(fx), fly = floorf(fy);
  const float ax = fx - flx, ay = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % w) + w) % w; x1 = ((x1 % w) + w) % w;
  y0 = ((y0 % h) + h) % h; y1 = ((y1 % h) + h) % h;
  const uchar4* t = S.tex_data + off;
  const uchar4 q00 = __ldg(t + y0 * w + x0), q10 = __ldg(t + y0 * w + x1);
  const uchar4 q01 = __ldg(t + y1 * w + x0), q11 = __ldg(t + y1 * w + x1);
  const float k = 1.0f / 255.0f;
  float4 r;
#define CRT_BILERP(f) (((float)q00.f * k * (1.0f - ax) + (float)q10.f * k * ax) * (1.0f - ay) + \
                       ((float)q01.f * k * (1.0f - ax) + (float)q11.f * k * ax) * ay)
  r.x = CRT_BILERP(x); r.y = CRT_BILERP(y); r.z = CRT_BILERP(z); r.w = CRT_BILERP(w);
#undef CRT_BILERP
  return r;
}

// IntersectLight, SURVEY A.7.
__device__ __forceinline__ v3 intersect_light(const DeviceScene& S, const DeviceParams& P, v3 o, v3 d, int depth,
                                              float hit_dist, float& pdf_out)
{
  v3 total = V(0, 0, 0);
  float pdf = 0.0f;
  const float inv_n = S.n_lights ? 1.0f / (float)S.n_lights : 0.0f;
  const bool miss = hit_dist == CRT_MAXFLOAT;
  for (uint32_t i = 0; i < S.n_lights; ++i) {
    const float4 e = __ldg(S.lights + 2 * i), p = __ldg(S.lights + 2 * i + 1);
    if (p.w != 0.0f) {
      v3 to = vsub(V(p.x, p.y, p.z), o);
      float dist = sqrtf(dot3(to, to));
      if (dist < hit_dist) {
        float cos_max = 1.0f / sqrtf(1.0f + (e.w * e.w) / (dist * dist));
        if (cos_max < 1.0f && dot3(d, vscale(to, 1.0f / dist)) >= cos_max) {
          hit_dist = dist;
          total = V(e.x, e.y, e.z);
          pdf = inv_n * cone_pdf(cos_max);
        }
      }
    } else if (hit_dist == CRT_MAXFLOAT) {
      if (e.w < 1.0f && dot3(d, V(p.x, p.y, p.z)) >= e.w) {
        total = vadd(total, V(e.x, e.y, e.z));
        pdf += inv_n * cone_pdf(e.w);
      }
    }
  }
  if (pdf == 0.0f && miss && hit_dist == CRT_MAXFLOAT) {
    if (depth == 0 && !P.env_as_background) total = V(P.background[0], P.background[1], P.background[2]);
    else total = env_lookup(S, d);
  }
  pdf_out = pdf;
  return total;
}

// SampleLight, SURVEY A.7.
__device__ __forceinline__ v3 sample_light(v3 to_light, float dist, bool infinite, float smooth, float& pdf, uint32_t& rng)
{
  Frame f = build_frame(vscale(to_light, 1.0f / dist));
  float cos_max = infinite ? smooth : 1.0f / sqrtf(1.0f + (smooth * smooth) / (dist * dist));
  float k1 = rand_float(rng);
  float k2 = rand_float(rng);
  float tz = 1.0f - k2 * (1.0f - cos_max);
  float sn, cs;
  sincos2pi(k1, sn, cs);
  float r = sqrtf(maxf(1.0f - tz * tz, 0.0f));
  pdf = (cos_max < 1.0f) ? pdf * cone_pdf(cos_max) : CRT_MAXFLOAT;
  return normalize3(from_local(V(cs * r, sn * r, tz), f));
}

// ------------------------------------------------------------------ queue helpers

// Appends one entry per lane with pred set; one atomicAdd per warp.  Must be
// called by all 32 lanes.
__device__ __forceinline__ uint32_t warp_push(uint32_t* counter, bool pred)
{
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0 && mask) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void flush_counters(Counters* g, const Counters& c)
{
  // warp-reduce then one atomic per warp and field
  unsigned long long v[14] = { c.rays_nearest, c.rays_any, c.n_inner, c.n_leaf, c.n_tri, c.n_switch, c.shaded_hits, c.samples,
                               c.n_inner_any, c.n_leaf_any, c.n_tri_any, c.n_switch_any, c.n_boxes, c.n_boxes_any };
  unsigned long long* out = reinterpret_cast<unsigned long long*>(g);
#pragma unroll
  for (int k = 0; k < 14; ++k) {
    unsigned long long x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, x);
  }
}

// ------------------------------------------------------------------ kernels

// GenerateRay: camera ray of one sample of pixel (px, py); rng is the generator state after the
// pixel jitter (and lens) draws.
__device__ __forceinline__ void camera_ray(const DeviceParams& P, uint32_t px, uint32_t py, uint32_t frame_seed,
                                           v3& o, v3& d, uint32_t& rng)
{
  rng = seed_rand(frame_seed, px, py, P.width, P.rng_radius);
  const float jx = rand_float(rng);
  const float jy = rand_float(rng);
  float la = 0.0f, lb = 0.0f;
  if (P.aperture_radius > 0.0f) { la = rand_float(rng); lb = rand_float(rng); }
  const float fx = ((float)px + jx) / (float)P.width;
  const float fy = ((float)py + jy) / (float)P.height;
  const float sx = fmaf(fx, 2.0f, -1.0f) * P.hw;
  const float sy = fmaf(fy, 2.0f, -1.0f) * P.hh;
  const v3 eye = V(P.eye[0], P.eye[1], P.eye[2]);
  const v3 cu = V(P.cu[0], P.cu[1], P.cu[2]), cv = V(P.cv[0], P.cv[1], P.cv[2]), cw = V(P.cw[0], P.cw[1], P.cw[2]);
  if (P.is_ortho) {
    o = vadd(eye, vadd(vscale(cu, sx), vscale(cv, sy)));
    d = cw;
  } else {
    o = eye;
    d = normalize3(vadd(cw, vadd(vscale(cu, sx), vscale(cv, sy))));
  }
  if (P.aperture_radius > 0.0f) {
    const float ft = P.focal_dist / dot3(d, cw);
    const v3 focus = vadd(o, vscale(d, ft));
    float sn, cs;
    sincos2pi(lb, sn, cs);
    const float r = sqrtf(la) * P.aperture_radius;
    o = vadd(o, vadd(vscale(cu, r * cs), vscale(cv, r * sn)));
    d = normalize3(vsub(focus, o));
  }
}

// the camera ray of one sample of pixel (px, py) written into path slot `slot`
__device__ __forceinline__ void generate_path(const PathState& st, const DeviceParams& P, uint32_t slot, uint32_t px, uint32_t py,
                                              uint32_t frame_seed)
{
  v3 o, d;
  uint32_t rng;
  camera_ray(P, px, py, frame_seed, o, d, rng);
  st_stream(&st.ray_o[slot], make_float4(o.x, o.y, o.z, 1.0f));
  st_stream(&st.ray_d[slot], make_float4(d.x, d.y, d.z, __int_as_float(0)));
  st_stream(&st.thr[slot], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng)));
  st_stream(&st.rad[slot], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
}

// GenerateRay for every path slot of the batch.  Slot layout: sample k
// of the batch occupies slots [k*tiles*32, (k+1)*tiles*32); inside, each warp owns
// an 8x4 pixel tile so primary rays of a warp are coherent.
__global__ void __launch_bounds__(256)
k_generate(PathState st, DeviceParams P, const uint32_t* __restrict__ frame_seeds, uint32_t n_batch)
{
  const uint32_t per_sample = P.n_tiles * 32u;
  const uint32_t total = per_sample * n_batch;
  const uint32_t stride = gridDim.x * blockDim.x;
  const bool aligned = (P.width & 7u) == 0 && (P.height & 3u) == 0;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < total; base += stride) {
    const uint32_t slot = base + (threadIdx.x & 31u);
    uint32_t px, py, k;
    slot_to_sample(P, slot, per_sample, px, py, k);
    const bool valid = slot < total && px < P.width && py < P.height;
    if (valid) generate_path(st, P, slot, px, py, __ldg(frame_seeds + k));
    if (aligned) {
      st.queue[0][slot] = slot;             // every slot is a pixel: identity queue, no atomics
    } else {
      const uint32_t at = warp_push(st.n_active, valid);
      if (valid) st.queue[0][at] = slot;
    }
  }
  if (aligned && blockIdx.x == 0 && threadIdx.x == 0) st.n_active[0] = total;
}

struct ExtendPolicy {
  PathState st;
  const uint32_t* __restrict__ q;
  __device__ __forceinline__ uint32_t load(uint32_t i, v3& o, v3& d, float& tmax, bool&) const
  {
    const uint32_t slot = ld_stream(&q[i]);
    const float4 ro = ld_stream(&st.ray_o[slot]), rd = ld_stream(&st.ray_d[slot]);
    o = V(ro.x, ro.y, ro.z); d = V(rd.x, rd.y, rd.z); tmax = CRT_MAXFLOAT;
    return slot;
  }
  __device__ __forceinline__ void store(uint32_t slot, const Hit& hit, bool, bool) const
  {
    st_stream(&st.hit[slot], make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri)));
    st_stream(&st.hit_inst[slot], (int32_t)hit.inst);
  }
};

// SceneNearestHit for every active path.  PERSISTENT selects the per-lane-refill driver
// (grid = resident CTAs) or the static one-ray-per-loop-iteration form (kept for A/B).
template <bool COUNT, bool PERSISTENT, bool QUAD>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK, CRT_TRACE_MIN_BLOCKS)
k_extend(DeviceScene S, PathState st, int depth, Counters* gcnt)
{
  const uint32_t n = st.n_active[depth];
  const uint32_t* __restrict__ q = st.queue[depth & 1];
  Counters cnt = {};
  if (PERSISTENT) {
    ExtendPolicy pol{ st, q };
    trace_persistent<0, COUNT, QUAD>(S, n, st.work_extend + depth, cnt, pol);
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const uint32_t slot = ld_stream(&q[i]);
      const float4 o = ld_stream(&st.ray_o[slot]), d = ld_stream(&st.ray_d[slot]);
      Hit hit;
      traverse<false, COUNT>(S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), CRT_MAXFLOAT, hit, cnt);
      if (COUNT) cnt.rays_nearest++;
      st_stream(&st.hit[slot], make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri)));
      st_stream(&st.hit_inst[slot], (int32_t)hit.inst);
    }
  }
  if (COUNT) flush_counters(gcnt, cnt);
}

// Depth 0 without a generate pass: the camera ray of slot i is a pure function of (pixel, frame seed), so the
// traversal kernel computes it when a lane takes slot i, and k_shade<.., FIRST> computes it again instead of
// reading four float4 of path state that would only hold (ray, 1, 0).  Saves the k_generate launch and about
// 160 bytes of HBM traffic per path.  Requires a tile-aligned resolution (every slot is a pixel).
struct PrimaryPolicy {
  PathState st;
  const DeviceParams& P;
  const uint32_t* __restrict__ seeds;
  uint32_t per_sample;
  __device__ __forceinline__ uint32_t load(uint32_t slot, v3& o, v3& d, float& tmax, bool&) const
  {
    uint32_t px, py, k, rng;
    slot_to_sample(P, slot, per_sample, px, py, k);
    camera_ray(P, px, py, __ldg(seeds + k), o, d, rng);
    tmax = CRT_MAXFLOAT;
    return slot;
  }
  __device__ __forceinline__ void store(uint32_t slot, const Hit& hit, bool, bool) const
  {
    st_stream(&st.hit[slot], make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri)));
    st_stream(&st.hit_inst[slot], (int32_t)hit.inst);
  }
};

template <bool COUNT, bool QUAD>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK, CRT_TRACE_MIN_BLOCKS)
k_extend_primary(DeviceScene S, PathState st, DeviceParams P, const uint32_t* __restrict__ seeds, uint32_t n_batch, Counters* gcnt)
{
  const uint32_t per_sample = P.n_tiles * 32u;
  const uint32_t n = per_sample * n_batch;
  Counters cnt = {};
  PrimaryPolicy pol{ st, P, seeds, per_sample };
  trace_persistent<0, COUNT, QUAD>(S, n, st.work_extend, cnt, pol);
  if (blockIdx.x == 0 && threadIdx.x == 0) st.n_active[0] = n;
  if (COUNT) flush_counters(gcnt, cnt);
}

// The same for the binary tree in lockstep: a warp takes the 32 slots of one 8x4 pixel tile and walks them together
// (no per-lane refill).  Camera rays of a tile visit nearly the same nodes, so staying in step keeps most lanes active
// (the persistent driver mixes rays at different stages in a warp: 14.8 of 32 lanes, issue slots 78 % busy).
#ifndef CRT_PRIMARY_MIN_BLOCKS
#define CRT_PRIMARY_MIN_BLOCKS 9      // no per-lane refill state to keep: 56 registers with 20 B of spills (3.13 ms per step at 9 CTAs, 3.25 at 8)
#endif
template <bool COUNT>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK, CRT_PRIMARY_MIN_BLOCKS)
k_extend_primary_lockstep(DeviceScene S, PathState st, DeviceParams P, const uint32_t* __restrict__ seeds, uint32_t n_batch, Counters* gcnt)
{
  const uint32_t per_sample = P.n_tiles * 32u;
  const uint32_t n = per_sample * n_batch;
  Counters cnt = {};
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) {
    uint32_t px, py, k;
    slot_to_sample(P, slot, per_sample, px, py, k);
    v3 o, d;
    uint32_t rng;
    camera_ray(P, px, py, __ldg(seeds + k), o, d, rng);
    Hit hit;
    traverse<false, COUNT>(S, o, d, CRT_MAXFLOAT, hit, cnt);
    if (COUNT) cnt.rays_nearest++;
    st_stream(&st.hit[slot], make_float4(hit.t, hit.u, hit.v, __int_as_float(hit.tri)));
    st_stream(&st.hit_inst[slot], (int32_t)hit.inst);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) st.n_active[0] = n;
  if (COUNT) flush_counters(gcnt, cnt);
}

// One bounce of PathTrace for one path whose closest hit is `hh` (t, u, v, triangle slot):
// implicit light / environment hit with MIS, emission, next-event estimation (fills the shadow ray), Beer-Lambert
// absorption, layered-BSDF sampling, termination / Russian roulette, continuation ray.  Shared by the wavefront
// kernel k_shade and by the per-path kernel k_tail, so both run the same arithmetic in the same order.
// The instance id of the hit is read from *inst_src when given (wavefront: st.hit_inst), else it is inst_reg.
template <bool COUNT, bool TEX, bool LEAN>
__device__ __forceinline__ void shade_bounce(const DeviceScene& S, const DeviceParams& P, const int depth, const bool two_sided,
                                             const float eps, const float4 hh, const int32_t* inst_src, const int32_t inst_reg,
                                             v3& org, v3& dir, v3& thr, float& imp_pdf, uint32_t& rng, bool& inside, v3& radiance,
                                             bool& want_shadow, v3& sh_o, v3& sh_d, v3& sh_c, float& sh_tmax, bool& want_next,
                                             Counters& cnt, const float4* smem_mats = nullptr)
{
  const int32_t tri = __float_as_int(hh.w);
  const bool found = tri >= 0;

  float exp_pdf;
  const v3 le = intersect_light(S, P, org, dir, depth, hh.x, exp_pdf);
  if (any_gt(le, 0.0f) || !found) {
    const float mis = (depth == 0 || imp_pdf == CRT_MAXFLOAT) ? 1.0f
                    : imp_pdf * imp_pdf / (exp_pdf * exp_pdf + imp_pdf * imp_pdf);
    radiance = vadd(radiance, vscale(vmul(thr, le), mis));
  } else {
    const int32_t inst = inst_src ? ld_stream(inst_src) : inst_reg;
    const float4* ir = S.inst + 4 * (size_t)inst;
    float4 m0, m1, m2, m3;
    ld_record64(ir, m0, m1, m2, m3);
    const v3 c0 = V(m0.x, m1.x, m2.x), c1 = V(m0.y, m1.y, m2.y), c2 = V(m0.z, m1.z, m2.z);
    // geometric normal from the stored vertices (same expression as tri_test's nn)
    const float4* tv = S.tri_verts + kTriStride * (size_t)tri;
    float4 a, b, c;
    ld_triangle(tv, a, b, c);
    const v3 p0 = V(a.x, a.y, a.z), p1 = V(b.x, b.y, b.z), p2 = V(c.x, c.y, c.z);
    const v3 nraw = cross3(vsub(p0, p2), vsub(p1, p0));
    const v3 ng = normalize3(V(dot3(c0, nraw), dot3(c1, nraw), dot3(c2, nraw)));
    org = vadd(org, vscale(dir, hh.x));
    // SmoothNormal, SURVEY A.4
    const float4* tn = S.tri_nrm + 3 * (size_t)tri;
    const v3 n0 = ld_rgb(tn, 0), n1 = ld_rgb(tn, 1), n2 = ld_rgb(tn, 2);
    v3 ns = vadd(vadd(vscale(n1, hh.y), vscale(n2, hh.z)), vscale(n0, (1.0f - hh.y) - hh.z));
    ns = normalize3(ns);
    ns = normalize3(V(dot3(c0, ns), dot3(c1, ns), dot3(c2, ns)));
    const Frame frame = build_frame(ns);

    const uint32_t mat_id = (uint32_t)__float_as_int(m3.y);
    Bsdf B;
    v3 mat_le, absorp;
    float absorp_k, kd_w = 0.0f, kt_w = 0.0f, le_w = 0.0f;   // texture id + 1, S scale, T scale
    if (mat_id < S.n_mats) {
      float4 kc, kd, ks, kt, le4, fc, fb, ab;
      if (smem_mats) {   // warp-uniform: the whole table is staged in shared memory (small scenes)
        const float4* mp = smem_mats + 8 * (size_t)mat_id;
        kc = mp[0]; kd = mp[1]; ks = mp[2]; kt = mp[3]; le4 = mp[4]; fc = mp[5]; fb = mp[6]; ab = mp[7];
      } else {
        const float4* mp = S.mats + 8 * (size_t)mat_id;
        kc = __ldg(mp); kd = __ldg(mp + 1); ks = __ldg(mp + 2); kt = __ldg(mp + 3);
        le4 = __ldg(mp + 4); fc = __ldg(mp + 5); fb = __ldg(mp + 6); ab = __ldg(mp + 7);
      }
      B.Kc = V(kc.x, kc.y, kc.z); B.Kc_w = kc.w;
      B.Kd = V(kd.x, kd.y, kd.z);
      B.Ks = V(ks.x, ks.y, ks.z); B.Ks_w = ks.w;
      B.Kt = V(kt.x, kt.y, kt.z);
      B.Fc = V(fc.x, fc.y, fc.z); B.Fb = V(fb.x, fb.y, fb.z);
      mat_le = V(le4.x, le4.y, le4.z);
      absorp = V(ab.x, ab.y, ab.z); absorp_k = ab.w;
      kd_w = kd.w; kt_w = kt.w; le_w = le4.w;
    } else {   // default grey diffuse (same record as the oracle's k_default_bsdf)
      B.Kc = V(0, 0, 0); B.Kc_w = 0.0f; B.Kd = V(0.8f, 0.8f, 0.8f); B.Ks = V(0, 0, 0); B.Ks_w = 0.0f;
      B.Kt = V(0, 0, 0); B.Fc = V(-1.0f, 0.0f, 0.0f); B.Fb = V(-1.0f, 0.0f, 1.0f);
      mat_le = V(0, 0, 0); absorp = V(0, 0, 0); absorp_k = 0.0f;
    }
    if (COUNT) cnt.shaded_hits++;

    // base-colour texture (USE_TEXTURES path of PathTrace; SmoothUV)
    if (TEX) {   // instantiated only for scenes that have textures
      if (kd_w >= 1.0f && (uint32_t)kd_w - 1u < S.n_tex) {
        const float2* tu = S.tri_uv + 3 * (size_t)tri;
        const float2 uv0 = __ldg(tu), uv1 = __ldg(tu + 1), uv2 = __ldg(tu + 2);
        const float w0 = (1.0f - hh.y) - hh.z;
        const float su = (uv1.x * hh.y + uv2.x * hh.z) + uv0.x * w0;
        const float sv = (uv1.y * hh.y + uv2.y * hh.z) + uv0.y * w0;
        const float ss = kt_w != 0.0f ? kt_w : 1.0f, ts = le_w != 0.0f ? le_w : 1.0f;
        const float4 tc = tex_lookup(S, (uint32_t)kd_w - 1u, su * ss, sv * ts);
        B.Kd = vmul(B.Kd, vscale(V(tc.x * tc.x, tc.y * tc.y, tc.z * tc.z), tc.w));
        if (tc.w != 1.0f) {
          const float ia = 1.0f - tc.w;
          B.Kt = V(ia + tc.w * B.Kt.x, ia + tc.w * B.Kt.y, ia + tc.w * B.Kt.z);
        }
      }
    }

    const v3 wo = to_local(V(-dir.x, -dir.y, -dir.z), frame);
    if (LEAN) {   // host guarantee: no coat, no transmission anywhere in the material table (and so never inside a medium)
      B.Kc = V(0, 0, 0); B.Kc_w = 0.0f; B.Kt = V(0, 0, 0);
      inside = false;
    }
    radiance = vadd(radiance, vmul(thr, mat_le));

    const v3 nee_k = vadd(B.Kd, vadd(B.Ks_w > CRT_FLT_EPS ? B.Ks : V(0, 0, 0), B.Kc_w > CRT_FLT_EPS ? B.Kc : V(0, 0, 0)));
    if (S.n_lights > 0 && dot3(nee_k, thr) > 0.0f) {
      exp_pdf = 1.0f / (float)S.n_lights;
      int li = (int)(rand_float(rng) * (float)S.n_lights);
      if (li > (int)S.n_lights - 1) li = (int)S.n_lights - 1;
      const float4 le_w = __ldg(S.lights + 2 * li), lp = __ldg(S.lights + 2 * li + 1);
      const bool infinite = lp.w == 0.0f;
      const v3 to = infinite ? V(lp.x, lp.y, lp.z) : vsub(V(lp.x, lp.y, lp.z), org);
      const float dist = sqrtf(dot3(to, to));
      const v3 ldir = sample_light(to, dist, infinite, le_w.w, exp_pdf, rng);
      const v3 wl = to_local(ldir, frame);
      const float bpdf = bsdf_pdf_layered<LEAN>(B, wo, wl, thr);
      imp_pdf = bpdf;
      const float mis = (exp_pdf == CRT_MAXFLOAT) ? 1.0f : exp_pdf / (exp_pdf * exp_pdf + bpdf * bpdf);
      const v3 contrib = vscale(vmul(V(le_w.x, le_w.y, le_w.z), eval_bsdf_layered<LEAN>(B, wl, wo, two_sided)), mis);
      if (any_gt(contrib, CRT_MIN_CONTRIBUTION)) {
        const float side = dot3(ng, ldir) >= 0.0f ? eps : -eps;
        sh_o = vadd(vadd(org, vscale(ldir, eps)), vscale(ng, side));
        sh_d = ldir;
        sh_tmax = infinite ? CRT_MAXFLOAT : dist;
        sh_c = vmul(thr, contrib);
        want_shadow = true;
      }
    }

    if (inside) {
      thr = vmul(thr, V(exp_poly(-hh.x * absorp_k * (1.0f - absorp.x)),
                        exp_poly(-hh.x * absorp_k * (1.0f - absorp.y)),
                        exp_poly(-hh.x * absorp_k * (1.0f - absorp.z))));
    }

    v3 wi;
    imp_pdf = sample_bsdf_layered<LEAN>(B, wo, wi, thr, inside, rng, two_sided);

    float survive = any_gt(thr, CRT_MIN_THROUGHPUT) ? 1.0f : 0.0f;
    const bool rr_on = P.russian_roulette && depth >= 3;
    if (rr_on) survive = minf(fmaf(0.0722f, thr.z, fmaf(0.7152f, thr.y, 0.2126f * thr.x)), 0.95f);
    const bool dead = rand_float(rng) > survive || all_lt(thr, CRT_MIN_THROUGHPUT);
    if (!dead && depth + 1 < P.max_depth) {
      if (rr_on) thr = vscale(thr, 1.0f / survive);
      dir = normalize3(from_local(wi, frame));
      const float side = dot3(ng, dir) >= 0.0f ? eps : -eps;
      org = vadd(vadd(org, vscale(dir, eps)), vscale(ng, side));
      want_next = true;
    }
  }
}

// One bounce of PathTrace for every active path: implicit
// light / environment hit with MIS, emission, next-event estimation (emits a
// shadow ray), Beer-Lambert absorption, layered-BSDF sampling, termination /
// Russian roulette, continuation ray.
#ifndef CRT_SHADE_MIN_BLOCKS
#define CRT_SHADE_MIN_BLOCKS 8
#endif
// SORT (bounces after the first): a CTA takes kShadeBatch queue entries at a time and files them in shared memory by
// shading class -- diffuse only / diffuse + glossy / coated / transmissive, taken from the material of the hit
// (DeviceScene::mat_class), and "missed" last -- each class starting on a warp boundary, then shades the list in that
// order, so that the lanes of a warp run the same lobes of the layered BSDF.  Without a class table the classes
// are just "hit" and "missed".  Only the order changes.
#ifndef CRT_SHADE_BATCH
#define CRT_SHADE_BATCH 512
#endif
constexpr uint32_t kShadeBatch = CRT_SHADE_BATCH;
constexpr int kClassMiss = 0, kClassDiffuse = 1, kClassGlossy = 2, kClassCoat = 3, kClassTransmissive = 4, kNumClasses = 5;
constexpr uint32_t kShadeListSize = kShadeBatch + 32u * (kNumClasses - 1);
// material table in shared memory (north_star: "the material table staged in shared memory"): tables of up to
// kSmemMats records (4 KB) are copied once per CTA; larger ones (C2: 1001 materials = 125 KB) stay in L1 / L2
constexpr uint32_t kSmemMats = 32;
template <bool COUNT, bool TEX, bool FIRST, bool SORT, bool LEAN, bool CLASSES>
__global__ void __launch_bounds__(128, CRT_SHADE_MIN_BLOCKS)
k_shade(DeviceScene S, DeviceParams P, PathState st, int depth, Counters* gcnt, const uint32_t* __restrict__ seeds, uint32_t tail_max)
{
  const uint32_t n = st.n_active[depth];
  if (!FIRST && n <= tail_max) return;     // k_tail(depth) has carried these paths to their end (same test there)
  const uint32_t* __restrict__ q = st.queue[depth & 1];
  uint32_t* __restrict__ qn = st.queue[(depth + 1) & 1];
  Counters cnt = {};
  const bool two_sided = P.two_sided != 0;
  const float eps = S.scene_eps;
  __shared__ uint32_t s_list[SORT ? (CLASSES ? kShadeListSize : kShadeBatch) : 1];
  __shared__ uint32_t s_tmp[SORT && CLASSES ? kShadeBatch : 1];
  __shared__ uint32_t s_cnt[kNumClasses], s_cur[kNumClasses], s_off[kNumClasses + 1];
  __shared__ float4 s_mats[LEAN ? 1 : 8 * kSmemMats];
  const float4* smem_mats = nullptr;
  if (!LEAN && S.mats_in_smem) {
    for (uint32_t k = threadIdx.x; k < 8u * S.n_mats; k += blockDim.x) s_mats[k] = __ldg(S.mats + k);
    __syncthreads();
    smem_mats = s_mats;
  }
  const uint32_t batch = SORT ? kShadeBatch : 128u;
  for (uint32_t bbase = blockIdx.x * batch; bbase < n; bbase += gridDim.x * batch) {
   uint32_t n_iter = 1, n_hit = 0, n_hit_pad = 0, n_miss = 0;
   if (SORT && CLASSES) {
     __syncthreads();                       // the previous batch's list is no longer read
     if (threadIdx.x < kNumClasses) { s_cnt[threadIdx.x] = 0; s_cur[threadIdx.x] = 0; }
     __syncthreads();
     const uint32_t lane = threadIdx.x & 31u;
     const unsigned below = (1u << lane) - 1u;
     // pass 1: slot and class of every entry of the batch (slot | class << 29 in s_tmp: slots are < 2^29), class sizes
#pragma unroll 1
     for (uint32_t k = 0; k < kShadeBatch / 128u; ++k) {
       const uint32_t e = bbase + k * 128u + threadIdx.x;
       uint32_t sl = 0;
       int cls = -1;
       if (e < n) {
         sl = ld_stream(&q[e]);
         cls = kClassMiss;
         if (__float_as_int(ld_stream(&st.hit[sl]).w) >= 0) {
           cls = kClassDiffuse;
           {
             const int32_t inst = ld_stream(&st.hit_inst[sl]);
             const uint32_t mid = (uint32_t)__float_as_int(__ldg(S.inst + 4 * (size_t)inst + 3).y);
             if (mid < S.n_mats) cls = (int)__ldg(S.mat_class + mid);
           }
         }
       }
       s_tmp[k * 128u + threadIdx.x] = cls < 0 ? 0xffffffffu : (sl | ((uint32_t)cls << 29));
#pragma unroll
       for (int c = 0; c < kNumClasses; ++c) {
         const unsigned m = __ballot_sync(0xffffffffu, cls == c);
         if (lane == 0 && m) atomicAdd(&s_cnt[c], (uint32_t)__popc(m));
       }
     }
     __syncthreads();
     if (threadIdx.x == 0) {                // class order in the list: 1, 2, 3, 4, then the misses; each on a warp boundary
       uint32_t o = 0;
       for (int c = 1; c <= kNumClasses; ++c) { const int cc = c % kNumClasses; s_off[cc] = o; o += (s_cnt[cc] + 31u) & ~31u; }
       s_off[kNumClasses] = o;
     }
     __syncthreads();
     // pass 2: scatter (each thread re-reads the entries it wrote)
#pragma unroll 1
     for (uint32_t k = 0; k < kShadeBatch / 128u; ++k) {
       const uint32_t v = s_tmp[k * 128u + threadIdx.x];
       const int cls = v == 0xffffffffu ? -1 : (int)(v >> 29);
#pragma unroll
       for (int c = 0; c < kNumClasses; ++c) {
         const unsigned m = __ballot_sync(0xffffffffu, cls == c);
         uint32_t base = 0;
         if (lane == 0 && m) base = atomicAdd(&s_cur[c], (uint32_t)__popc(m));
         base = __shfl_sync(0xffffffffu, base, 0);
         if (cls == c) s_list[s_off[c] + base + __popc(m & below)] = v & 0x1fffffffu;
       }
     }
     __syncthreads();
     n_iter = (s_off[kNumClasses] + 127u) / 128u;
   }
   if (SORT && !CLASSES) {                  // two classes, one pass: hits from the front of the list, misses from the back
     __syncthreads();
     if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
     __syncthreads();
     const uint32_t lane = threadIdx.x & 31u;
#pragma unroll 1
     for (uint32_t k = 0; k < kShadeBatch / 128u; ++k) {
       const uint32_t e = bbase + k * 128u + threadIdx.x;
       const bool ok = e < n;
       uint32_t sl = 0;
       bool found = false;
       if (ok) { sl = ld_stream(&q[e]); found = __float_as_int(ld_stream(&st.hit[sl]).w) >= 0; }
       const unsigned mh = __ballot_sync(0xffffffffu, ok && found), mm = __ballot_sync(0xffffffffu, ok && !found);
       uint32_t bh = 0, bm = 0;
       if (lane == 0) { if (mh) bh = atomicAdd(&s_cnt[0], (uint32_t)__popc(mh)); if (mm) bm = atomicAdd(&s_cnt[1], (uint32_t)__popc(mm)); }
       bh = __shfl_sync(0xffffffffu, bh, 0); bm = __shfl_sync(0xffffffffu, bm, 0);
       const unsigned below = (1u << lane) - 1u;
       if (ok && found) s_list[bh + __popc(mh & below)] = sl;
       if (ok && !found) s_list[kShadeBatch - 1u - (bm + __popc(mm & below))] = sl;
     }
     __syncthreads();
     n_hit = s_cnt[0]; n_miss = s_cnt[1];
     n_hit_pad = (n_hit + 31u) & ~31u;      // the misses start on a warp boundary
     n_iter = (n_hit_pad + n_miss + 127u) / 128u;
   }
#pragma unroll 1
   for (uint32_t it = 0; it < n_iter; ++it) {
    uint32_t i = bbase + threadIdx.x;
    bool valid = i < n;
    uint32_t slot = 0;
    if (SORT && CLASSES) {
      const uint32_t e = it * 128u + threadIdx.x;
      valid = false;
#pragma unroll
      for (int c = 0; c < kNumClasses; ++c) valid = valid || (e >= s_off[c] && e < s_off[c] + s_cnt[c]);
      if (valid) slot = s_list[e];
    }
    if (SORT && !CLASSES) {
      const uint32_t e = it * 128u + threadIdx.x;
      valid = e < n_hit || (e >= n_hit_pad && e < n_hit_pad + n_miss);
      if (valid) slot = e < n_hit ? s_list[e] : s_list[kShadeBatch - 1u - (e - n_hit_pad)];
    }
    bool want_shadow = false, want_next = false;
    v3 sh_o = V(0, 0, 0), sh_d = V(0, 0, 0), sh_c = V(0, 0, 0);
    float sh_tmax = 0.0f;
    v3 org = V(0, 0, 0), dir = V(0, 0, 1), thr = V(0, 0, 0);
    float imp_pdf = 1.0f;
    uint32_t rng = 0;
    bool inside = false;
    if (valid) {
      float4 hh;
      if (FIRST) {
        // depth 0 after k_extend_primary: slot i is pixel sample i; its state is (camera ray, throughput 1, radiance 0)
        slot = i;
        uint32_t px, py, k;
        slot_to_sample(P, slot, P.n_tiles * 32u, px, py, k);
        camera_ray(P, px, py, __ldg(seeds + k), org, dir, rng);
        thr = V(1.0f, 1.0f, 1.0f);
        hh = ld_stream(&st.hit[slot]);
      } else {
        if (!SORT) slot = ld_stream(&q[i]);
        const float4 ro = ld_stream(&st.ray_o[slot]), rd = ld_stream(&st.ray_d[slot]), tw = ld_stream(&st.thr[slot]);
        hh = ld_stream(&st.hit[slot]);
        org = V(ro.x, ro.y, ro.z); dir = V(rd.x, rd.y, rd.z);
        imp_pdf = ro.w;
        inside = (__float_as_int(rd.w) & 1) != 0;
        thr = V(tw.x, tw.y, tw.z);
        rng = __float_as_uint(tw.w);
      }
      // A bounce adds to the path's radiance at most once (a light / environment hit or the surface's emission), and
      // for most paths it adds nothing: the bounce is shaded against a zero, and the 16-byte radiance record is read
      // and rewritten only when there is something to add (x + 0 == x bit for bit, so skipping the add changes nothing).
      v3 radiance = V(0.0f, 0.0f, 0.0f);
      shade_bounce<COUNT, TEX, LEAN>(S, P, depth, two_sided, eps, hh, &st.hit_inst[slot], -1, org, dir, thr, imp_pdf, rng, inside, radiance,
                                     want_shadow, sh_o, sh_d, sh_c, sh_tmax, want_next, cnt, smem_mats);
      if (FIRST) {                               // this write initialises the slot
        st_stream(&st.rad[slot], make_float4(radiance.x, radiance.y, radiance.z, 0.0f));
      } else if (radiance.x != 0.0f || radiance.y != 0.0f || radiance.z != 0.0f) {     // also true for NaN
        float4 rr = ld_stream(&st.rad[slot]);
        rr.x += radiance.x; rr.y += radiance.y; rr.z += radiance.z;
        st_stream(&st.rad[slot], rr);
      }
    }
    // ---- warp-aggregated queue compaction
    const uint32_t sh_at = warp_push(st.n_shadow + depth, want_shadow);
    if (want_shadow) {
      st_stream(&st.sh_o[sh_at], make_float4(sh_o.x, sh_o.y, sh_o.z, sh_tmax));
      st_stream(&st.sh_d[sh_at], make_float4(sh_d.x, sh_d.y, sh_d.z, __uint_as_float(slot)));
      st_stream(&st.sh_c[sh_at], make_float4(sh_c.x, sh_c.y, sh_c.z, 0.0f));
    }
    const uint32_t nx_at = warp_push(st.n_active + depth + 1, want_next);
    if (want_next) {
      st_stream(&qn[nx_at], slot);
      st_stream(&st.ray_o[slot], make_float4(org.x, org.y, org.z, imp_pdf));
      st_stream(&st.ray_d[slot], make_float4(dir.x, dir.y, dir.z, __int_as_float(inside ? 1 : 0)));
      st_stream(&st.thr[slot], make_float4(thr.x, thr.y, thr.z, __uint_as_float(rng)));
    }
   }
  }
  if (COUNT) flush_counters(gcnt, cnt);
}

// Per-path kernel for the thin end of a wave.  After a few bounces a wave has few paths left, and every wavefront
// launch then lasts as long as its longest single ray (about 100 us of dependent node fetches, whatever the ray
// count): eight such launches are a quarter of a frame at the reference's cadence of one sample per Redraw().
// k_tail(depth) is enqueued before k_shade(depth) from the second bounce on; it does nothing while more than
// `tail_max` paths are active, and otherwise every lane takes one path and carries it to its end -- shade,
// shadow ray (any hit), continuation ray (closest hit), shade ... -- with the functions the wavefront kernels use
// (shade_bounce, traverse), in the order the wavefront applies them, so every path ends with the same bits.
// k_shade(depth) and everything after it then find nothing to do (k_shade makes the same n <= tail_max test).
#ifndef CRT_TAIL_MIN_BLOCKS
#define CRT_TAIL_MIN_BLOCKS 4
#endif
template <bool COUNT, bool TEX, bool LEAN>
__global__ void __launch_bounds__(128, CRT_TAIL_MIN_BLOCKS)
k_tail(DeviceScene S, DeviceParams P, PathState st, int depth0, uint32_t tail_max, Counters* gcnt)
{
  const uint32_t n = st.n_active[depth0];
  if (n == 0 || n > tail_max) return;
  const uint32_t* __restrict__ q = st.queue[depth0 & 1];
  Counters cnt = {};
  const bool two_sided = P.two_sided != 0;
  const float eps = S.scene_eps;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = ld_stream(&q[i]);
    const float4 ro = ld_stream(&st.ray_o[slot]), rd = ld_stream(&st.ray_d[slot]), tw = ld_stream(&st.thr[slot]);
    float4 hh = ld_stream(&st.hit[slot]);
    float4 rr = ld_stream(&st.rad[slot]);
    v3 org = V(ro.x, ro.y, ro.z), dir = V(rd.x, rd.y, rd.z), thr = V(tw.x, tw.y, tw.z);
    float imp_pdf = ro.w;
    bool inside = (__float_as_int(rd.w) & 1) != 0;
    uint32_t rng = __float_as_uint(tw.w);
    v3 radiance = V(rr.x, rr.y, rr.z);
    const int32_t* inst_src = &st.hit_inst[slot];
    int32_t inst_reg = -1;
    for (int depth = depth0;; ++depth) {
      bool want_shadow = false, want_next = false;
      v3 sh_o = V(0, 0, 0), sh_d = V(0, 0, 0), sh_c = V(0, 0, 0);
      float sh_tmax = 0.0f;
      shade_bounce<COUNT, TEX, LEAN>(S, P, depth, two_sided, eps, hh, inst_src, inst_reg, org, dir, thr, imp_pdf, rng, inside, radiance,
                                     want_shadow, sh_o, sh_d, sh_c, sh_tmax, want_next, cnt);
      if (want_shadow) {
        Hit sh;
        const bool occluded = traverse<true, COUNT>(S, sh_o, sh_d, sh_tmax, sh, cnt);
        if (COUNT) cnt.rays_any++;
        if (!occluded) { radiance.x += sh_c.x; radiance.y += sh_c.y; radiance.z += sh_c.z; }
      }
      if (!want_next) break;               // want_next implies depth + 1 < max_depth
      Hit h;
      traverse<false, COUNT>(S, org, dir, CRT_MAXFLOAT, h, cnt);
      if (COUNT) cnt.rays_nearest++;
      hh = make_float4(h.t, h.u, h.v, __int_as_float(h.tri));
      inst_src = nullptr;
      inst_reg = h.inst;
    }
    if (radiance.x != rr.x || radiance.y != rr.y || radiance.z != rr.z) {
      rr.x = radiance.x; rr.y = radiance.y; rr.z = radiance.z;
      st_stream(&st.rad[slot], rr);
    }
  }
  if (COUNT) flush_counters(gcnt, cnt);
}

struct ConnectPolicy {
  PathState st;
  __device__ __forceinline__ uint32_t load(uint32_t i, v3& o, v3& d, float& tmax, bool&) const
  {
    const float4 ro = ld_stream(&st.sh_o[i]), rd = ld_stream(&st.sh_d[i]);
    o = V(ro.x, ro.y, ro.z); d = V(rd.x, rd.y, rd.z); tmax = ro.w;
    return i;
  }
  __device__ __forceinline__ void store(uint32_t i, const Hit&, bool occluded, bool) const
  {
    if (occluded) return;
    const uint32_t slot = __float_as_uint(ld_stream(&st.sh_d[i]).w);
    const float4 c = ld_stream(&st.sh_c[i]);
    float4 r = ld_stream(&st.rad[slot]);
    r.x += c.x; r.y += c.y; r.z += c.z;
    st_stream(&st.rad[slot], r);
  }
};

// SceneAnyHit for the shadow rays of this bounce; visible => add the contribution.
template <bool COUNT, bool PERSISTENT, bool QUAD>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK, CRT_TRACE_MIN_BLOCKS)
k_connect(DeviceScene S, PathState st, int depth, Counters* gcnt)
{
  const uint32_t n = st.n_shadow[depth];
  if (n == 0) return;
  Counters cnt = {};
  if (PERSISTENT) {
    ConnectPolicy pol{ st };
    trace_persistent<1, COUNT, QUAD>(S, n, st.work_connect + depth, cnt, pol);
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const float4 o = ld_stream(&st.sh_o[i]), d = ld_stream(&st.sh_d[i]);
      Hit hit;
      const bool occluded = traverse<true, COUNT>(S, V(o.x, o.y, o.z), V(d.x, d.y, d.z), o.w, hit, cnt);
      if (COUNT) cnt.rays_any++;
      if (!occluded) {
        const uint32_t slot = __float_as_uint(d.w);
        const float4 c = ld_stream(&st.sh_c[i]);
        float4 r = ld_stream(&st.rad[slot]);
        r.x += c.x; r.y += c.y; r.z += c.z;
        st_stream(&st.rad[slot], r);
      }
    }
  }
  if (COUNT) flush_counters(gcnt, cnt);
}

// Fused launch: the shadow rays of bounce depth-1 (any hit) and the continuation rays of bounce
// depth (closest hit) are independent, so one persistent launch walks both queues; this halves the
// number of latency-bound launch tails per bounce.  Index space: [0, n_ext) continuation rays,
// [n_ext, n_ext + n_sh) shadow rays.
struct DualPolicy {
  ExtendPolicy ext;
  ConnectPolicy con;
  uint32_t n_ext;
  __device__ __forceinline__ uint32_t load(uint32_t i, v3& o, v3& d, float& tmax, bool& any) const
  {
    any = i >= n_ext;
    return any ? con.load(i - n_ext, o, d, tmax, any) : ext.load(i, o, d, tmax, any);
  }
  __device__ __forceinline__ void store(uint32_t token, const Hit& hit, bool found, bool any) const
  {
    if (any) con.store(token, hit, found, true);
    else ext.store(token, hit, found, false);
  }
};

template <bool COUNT, bool QUAD>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK, CRT_TRACE_MIN_BLOCKS)
k_trace_dual(DeviceScene S, PathState st, int depth, Counters* gcnt)
{
  const uint32_t n_ext = st.n_active[depth];
  const uint32_t n_sh = st.n_shadow[depth - 1];
  if (n_ext + n_sh == 0) return;           // the wave has ended (or k_tail has taken it over): skip the work-counter atomics
  Counters cnt = {};
  DualPolicy pol{ ExtendPolicy{ st, st.queue[depth & 1] }, ConnectPolicy{ st }, n_ext };
  trace_persistent<2, COUNT, QUAD>(S, n_ext + n_sh, st.work_extend + depth, cnt, pol);
  if (COUNT) flush_counters(gcnt, cnt);
}

// Accumulation: NaN -> 0, clamp, add the batch's samples of each pixel
// in sample order (deterministic), count in .w.
__global__ void __launch_bounds__(256)
k_resolve(PathState st, DeviceParams P, float4* __restrict__ accum, uint32_t n_batch, Counters* gcnt)
{
  const uint32_t per_sample = P.n_tiles * 32u;
  for (uint32_t in = blockIdx.x * blockDim.x + threadIdx.x; in < per_sample; in += gridDim.x * blockDim.x) {
    const uint32_t tile = P.tile0 + (in >> 5), lane = in & 31u;
    const uint32_t px = (tile % P.tiles_x) * 8u + (lane & 7u);
    const uint32_t py = (tile / P.tiles_x) * 4u + (lane >> 3);
    if (px >= P.width || py >= P.height) continue;
    float4 a = accum[(size_t)py * P.width + px];
    for (uint32_t k = 0; k < n_batch; ++k) {
      const float4 c = ld_stream(&st.rad[sample_to_slot(P, in, k, per_sample)]);
      a.x += (c.x != c.x) ? 0.0f : minf(c.x, P.max_radiance);
      a.y += (c.y != c.y) ? 0.0f : minf(c.y, P.max_radiance);
      a.z += (c.z != c.z) ? 0.0f : minf(c.z, P.max_radiance);
      a.w += 1.0f;
    }
    accum[(size_t)py * P.width + px] = a;
  }
  if (gcnt && blockIdx.x == 0 && threadIdx.x == 0 && P.tile0 == 0)      // once per wave (the part that holds tile 0)
    atomicAdd(&gcnt->samples, (unsigned long long)P.width * P.height * n_batch);
}

// ------------------------------------------------------------------ adaptive screen sampling
//
// Graphic3d_RenderingParams::AdaptiveScreenSampling (SettingsWidget.cxx:427-478; OCCT's
// OpenGl_TileSampler draws NbRayTracingTiles tiles per frame with probability proportional
// to a per-tile error estimate).  Here the same expected distribution is dealt out by
// systematic sampling in integer arithmetic, on the device, with no host round trip:
//   wave budget B tile samples; tile weight w = clamp(err, W0/(8 NT), 4 W0/NT) + 1;
//   cum[j] = floor((C_j * B + off) / Wt), C = exclusive prefix of w, off a Weyl offset < Wt;
//   tile j receives k_j = cum[j+1] - cum[j] more samples for each of its pixels.
// err[j] = sum over the tile's pixels of floor(4096 * |sqrt(min(L_all,1)) - sqrt(min(L_even,1))|),
// the display-space distance between the mean of all samples and of the even-numbered ones.
// Everything is integer or in a fixed float order, so the oracle's orc_render_adaptive
// reproduces it bit for bit.

constexpr uint32_t kAdaptiveTile = 32u;          // pixels per tile side (OCCT RayTracingTileSize)
constexpr uint32_t kAdaptiveSlots = 1024u;       // path slots of one tile sample

struct AdaptiveState {
  uint32_t* count;      // samples per pixel of each tile so far
  uint32_t* err;        // fixed-point error estimate per tile
  uint32_t* cum;        // [nt + 1] exclusive prefix of this wave's tile samples
  uint32_t* qoff;       // [nt + 1] exclusive prefix of this wave's paths (queue offsets)
  float* even;          // per pixel: luminance sum of the even-numbered samples
  uint32_t ntx, nty, nt;
  uint32_t first_parity;  // parity of the first sample index of the accumulation
  // Several GPUs on one accumulation (crt_group): every member holds the same count / err arrays and runs the same
  // allocation; of the k new samples of a tile -- global sample indices count .. count + k - 1 -- member `rank` renders
  // the ones congruent to rank modulo n_members.  cum_own is the exclusive prefix of this member's tile samples
  // (== cum for one member); the error estimate is then rebuilt from all members' sums by k_adaptive_error_peers.
  uint32_t* cum_own;
  uint32_t rank, n_members;
};

// number of sample indices g in [0, x) with g = rank (mod n)
__device__ __forceinline__ uint32_t adaptive_own_below(uint32_t x, uint32_t rank, uint32_t n) { return (x + n - 1u - rank) / n; }

__device__ __forceinline__ uint32_t adaptive_valid(const AdaptiveState& A, const DeviceParams& P, uint32_t j, uint32_t& vw, uint32_t& vh)
{
  const uint32_t tx = j % A.ntx, ty = j / A.ntx;
  vw = min(kAdaptiveTile, P.width - tx * kAdaptiveTile);
  vh = min(kAdaptiveTile, P.height - ty * kAdaptiveTile);
  return vw * vh;
}

// exclusive block scan of one 64-bit value per thread (1024 threads); returns the prefix, total in `total`
__device__ __forceinline__ unsigned long long block_scan_u64(unsigned long long v, unsigned long long* smem, unsigned long long& total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();                 // smem may still be read from a previous call
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = smem[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    smem[32 + lane] = w;           // inclusive scan of warp totals
  }
  __syncthreads();
  total = smem[32 + 31];
  const unsigned long long warp_base = warp ? smem[32 + warp - 1] : 0ull;
  return warp_base + x - v;
}

// One CTA of 1024 threads; each thread owns a contiguous run of tiles.
__global__ void __launch_bounds__(1024)
k_adaptive_allocate(AdaptiveState A, DeviceParams P, uint32_t budget, uint32_t wave)
{
  __shared__ unsigned long long smem[64];
  const uint32_t per = (A.nt + blockDim.x - 1) / blockDim.x;
  const uint32_t j0 = min(threadIdx.x * per, A.nt), j1 = min(j0 + per, A.nt);
  unsigned long long part = 0, W0 = 0, Wt = 0;
  for (uint32_t j = j0; j < j1; ++j) part += A.err[j];
  block_scan_u64(part, smem, W0);
  const unsigned long long lo = W0 / (8ull * A.nt), hi = (4ull * W0) / A.nt;
  part = 0;
  for (uint32_t j = j0; j < j1; ++j) part += min(max((unsigned long long)A.err[j], lo), hi) + 1ull;
  unsigned long long C = block_scan_u64(part, smem, Wt);
  const unsigned long long off = ((unsigned long long)((wave * 40503u) & 0xffffu) * Wt) >> 16;
  // tile samples and paths of this thread's run
  unsigned long long paths = 0, own = 0;
  {
    unsigned long long c = C;
    for (uint32_t j = j0; j < j1; ++j) {
      const unsigned long long w = min(max((unsigned long long)A.err[j], lo), hi) + 1ull;
      const uint32_t a = (uint32_t)((c * budget + off) / Wt), b = (uint32_t)(((c + w) * budget + off) / Wt);
      uint32_t vw, vh;
      const uint32_t cj = A.count[j];
      const uint32_t mine = adaptive_own_below(cj + (b - a), A.rank, A.n_members) - adaptive_own_below(cj, A.rank, A.n_members);
      paths += (unsigned long long)mine * adaptive_valid(A, P, j, vw, vh);
      own += mine;
      A.cum[j] = a;
      c += w;
    }
  }
  if (A.n_members > 1) {                         // prefix of this member's tile samples (one member: cum_own aliases cum)
    unsigned long long own_total = 0;
    unsigned long long o = block_scan_u64(own, smem, own_total);
    unsigned long long c = C;
    for (uint32_t j = j0; j < j1; ++j) {
      const unsigned long long w = min(max((unsigned long long)A.err[j], lo), hi) + 1ull;
      const uint32_t a = (uint32_t)((c * budget + off) / Wt), b = (uint32_t)(((c + w) * budget + off) / Wt);
      const uint32_t cj = A.count[j];
      A.cum_own[j] = (uint32_t)o;
      o += adaptive_own_below(cj + (b - a), A.rank, A.n_members) - adaptive_own_below(cj, A.rank, A.n_members);
      c += w;
    }
    if (threadIdx.x == 0) A.cum_own[A.nt] = (uint32_t)own_total;
  }
  unsigned long long total_paths = 0;
  unsigned long long q = block_scan_u64(paths, smem, total_paths);
  {
    unsigned long long c = C;
    for (uint32_t j = j0; j < j1; ++j) {
      const unsigned long long w = min(max((unsigned long long)A.err[j], lo), hi) + 1ull;
      const uint32_t a = (uint32_t)((c * budget + off) / Wt), b = (uint32_t)(((c + w) * budget + off) / Wt);
      uint32_t vw, vh;
      const uint32_t cj = A.count[j];
      A.qoff[j] = (uint32_t)q;
      q += (unsigned long long)(adaptive_own_below(cj + (b - a), A.rank, A.n_members) - adaptive_own_below(cj, A.rank, A.n_members))
           * adaptive_valid(A, P, j, vw, vh);
      c += w;
    }
  }
  if (threadIdx.x == 0) { A.cum[A.nt] = budget; A.qoff[A.nt] = (uint32_t)total_paths; }
}

// GenerateRay for the wave's tile samples.  Slot = tile sample * 1024 + (8x4 sub-tile * 32 + lane),
// so a warp still owns an 8x4 pixel block; the queue is written at computed offsets (no atomics).
__global__ void __launch_bounds__(256)
k_generate_adaptive(PathState st, DeviceParams P, AdaptiveState A, const uint32_t* __restrict__ seeds, uint32_t budget)
{
  (void)budget;
  const uint32_t total = A.cum_own[A.nt] * kAdaptiveSlots;      // this member's tile samples (one member: the budget)
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += stride) {
    const uint32_t ts = slot >> 10, within = slot & 1023u;
    uint32_t lo = 0, hi = A.nt;                 // last j with cum_own[j] <= ts  (warp-uniform search)
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(A.cum_own + mid) <= ts) lo = mid; else hi = mid;
    }
    const uint32_t j = lo, tl = ts - __ldg(A.cum_own + j);
    // the tl-th sample of this member in the tile: global index count + t with (count + t) = rank (mod n_members)
    const uint32_t t = (A.rank + A.n_members - __ldg(A.count + j) % A.n_members) % A.n_members + tl * A.n_members;
    const uint32_t sub = within >> 5, lane = within & 31u;
    const uint32_t lx = (sub & 3u) * 8u + (lane & 7u), ly = (sub >> 2) * 4u + (lane >> 3);
    uint32_t vw, vh;
    const uint32_t nv = adaptive_valid(A, P, j, vw, vh);
    if (lx < vw && ly < vh) {
      const uint32_t px = (j % A.ntx) * kAdaptiveTile + lx, py = (j / A.ntx) * kAdaptiveTile + ly;
      generate_path(st, P, slot, px, py, __ldg(seeds + __ldg(A.count + j) + t));
      const uint32_t pos = (nv == kAdaptiveSlots) ? within : ly * vw + lx;
      st.queue[0][__ldg(A.qoff + j) + tl * nv + pos] = slot;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) st.n_active[0] = A.qoff[A.nt];
}

__device__ __forceinline__ float luminance(float r, float g, float b)
{
  return fmaf(0.0722f, b, fmaf(0.7152f, g, 0.2126f * r));
}

// Accumulation of an adaptive wave: one CTA per tile, 4 pixels per thread; adds the tile's k new
// samples of every pixel in sample order, refreshes the tile's error estimate and sample count.
__global__ void __launch_bounds__(256)
k_resolve_adaptive(PathState st, DeviceParams P, AdaptiveState A, float4* __restrict__ accum, Counters* gcnt)
{
  __shared__ uint32_t s_err[8];
  for (uint32_t j = blockIdx.x; j < A.nt; j += gridDim.x) {
    const uint32_t c0 = A.cum_own[j], k = A.cum_own[j + 1] - c0;      // this member's new samples of the tile
    if (k == 0) continue;                        // block-uniform
    const bool alone = A.n_members == 1u;        // several members: err / count are rebuilt from all sums afterwards
    const uint32_t n_old = A.count[j], n_new = n_old + k;
    const uint32_t t_first = (A.rank + A.n_members - n_old % A.n_members) % A.n_members;
    const uint32_t n_even = A.first_parity ? n_new / 2u : (n_new + 1u) / 2u;
    uint32_t vw, vh;
    const uint32_t nv = adaptive_valid(A, P, j, vw, vh);
    uint32_t e_sum = 0;
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t q = threadIdx.x + 256u * i;
      const uint32_t lx = q & 31u, ly = q >> 5;
      if (lx >= vw || ly >= vh) continue;
      const uint32_t within = ((ly >> 2) * 4u + (lx >> 3)) * 32u + (ly & 3u) * 8u + (lx & 7u);
      const size_t pix = (size_t)((j / A.ntx) * kAdaptiveTile + ly) * P.width + (j % A.ntx) * kAdaptiveTile + lx;
      float4 a = accum[pix];
      float ev = A.even[pix];
      for (uint32_t t = 0; t < k; ++t) {
        const float4 c = ld_stream(&st.rad[(size_t)(c0 + t) * kAdaptiveSlots + within]);
        const float r = (c.x != c.x) ? 0.0f : minf(c.x, P.max_radiance);
        const float g = (c.y != c.y) ? 0.0f : minf(c.y, P.max_radiance);
        const float b = (c.z != c.z) ? 0.0f : minf(c.z, P.max_radiance);
        a.x += r; a.y += g; a.z += b; a.w += 1.0f;
        if (((A.first_parity + n_old + t_first + t * A.n_members) & 1u) == 0u) ev += luminance(r, g, b);
      }
      accum[pix] = a;
      A.even[pix] = ev;
      if (alone && n_new >= 2u && n_even > 0u) {
        const float l_all = luminance(a.x, a.y, a.z) / (float)n_new;
        const float l_even = ev / (float)n_even;
        const float e = fabsf(sqrtf(minf(maxf(l_all, 0.0f), 1.0f)) - sqrtf(minf(maxf(l_even, 0.0f), 1.0f)));
        e_sum += (uint32_t)(e * 4096.0f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o);
    __syncthreads();                             // everyone has read count[j]; s_err free again
    if ((threadIdx.x & 31u) == 0u) s_err[threadIdx.x >> 5] = e_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t e = 0;
      for (int w = 0; w < 8; ++w) e += s_err[w];
      if (alone) { A.err[j] = e; A.count[j] = n_new; }
      if (gcnt) atomicAdd(&gcnt->samples, (unsigned long long)nv * k);
    }
  }
}

// Adaptive sampling over several GPUs (crt_group), the exchange step of a wave: member `rank` rebuilds the error
// estimate and the sample count of the tiles j = rank (mod n) from ALL members' sums -- the float4 accumulation
// buffers and the even-sample planes, read through peer addresses and added in member order -- and writes the two
// numbers into every member's arrays, so that all members run the next allocation from the same global estimate.
constexpr int kMaxGroupAdaptive = 16;
struct PeerAdaptive {
  const float4* accum[kMaxGroupAdaptive];
  const float* even[kMaxGroupAdaptive];
  uint32_t* err[kMaxGroupAdaptive];
  uint32_t* count[kMaxGroupAdaptive];
  int n;
};

__global__ void __launch_bounds__(256)
k_adaptive_error_peers(PeerAdaptive G, DeviceParams P, AdaptiveState A)
{
  __shared__ uint32_t s_err[8];
  for (uint32_t j = A.rank + blockIdx.x * A.n_members; j < A.nt; j += gridDim.x * A.n_members) {
    const uint32_t k = A.cum[j + 1] - A.cum[j];  // new samples of the tile in this wave, over all members
    if (k == 0) continue;                        // block-uniform
    const uint32_t n_new = A.count[j] + k;
    const uint32_t n_even = A.first_parity ? n_new / 2u : (n_new + 1u) / 2u;
    uint32_t vw, vh;
    adaptive_valid(A, P, j, vw, vh);
    uint32_t e_sum = 0;
    for (uint32_t i = 0; i < 4; ++i) {
      const uint32_t q = threadIdx.x + 256u * i;
      const uint32_t lx = q & 31u, ly = q >> 5;
      if (lx >= vw || ly >= vh) continue;
      const size_t pix = (size_t)((j / A.ntx) * kAdaptiveTile + ly) * P.width + (j % A.ntx) * kAdaptiveTile + lx;
      float4 a = G.accum[0][pix];
      float ev = G.even[0][pix];
      for (int m = 1; m < G.n; ++m) {
        const float4 b = G.accum[m][pix];
        a.x += b.x; a.y += b.y; a.z += b.z;
        ev += G.even[m][pix];
      }
      if (n_new >= 2u && n_even > 0u) {
        const float l_all = luminance(a.x, a.y, a.z) / (float)n_new;
        const float l_even = ev / (float)n_even;
        const float e = fabsf(sqrtf(minf(maxf(l_all, 0.0f), 1.0f)) - sqrtf(minf(maxf(l_even, 0.0f), 1.0f)));
        e_sum += (uint32_t)(e * 4096.0f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o);
    __syncthreads();                             // s_err free again
    if ((threadIdx.x & 31u) == 0u) s_err[threadIdx.x >> 5] = e_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t e = 0;
      for (int w = 0; w < 8; ++w) e += s_err[w];
      for (int m = 0; m < G.n; ++m) { G.err[m][j] = e; G.count[m][j] = n_new; }
    }
  }
}

// Display.fs restated.
__device__ __forceinline__ float filmic(float c)
{
  float f = fmaf(1.425f, c, 0.05f);
  return (fmaf(c, f, 0.004f)) / (fmaf(c, f + 0.55f, 0.0491f)) - 0.0821f;
}

// one pixel of the Display pass: mean of the sums, exposure, optional filmic curve, gamma 2, RGB8 rounding
__device__ __forceinline__ void display_pixel(const float4 a, uint32_t i, float exposure_scale, int tone_map, float wp_curve,
                                              uint8_t* __restrict__ rgb8, float* __restrict__ rgb32f)
{
  const float inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
  const float m[3] = { a.x * inv, a.y * inv, a.z * inv };
  if (rgb32f) { rgb32f[3 * (size_t)i] = m[0]; rgb32f[3 * (size_t)i + 1] = m[1]; rgb32f[3 * (size_t)i + 2] = m[2]; }
  if (rgb8) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = m[c] * exposure_scale;
      if (tone_map) x = filmic(x) / wp_curve;
      x = sqrtf(maxf(x, 0.0f));
      x = minf(x, 1.0f);
      rgb8[3 * (size_t)i + c] = (uint8_t)(int)fmaf(x, 255.0f, 0.5f);
    }
  }
}

__global__ void __launch_bounds__(256)
k_display(const float4* __restrict__ accum, uint32_t n_pixels, float exposure_scale, int tone_map, float wp_curve,
          uint8_t* __restrict__ rgb8, float* __restrict__ rgb32f)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x)
    display_pixel(accum[i], i, exposure_scale, tone_map, wp_curve, rgb8, rgb32f);
}

// Multi-GPU Redraw (crt_group): the exchange step and the Display pass in ONE kernel.  Every GPU of the group holds
// the sample sums of its own sample range; GPU r of the group runs this kernel over rows [first, first + count) of
// the frame, loads the float4 sums of that slice from every member's accumulation buffer through NVLink peer
// addresses (plain ld.global on a peer pointer), adds them in member order -- a fixed order, so the image does not
// depend on which GPU reduced which slice -- and writes the tone-mapped pixels.  No intermediate sum buffer, no
// second pass: 16 B x members read per pixel, 3 B written.
constexpr int kMaxGroup = 16;
struct PeerAccums { const float4* p[kMaxGroup]; int n; };

__global__ void __launch_bounds__(256)
k_display_peers(PeerAccums A, uint32_t first, uint32_t count, float exposure_scale, int tone_map, float wp_curve,
                uint8_t* __restrict__ rgb8, float* __restrict__ rgb32f, float4* __restrict__ sum_out)
{
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const uint32_t i = first + k;
    float4 a = A.p[0][i];
    for (int r = 1; r < A.n; ++r) {
      const float4 b = A.p[r][i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (sum_out) sum_out[i] = a;
    display_pixel(a, i, exposure_scale, tone_map, wp_curve, rgb8, rgb32f);
  }
}

// crt_wavefront_rays: copies the rays of one bounce out of the path state (parity hook, not on the render path)
__global__ void __launch_bounds__(256)
k_gather_rays(PathState st, const uint32_t* __restrict__ q, uint32_t n, int shadow, float4* __restrict__ org, float4* __restrict__ dir)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (shadow) {
      const float4 o = st.sh_o[i], d = st.sh_d[i];
      org[i] = make_float4(o.x, o.y, o.z, 0.0f);
      dir[i] = make_float4(d.x, d.y, d.z, o.w);
    } else {
      const uint32_t slot = q[i];
      const float4 o = st.ray_o[slot], d = st.ray_d[slot];
      org[i] = make_float4(o.x, o.y, o.z, 0.0f);
      dir[i] = make_float4(d.x, d.y, d.z, CRT_MAXFLOAT);
    }
  }
}

template <bool ANY>
struct TracePolicy {
  const float4* __restrict__ org;
  const float4* __restrict__ dir;
  float4* __restrict__ hit4;
  int32_t* __restrict__ hit_inst;
  const float4* __restrict__ tri_verts;
  __device__ __forceinline__ uint32_t load(uint32_t i, v3& o, v3& d, float& tmax, bool&) const
  {
    const float4 ro = org[i], rd = dir[i];
    o = V(ro.x, ro.y, ro.z); d = V(rd.x, rd.y, rd.z); tmax = rd.w;
    return i;
  }
  __device__ __forceinline__ void store(uint32_t i, const Hit& hit, bool f, bool = false) const
  {
    int32_t prim = -1;
    if (f) prim = ANY ? 0 : __float_as_int(__ldg(tri_verts + kTriStride * (size_t)hit.tri).w);
    hit4[i] = make_float4(hit.t, hit.u, hit.v, __int_as_float(prim));
    if (hit_inst) hit_inst[i] = hit.inst;
  }
};

// Batch SceneNearestHit / SceneAnyHit on caller rays (parity hook crt_trace).
template <bool ANY, bool COUNT, bool PERSISTENT, bool QUAD>
__global__ void __launch_bounds__(CRT_TRACE_BLOCK)
k_trace(DeviceScene S, const float4* __restrict__ org, const float4* __restrict__ dir, uint32_t n,
        float4* __restrict__ hit4, int32_t* __restrict__ hit_inst, uint32_t* work, Counters* gcnt)
{
  Counters cnt = {};
  TracePolicy<ANY> pol{ org, dir, hit4, hit_inst, S.tri_verts };
  if (PERSISTENT) {
    trace_persistent<ANY ? 1 : 0, COUNT, QUAD>(S, n, work, cnt, pol);
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      v3 o, d; float tmax; bool unused = false;
      pol.load(i, o, d, tmax, unused);
      Hit hit;
      const bool f = traverse<ANY, COUNT>(S, o, d, tmax, hit, cnt);
      if (COUNT) { if (ANY) cnt.rays_any++; else cnt.rays_nearest++; }
      pol.store(i, hit, f);
    }
  }
  if (COUNT) flush_counters(gcnt, cnt);
}

}  // namespace crt
