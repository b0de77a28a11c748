/*
 * cadrays_oracle.c -- CPU ORACLE (test infrastructure, never shipped, never on
 * the product path).  PARITY UNPINNED -- see cadrays_oracle.h.
 *
 * Scalar fp32 restatement of the OCCT path tracer that CADRays calls through
 * V3d_View::Redraw() (src/Launcher/AppViewer.cxx:1047).  OCCT itself is absent
 * (unpinned external dependency), so the functions below follow SURVEY.md
 * Appendix A paragraph by paragraph and the parameter semantics recoverable
 * from CADRays' own sources, which are cited per function.
 *
 * Arithmetic contract (what makes GPU-vs-oracle comparison bit-exact):
 *   - compile with -O2 -mfma -ffp-contract=off: no implicit contraction, fused
 *     multiply-adds only where fmaf() is written;
 *   - +,-,*,/,sqrtf are IEEE-754 correctly rounded here and on the GPU
 *     (nvcc -fmad=false, default -prec-div/-prec-sqrt);
 *   - no libm transcendental is used on the per-sample path: sin/cos of 2*pi*x,
 *     exp, atan2 and acos are fixed polynomials defined in this file;
 *   - dot3/cross3 have the fixed fma association written below.
 */
#include "cadrays_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXFLOAT   1.0e15f          /* "MAXFLOAT (1e15)" */
#define ORC_FLT_EPS    1.0e-5f          /* roughness / weight threshold, SURVEY A.6 */
#define ORC_PI         3.14159265358979f
#define ORC_2PI        6.28318530717959f
#define ORC_INV_PI     0.318309886183791f
#define ORC_INV_2PI    0.159154943091895f
#define ORC_MIN_THROUGHPUT   1.0e-3f    /* SURVEY A.7 */
#define ORC_MIN_CONTRIBUTION 1.0e-2f    /* SURVEY A.7 */
#define ORC_STACK      128

typedef struct { float x, y, z; } v3;

/* ------------------------------------------------------------------ vec ops */
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
static inline v3 cross3(v3 a, v3 b)
{
  return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
static inline v3 normalize3(v3 a) { return vscale(a, 1.0f / sqrtf(dot3(a, a))); }
static inline float minf(float a, float b) { return a < b ? a : b; }
static inline float maxf(float a, float b) { return a > b ? a : b; }

/* ---------------------------------------------------------- fixed polynomials */

/* sin and cos of 2*pi*x, x in [0,1].  Quadrant reduction + Taylor-form Horner. */
void orc_sincos2pi(float x, float* s, float* c)
{
  float y = x * 4.0f;
  int   q = (int)(y + 0.5f);
  float r = y - (float)q;                 /* [-0.5, 0.5] */
  float a = r * 1.57079632679490f;        /* [-pi/4, pi/4] */
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
    case 0:  *s =  ps; *c =  pc; break;
    case 1:  *s =  pc; *c = -ps; break;
    case 2:  *s = -ps; *c = -pc; break;
    default: *s = -pc; *c =  ps; break;
  }
}

/* e^x, clamped to [-87, 88]. */
float orc_exp(float x)
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
  union { uint32_t u; float f; } sc;
  sc.u = (uint32_t)((int)n + 127) << 23;
  return p * sc.f;
}

/* atan2 by a degree-11 odd polynomial on [0,1] (max error about 1e-5 rad). */
float orc_atan2(float y, float x)
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
  if (x < 0.0f) r = ORC_PI - r;
  if (y < 0.0f) r = -r;
  return r;
}

/* acos by sqrt(1-|x|) * cubic (max error about 7e-5 rad). */
float orc_acos(float x)
{
  float ax = minf(fabsf(x), 1.0f);
  float p = fmaf(ax, -0.0187293f, 0.0742610f);
  p = fmaf(ax, p, -0.2121144f);
  p = fmaf(ax, p, 1.5707288f);
  float r = sqrtf(1.0f - ax) * p;
  return x < 0.0f ? ORC_PI - r : r;
}

/* ------------------------------------------------------------------------ RNG */

/* math_BullardGenerator as recalled in SURVEY A.1: the seed of sample index k is
 * the k-th NextInt() >> 2 of a generator seeded with frame_seed0. */
uint32_t orc_bullard_frame_seed(uint32_t seed0, uint64_t sample_index)
{
  uint32_t hi = seed0, lo = seed0 ^ 0x49616E42u, out = 0;
  for (uint64_t k = 0; k <= sample_index; ++k) {
    hi = (hi >> 2) + (hi << 2);
    hi += lo;
    lo += hi;
    out = hi;
  }
  return out >> 2;
}

/* SeedRand, SURVEY A.8. radius = 8 in coherent mode (SettingsWidget.cxx:419-425) else 1. */
uint32_t orc_seed_rand(uint32_t frame_seed, uint32_t px, uint32_t py, uint32_t size_x, int radius)
{
  uint32_t s = (py / (uint32_t)radius) * size_x + px / (uint32_t)radius + frame_seed;
  s = (s + 0x479ab41du) + (s << 8);
  s = (s ^ 0xe4aa10ceu) ^ (s >> 5);
  s = (s + 0x9942f0a6u) - (s << 14);
  s = (s ^ 0x5aedd67du) ^ (s >> 3);
  s = (s + 0x17bea992u) + (s << 7);
  return s;
}

/* RandFloat, SURVEY A.8: xorshift32 (13,17,5), float(state) * 2^-32.  The
 * product is clamped below 1 (128 of 2^32 states round to 1.0f otherwise). */
float orc_rand_float(uint32_t* state)
{
  uint32_t s = *state;
  s ^= s << 13;
  s ^= s >> 17;
  s ^= s << 5;
  *state = s;
  return minf((float)s * 2.3283064365386963e-10f, 0.99999994f);
}

/* ---------------------------------------------------------------------- scene */

typedef struct {
  uint32_t magic, version, n_nodes, n_verts, n_tris, n_inst, n_top_nodes, flags;
  float    scene_min[3], scene_max[3], scene_eps;
  uint32_t reserved;
} blob_header;

struct orc_scene {
  blob_header    hdr;
  const int32_t* node_info;   /* ivec4 per node */
  const float*   node_min;    /* vec3 per node */
  const float*   node_max;
  const float*   vert_pos;    /* vec3 per vertex */
  const float*   vert_nrm;
  const float*   vert_uv;
  const int32_t* tris;        /* ivec4: v0,v1,v2 (mesh-local), caller's triangle index */
  const float*   inst_inv;    /* 4 vec4 rows of the inverse (world->object) matrix */
  const int32_t* inst_meta;   /* material id, mesh id, bottom root node, 0 */
  uint8_t*       storage;
  int32_t*       inst_geom;   /* per instance: vertex offset, triangle offset, triangle count */
  crt_bsdf*      mats;   uint32_t n_mats;
  /* lights in shader form: emission, w = cosMax | radius; xyz = to-light dir | pos */
  float*         lights; uint32_t n_lights;   /* 8 floats per light */
  float*         env;    uint32_t env_w, env_h;
  /* base-colour textures (Graphic3d_Texture2Dmanual on the aspect, AisMesh.cxx:343-345): RGBA8 texels
   * of all textures back to back (rows top-down as in the image file), table = offset, w, h per texture */
  uint8_t*       tex_data; uint32_t* tex_table; uint32_t n_tex;
  crt_params     params;
  crt_camera     cam;
  v3             cam_u, cam_v, cam_w;
  float          cam_hw, cam_hh;
};

static size_t align16(size_t x) { return (x + 15u) & ~(size_t)15u; }

orc_scene* orc_scene_from_blob(const void* blob, size_t size)
{
  if (!blob || size < sizeof(blob_header)) return NULL;
  orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
  s->storage = (uint8_t*)malloc(size);
  memcpy(s->storage, blob, size);
  memcpy(&s->hdr, s->storage, sizeof(blob_header));
  if (s->hdr.magic != 0x42545243u || s->hdr.version != 1u) { orc_scene_free(s); return NULL; }
  size_t off = sizeof(blob_header);
  const blob_header* h = &s->hdr;
  s->node_info = (const int32_t*)(s->storage + off); off = align16(off + (size_t)16 * h->n_nodes);
  s->node_min  = (const float*)(s->storage + off);   off = align16(off + (size_t)12 * h->n_nodes);
  s->node_max  = (const float*)(s->storage + off);   off = align16(off + (size_t)12 * h->n_nodes);
  s->vert_pos  = (const float*)(s->storage + off);   off = align16(off + (size_t)12 * h->n_verts);
  s->vert_nrm  = (const float*)(s->storage + off);   off = align16(off + (size_t)12 * h->n_verts);
  s->vert_uv   = (const float*)(s->storage + off);   off = align16(off + (size_t)8 * h->n_verts);
  s->tris      = (const int32_t*)(s->storage + off); off = align16(off + (size_t)16 * h->n_tris);
  s->inst_inv  = (const float*)(s->storage + off);   off = align16(off + (size_t)64 * h->n_inst);
  s->inst_meta = (const int32_t*)(s->storage + off); off = align16(off + (size_t)16 * h->n_inst);
  if (off > size) { orc_scene_free(s); return NULL; }
  /* per-instance geometry ranges, from the top-level leaf records */
  s->inst_geom = (int32_t*)calloc(3 * (size_t)(h->n_inst ? h->n_inst : 1), sizeof(int32_t));
  for (uint32_t q = 0; q < h->n_top_nodes; ++q) {
    const int32_t* info = s->node_info + 4 * q;
    if (info[0] <= 0) continue;
    int32_t k = info[0] - 1, root = info[1], mx = -1;
    int st[ORC_STACK]; int hd = 0; st[0] = root;
    while (hd >= 0) {
      const int32_t* ni = s->node_info + 4 * st[hd--];
      if (ni[0] == 0) {
        if (h->flags & 2u) { for (int c = 0; c <= ni[2]; ++c) st[++hd] = root + ni[1] + c; }
        else { st[++hd] = root + ni[1]; st[++hd] = root + ni[2]; }
      }
      else if (ni[2] > mx) mx = ni[2];
    }
    s->inst_geom[3 * k] = info[2]; s->inst_geom[3 * k + 1] = info[3]; s->inst_geom[3 * k + 2] = mx + 1;
  }
  crt_params p;
  memset(&p, 0, sizeof p);
  p.max_depth = 8; p.max_radiance = 50.0f; p.focal_dist = 1.0f; p.white_point = 1.0f;
  p.env_as_background = 1; p.frame_seed0 = 1; p.russian_roulette = 1;
  orc_set_params(s, &p);
  return s;
}

void orc_scene_free(orc_scene* s)
{
  if (!s) return;
  free(s->storage); free(s->inst_geom); free(s->mats); free(s->lights); free(s->env);
  free(s->tex_data); free(s->tex_table); free(s);
}

void orc_set_materials(orc_scene* s, const crt_bsdf* m, uint32_t n)
{
  free(s->mats);
  s->mats = (crt_bsdf*)malloc(sizeof(crt_bsdf) * (n ? n : 1));
  if (n) memcpy(s->mats, m, sizeof(crt_bsdf) * n);
  s->n_mats = n;
}

/* Light table in the shader's form: record 0 = (rgb*intensity,
 * w = cos(smoothness) for directional | radius for positional), record 1 =
 * (to-light direction | position, w = 0 | 1).  crt_light.posdir of a directional
 * light is the direction the light travels ("vlight ... direction",
 * Materials.tcl:203), so the to-light vector is its negated normalisation. */
void orc_set_lights(orc_scene* s, const crt_light* l, uint32_t n)
{
  free(s->lights);
  s->lights = (float*)malloc(sizeof(float) * 8 * (n ? n : 1));
  s->n_lights = n;
  for (uint32_t i = 0; i < n; ++i) {
    float* r = s->lights + 8 * i;
    r[0] = l[i].emission[0]; r[1] = l[i].emission[1]; r[2] = l[i].emission[2];
    if (l[i].is_point) {
      r[3] = l[i].smoothness;
      r[4] = l[i].posdir[0]; r[5] = l[i].posdir[1]; r[6] = l[i].posdir[2]; r[7] = 1.0f;
    } else {
      r[3] = cosf(l[i].smoothness);
      v3 d = normalize3(V(l[i].posdir[0], l[i].posdir[1], l[i].posdir[2]));
      r[4] = -d.x; r[5] = -d.y; r[6] = -d.z; r[7] = 0.0f;
    }
  }
}

void orc_set_envmap_rgb32f(orc_scene* s, const float* rgb, uint32_t w, uint32_t h)
{
  free(s->env); s->env = NULL; s->env_w = s->env_h = 0;
  if (!rgb || !w || !h) return;
  s->env = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
  memcpy(s->env, rgb, sizeof(float) * 3 * (size_t)w * h);
  s->env_w = w; s->env_h = h;
}

/* OCCT linearises 8-bit environment texels by squaring (gamma 2, the inverse of
 * Display.fs' sqrt) -- SURVEY A.7/A.9. */
void orc_set_envmap_rgb8(orc_scene* s, const uint8_t* rgb, uint32_t w, uint32_t h)
{
  free(s->env); s->env = NULL; s->env_w = s->env_h = 0;
  if (!rgb || !w || !h) return;
  size_t n = 3 * (size_t)w * h;
  s->env = (float*)malloc(sizeof(float) * n);
  for (size_t i = 0; i < n; ++i) { float c = (float)rgb[i] * (1.0f / 255.0f); s->env[i] = c * c; }
  s->env_w = w; s->env_h = h;
}

/* Textures: n RGBA8 images concatenated in `texels`, sizes[2k], sizes[2k+1] = width, height. */
void orc_set_textures(orc_scene* s, const uint8_t* texels, const uint32_t* sizes, uint32_t n)
{
  free(s->tex_data); free(s->tex_table);
  s->tex_data = NULL; s->tex_table = NULL; s->n_tex = 0;
  if (!n) return;
  size_t total = 0;
  s->tex_table = (uint32_t*)malloc(sizeof(uint32_t) * 3 * n);
  for (uint32_t k = 0; k < n; ++k) {
    s->tex_table[3 * k] = (uint32_t)total; s->tex_table[3 * k + 1] = sizes[2 * k]; s->tex_table[3 * k + 2] = sizes[2 * k + 1];
    total += (size_t)sizes[2 * k] * sizes[2 * k + 1];
  }
  s->tex_data = (uint8_t*)malloc(4 * total);
  memcpy(s->tex_data, texels, 4 * total);
  s->n_tex = n;
}

void orc_set_params(orc_scene* s, const crt_params* p) { s->params = *p; }

/* Camera basis (Graphic3d_Camera eye/dir/up/FOVy/aspect, AppViewer.cxx:993-1042;
 * SURVEY A.2 corner rays written in closed form). */
void orc_set_camera(orc_scene* s, const crt_camera* c)
{
  s->cam = *c;
  v3 w = normalize3(V(c->dir[0], c->dir[1], c->dir[2]));
  v3 u = normalize3(cross3(w, V(c->up[0], c->up[1], c->up[2])));
  v3 v = cross3(u, w);
  s->cam_u = u; s->cam_v = v; s->cam_w = w;
  if (c->is_ortho) {
    s->cam_hh = c->ortho_scale * 0.5f;
  } else {
    s->cam_hh = tanf(c->fovy_deg * 0.5f * 0.0174532925199433f);
  }
  s->cam_hw = s->cam_hh * c->aspect;
}

float orc_scene_epsilon(const orc_scene* s) { return s->hdr.scene_eps; }

/* GenerateRay + thin lens, SURVEY A.2.  px,py in [0,1] (bottom-left origin). */
void orc_camera_ray(const orc_scene* s, float px, float py, float lens_a, float lens_b,
                    float org[3], float dir[3])
{
  float sx = fmaf(px, 2.0f, -1.0f) * s->cam_hw;
  float sy = fmaf(py, 2.0f, -1.0f) * s->cam_hh;
  v3 eye = V(s->cam.eye[0], s->cam.eye[1], s->cam.eye[2]);
  v3 o, d;
  if (s->cam.is_ortho) {
    o = vadd(eye, vadd(vscale(s->cam_u, sx), vscale(s->cam_v, sy)));
    d = s->cam_w;
  } else {
    o = eye;
    d = normalize3(vadd(s->cam_w, vadd(vscale(s->cam_u, sx), vscale(s->cam_v, sy))));
  }
  if (s->params.aperture_radius > 0.0f) {
    float ft = s->params.focal_dist / dot3(d, s->cam_w);
    v3 focus = vadd(o, vscale(d, ft));
    float sn, cs;
    orc_sincos2pi(lens_b, &sn, &cs);
    float r = sqrtf(lens_a) * s->params.aperture_radius;
    o = vadd(o, vadd(vscale(s->cam_u, r * cs), vscale(s->cam_v, r * sn)));
    d = normalize3(vsub(focus, o));
  }
  org[0] = o.x; org[1] = o.y; org[2] = o.z;
  dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
}

/* ------------------------------------------------------------------ traversal */

typedef struct { float t, u, v; int tri; int inst; int voff; v3 n; } hit_t;
typedef struct { v3 o, d, inv, oinv; } ray_t;

/* InverseDirection, SURVEY A.3: sign(d) / max(|d|, 2^-80). */
static inline float inv_dir(float d)
{
  float a = 1.0f / maxf(fabsf(d), 8.271806125530277e-25f);
  return d < 0.0f ? -a : a;
}

static inline void ray_setup(ray_t* r, v3 o, v3 d)
{
  r->o = o; r->d = d;
  r->inv = V(inv_dir(d.x), inv_dir(d.y), inv_dir(d.z));
  r->oinv = V(-(o.x * r->inv.x), -(o.y * r->inv.y), -(o.z * r->inv.z));
}

/* Slab test in the form t = fma(bound, inv, -(o*inv)).
 * Returns 1 iff max(tEnter,0) <= min(tExit, tbest); *tenter gets the raw entry. */
static inline int slab(const ray_t* r, const float* bmin, const float* bmax, float tbest, float* tenter)
{
  float x0 = fmaf(bmin[0], r->inv.x, r->oinv.x), x1 = fmaf(bmax[0], r->inv.x, r->oinv.x);
  float y0 = fmaf(bmin[1], r->inv.y, r->oinv.y), y1 = fmaf(bmax[1], r->inv.y, r->oinv.y);
  float z0 = fmaf(bmin[2], r->inv.z, r->oinv.z), z1 = fmaf(bmax[2], r->inv.z, r->oinv.z);
  float te = maxf(maxf(minf(x0, x1), minf(y0, y1)), minf(z0, z1));
  float tx = minf(minf(maxf(x0, x1), maxf(y0, y1)), maxf(z0, z1));
  *tenter = te;
  return maxf(te, 0.0f) <= minf(tx, tbest);
}

/* IntersectTriangle, SURVEY A.4.  Returns 1 and t,u,v,n when the ray hits. */
static inline int tri_test(v3 o, v3 d, v3 p0, v3 p1, v3 p2, float* t, float* u, float* v, v3* n)
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
  if (tt >= 0.0f && uu >= 0.0f && vv >= 0.0f && uu + vv <= 1.0f) {
    *t = tt; *u = uu; *v = vv; *n = nn;
    return 1;
  }
  return 0;
}

static inline v3 ld3(const float* p, int i) { return V(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

/* world -> object by the instance's inverse matrix rows, direction not renormalised
 * ("so t stays in world units"). */
static inline v3 xf_point(const float* m, v3 p)
{
  return V(fmaf(m[2], p.z, fmaf(m[1], p.y, m[0] * p.x)) + m[3],
           fmaf(m[6], p.z, fmaf(m[5], p.y, m[4] * p.x)) + m[7],
           fmaf(m[10], p.z, fmaf(m[9], p.y, m[8] * p.x)) + m[11]);
}
static inline v3 xf_vector(const float* m, v3 p)
{
  return V(fmaf(m[2], p.z, fmaf(m[1], p.y, m[0] * p.x)),
           fmaf(m[6], p.z, fmaf(m[5], p.y, m[4] * p.x)),
           fmaf(m[10], p.z, fmaf(m[9], p.y, m[8] * p.x)));
}

/* optional per-ray event recorder (analysis tooling: feeds tests/analysis/simt_model.py) */
static __thread uint8_t* g_ev = NULL;
static __thread int g_ev_n = 0, g_ev_cap = 0;
#define ORC_EVENT(code) do { if (g_ev && g_ev_n < g_ev_cap) g_ev[g_ev_n++] = (uint8_t)(code); } while (0)

/* SceneNearestHit / SceneAnyHit, SURVEY A.3. */
static int traverse(const orc_scene* s, v3 org, v3 dir, float tmax, int any_hit, hit_t* hit,
                    crt_stats* st)
{
  hit->t = tmax; hit->tri = -1; hit->inst = -1; hit->voff = 0; hit->u = hit->v = 0.0f; hit->n = V(0, 0, 0);
  if (s->hdr.n_nodes == 0) return 0;
  /* degenerate rays (zero or NaN direction, NaN origin) miss by definition */
  if (!(dot3(dir, dir) > 0.0f) || !(dot3(org, org) >= 0.0f)) return 0;
  int stack[ORC_STACK];
  int head = -1, stop = -1;
  int node = 0, node_off = 0, vert_off = 0, tri_off = 0, inst = -1;
  ray_t world, cur;
  ray_setup(&world, org, dir);
  cur = world;
  uint64_t n_inner = 0, n_leaf = 0, n_tri = 0, n_switch = 0, n_boxes = 0;
  int found = 0;
  for (;;) {
    const int32_t* info = s->node_info + 4 * node;
    if (info[0] == 0) {                                   /* inner node */
      ++n_inner; ORC_EVENT(1);
      if (s->hdr.flags & 2u) {
        /* QUAD_BVH: up to 4 contiguous children, all tested, sorted by entry distance with
         * the 5-comparator network (0,1)(2,3)(0,2)(1,3)(1,2) (swap when the later entry is strictly nearer),
         * pushed far-to-near */
        const int first = node_off + info[1], k = info[2] + 1;
        float te[4]; int id[4];
        for (int c = 0; c < 4; ++c) {
          te[c] = 3.0e38f; id[c] = -1;
          if (c < k) {
            float t;
            if (slab(&cur, s->node_min + 3 * (first + c), s->node_max + 3 * (first + c), hit->t, &t)) { te[c] = t; id[c] = first + c; }
          }
        }
        n_boxes += (uint64_t)k;
#define ORC_CSWAP(i, j) do { if (te[j] < te[i]) { float tf = te[i]; te[i] = te[j]; te[j] = tf; int ti = id[i]; id[i] = id[j]; id[j] = ti; } } while (0)
        ORC_CSWAP(0, 1); ORC_CSWAP(2, 3); ORC_CSWAP(0, 2); ORC_CSWAP(1, 3); ORC_CSWAP(1, 2);
#undef ORC_CSWAP
        for (int c = 3; c >= 1; --c) if (id[c] >= 0) stack[++head] = id[c];
        if (id[0] >= 0) { node = id[0]; continue; }
      } else {
        int l = node_off + info[1], r = node_off + info[2];
        float tl, tr;
        int hl = slab(&cur, s->node_min + 3 * l, s->node_max + 3 * l, hit->t, &tl);
        int hr = slab(&cur, s->node_min + 3 * r, s->node_max + 3 * r, hit->t, &tr);
        n_boxes += 2;
        if (hl && hr) {
          int nearer = (tr < tl) ? r : l;
          int farther = (tr < tl) ? l : r;
          stack[++head] = farther;
          node = nearer;
          continue;
        }
        if (hl) { node = l; continue; }
        if (hr) { node = r; continue; }
      }
    } else if (info[0] < 0) {                             /* bottom-level leaf */
      ++n_leaf; ORC_EVENT(3 + (info[2] - info[1] + 1));
      for (int i = info[1]; i <= info[2]; ++i) {
        const int32_t* tr = s->tris + 4 * (tri_off + i);
        ++n_tri;
        float t, u, v; v3 n;
        if (tri_test(cur.o, cur.d, ld3(s->vert_pos, vert_off + tr[0]), ld3(s->vert_pos, vert_off + tr[1]),
                     ld3(s->vert_pos, vert_off + tr[2]), &t, &u, &v, &n) && t < hit->t) {
          hit->t = t; hit->u = u; hit->v = v; hit->n = n; hit->tri = tri_off + i; hit->inst = inst; hit->voff = vert_off;
          found = 1;
          if (any_hit) goto done;
        }
      }
    } else {                                              /* top-level leaf: enter instance */
      ++n_switch; ORC_EVENT(2);
      inst = info[0] - 1; node_off = info[1]; vert_off = info[2]; tri_off = info[3];
      const float* m = s->inst_inv + 16 * inst;
      ray_setup(&cur, xf_point(m, world.o), xf_vector(m, world.d));
      node = node_off;
      stop = head;
      continue;
    }
    /* pop */
    if (head < 0) break;
    if (head == stop) { cur = world; node_off = 0; stop = -1; }
    node = stack[head--];
  }
done:
  if (st) {
    if (any_hit) { st->n_inner_any += n_inner; st->n_leaf_any += n_leaf; st->n_tri_any += n_tri; st->n_switch_any += n_switch; st->n_boxes_any += n_boxes; }
    else { st->n_inner += n_inner; st->n_leaf += n_leaf; st->n_tri += n_tri; st->n_switch += n_switch; st->n_boxes += n_boxes; }
  }
  return found;
}

void orc_trace(const orc_scene* s, const float* org, const float* dir, const float* tmax,
               uint32_t n, int any_hit,
               int32_t* prim, int32_t* inst, float* t, float* u, float* v, crt_stats* stats)
{
  for (uint32_t i = 0; i < n; ++i) {
    hit_t h;
    int f = traverse(s, ld3(org, i), ld3(dir, i), tmax ? tmax[i] : ORC_MAXFLOAT, any_hit, &h, stats);
    if (stats) { if (any_hit) stats->rays_any++; else stats->rays_nearest++; }
    if (any_hit) {
      if (prim) prim[i] = f ? 0 : -1;
      if (inst) inst[i] = f ? h.inst : -1;
    } else {
      if (prim) prim[i] = f ? s->tris[4 * h.tri + 3] : -1;
      if (inst) inst[i] = h.inst;
    }
    if (t) t[i] = h.t;
    if (u) u[i] = h.u;
    if (v) v[i] = h.v;
  }
}

/* Event strings of the traversal of each ray: 1 = inner node, 2 = instance switch, 3+k = leaf with k
 * triangles.  events: n x cap bytes, lens: n ints.  Analysis tooling only. */
void orc_trace_events(const orc_scene* s, const float* org, const float* dir, const float* tmax, uint32_t n,
                      int any_hit, uint8_t* events, int cap, int32_t* lens)
{
  for (uint32_t i = 0; i < n; ++i) {
    hit_t h;
    g_ev = events + (size_t)i * cap; g_ev_n = 0; g_ev_cap = cap;
    traverse(s, ld3(org, i), ld3(dir, i), tmax ? tmax[i] : ORC_MAXFLOAT, any_hit, &h, NULL);
    lens[i] = g_ev_n;
    g_ev = NULL;
  }
}

void orc_trace_brute(const orc_scene* s, const float* org, const float* dir, const float* tmax,
                     uint32_t n, int any_hit,
                     int32_t* prim, int32_t* inst, float* t, float* u, float* v)
{
  for (uint32_t i = 0; i < n; ++i) {
    v3 o = ld3(org, i), d = ld3(dir, i);
    float best = tmax ? tmax[i] : ORC_MAXFLOAT, bu = 0, bv = 0;
    int btri = -1, binst = -1;
    for (uint32_t k = 0; k < s->hdr.n_inst; ++k) {
      const float* m = s->inst_inv + 16 * k;
      v3 lo = xf_point(m, o), ld = xf_vector(m, d);
      int32_t voff = s->inst_geom[3 * k], toff = s->inst_geom[3 * k + 1], tcount = s->inst_geom[3 * k + 2];
      for (int32_t j = 0; j < tcount; ++j) {
        const int32_t* tr = s->tris + 4 * (toff + j);
        float tt, uu, vv; v3 nn;
        if (tri_test(lo, ld, ld3(s->vert_pos, voff + tr[0]), ld3(s->vert_pos, voff + tr[1]),
                     ld3(s->vert_pos, voff + tr[2]), &tt, &uu, &vv, &nn) && tt < best) {
          best = tt; bu = uu; bv = vv; btri = toff + j; binst = (int)k;
        }
      }
    }
    if (any_hit) { if (prim) prim[i] = btri >= 0 ? 0 : -1; }
    else if (prim) prim[i] = btri >= 0 ? s->tris[4 * btri + 3] : -1;
    if (inst) inst[i] = binst;
    if (t) t[i] = best;
    if (u) u[i] = bu;
    if (v) v[i] = bv;
  }
}

/* ----------------------------------------------------------------------- BSDF */

typedef struct {
  v3 Kc; float Kc_w;
  v3 Kd;
  v3 Ks; float Ks_w;
  v3 Kt;
  v3 Fc;   /* FresnelCoat xyz */
  v3 Fb;   /* FresnelBase xyz */
} bsdf_t;

static bsdf_t bsdf_load(const crt_bsdf* b)
{
  bsdf_t r;
  r.Kc = V(b->Kc[0], b->Kc[1], b->Kc[2]); r.Kc_w = b->Kc[3];
  r.Kd = V(b->Kd[0], b->Kd[1], b->Kd[2]);
  r.Ks = V(b->Ks[0], b->Ks[1], b->Ks[2]); r.Ks_w = b->Ks[3];
  r.Kt = V(b->Kt[0], b->Kt[1], b->Kt[2]);
  r.Fc = V(b->FresnelCoat[0], b->FresnelCoat[1], b->FresnelCoat[2]);
  r.Fb = V(b->FresnelBase[0], b->FresnelBase[1], b->FresnelBase[2]);
  return r;
}

static float fresnel_dielectric4(float cos_i, float cos_t, float eta_i, float eta_t)
{
  float parl = (eta_t * cos_i - eta_i * cos_t) / (eta_t * cos_i + eta_i * cos_t);
  float perp = (eta_i * cos_i - eta_t * cos_t) / (eta_i * cos_i + eta_t * cos_t);
  return (parl * parl + perp * perp) * 0.5f;
}

/* handles inside/outside by the sign of cos_i and total internal reflection */
static float fresnel_dielectric(float cos_i, float index)
{
  float eta_i = cos_i > 0.0f ? 1.0f : index;
  float eta_t = cos_i > 0.0f ? index : 1.0f;
  float sin_t2 = (eta_i * eta_i) / (eta_t * eta_t) * (1.0f - cos_i * cos_i);
  if (sin_t2 < 1.0f) return fresnel_dielectric4(fabsf(cos_i), sqrtf(1.0f - sin_t2), eta_i, eta_t);
  return 1.0f;
}

static float fresnel_conductor(float cos_i, float eta, float k)
{
  float tmp = 2.0f * eta * cos_i;
  float tmp1 = eta * eta + k * k;
  float s_perp = (tmp1 - tmp + cos_i * cos_i) / (tmp1 + tmp + cos_i * cos_i);
  float tmp2 = tmp1 * cos_i * cos_i;
  float s_parl = (tmp2 - tmp + 1.0f) / (tmp2 + tmp + 1.0f);
  return (s_perp + s_parl) * 0.5f;
}

/* fresnelMedia, SURVEY A.5/A.6: type coded in the sign of x (Graphic3d_Fresnel::
 * Serialize, MaterialEditor.cxx:209-255; export ImportExport.cxx:204-227). */
static v3 fresnel_media(float cos_i, v3 f)
{
  if (f.x > -0.5f) {
    float m = 1.0f - fabsf(cos_i);
    float m2 = m * m;
    float m5 = m2 * m2 * m;
    return V(f.x + (1.0f - f.x) * m5, f.y + (1.0f - f.y) * m5, f.z + (1.0f - f.z) * m5);
  }
  if (f.x > -1.5f) return V(f.z, f.z, f.z);
  if (f.x > -2.5f) { float c = fresnel_conductor(fabsf(cos_i), f.y, f.z); return V(c, c, c); }
  { float c = fresnel_dielectric(cos_i, f.y); return V(c, c, c); }
}

void orc_fresnel(float cos_i, const float f[4], float out[3])
{
  v3 r = fresnel_media(cos_i, V(f[0], f[1], f[2]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* GGX normal distribution D(m) for roughness a. */
static float ggx_d(float mz, float a)
{
  float a2 = a * a;
  float q = fmaf(mz * mz, a2 - 1.0f, 1.0f);
  return a2 / (ORC_PI * q * q);
}

/* Smith G1 = 2 / (1 + sqrt(1 + a^2 tan^2)), SURVEY A.6. */
static float smith_g1(v3 dir, v3 m, float a)
{
  if (dot3(dir, m) * dir.z <= 0.0f) return 0.0f;
  float c2 = dir.z * dir.z;
  float tan2 = (1.0f - c2) / c2;
  return 2.0f / (1.0f + sqrtf(fmaf(a * a, tan2, 1.0f)));
}

/* f * cos_i of the glossy lobe. */
static v3 eval_ggx(v3 wi, v3 wo, v3 fresnel, float a)
{
  if (wi.z <= 0.0f || wo.z <= 0.0f) return V(0, 0, 0);
  v3 h = normalize3(vadd(wi, wo));
  float d = ggx_d(h.z, a);
  float g = smith_g1(wo, h, a) * smith_g1(wi, h, a);
  return vscale(fresnel_media(dot3(wo, h), fresnel), d * g / (4.0f * wo.z));
}

static float eval_lambert(v3 wi, v3 wo)
{
  return (wi.z <= 0.0f || wo.z <= 0.0f) ? 0.0f : wi.z * ORC_INV_PI;
}

/* EvalBsdfLayered, SURVEY A.6: base {Kd Lambert + Ks GGX*F_base} seen through
 * (1 - F_coat(wo)), plus coat Kc GGX * F_coat. */
static v3 eval_bsdf_layered(const bsdf_t* b, v3 wi, v3 wo, int two_sided)
{
  if (two_sided) { wi.z = fabsf(wi.z); wo.z = fabsf(wo.z); }
  v3 r = vscale(b->Kd, eval_lambert(wi, wo));
  if (b->Ks_w > ORC_FLT_EPS) r = vadd(r, vmul(b->Ks, eval_ggx(wi, wo, b->Fb, b->Ks_w)));
  v3 cf = fresnel_media(wo.z, b->Fc);
  r = vmul(r, V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z));
  if (b->Kc_w > ORC_FLT_EPS) r = vadd(r, vmul(b->Kc, eval_ggx(wi, wo, b->Fc, b->Kc_w)));
  return r;
}

static float ggx_pdf_term(float hz, float a, float wi_dot_h)
{
  return ggx_d(hz, a) * fabsf(hz) * 0.25f / wi_dot_h;
}

/* BsdfPdfLayered, SURVEY A.6: lobe-selection-weighted pdf of direction wi. */
static float bsdf_pdf_layered(const bsdf_t* b, v3 wo, v3 wi, v3 weight)
{
  v3 cf = fresnel_media(wo.z, b->Fc);
  v3 ct = V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z);
  float pc = dot3(vmul(b->Kc, cf), weight);
  float pd = dot3(vmul(b->Kd, ct), weight);
  float ps = dot3(vmul(b->Ks, ct), weight);
  float pt = dot3(vmul(b->Kt, ct), weight);
  float pdf = 0.0f;
  if (wi.z * wo.z > 0.0f) {
    v3 h = normalize3(vadd(wi, wo));
    float wh = dot3(wi, h);
    pdf = pd * fabsf(wi.z * ORC_INV_PI);
    if (b->Kc_w > ORC_FLT_EPS) pdf += pc * ggx_pdf_term(h.z, b->Kc_w, wh);
    if (b->Ks_w > ORC_FLT_EPS) pdf += ps * ggx_pdf_term(h.z, b->Ks_w, wh);
  }
  return pdf / ((pc + pd) + (ps + pt));
}

/* cosine-weighted hemisphere; two-sided flips onto wo's side. */
static v3 sample_lambert(v3 wo, v3* wi, float* pdf, uint32_t* rng, int two_sided)
{
  float k1 = orc_rand_float(rng);
  float k2 = orc_rand_float(rng);
  float sn, cs;
  orc_sincos2pi(k1, &sn, &cs);
  float r = sqrtf(k2);
  v3 w = V(cs * r, sn * r, sqrtf(1.0f - k2));
  if (two_sided && wo.z < 0.0f) w.z = -w.z;
  *wi = w;
  *pdf *= fabsf(w.z) * ORC_INV_PI;
  if (two_sided) return V(1, 1, 1);
  return wo.z >= 0.0f ? V(1, 1, 1) : V(0, 0, 0);
}

/* GGX microfacet sampling, SURVEY A.6: tan^2(theta_m) = a^2 k1/(1-k1), phi = 2 pi k2,
 * weight |wo.m| G / (|wo.z| |m.z|) * F. */
static v3 sample_ggx(v3 wo, v3* wi, v3 fresnel, float a, float* pdf, uint32_t* rng, int two_sided)
{
  float k1 = orc_rand_float(rng);
  float k2 = orc_rand_float(rng);
  float tan2 = a * a * k1 / (1.0f - k1);
  float cos_m = 1.0f / sqrtf(1.0f + tan2);
  float sin_m = sqrtf(maxf(1.0f - cos_m * cos_m, 0.0f));
  float sn, cs;
  orc_sincos2pi(k2, &sn, &cs);
  v3 m = V(cs * sin_m, sn * sin_m, cos_m);
  *pdf *= ggx_d(cos_m, a) * cos_m;
  int flip = two_sided && wo.z < 0.0f;
  if (flip) wo.z = -wo.z;
  float cos_d = dot3(wo, m);
  v3 w = V(fmaf(2.0f * cos_d, m.x, -wo.x), fmaf(2.0f * cos_d, m.y, -wo.y), fmaf(2.0f * cos_d, m.z, -wo.z));
  *wi = w;
  if (w.z <= 0.0f || wo.z <= 0.0f) return V(0, 0, 0);
  *pdf /= 4.0f * cos_d;
  float g = smith_g1(wo, m, a) * smith_g1(w, m, a);
  if (flip) wi->z = -w.z;
  return vscale(fresnel_media(cos_d, fresnel), (g * cos_d) / (wo.z * cos_m));
}

/* refraction through the coat interface; eta from the coat's dielectric IOR
 * (glass preset keeps its IOR in FresnelCoat, MaterialEditor.cxx:791-796). */
static v3 transmitted(float index, v3 wo)
{
  float eta = wo.z > 0.0f ? 1.0f / index : index;
  float sin_t2 = eta * eta * (1.0f - wo.z * wo.z);
  float cos_t = sqrtf(1.0f - minf(sin_t2, 1.0f));
  if (wo.z > 0.0f) cos_t = -cos_t;
  return normalize3(V(-eta * wo.x, -eta * wo.y, cos_t));
}

/* SampleBsdfLayered, SURVEY A.6.  Returns the pdf of the sampled direction
 * (ORC_MAXFLOAT for delta lobes), multiplies *weight by f*cos/pdf. */
static float sample_bsdf_layered(const bsdf_t* b, v3 wo, v3* wi, v3* weight, int* inside,
                                 uint32_t* rng, int two_sided)
{
  float pdf = 0.0f;
  v3 cf = fresnel_media(wo.z, b->Fc);
  v3 ct = V(1.0f - cf.x, 1.0f - cf.y, 1.0f - cf.z);
  float pc = dot3(vmul(b->Kc, cf), *weight);
  float pd = dot3(vmul(b->Kd, ct), *weight);
  float ps = dot3(vmul(b->Ks, ct), *weight);
  float pt = dot3(vmul(b->Kt, ct), *weight);
  float total = (pc + pd) + (ps + pt);
  float ksi = total * orc_rand_float(rng);
  *wi = V(0, 0, 1);
  if (ksi < pc) {                                   /* reflection from the coat */
    pdf = pc / total;
    *weight = vmul(*weight, vscale(b->Kc, 1.0f / pdf));
    if (b->Kc_w < ORC_FLT_EPS) {
      *weight = vmul(*weight, cf);
      *wi = V(-wo.x, -wo.y, wo.z);
      pdf = ORC_MAXFLOAT;
    } else {
      *weight = vmul(*weight, sample_ggx(wo, wi, b->Fc, b->Kc_w, &pdf, rng, two_sided));
    }
  } else if (ksi < total) {                         /* base layer, seen through the coat */
    *weight = vmul(*weight, ct);
    if (ksi < pc + pd) {                            /* diffuse */
      pdf = pd / total;
      *weight = vmul(*weight, vscale(b->Kd, 1.0f / pdf));
      *weight = vmul(*weight, sample_lambert(wo, wi, &pdf, rng, two_sided));
    } else if (ksi < (pc + pd) + ps) {              /* glossy / mirror */
      pdf = ps / total;
      *weight = vmul(*weight, vscale(b->Ks, 1.0f / pdf));
      if (b->Ks_w < ORC_FLT_EPS) {
        *weight = vmul(*weight, fresnel_media(wo.z, b->Fb));
        *wi = V(-wo.x, -wo.y, wo.z);
        pdf = ORC_MAXFLOAT;
      } else {
        *weight = vmul(*weight, sample_ggx(wo, wi, b->Fb, b->Ks_w, &pdf, rng, two_sided));
      }
    } else {                                        /* specular transmission */
      pdf = pt / total;
      *weight = vmul(*weight, vscale(b->Kt, 1.0f / pdf));
      float index = b->Fc.x > -2.5f ? 1.0f : b->Fc.y;
      *wi = transmitted(index, wo);
      *inside = !*inside;
      pdf = ORC_MAXFLOAT;
    }
  }
  if (!(total >= ORC_FLT_EPS)) *weight = V(0, 0, 0);
  return pdf;
}

void orc_bsdf_eval(const crt_bsdf* b, const float wi[3], const float wo[3], int two_sided, float out[3])
{
  bsdf_t B = bsdf_load(b);
  v3 r = eval_bsdf_layered(&B, V(wi[0], wi[1], wi[2]), V(wo[0], wo[1], wo[2]), two_sided);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

float orc_bsdf_pdf(const crt_bsdf* b, const float wo[3], const float wi[3], const float weight[3])
{
  bsdf_t B = bsdf_load(b);
  return bsdf_pdf_layered(&B, V(wo[0], wo[1], wo[2]), V(wi[0], wi[1], wi[2]), V(weight[0], weight[1], weight[2]));
}

float orc_bsdf_sample(const crt_bsdf* b, const float wo[3], float wi[3], float weight[3],
                      int* inside, uint32_t* rng, int two_sided)
{
  bsdf_t B = bsdf_load(b);
  v3 w = V(weight[0], weight[1], weight[2]), o;
  float pdf = sample_bsdf_layered(&B, V(wo[0], wo[1], wo[2]), &o, &w, inside, rng, two_sided);
  wi[0] = o.x; wi[1] = o.y; wi[2] = o.z;
  weight[0] = w.x; weight[1] = w.y; weight[2] = w.z;
  return pdf;
}

/* ------------------------------------------------------------ local space */

typedef struct { v3 x, y, z; } frame_t;

/* buildLocalSpace, SURVEY A.6: the larger of (n.z,0,-n.x) / (0,-n.z,n.y). */
static frame_t build_frame(v3 n)
{
  v3 ax = V(n.z, 0.0f, -n.x);
  v3 ay = V(0.0f, -n.z, n.y);
  float lx = dot3(ax, ax), ly = dot3(ay, ay);
  frame_t f;
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
static v3 to_local(v3 v, const frame_t* f) { return V(dot3(v, f->x), dot3(v, f->y), dot3(v, f->z)); }
static v3 from_local(v3 v, const frame_t* f)
{
  return vadd(vadd(vscale(f->x, v.x), vscale(f->y, v.y)), vscale(f->z, v.z));
}

/* ----------------------------------------------------------- lights / env */

static float cone_pdf(float cos_max) { return 1.0f / (ORC_2PI - cos_max * ORC_2PI); }

/* textureLod(sampler, st, 0) with GL_LINEAR / GL_REPEAT, t = 0 at the bottom row of the image.
 * Returns rgba in [0,1]. */
static void tex_lookup(const orc_scene* s, uint32_t tex, float u, float v, float out[4])
{
  const uint32_t off = s->tex_table[3 * tex];
  const int w = (int)s->tex_table[3 * tex + 1], h = (int)s->tex_table[3 * tex + 2];
  float fx = fmaf(u, (float)w, -0.5f);
  float fy = fmaf(1.0f - v, (float)h, -0.5f);
  float flx = floorf(fx), fly = floorf(fy);
  float ax = fx - flx, ay = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % w) + w) % w; x1 = ((x1 % w) + w) % w;
  y0 = ((y0 % h) + h) % h; y1 = ((y1 % h) + h) % h;
  const uint8_t* t = s->tex_data + 4 * (size_t)off;
  for (int c = 0; c < 4; ++c) {
    float c00 = (float)t[4 * (y0 * w + x0) + c] * (1.0f / 255.0f), c10 = (float)t[4 * (y0 * w + x1) + c] * (1.0f / 255.0f);
    float c01 = (float)t[4 * (y1 * w + x0) + c] * (1.0f / 255.0f), c11 = (float)t[4 * (y1 * w + x1) + c] * (1.0f / 255.0f);
    float top = c00 * (1.0f - ax) + c10 * ax;
    float bot = c01 * (1.0f - ax) + c11 * ax;
    out[c] = top * (1.0f - ay) + bot * ay;
  }
}

/* Latlong + bilinear fetch, SURVEY A.7 (Z-up, V3d_XposYnegZpos at AppViewer.cxx:610):
 * u = (atan2(d.y, d.x) + pi) / 2pi, v = acos(d.z) / pi, so +Z looks at the top
 * row (row 0) of the image.  Bilinear, wrap in u, clamp in v. */
static v3 env_lookup(const orc_scene* s, v3 d)
{
  if (!s->env) return V(0, 0, 0);
  float u = (orc_atan2(d.y, d.x) + ORC_PI) * ORC_INV_2PI;
  float v = orc_acos(d.z) * ORC_INV_PI;                 /* 0 at +Z = top row */
  float fx = fmaf(u, (float)s->env_w, -0.5f);
  float fy = fmaf(v, (float)s->env_h, -0.5f);
  float flx = floorf(fx), fly = floorf(fy);
  float ax = fx - flx, ay = fy - fly;
  int x0 = (int)flx, y0 = (int)fly;
  int w = (int)s->env_w, h = (int)s->env_h;
  int x1 = x0 + 1, y1 = y0 + 1;
  x0 = ((x0 % w) + w) % w; x1 = ((x1 % w) + w) % w;
  y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0);
  y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
  const float* e = s->env;
  v3 c00 = ld3(e, y0 * w + x0), c10 = ld3(e, y0 * w + x1);
  v3 c01 = ld3(e, y1 * w + x0), c11 = ld3(e, y1 * w + x1);
  v3 top = vadd(vscale(c00, 1.0f - ax), vscale(c10, ax));
  v3 bot = vadd(vscale(c01, 1.0f - ax), vscale(c11, ax));
  return vadd(vscale(top, 1.0f - ay), vscale(bot, ay));
}

/* IntersectLight, SURVEY A.7: implicit hit of light shapes along the ray up to
 * hit_t; on a miss with no light found returns the environment / background. */
static v3 intersect_light(const orc_scene* s, v3 o, v3 d, int depth, float hit_dist, float* pdf_out)
{
  v3 total = V(0, 0, 0);
  float pdf = 0.0f;
  float inv_n = s->n_lights ? 1.0f / (float)s->n_lights : 0.0f;
  int miss = hit_dist == ORC_MAXFLOAT;
  for (uint32_t i = 0; i < s->n_lights; ++i) {
    const float* L = s->lights + 8 * i;
    if (L[7] != 0.0f) {                                   /* positional: sphere of radius L[3] */
      v3 to = vsub(V(L[4], L[5], L[6]), o);
      float dist = sqrtf(dot3(to, to));
      if (dist < hit_dist) {
        float cos_max = 1.0f / sqrtf(1.0f + (L[3] * L[3]) / (dist * dist));
        if (cos_max < 1.0f && dot3(d, vscale(to, 1.0f / dist)) >= cos_max) {
          hit_dist = dist;
          total = V(L[0], L[1], L[2]);
          pdf = inv_n * cone_pdf(cos_max);
        }
      }
    } else if (hit_dist == ORC_MAXFLOAT) {                /* directional: cone of cos L[3] */
      if (L[3] < 1.0f && dot3(d, V(L[4], L[5], L[6])) >= L[3]) {
        total = vadd(total, V(L[0], L[1], L[2]));
        pdf += inv_n * cone_pdf(L[3]);
      }
    }
  }
  if (pdf == 0.0f && miss && hit_dist == ORC_MAXFLOAT) {
    if (depth == 0 && !(s->params.env_as_background && s->env)) {
      total = V(s->params.background[0], s->params.background[1], s->params.background[2]);
    } else {
      total = env_lookup(s, d);
    }
  }
  *pdf_out = pdf;
  return total;
}

/* SampleLight, SURVEY A.7: uniform cone around the light direction. */
static v3 sample_light(v3 to_light, float dist, int infinite, float smooth, float* pdf, uint32_t* rng)
{
  frame_t f = build_frame(vscale(to_light, 1.0f / dist));
  float cos_max = infinite ? smooth : 1.0f / sqrtf(1.0f + (smooth * smooth) / (dist * dist));
  float k1 = orc_rand_float(rng);
  float k2 = orc_rand_float(rng);
  float tz = 1.0f - k2 * (1.0f - cos_max);
  float sn, cs;
  orc_sincos2pi(k1, &sn, &cs);
  float r = sqrtf(maxf(1.0f - tz * tz, 0.0f));
  *pdf = (cos_max < 1.0f) ? *pdf * cone_pdf(cos_max) : ORC_MAXFLOAT;
  return normalize3(from_local(V(cs * r, sn * r, tz), &f));
}

/* ------------------------------------------------------------- path trace */

static const crt_bsdf k_default_bsdf = {
  { 0, 0, 0, 0 }, { 0.8f, 0.8f, 0.8f, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
  { -1, 0, 0, 0 }, { -1, 0, 1, 0 }, { 0, 0, 0, 0 }
};

static inline int any_gt(v3 a, float s) { return a.x > s || a.y > s || a.z > s; }
static inline int all_lt(v3 a, float s) { return a.x < s && a.y < s && a.z < s; }

/* PathTrace, SURVEY A.1/A.6/A.7.  One radiance sample along the given ray. */
static v3 path_trace(const orc_scene* s, v3 org, v3 dir, uint32_t* rng, crt_stats* st)
{
  v3 radiance = V(0, 0, 0), thr = V(1, 1, 1);
  int inside = 0;
  float exp_pdf = 1.0f, imp_pdf = 1.0f;
  const float eps = s->hdr.scene_eps;
  const int two_sided = s->params.two_sided;
  for (int depth = 0; depth < s->params.max_depth; ++depth) {
    hit_t hit;
    int found = traverse(s, org, dir, ORC_MAXFLOAT, 0, &hit, st);
    if (st) st->rays_nearest++;

    v3 le = intersect_light(s, org, dir, depth, hit.t, &exp_pdf);
    if (any_gt(le, 0.0f) || !found) {
      float mis = (depth == 0 || imp_pdf == ORC_MAXFLOAT) ? 1.0f
                : imp_pdf * imp_pdf / (exp_pdf * exp_pdf + imp_pdf * imp_pdf);
      radiance = vadd(radiance, vscale(vmul(thr, le), mis));
      break;
    }

    /* geometric normal to world space: rows of the inverse matrix, transposed */
    const float* m = s->inst_inv + 16 * hit.inst;
    v3 c0 = V(m[0], m[4], m[8]), c1 = V(m[1], m[5], m[9]), c2 = V(m[2], m[6], m[10]);
    v3 ng = normalize3(V(dot3(c0, hit.n), dot3(c1, hit.n), dot3(c2, hit.n)));
    org = vadd(org, vscale(dir, hit.t));

    /* SmoothNormal, SURVEY A.4: (1-u-v) n0 + u n1 + v n2 */
    const int32_t* tr = s->tris + 4 * hit.tri;
    int voff = hit.voff;
    v3 n0 = ld3(s->vert_nrm, voff + tr[0]), n1 = ld3(s->vert_nrm, voff + tr[1]), n2 = ld3(s->vert_nrm, voff + tr[2]);
    v3 ns = vadd(vadd(vscale(n1, hit.u), vscale(n2, hit.v)), vscale(n0, (1.0f - hit.u) - hit.v));
    ns = normalize3(ns);
    ns = normalize3(V(dot3(c0, ns), dot3(c1, ns), dot3(c2, ns)));
    frame_t frame = build_frame(ns);

    uint32_t mat_id = (uint32_t)s->inst_meta[4 * hit.inst];
    const crt_bsdf* mat = mat_id < s->n_mats ? s->mats + mat_id : &k_default_bsdf;
    bsdf_t B = bsdf_load(mat);
    if (st) st->shaded_hits++;

    /* base-colour texture (USE_TEXTURES path of PathTrace; SmoothUV): Kd *= rgb^2 * a,
     * alpha < 1 mixes in transmission.  Kd.w = texture index + 1 (0 = none), Kt.w / Le.w = S / T scale. */
    if (mat->Kd[3] >= 1.0f && (uint32_t)mat->Kd[3] - 1u < s->n_tex) {
      const float* uv0 = s->vert_uv + 2 * (voff + tr[0]);
      const float* uv1 = s->vert_uv + 2 * (voff + tr[1]);
      const float* uv2 = s->vert_uv + 2 * (voff + tr[2]);
      float w0 = (1.0f - hit.u) - hit.v;
      float tu = (uv1[0] * hit.u + uv2[0] * hit.v) + uv0[0] * w0;
      float tv = (uv1[1] * hit.u + uv2[1] * hit.v) + uv0[1] * w0;
      float ss = mat->Kt[3] != 0.0f ? mat->Kt[3] : 1.0f, ts = mat->Le[3] != 0.0f ? mat->Le[3] : 1.0f;
      float tc[4];
      tex_lookup(s, (uint32_t)mat->Kd[3] - 1u, tu * ss, tv * ts, tc);
      B.Kd = vmul(B.Kd, vscale(V(tc[0] * tc[0], tc[1] * tc[1], tc[2] * tc[2]), tc[3]));
      if (tc[3] != 1.0f) {
        float ia = 1.0f - tc[3];
        B.Kt = V(ia + tc[3] * B.Kt.x, ia + tc[3] * B.Kt.y, ia + tc[3] * B.Kt.z);
      }
    }

    v3 wo = to_local(V(-dir.x, -dir.y, -dir.z), &frame);

    /* self-emission */
    radiance = vadd(radiance, vmul(thr, V(mat->Le[0], mat->Le[1], mat->Le[2])));

    /* next-event estimation on one uniformly picked light */
    v3 nee_k = vadd(B.Kd, vadd(B.Ks_w > ORC_FLT_EPS ? B.Ks : V(0, 0, 0), B.Kc_w > ORC_FLT_EPS ? B.Kc : V(0, 0, 0)));
    if (s->n_lights > 0 && dot3(nee_k, thr) > 0.0f) {
      exp_pdf = 1.0f / (float)s->n_lights;
      int li = (int)(orc_rand_float(rng) * (float)s->n_lights);
      if (li > (int)s->n_lights - 1) li = (int)s->n_lights - 1;
      const float* L = s->lights + 8 * li;
      int infinite = L[7] == 0.0f;
      v3 to = infinite ? V(L[4], L[5], L[6]) : vsub(V(L[4], L[5], L[6]), org);
      float dist = sqrtf(dot3(to, to));
      v3 ldir = sample_light(to, dist, infinite, L[3], &exp_pdf, rng);
      v3 wl = to_local(ldir, &frame);
      imp_pdf = bsdf_pdf_layered(&B, wo, wl, thr);
      float mis = (exp_pdf == ORC_MAXFLOAT) ? 1.0f : exp_pdf / (exp_pdf * exp_pdf + imp_pdf * imp_pdf);
      v3 contrib = vscale(vmul(V(L[0], L[1], L[2]), eval_bsdf_layered(&B, wl, wo, two_sided)), mis);
      if (any_gt(contrib, ORC_MIN_CONTRIBUTION)) {
        float side = dot3(ng, ldir) >= 0.0f ? eps : -eps;
        v3 so = vadd(vadd(org, vscale(ldir, eps)), vscale(ng, side));
        hit_t sh;
        int occluded = traverse(s, so, ldir, infinite ? ORC_MAXFLOAT : dist, 1, &sh, st);
        if (st) st->rays_any++;
        if (!occluded) radiance = vadd(radiance, vmul(thr, contrib));
      }
    }

    /* Beer-Lambert attenuation of the segment just travelled inside a medium */
    if (inside) {
      float k = mat->Absorption[3];
      thr = vmul(thr, V(orc_exp(-hit.t * k * (1.0f - mat->Absorption[0])),
                        orc_exp(-hit.t * k * (1.0f - mat->Absorption[1])),
                        orc_exp(-hit.t * k * (1.0f - mat->Absorption[2]))));
    }

    v3 wi;
    imp_pdf = sample_bsdf_layered(&B, wo, &wi, &thr, &inside, rng, two_sided);

    float survive = any_gt(thr, ORC_MIN_THROUGHPUT) ? 1.0f : 0.0f;
    if (s->params.russian_roulette && depth >= 3)
      survive = minf(fmaf(0.0722f, thr.z, fmaf(0.7152f, thr.y, 0.2126f * thr.x)), 0.95f);
    if (orc_rand_float(rng) > survive || all_lt(thr, ORC_MIN_THROUGHPUT)) break;
    if (s->params.russian_roulette && depth >= 3) thr = vscale(thr, 1.0f / survive);

    dir = normalize3(from_local(wi, &frame));
    float side = dot3(ng, dir) >= 0.0f ? eps : -eps;
    org = vadd(vadd(org, vscale(dir, eps)), vscale(ng, side));
  }
  return radiance;
}

/* one sample of pixel (x, y): GenerateRay + PathTrace, then NaN -> 0 and the radiance clamp */
static v3 render_sample(const orc_scene* s, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint32_t frame_seed,
                        int radius, crt_stats* st)
{
  uint32_t rng = orc_seed_rand(frame_seed, x, y, w, radius);
  float jx = orc_rand_float(&rng);
  float jy = orc_rand_float(&rng);
  float la = 0.0f, lb = 0.0f;
  if (s->params.aperture_radius > 0.0f) { la = orc_rand_float(&rng); lb = orc_rand_float(&rng); }
  float o[3], d[3];
  orc_camera_ray(s, ((float)x + jx) / (float)w, ((float)y + jy) / (float)h, la, lb, o, d);
  v3 c = path_trace(s, V(o[0], o[1], o[2]), V(d[0], d[1], d[2]), &rng, st);
  float mr = s->params.max_radiance;
  c.x = (c.x != c.x) ? 0.0f : minf(c.x, mr);
  c.y = (c.y != c.y) ? 0.0f : minf(c.y, mr);
  c.z = (c.z != c.z) ? 0.0f : minf(c.z, mr);
  return c;
}

void orc_render(const orc_scene* s, uint32_t w, uint32_t h, uint64_t first_sample,
                uint32_t n_samples, float* accum4, int nthreads, crt_stats* stats)
{
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  int radius = s->params.coherent_rng ? 8 : 1;
  for (uint32_t k = 0; k < n_samples; ++k) {
    uint32_t frame_seed = orc_bullard_frame_seed(s->params.frame_seed0, first_sample + k);
    crt_stats total;
    memset(&total, 0, sizeof total);
#pragma omp parallel
    {
      crt_stats local;
      memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
      for (int64_t y = 0; y < (int64_t)h; ++y) {
        for (uint32_t x = 0; x < w; ++x) {
          v3 c = render_sample(s, x, (uint32_t)y, w, h, frame_seed, radius, stats ? &local : NULL);
          float* a = accum4 + 4 * ((size_t)y * w + x);
          a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += 1.0f;
        }
      }
      if (stats) {
#pragma omp critical
        {
          total.rays_nearest += local.rays_nearest; total.rays_any += local.rays_any;
          total.n_inner += local.n_inner; total.n_leaf += local.n_leaf; total.n_tri += local.n_tri;
          total.n_switch += local.n_switch; total.shaded_hits += local.shaded_hits;
          total.n_inner_any += local.n_inner_any; total.n_leaf_any += local.n_leaf_any;
          total.n_tri_any += local.n_tri_any; total.n_switch_any += local.n_switch_any;
          total.n_boxes += local.n_boxes; total.n_boxes_any += local.n_boxes_any;
        }
      }
    }
    if (stats) {
      stats->rays_nearest += total.rays_nearest; stats->rays_any += total.rays_any;
      stats->n_inner += total.n_inner; stats->n_leaf += total.n_leaf; stats->n_tri += total.n_tri;
      stats->n_switch += total.n_switch; stats->shaded_hits += total.shaded_hits;
      stats->n_inner_any += total.n_inner_any; stats->n_leaf_any += total.n_leaf_any;
      stats->n_tri_any += total.n_tri_any; stats->n_switch_any += total.n_switch_any;
      stats->n_boxes += total.n_boxes; stats->n_boxes_any += total.n_boxes_any;
      stats->samples += (uint64_t)w * h;
    }
  }
}

/* Adaptive screen sampling (Graphic3d_RenderingParams::AdaptiveScreenSampling, SettingsWidget.cxx:427-478),
 * the specification the CUDA kernels k_adaptive_allocate / k_generate_adaptive / k_resolve_adaptive follow:
 *   tiles of 32x32 pixels, NT of them; per wave a budget of B tile samples;
 *   W0 = sum err;  w_j = min(max(err_j, W0 / (8 NT)), 4 W0 / NT) + 1;  C = exclusive prefix of w, Wt its total;
 *   off = (((wave * 40503) & 0xffff) * Wt) >> 16;  cum_j = (C_j * B + off) / Wt;  k_j = cum_{j+1} - cum_j;
 *   every pixel of tile j receives samples count_j .. count_j + k_j - 1 of its stream (added in that order);
 *   even_p accumulates the luminance of the even-numbered samples;
 *   err_j = sum_p floor(4096 * |sqrt(clamp01(L_all)) - sqrt(clamp01(L_even))|) for the tiles sampled in the wave.
 * State (tile_count, tile_err: NT each; even: w*h; *wave) belongs to the caller and starts zeroed. */
static float luminance(float r, float g, float b) { return fmaf(0.0722f, b, fmaf(0.7152f, g, 0.2126f * r)); }

void orc_adaptive_allocate(const uint32_t* tile_err, uint32_t nt, uint32_t budget, uint32_t wave, uint32_t* cum)
{
  uint64_t W0 = 0, Wt = 0, C = 0;
  for (uint32_t j = 0; j < nt; ++j) W0 += tile_err[j];
  uint64_t lo = W0 / (8ull * nt), hi = (4ull * W0) / nt;
  for (uint32_t j = 0; j < nt; ++j) {
    uint64_t e = tile_err[j];
    Wt += (e < lo ? lo : (e > hi ? hi : e)) + 1ull;
  }
  uint64_t off = ((uint64_t)((wave * 40503u) & 0xffffu) * Wt) >> 16;
  for (uint32_t j = 0; j < nt; ++j) {
    uint64_t e = tile_err[j];
    cum[j] = (uint32_t)((C * budget + off) / Wt);
    C += (e < lo ? lo : (e > hi ? hi : e)) + 1ull;
  }
  cum[nt] = budget;
}

void orc_render_adaptive(const orc_scene* s, uint32_t w, uint32_t h, uint64_t first_sample, uint64_t tile_samples,
                         uint64_t wave_cap, float* accum4, uint32_t* tile_count, uint32_t* tile_err, float* even,
                         uint32_t* wave, int nthreads)
{
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
  const uint32_t T = 32;
  const uint32_t ntx = (w + T - 1) / T, nty = (h + T - 1) / T, nt = ntx * nty;
  const int radius = s->params.coherent_rng ? 8 : 1;
  const uint32_t parity = (uint32_t)(first_sample & 1u);
  uint32_t* cum = (uint32_t*)malloc(sizeof(uint32_t) * (nt + 1));
  while (tile_samples > 0) {
    uint32_t budget = (uint32_t)(tile_samples < wave_cap ? tile_samples : wave_cap);
    orc_adaptive_allocate(tile_err, nt, budget, *wave, cum);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t j = 0; j < (int64_t)nt; ++j) {
      uint32_t k = cum[j + 1] - cum[j];
      if (!k) continue;
      uint32_t n_old = tile_count[j], n_new = n_old + k;
      uint32_t n_even = parity ? n_new / 2u : (n_new + 1u) / 2u;
      uint32_t x0 = ((uint32_t)j % ntx) * T, y0 = ((uint32_t)j / ntx) * T;
      uint32_t e_sum = 0;
      uint32_t* seeds = (uint32_t*)malloc(sizeof(uint32_t) * k);
      for (uint32_t t = 0; t < k; ++t) seeds[t] = orc_bullard_frame_seed(s->params.frame_seed0, first_sample + n_old + t);
      for (uint32_t y = y0; y < y0 + T && y < h; ++y) {
        for (uint32_t x = x0; x < x0 + T && x < w; ++x) {
          float* a = accum4 + 4 * ((size_t)y * w + x);
          float ev = even[(size_t)y * w + x];
          for (uint32_t t = 0; t < k; ++t) {
            v3 c = render_sample(s, x, y, w, h, seeds[t], radius, NULL);
            a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += 1.0f;
            if (((parity + n_old + t) & 1u) == 0u) ev += luminance(c.x, c.y, c.z);
          }
          even[(size_t)y * w + x] = ev;
          if (n_new >= 2u && n_even > 0u) {
            float l_all = luminance(a[0], a[1], a[2]) / (float)n_new;
            float l_even = ev / (float)n_even;
            float e = fabsf(sqrtf(minf(maxf(l_all, 0.0f), 1.0f)) - sqrtf(minf(maxf(l_even, 0.0f), 1.0f)));
            e_sum += (uint32_t)(e * 4096.0f);
          }
        }
      }
      free(seeds);
      tile_err[j] = e_sum;
      tile_count[j] = n_new;
    }
    (*wave)++;
    tile_samples -= budget;
  }
  free(cum);
}

/* Display.fs restated: mean * 2^exposure, optional filmic curve with
 * white point, gamma 2 (sqrt), RGB8, bottom-up rows. */
static float filmic(float c)
{
  float f = fmaf(1.425f, c, 0.05f);
  return (fmaf(c, f, 0.004f)) / (fmaf(c, f + 0.55f, 0.0491f)) - 0.0821f;
}

void orc_display(const orc_scene* s, const float* accum4, uint32_t w, uint32_t h, uint8_t* rgb8)
{
  float ex = exp2f(s->params.exposure);
  float wp = s->params.tone_map ? filmic(s->params.white_point) : 1.0f;
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const float* a = accum4 + 4 * i;
    float inv = a[3] > 0.0f ? 1.0f / a[3] : 0.0f;
    for (int c = 0; c < 3; ++c) {
      float x = a[c] * inv * ex;
      if (s->params.tone_map) x = filmic(x) / wp;
      x = sqrtf(maxf(x, 0.0f));
      x = minf(x, 1.0f);
      rgb8[3 * i + c] = (uint8_t)(int)fmaf(x, 255.0f, 0.5f);
    }
  }
}

void orc_hdr(const float* accum4, uint32_t w, uint32_t h, float* rgb32f)
{
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const float* a = accum4 + 4 * i;
    float inv = a[3] > 0.0f ? 1.0f / a[3] : 0.0f;
    rgb32f[3 * i + 0] = a[0] * inv; rgb32f[3 * i + 1] = a[1] * inv; rgb32f[3 * i + 2] = a[2] * inv;
  }
}
