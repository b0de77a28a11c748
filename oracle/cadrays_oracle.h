/*
 * cadrays_oracle.h -- CPU restatement (ORACLE) of the OCCT path-tracing hot path
 * that CADRays drives.  TEST INFRASTRUCTURE ONLY: nothing in the product
 * (cadrays_b200/, libcadrays_b200.so) may include, link or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and there only as the checker or the reported CPU baseline.
 *
 * PARITY UNPINNED: the algorithm lives in Open CASCADE Technology (TKOpenGl +
 * src/Shaders, the .fs files), which is an unpinned external dependency of CADRays
 * ("Current OCCT development snapshot", README.md:51) and is absent from
 * /root/reference.  The reference ships no golden vectors for this path
 * (testing/CADRays_Testing.py compares against an unshipped template folder).
 * This file restates the published behaviour of that path tracer as described in
 * SURVEY.md Appendix A and as recoverable from CADRays' call sites; each function
 * cites the call site / appendix paragraph it follows.
 */
#ifndef CADRAYS_ORACLE_H
#define CADRAYS_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/cadrays_b200.h"   /* POD boundary types only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* Parses the blob written by crt_bvh_export (layout: DESIGN.md "BVH blob"). */
orc_scene* orc_scene_from_blob(const void* blob, size_t size);
void       orc_scene_free(orc_scene* s);

void orc_set_materials(orc_scene* s, const crt_bsdf* m, uint32_t n);
void orc_set_lights(orc_scene* s, const crt_light* l, uint32_t n);
void orc_set_envmap_rgb32f(orc_scene* s, const float* rgb, uint32_t w, uint32_t h);
void orc_set_envmap_rgb8(orc_scene* s, const uint8_t* rgb, uint32_t w, uint32_t h);
void orc_set_textures(orc_scene* s, const uint8_t* rgba8_texels, const uint32_t* sizes_wh, uint32_t n);
void orc_set_params(orc_scene* s, const crt_params* p);
void orc_set_camera(orc_scene* s, const crt_camera* c);

/* SceneNearestHit / SceneAnyHit over the blob's two-level BVH.
 * Same output convention as crt_trace.  stats may be NULL. */
void orc_trace(const orc_scene* s, const float* org, const float* dir, const float* tmax,
               uint32_t n, int any_hit,
               int32_t* prim, int32_t* inst, float* t, float* u, float* v, crt_stats* stats);
/* Brute force over every instance and triangle with the same triangle test;
 * validates the BVH (builder and traversal) itself. */
void orc_trace_brute(const orc_scene* s, const float* org, const float* dir, const float* tmax,
                     uint32_t n, int any_hit,
                     int32_t* prim, int32_t* inst, float* t, float* u, float* v);

void orc_trace_events(const orc_scene* s, const float* org, const float* dir, const float* tmax, uint32_t n,
                      int any_hit, uint8_t* events, int cap, int32_t* lens);

/* PathTrace + accumulation: adds samples
 * [first_sample, first_sample + n_samples) of every pixel to accum4
 * (float4 per pixel, rgb = radiance sum, a = sample count; bottom-up rows).
 * nthreads <= 0: all OpenMP threads. */
void orc_render(const orc_scene* s, uint32_t w, uint32_t h, uint64_t first_sample,
                uint32_t n_samples, float* accum4, int nthreads, crt_stats* stats);

/* Adaptive screen sampling, the specification of the CUDA path's crt_params.adaptive_sampling (see the .c file).
 * cum has nt + 1 entries. */
void orc_adaptive_allocate(const uint32_t* tile_err, uint32_t nt, uint32_t budget, uint32_t wave, uint32_t* cum);
void orc_render_adaptive(const orc_scene* s, uint32_t w, uint32_t h, uint64_t first_sample, uint64_t tile_samples,
                         uint64_t wave_cap, float* accum4, uint32_t* tile_count, uint32_t* tile_err, float* even,
                         uint32_t* wave, int nthreads);
/* Display.fs restated: mean -> exposure -> optional filmic -> gamma 2 -> RGB8. */
void orc_display(const orc_scene* s, const float* accum4, uint32_t w, uint32_t h, uint8_t* rgb8);
void orc_hdr(const float* accum4, uint32_t w, uint32_t h, float* rgb32f);

/* ---- unit hooks for property tests ---- */
void     orc_sincos2pi(float x, float* s, float* c);
float    orc_exp(float x);
float    orc_atan2(float y, float x);
float    orc_acos(float x);
uint32_t orc_bullard_frame_seed(uint32_t seed0, uint64_t sample_index);
uint32_t orc_seed_rand(uint32_t frame_seed, uint32_t px, uint32_t py, uint32_t size_x, int radius);
float    orc_rand_float(uint32_t* state);
void     orc_fresnel(float cos_i, const float f[4], float out[3]);
/* local frame: z = shading normal.  eval returns f*|cos_i| (rgb). */
void     orc_bsdf_eval(const crt_bsdf* b, const float wi[3], const float wo[3], int two_sided, float out[3]);
float    orc_bsdf_pdf(const crt_bsdf* b, const float wo[3], const float wi[3], const float weight[3]);
/* draws from *rng; returns pdf, updates weight/inside. */
float    orc_bsdf_sample(const crt_bsdf* b, const float wo[3], float wi[3], float weight[3],
                         int* inside, uint32_t* rng, int two_sided);
void     orc_camera_ray(const orc_scene* s, float px, float py, float lens_a, float lens_b,
                        float org[3], float dir[3]);
float    orc_scene_epsilon(const orc_scene* s);

#ifdef __cplusplus
}
#endif
#endif
