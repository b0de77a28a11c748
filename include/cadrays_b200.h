/*
 * cadrays_b200.h -- C-ABI of libcadrays_b200.so
 *
 * B200-native (sm_100a) replacement for the path-tracing hot path that the
 * CADRays application drives inside Open CASCADE Technology (OCCT):
 * progressive PathTrace loop, double-layer Graphic3d_BSDF sampling, direct
 * light sampling and two-level BVH traversal.
 *
 * CADRays has no FFI seam for its renderer; it reaches the path tracer only
 * through OCCT C++ calls.  Every entry point below therefore cites the
 * reference call site it stands in for (paths relative to the CADRays
 * repository, "file:line").  An OCCT-enabled host would call these from inside
 * OpenGl_View::raytrace(); INTEGRATION.md shows the adapter.
 *
 * Conventions
 *   - every function returns int: 0 = CRT_OK, negative = crt_status error class
 *     (reference convention: TCL commands return 1 and print a message,
 *     src/ImportExport/ImportExportPlugin.cxx:53-65; BufferDump returns bool,
 *     src/Launcher/AppGui.cxx:430-433).  No exception crosses the ABI.
 *   - crt_last_error() returns a thread-local UTF-8 message for the last
 *     failing call made on the calling thread.
 *   - plain pointers and sizes only; the caller keeps ownership of all host
 *     arrays, the library copies before returning.
 *   - a context is single-owner and not thread-safe (the reference renders from
 *     the one thread that owns the GL context, src/Launcher/AppViewer.cxx:593).
 *   - there is NO CPU fallback: crt_create fails with CRT_ERR_NO_DEVICE when no
 *     sm_100-class CUDA device is usable.
 */
#ifndef CADRAYS_B200_H
#define CADRAYS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRT_ABI_VERSION 3

typedef enum crt_status {
  CRT_OK               =  0,
  CRT_ERR_INVALID_ARG  = -1,  /* null pointer, out-of-range id, bad size        */
  CRT_ERR_NO_DEVICE    = -2,  /* no usable CUDA device / wrong architecture     */
  CRT_ERR_CUDA         = -3,  /* a CUDA runtime call failed                     */
  CRT_ERR_OUT_OF_MEMORY= -4,
  CRT_ERR_STATE        = -5,  /* call order violated (e.g. render before commit)*/
  CRT_ERR_FORMAT       = -6   /* malformed BVH blob                             */
} crt_status;

typedef struct crt_context crt_context;

/* ---------------------------------------------------------------------------
 * Graphic3d_BSDF as CADRays fills it (src/Launcher/MaterialEditor.cxx:281-331,
 * src/ImportExport/ImportExport.cxx:155-231).  128 bytes, 8 x vec4.
 *   Kc  rgb = coat specular weight,   w = coat roughness
 *   Kd  rgb = base diffuse weight,    w = base-colour texture: id + 1, 0 = none (texture id in OCCT)
 *   Ks  rgb = base specular weight,   w = base roughness
 *   Kt  rgb = specular transmission,  w = texture S scale (0 = 1; "rttexture -scale S T")
 *   Le  rgb = emitted radiance,       w = texture T scale (0 = 1)
 *   FresnelCoat / FresnelBase: Graphic3d_Fresnel::Serialize() encoding
 *       Schlick    ( r,  g,  b, .)  with r >= 0
 *       Constant   (-1,  ., f,  .)
 *       Conductor  (-2,  n, k,  .)
 *       Dielectric (-3, ior, ., .)
 *     (src/Launcher/MaterialEditor.cxx:209-255, ImportExport.cxx:204-227)
 *   Absorption rgb = transmitted colour, w = Beer-Lambert coefficient >= 0
 * ------------------------------------------------------------------------- */
typedef struct crt_bsdf {
  float Kc[4];
  float Kd[4];
  float Ks[4];
  float Kt[4];
  float Le[4];
  float FresnelCoat[4];
  float FresnelBase[4];
  float Absorption[4];
} crt_bsdf;

/* Light source (V3d_DirectionalLight / V3d_PositionalLight as CADRays edits
 * them, src/Launcher/LightSourcesEditor.cxx:242-310; TCL "vlight add ...",
 * data/scripts/CornellBox.tcl:11-14, data/scripts/Materials.tcl:202-203).
 *   emission   colour * intensity
 *   smoothness directional: cone half-angle in radians (<= pi/2)
 *              positional : sphere radius in world units
 *   posdir     directional: direction the light travels (as "vlight direction")
 *              positional : world position
 *   is_point   0 = directional (infinite), 1 = positional */
typedef struct crt_light {
  float   emission[3];
  float   smoothness;
  float   posdir[3];
  int32_t is_point;
} crt_light;

/* The Graphic3d_RenderingParams fields CADRays drives
 * (src/Launcher/SettingsWidget.cxx:65-90,217-229,263-478). */
typedef struct crt_params {
  int32_t  max_depth;        /* RaytracingDepth (1..32), SettingsWidget.cxx:310-316 */
  float    max_radiance;     /* RadianceClampingValue, :318-326                     */
  int32_t  two_sided;        /* TwoSidedBsdfModels, :328-334                        */
  int32_t  coherent_rng;     /* CoherentPathTracingMode (8x8 shared seeds), :419-425*/
  float    aperture_radius;  /* CameraApertureRadius, :217-222                      */
  float    focal_dist;       /* CameraFocalPlaneDist, :224-229; AppGui.cxx:91       */
  int32_t  tone_map;         /* 0 = Disabled, 1 = Filmic, :348-403                  */
  float    white_point;      /* WhitePoint                                          */
  float    exposure;         /* Exposure (stops)                                    */
  int32_t  env_as_background;/* UseEnvironmentMapBackground, LightSourcesEditor.cxx:359-364 */
  uint32_t frame_seed0;      /* seed of the per-frame generator (math_BullardGenerator) */
  int32_t  russian_roulette; /* 1 = roulette after depth 3             */
  float    background[3];    /* colour shown by primary rays that miss when the map is hidden */
  int32_t  samples_per_batch;/* samples per pixel kept in flight per wave; 0 = auto */
  int32_t  bvh_width;        /* 0 or 2 = binary BVH (default); 4 = OCCT's optional QUAD_BVH collapse.
                                Changing it rebuilds the scene at the next crt_commit. */
  int32_t  adaptive_sampling;/* AdaptiveScreenSampling, SettingsWidget.cxx:70,427-436; `vrenderparams -iss`
                                (CornellBox.tcl:78-79): 32x32-pixel screen tiles are sampled in proportion to
                                their estimated visual error (see crt_render) */
  int32_t  adaptive_tiles;   /* NbRayTracingTiles ("GPU load", SettingsWidget.cxx:72,471-476): tile samples one
                                crt_render sample unit spends in adaptive mode; 0 = one per screen tile */
} crt_params;

/* Graphic3d_Camera as CADRays sets it (src/Launcher/AppViewer.cxx:947,993-1042;
 * src/Launcher/SettingsWidget.cxx:185,202-210). */
typedef struct crt_camera {
  float   eye[3];
  float   dir[3];       /* view direction, need not be normalised */
  float   up[3];
  float   fovy_deg;     /* perspective only */
  float   aspect;       /* width / height   */
  int32_t is_ortho;
  float   ortho_scale;  /* full view height in world units (vviewparams -size) */
} crt_camera;

/* Work counters of the traversal kernels since the last crt_stats_reset.
 * Filled only while crt_stats_enable(ctx, 1) is active (instrumented build of
 * the SAME kernels); used to compute the algorithmic bytes of SURVEY 8(d):
 *   bytes = 16*n_inner + 24*n_boxes + 16*n_leaf + 52*n_tri + 64*n_switch  (+ shading)
 *         = 64*n_inner + ... for binary trees (2 boxes per inner visit). */
typedef struct crt_stats {
  uint64_t rays_nearest;
  uint64_t rays_any;
  uint64_t n_inner;       /* closest-hit rays: inner-node visits              */
  uint64_t n_leaf;        /*                   leaf visits                    */
  uint64_t n_tri;         /*                   triangle tests                 */
  uint64_t n_switch;      /*                   instance (level) switches      */
  uint64_t shaded_hits;   /* surface interactions that fetched a BSDF record */
  uint64_t samples;       /* finished path samples                            */
  uint64_t n_inner_any;   /* the same four for any-hit (shadow) rays          */
  uint64_t n_leaf_any;
  uint64_t n_tri_any;
  uint64_t n_switch_any;
  uint64_t n_boxes;       /* child boxes tested by closest-hit rays (2 per binary node, 2..4 per quad node) */
  uint64_t n_boxes_any;   /* the same for any-hit rays */
} crt_stats;

/* -------------------------------- lifetime -------------------------------- */

/* Stands in for creating the OCCT driver/viewer/view
 * (src/Launcher/AppViewer.cxx:601-638).  One context per GPU. */
int  crt_create(int device_ordinal, crt_context** out_ctx);
/* Scene assembly without a GPU: meshes/instances + crt_commit (host BVH build, as
 * OCCT does on the CPU) + crt_bvh_export work; every entry point that needs the
 * device returns CRT_ERR_NO_DEVICE.  Used to test the host logic. */
int  crt_create_host_only(crt_context** out_ctx);
void crt_destroy(crt_context* ctx);
const char* crt_last_error(void);
int  crt_abi_version(void);

/* ------------------------------- scene input ------------------------------ */

/* Graphic3d_ArrayOfTriangles(nVerts, nTris*3, normals, no colours, texels) with
 * AddVertex(pos, normal, uv) / AddEdge(index) -- src/ImportExport/AisMesh.cxx:372-413.
 * idx is 0-based here (AddEdge is 1-based).  nrm may be NULL (geometric normals
 * are used), uv may be NULL. */
int crt_mesh_create(crt_context* ctx, const float* pos, const float* nrm, const float* uv,
                    uint32_t n_verts, const uint32_t* idx, uint32_t n_tris, uint32_t* out_mesh_id);

/* One displayed AIS object: a mesh, its gp_Trsf (SetLocation,
 * src/ImGui/ImRaytraceControls.cxx:88, src/ImportExport/ImportExportPlugin.cxx:946-948)
 * and its material aspect (one per group, AisMesh.cxx:351).
 * xf is a row-major 3x4 object-to-world matrix; NULL = identity. */
int crt_instance_add(crt_context* ctx, uint32_t mesh_id, const float xf[12],
                     uint32_t material_id, uint32_t* out_inst_id);
int crt_instance_set_transform(crt_context* ctx, uint32_t inst_id, const float xf[12]);
int crt_instance_set_material(crt_context* ctx, uint32_t inst_id, uint32_t material_id);
/* AIS_InteractiveContext::Erase / Display of an object that stays in the scene tree (the eye toggle of CADRays'
 * scene panel; `verase` / `vdisplay` in scripts): a hidden instance keeps its id, transform and material but is left
 * out of the top-level tree at the next crt_commit.  Bottom-level trees are untouched. */
int crt_instance_set_visible(crt_context* ctx, uint32_t inst_id, int visible);
/* vclear (data/scripts/CornellBox.tcl:8): drops meshes and instances. */
int crt_scene_clear(crt_context* ctx);

/* Graphic3d_MaterialAspect::SetBSDF + SetMaterial
 * (src/Launcher/MaterialEditor.cxx:331-337, src/ImportExport/Utils.cxx:83-93). */
int crt_materials_set(crt_context* ctx, const crt_bsdf* bsdfs, uint32_t n);

/* Graphic3d_AspectFillArea3d::SetTextureMap(Graphic3d_Texture2Dmanual) + SetTextureMapOn
 * (src/ImportExport/AisMesh.cxx:343-345, src/ImportExport/ImportExportPlugin.cxx:737-746 "rttexture").
 * RGBA8, rows top-down as in the image file; sampled bilinearly with repeat wrap at the hit's
 * interpolated texel coordinates (SmoothUV); Kd *= rgb^2 * a, a < 1 mixes in transmission.
 * A material refers to texture id through crt_bsdf.Kd[3] = id + 1. */
int crt_texture_create(crt_context* ctx, const uint8_t* rgba8, uint32_t w, uint32_t h, uint32_t* out_texture_id);
int crt_textures_clear(crt_context* ctx);

/* V3d_Viewer::SetLightOn / DelLight / UpdateLights
 * (src/Launcher/LightSourcesEditor.cxx:404-412). */
int crt_lights_set(crt_context* ctx, const crt_light* lights, uint32_t n);

/* V3d_View::SetTextureEnv(Graphic3d_TextureEnv(file))
 * (src/Launcher/LightSourcesEditor.cxx:353-354, AppGui.cxx:963).  Lat-long map,
 * row 0 = top.  rgb8 texels are linearised as (c/255)^2; rgb32f is taken as
 * linear radiance.  w = h = 0 removes the map. */
int crt_envmap_set_rgb8(crt_context* ctx, const uint8_t* rgb, uint32_t w, uint32_t h);
int crt_envmap_set_rgb32f(crt_context* ctx, const float* rgb, uint32_t w, uint32_t h);

/* View()->ChangeRenderingParams() (src/Launcher/SettingsWidget.cxx:65-90). */
int crt_params_default(crt_params* out);
int crt_params_set(crt_context* ctx, const crt_params* p);
/* Graphic3d_Camera setters (src/Launcher/AppViewer.cxx:993-1042). */
int crt_camera_set(crt_context* ctx, const crt_camera* cam);
/* FBO resize (src/Launcher/AppViewer.cxx:959-971). */
int crt_resize(crt_context* ctx, uint32_t width, uint32_t height);

/* Explicit form of OCCT's implicit state-counter invalidation
 * (updateRaytraceGeometry / uploadRaytraceData inside Redraw): builds the
 * two-level BVH on the host, uploads, resets accumulation.  No-op when clean. */
int crt_commit(crt_context* ctx);

/* --------------------------------- render --------------------------------- */

/* V3d_View::Redraw() (src/Launcher/AppViewer.cxx:1047): the reference adds one
 * sample per pixel per call; this adds n_samples.  Synchronous.  Sample s of
 * pixel p uses the random stream (frame_seed0, first_sample + s, p) regardless
 * of batching, so disjoint sample ranges on several GPUs union to the
 * single-GPU stream set. */
int crt_render(crt_context* ctx, uint32_t n_samples, uint64_t* out_total_samples);
/* Adaptive screen sampling (crt_params.adaptive_sampling; OCCT's OpenGl_TileSampler driven by
 * Graphic3d_RenderingParams::AdaptiveScreenSampling / NbRayTracingTiles, SettingsWidget.cxx:427-478).
 * With it on, one crt_render sample unit spends `adaptive_tiles` tile samples (one tile sample =
 * one more sample for every pixel of one 32x32 tile) instead of one sample for every pixel, and
 * each wave hands them to the tiles in proportion to their error estimate, clamped to
 * [1/8, 4] x the mean.  A pixel with n samples holds exactly the first n samples of its
 * non-adaptive stream.  Per-pixel sample counts are in the accumulator's 4th channel;
 * crt_adaptive_tiles_get copies the per-tile counts and error estimates (row-major tiles,
 * CRT_ADAPTIVE_TILE pixels square; either array may be NULL) - what ShowSamplingTiles
 * (SettingsWidget.cxx:443-449) displays.  Do not all-reduce a bound accumulator in place while
 * adaptive sampling runs: the error estimate is kept from this context's own samples. */
#define CRT_ADAPTIVE_TILE 32
int crt_adaptive_tiles_get(crt_context* ctx, uint32_t* counts, uint32_t* errors, uint32_t capacity,
                           uint32_t* out_tiles_x, uint32_t* out_tiles_y);
/* As crt_render but returns after enqueueing on the context stream. */
int crt_render_async(crt_context* ctx, uint32_t n_samples);
int crt_sync(crt_context* ctx);
/* Accumulation restart (OCCT resets myAccumFrames on camera/scene change,
 * src/Launcher/AppViewer.cxx:979-984); the next sample index becomes first_sample. */
int crt_reset_accumulation(crt_context* ctx, uint64_t first_sample);
/* Moves the sample cursor without clearing the buffer: the next crt_render starts
 * at sample index `next_sample` (a rank that owns several disjoint sample ranges). */
int crt_set_next_sample(crt_context* ctx, uint64_t next_sample);

/* ------------------------------- pixels out ------------------------------- */

/* Graphic3d_CView::BufferDump(Image_PixMap&, Graphic3d_BT_RGB)
 * (src/Launcher/AppViewer.cxx:1259-1262): tone-mapped RGB8, bottom-up rows.
 * stride_bytes = 0 means width*3. */
int crt_read_ldr(crt_context* ctx, uint8_t* rgb8, size_t stride_bytes);
/* BufferDump(..., Graphic3d_BT_RGB_RayTraceHdrLeft) into ImgRGBF
 * (src/Launcher/AppGui.cxx:345-350,430): mean radiance, bottom-up rows.
 * stride_bytes = 0 means width*12. */
int crt_read_hdr(crt_context* ctx, float* rgb32f, size_t stride_bytes);
/* Device pointer of the float4 accumulation buffer (rgb = radiance SUM,
 * a = sample count), width*height*16 bytes, for zero-copy / NCCL all-reduce. */
int crt_accum_device_ptr(crt_context* ctx, void** out_ptr, size_t* out_bytes);
/* Replace the accumulation buffer by caller-owned device memory (e.g. a torch
 * tensor that NCCL reduces).  NULL restores the internal buffer. */
int crt_accum_bind(crt_context* ctx, void* device_ptr, size_t bytes);
/* Tone-map an external float4 sum buffer (e.g. the all-reduced one) into the
 * context's LDR/HDR read-back path. */
int crt_read_ldr_from(crt_context* ctx, const void* device_accum, uint8_t* rgb8, size_t stride_bytes);

/* ------------------------- several GPUs, one host process ------------------------ */

/* CADRays is one C++ process that calls V3d_View::Redraw() (src/Launcher/AppViewer.cxx:1047) and
 * BufferDump (AppViewer.cxx:1259-1262) from the thread that owns the view.  A group lets that one call
 * drive every GPU of the box: `primary` (the context the host already fills with meshes, materials,
 * lights, parameters and camera -- it stays the only one the host talks to) becomes member 0, and the
 * group creates one replica context per further entry of `devices` (devices[0] must be the primary's
 * device; an ordinal may repeat, which puts several members on one GPU -- meant for testing).
 *   crt_group_commit   = crt_commit for all members: the BVH is built and converted once on the host and the
 *                        same device layout is uploaded to every GPU in parallel; an instance-only edit
 *                        re-uploads top-level nodes + instance records only, on every member.
 *   crt_group_render   = Redraw: the next n_samples sample indices are dealt out in contiguous blocks, member r
 *                        takes n/N (+1 for r < n mod N); synchronous.  Sample s of pixel p is the same random
 *                        stream on any GPU, so the members' union is the single-GPU sample set.
 *   crt_group_read_*   = BufferDump of the combined frame: by default one fused kernel per member reads its
 *                        block of rows from every member's accumulation buffer over NVLink peer addresses, adds
 *                        the sums in member order and tone-maps (exchange + Display in one pass, image
 *                        independent of the member count up to float summation order); with
 *                        CRT_GROUP_REDUCE=nccl in the environment, or without peer access, ncclReduce to member
 *                        0 + the ordinary Display pass (libnccl.so.2 is bound with dlopen at group creation).
 * Setters are called on `primary` only; every crt_group_* call replicates what changed since the last one
 * (a camera or parameter change restarts the accumulation on all members, as on one context).  Do not call
 * crt_render / crt_commit on the primary directly while it belongs to a group.  With adaptive screen sampling
 * (crt_params.adaptive_sampling) every member runs the same tile allocation from the same global error estimate and
 * renders the tile samples whose global sample index is congruent to its rank modulo the member count; after each
 * wave the members rebuild the estimate from all members' sums over peer addresses (needs peer access). */
typedef struct crt_group crt_group;
int  crt_group_create(crt_context* primary, const int* devices, int n_devices, crt_group** out_group);
void crt_group_destroy(crt_group* group);            /* destroys the replicas; `primary` stays with the caller */
int  crt_group_size(const crt_group* group);
int  crt_group_member(crt_group* group, int rank, crt_context** out_ctx);   /* for statistics / timing queries */
int  crt_group_commit(crt_group* group);
int  crt_group_render(crt_group* group, uint32_t n_samples, uint64_t* out_total_samples);
int  crt_group_reset_accumulation(crt_group* group, uint64_t first_sample);
int  crt_group_read_ldr(crt_group* group, uint8_t* rgb8, size_t stride_bytes);
int  crt_group_read_hdr(crt_group* group, float* rgb32f, size_t stride_bytes);
/* peer access between all members (1/0), whether the NCCL path is in use, device time (ms, slowest member) of
 * the exchange + Display kernel of the last crt_group_read_*; any pointer may be NULL */
int  crt_group_info(crt_group* group, int* out_peer_access, int* out_uses_nccl, double* out_last_reduce_ms);

/* ------------------------------ parity hooks ------------------------------ */

/* Batch SceneNearestHit / SceneAnyHit on caller rays (host arrays of n x 3
 * floats, tmax n floats or NULL = infinity).  Outputs (host, any may be NULL):
 * prim = caller's triangle index inside its mesh (-1 = miss), inst = instance id,
 * t, u, v.  any_hit != 0: prim is 0 for "occluded", -1 for "visible". */
int crt_trace(crt_context* ctx, const float* org, const float* dir, const float* tmax,
              uint32_t n, int any_hit,
              int32_t* prim, int32_t* inst, float* t, float* u, float* v);
/* Same with DEVICE pointers (SoA: org/dir as float4 {x,y,z,tmax-in-dir.w}); no
 * copies, asynchronous on the context stream.  Used by the bench's
 * inputs-resident traversal line. */
int crt_trace_device(crt_context* ctx, const void* org4, const void* dir4, uint32_t n,
                     int any_hit, void* hit4 /* float4 t,u,v,prim-bits */, void* inst_i32);

/* The rays the LAST wave of crt_render traced at bounce `depth` (read back from the path
 * state; n x 3 floats each, tmax n floats; any output may be NULL; out_n = how many exist, at
 * most `capacity` are copied).  kind 0: continuation rays entering bounce `depth` (closest-hit
 * queries), kind 1: the shadow rays bounce `depth` emitted (any-hit queries, tmax = light
 * distance).  The wavefront reuses its buffers from bounce to bounce, so the rays of bounce
 * `depth` are intact only when the wave ran with crt_params.max_depth == depth + 1.
 * Parity tests feed these real secondary rays to crt_trace and to the oracle. */
int crt_wavefront_rays(crt_context* ctx, int depth, int kind, float* org, float* dir, float* tmax,
                       uint32_t capacity, uint32_t* out_n);

/* Flattened two-level BVH + geometry exactly as the kernels walk it, so the CPU
 * oracle can traverse identical bytes.  Call with buf = NULL to query the size. */
int crt_bvh_export(crt_context* ctx, void* buf, size_t capacity, size_t* out_size);
int crt_bvh_import(crt_context* ctx, const void* buf, size_t size);

/* --------------------------------- metrics -------------------------------- */
int crt_stats_enable(crt_context* ctx, int on);
int crt_stats_reset(crt_context* ctx);
int crt_stats_get(crt_context* ctx, crt_stats* out);
/* Device time (ms, CUDA events on the context stream) spent in each kernel
 * family since crt_stats_reset: [0] generate, [1] extend (nearest hit),
 * [2] shade, [3] connect (any hit), [4] resolve+display, [5] whole render calls.
 * Only collected while crt_timing_enable(ctx, 1).  While it is on, a wave runs unsplit on one stream (normally its
 * two halves share the GPU on two streams): a kernel's duration is only defined while kernels do not overlap, so timed
 * renders are a few per cent slower than untimed ones. */
int crt_timing_enable(crt_context* ctx, int on);
/* Kernels this context has enqueued since crt_stats_reset (always counted). */
int crt_launch_count(crt_context* ctx, uint64_t* out_launches);
int crt_timing_get(crt_context* ctx, double ms[6], uint64_t launches[6]);
/* Bytes of the committed scene in device memory: `traversal` = what SceneNearestHit / SceneAnyHit read (nodes,
 * triangle vertices, instance records), `total` adds the shading-side arrays (vertex normals, texels, materials,
 * lights, textures, environment).  The bench sizes its L2 / HBM probe with the first. */
int crt_scene_bytes(crt_context* ctx, size_t* out_traversal, size_t* out_total);
/* How many crt_commit calls re-uploaded only the top-level nodes and the instance records (an object moved,
 * changed its material or -- with the number of visible objects unchanged -- its visibility:
 * ImRaytraceControls.cxx:88, MaterialEditor.cxx:522-523) instead of the whole device layout. */
int crt_commit_stats(crt_context* ctx, uint64_t* out_top_level_patches);
/* CUDA stream of the context as an opaque cudaStream_t. */
int crt_stream(crt_context* ctx, void** out_stream);

#ifdef __cplusplus
}
#endif
#endif /* CADRAYS_B200_H */
