// host_scene.hpp -- host-side scene store, two-level BVH build, blob and device layout.
//
// Stands in for OCCT's OpenGl_RaytraceGeometry / OpenGl_TriangleSet /
// BVH_BinnedBuilder, which CADRays triggers
// implicitly from V3d_View::Redraw() (src/Launcher/AppViewer.cxx:1047) after
// AisMesh::Compute hands over a Graphic3d_ArrayOfTriangles
// (src/ImportExport/AisMesh.cxx:372-423).  CPU code, as in the reference.
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <utility>
#include <vector>

namespace crt {

struct Mesh {
  std::vector<float>    pos;   // 3 per vertex
  std::vector<float>    nrm;   // 3 per vertex (geometric normals synthesised when absent)
  std::vector<float>    uv;    // 2 per vertex (zeros when absent)
  std::vector<uint32_t> idx;   // 3 per triangle, 0-based
  bool has_uv = false;
};

struct Instance {
  uint32_t mesh;
  uint32_t material;
  float    xf[12];             // row-major 3x4 object -> world
  bool     visible = true;     // hidden instances keep their id and records but are left out of the top-level tree
  // world box of the transformed vertices, kept with the transform and mesh it was computed for: a commit after
  // one object moved transforms the vertices of that object only
  mutable bool     box_valid = false;
  mutable uint32_t box_mesh = 0;
  mutable float    box_xf[12] = { 0 };
  mutable float    box_lo[3] = { 0 }, box_hi[3] = { 0 };
};

// Bottom-level tree of one mesh, kept across commits: moving an object or changing its material
// (ImRaytraceControls.cxx:88 SetLocation, MaterialEditor.cxx:331-337) only rebuilds the top-level tree.
struct BottomTreeNode {
  float lo[3], hi[3];
  int32_t a, b;   // inner: child node indices; leaf: first/last primitive (inclusive)
  bool leaf;
};
struct BottomTree {
  std::vector<BottomTreeNode> nodes;
  std::vector<uint32_t> order;   // BVH order -> caller's triangle index
  int depth = 0;
  bool built = false;
  uint32_t node_off = 0, vert_off = 0, tri_off = 0;   // offsets inside the blob being assembled
};

struct HostScene {
  std::vector<Mesh>     meshes;
  std::vector<Instance> instances;
  mutable std::vector<BottomTree> tree_cache;   // parallel to meshes
  mutable uint64_t trees_built = 0;             // statistics: bottom trees built so far
  // what the mesh sections of the last blob were built from (mesh ids and sizes, instance -> mesh map, tree
  // width, top-level node count): when an edit only moves instances or changes their materials, build_blob
  // rewrites the header, the top-level nodes and the instance records of the caller's blob and leaves the
  // vertex / triangle / bottom-node sections (almost all of its bytes) alone
  mutable std::vector<uint64_t> blob_signature;
  mutable uint64_t blobs_patched = 0;           // statistics: how often that shortcut was taken
};

// "CRTB" blob, version 1 -- layout documented in DESIGN.md.  64-byte header
// followed by 16-byte aligned sections.
struct BlobHeader {
  uint32_t magic, version, n_nodes, n_verts, n_tris, n_inst, n_top_nodes, flags;
  float    scene_min[3], scene_max[3], scene_eps;
  uint32_t reserved;
};
static_assert(sizeof(BlobHeader) == 64, "blob header is 64 bytes");
constexpr uint32_t kBlobMagic = 0x42545243u;

struct BlobView {
  BlobHeader     hdr;
  const int32_t* node_info;
  const float*   node_min;
  const float*   node_max;
  const float*   vert_pos;
  const float*   vert_nrm;
  const float*   vert_uv;
  const int32_t* tris;
  const float*   inst_inv;
  const int32_t* inst_meta;
};

// Builder constants (binned SAH, depth 32; top level leaf size 1).  OCCT's BVH_BinnedBuilder stops at 5
// triangles per leaf; here the default is 2: a triangle test costs the traversal kernels about 1.8 node visits (it runs
// with half the lanes of the node loop), and the smaller leaves measured 2.3 % less traversal time on config C2, 2.2 %
// on C5 instanced, 0.5 % on C5 flattened for 30 % more node memory (profiles/README.md).  CRT_LEAF_SIZE=5 restores
// OCCT's value; hits do not depend on it except for exact-distance ties.
constexpr int kBottomLeafSize = 2;
constexpr int kTopLeafSize    = 1;
constexpr int kMaxTreeDepth   = 32;
constexpr int kBottomBins     = 48;
constexpr int kTopBins        = 32;

// Builds bottom BVHs per mesh (shared by all its instances), the top BVH over
// instance world boxes, and serialises everything.  Returns false + message on
// invalid input.
// bvh_width: 2 = binary trees (default), 4 = OCCT's optional 4-wide collapse on both levels (blob flag bit 1:
// inner node info = (0, first child, child count - 1, 0), children contiguous).
bool build_blob(const HostScene& scene, std::vector<uint8_t>& blob, std::string& err, int bvh_width = 2);
bool parse_blob(const void* data, size_t size, BlobView& view, std::string& err);

// Reference encodings of child / root references in the device layout.
constexpr uint32_t kRefLeafBit = 0x80000000u;
constexpr uint32_t kRefInstBit = 0x40000000u;
constexpr int32_t  kRefNone    = 0x7fffffff;       // empty scene

struct f4 { float x, y, z, w; };

// float4 records per triangle in tri_verts: 3 = packed 48 B, 4 = padded to 64 B so that a triangle is
// one aligned 64-byte record (two 256-bit loads).
#ifndef CRT_TRI_STRIDE
#define CRT_TRI_STRIDE 3
#endif
constexpr size_t kTriStride = CRT_TRI_STRIDE;

// What the kernels walk.  Derived from the blob, never from HostScene, so that
// crt_bvh_import and crt_commit share one path.
struct DeviceLayout {
  // inner nodes only, 64 B each:
  //   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = same for child 1
  //   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (ref0, ref1, 0, 0) as int bits
  std::vector<f4> nodes;
  // kTriStride x float4 per triangle in blob order: v0.w = caller's triangle index bits,
  // v1.w = 1 if last triangle of its leaf, v2.w = 0 (+ one pad float4 when kTriStride == 4)
  std::vector<f4> tri_verts;
  // 3 x float4 per triangle: vertex normals
  std::vector<f4> tri_nrm;
  // 6 floats per triangle: (u,v) of the three vertices (zeros when the mesh has no texels)
  std::vector<float> tri_uv;
  // 4 x float4 per instance: 3 rows of the inverse matrix, then (rootRef, material, 0, 0) bits
  std::vector<f4> inst;
  int32_t top_root = kRefNone;
  uint32_t n_tris = 0, n_inst = 0;
  uint32_t n_top_inner = 0;    // nodes[0 .. n_top_inner) = top-level tree in breadth-first order (binary layout)
  // quad layout (blob flag bit 1): 8 x float4 per inner node = 4 child boxes (lo.xyz, hi.xyz each, 24 floats),
  // then (ref0..ref3) as int bits (kRefNone = no such child), then one pad float4
  bool quad = false;
  int max_depth_top = 0, max_depth_bottom = 0;
  // blob index of a mesh's bottom root node -> device reference of its converted tree (kept so that a commit
  // after an instance-only edit can re-emit the top level without touching the bottom trees)
  std::vector<std::pair<int32_t, int32_t>> mesh_root_ref;
};

bool build_device_layout(const BlobView& v, DeviceLayout& out, std::string& err);
// Top-level nodes + instance records only (out.nodes = the n_top_inner top nodes, out.inst, out.top_root), for a
// blob that differs from the one `prev` was built from in its header, top-level nodes and instance records only.
// Returns false when the shortcut does not apply (4-wide layout, different top-level node count, unknown mesh):
// the caller then builds the full layout.
bool build_device_layout_top(const BlobView& v, const DeviceLayout& prev, DeviceLayout& out, std::string& err);

}  // namespace crt
