// host_scene.cpp -- two-level BVH build (binned SAH), blob writer/parser and
// device layout derivation.  See host_scene.hpp for what this replaces.
#include "host_scene.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <cstdio>
#include <unordered_map>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace crt {

namespace {

struct Prim {
  float lo[3], hi[3], c[3];
  uint32_t id;
};

using TreeNode = BottomTreeNode;

inline float half_area(const float* lo, const float* hi)
{
  float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  return dx * dy + dy * dz + dz * dx;
}

struct Box {
  float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
  float hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  void grow(const float* l, const float* h)
  {
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h[k]); }
  }
  void grow_pt(const float* p)
  {
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
  }
};

// Binned SAH builder in the manner of BVH_BinnedBuilder: every node
// with more than leaf_size primitives and depth < kMaxTreeDepth is split at the
// cheapest of (nbins-1) planes per axis; primitives are binned by centroid.
// A degenerate split falls back to an index median.  The two children of a node
// are allocated next to each other.
//
// build_range builds the subtree over prims[lo, hi) into `nodes` (nodes[0] = its root, child indices local to
// `nodes`, leaf ranges in global primitive indices).  With stop_count > 0 a node of at most stop_count primitives
// (other than the root) is not subdivided but recorded in `tasks` and marked a = -1 - task index: the upper part of
// a large tree is built serially this way, the recorded subtrees in parallel, and stitch_tree() renumbers
// everything into exactly the node order the serial algorithm produces (children allocated when the parent is
// processed, depth first, left first) -- the blob does not depend on the number of threads.
struct BuildTask { int lo, hi, depth; };

inline int env_int(const char* name, int dflt)
{
  const char* v = std::getenv(name);
  return v ? std::max(1, std::atoi(v)) : dflt;
}


// Runs fn(part, begin, end) on `parts` contiguous slices of [lo, hi), each on its own std::thread.  Used for the
// passes over the few very large nodes at the top of a big tree; plain threads rather than OpenMP regions because
// the regions are short and frequent, and spinning OpenMP workers between them slowed the serial code in between
// several-fold on the hosts measured.
template <class F>
void parallel_slices(int lo, int hi, int parts, F fn)
{
  std::vector<std::thread> pool;
  const long n = (long)hi - lo;
  for (int p = 1; p < parts; ++p)
    pool.emplace_back([=] { fn(p, lo + (int)(n * p / parts), lo + (int)(n * (p + 1) / parts)); });
  fn(0, lo, lo + (int)(n / parts));
  for (std::thread& t : pool) t.join();
}

inline int build_threads()
{
  static const int n = [] {
    int v = (int)std::thread::hardware_concurrency();
#ifdef _OPENMP
    v = std::min(v, omp_get_max_threads());
#endif
    return std::max(1, std::min(v, 32));
  }();
  return n;
}
constexpr int kWideNodeMin = 65536;   // primitives; see build_range

void build_range(std::vector<Prim>& prims, int lo, int hi, int depth0, int leaf_size, int nbins, std::vector<TreeNode>& nodes,
                 int& max_depth, int stop_count, std::vector<BuildTask>* tasks)
{
  nodes.clear();
  max_depth = depth0;
  if (hi <= lo) return;
  struct Work { int node, lo, hi, depth; };
  std::vector<Work> stack;
  nodes.push_back(TreeNode{});
  stack.push_back({ 0, lo, hi, depth0 });
  std::vector<int> bin_count(nbins);
  std::vector<Box> bin_box(nbins);
  std::vector<float> right_area(nbins);
  std::vector<int> right_count(nbins);

  while (!stack.empty()) {
    Work w = stack.back();
    stack.pop_back();
    max_depth = std::max(max_depth, w.depth);
    // the few nodes at the top of a large tree hold most of the primitives: their two passes (bounds, binning) run
    // on all threads with per-thread partial results; min / max / integer counts do not depend on the order
    const bool wide = (w.hi - w.lo) >= kWideNodeMin;
    Box nb, cb;
    if (wide) {
      const int parts = build_threads();
      std::vector<Box> pnb(parts), pcb(parts);
      parallel_slices(w.lo, w.hi, parts, [&](int p, int a, int b) {
        Box lnb, lcb;
        for (int i = a; i < b; ++i) { lnb.grow(prims[i].lo, prims[i].hi); lcb.grow_pt(prims[i].c); }
        pnb[p] = lnb; pcb[p] = lcb;
      });
      for (int p = 0; p < parts; ++p) { nb.grow(pnb[p].lo, pnb[p].hi); cb.grow(pcb[p].lo, pcb[p].hi); }
    } else {
      for (int i = w.lo; i < w.hi; ++i) { nb.grow(prims[i].lo, prims[i].hi); cb.grow_pt(prims[i].c); }
    }
    TreeNode nd;
    std::memcpy(nd.lo, nb.lo, 12);
    std::memcpy(nd.hi, nb.hi, 12);
    const int count = w.hi - w.lo;
    if (stop_count > 0 && w.node != 0 && count <= stop_count && count > leaf_size && w.depth < kMaxTreeDepth) {
      nd.leaf = false; nd.a = -1 - (int32_t)tasks->size(); nd.b = 0;
      nodes[w.node] = nd;
      tasks->push_back({ w.lo, w.hi, w.depth });
      continue;
    }
    if (count <= leaf_size || w.depth >= kMaxTreeDepth) {
      nd.leaf = true; nd.a = w.lo; nd.b = w.hi - 1;
      nodes[w.node] = nd;
      continue;
    }
    int best_axis = -1, best_split = -1;
    float best_cost = FLT_MAX;
    for (int axis = 0; axis < 3; ++axis) {
      const float cmin = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
      if (!(ext > 0.0f)) continue;
      const float scale = (float)nbins / ext;
      std::fill(bin_count.begin(), bin_count.end(), 0);
      std::fill(bin_box.begin(), bin_box.end(), Box{});
      if (wide) {
        const int parts = build_threads();
        std::vector<std::vector<int>> pc(parts, std::vector<int>(nbins, 0));
        std::vector<std::vector<Box>> pb(parts, std::vector<Box>(nbins));
        parallel_slices(w.lo, w.hi, parts, [&](int p, int a, int e) {
          std::vector<int> lc(nbins, 0);
          std::vector<Box> lb(nbins);
          for (int i = a; i < e; ++i) {
            int b = std::min(nbins - 1, (int)((prims[i].c[axis] - cmin) * scale));
            lc[b]++;
            lb[b].grow(prims[i].lo, prims[i].hi);
          }
          pc[p] = lc; pb[p] = lb;
        });
        for (int p = 0; p < parts; ++p)
          for (int b = 0; b < nbins; ++b) { bin_count[b] += pc[p][b]; if (pc[p][b]) bin_box[b].grow(pb[p][b].lo, pb[p][b].hi); }
      } else {
        for (int i = w.lo; i < w.hi; ++i) {
          int b = std::min(nbins - 1, (int)((prims[i].c[axis] - cmin) * scale));
          bin_count[b]++;
          bin_box[b].grow(prims[i].lo, prims[i].hi);
        }
      }
      Box acc;
      int cnt = 0;
      for (int b = nbins - 1; b > 0; --b) {
        if (bin_count[b]) acc.grow(bin_box[b].lo, bin_box[b].hi);
        cnt += bin_count[b];
        right_area[b] = cnt ? half_area(acc.lo, acc.hi) : 0.0f;
        right_count[b] = cnt;
      }
      acc = Box{};
      cnt = 0;
      for (int b = 0; b < nbins - 1; ++b) {
        if (bin_count[b]) acc.grow(bin_box[b].lo, bin_box[b].hi);
        cnt += bin_count[b];
        if (cnt == 0 || right_count[b + 1] == 0) continue;
        float cost = half_area(acc.lo, acc.hi) * (float)cnt + right_area[b + 1] * (float)right_count[b + 1];
        if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = b; }
      }
    }
    {   // experiment: SAH termination (CRT_SAH_LEAF = largest leaf, CRT_SAH_CT = node cost in units of 1/16 triangle test)
      static const int sah_leaf = std::getenv("CRT_SAH_LEAF") ? std::atoi(std::getenv("CRT_SAH_LEAF")) : 0;
      static const float sah_ct = std::getenv("CRT_SAH_CT") ? (float)std::atoi(std::getenv("CRT_SAH_CT")) / 16.0f : 1.0f;
      if (sah_leaf > 0 && count <= sah_leaf && best_axis >= 0) {
        const float pa = half_area(nb.lo, nb.hi);
        if ((float)count * pa <= best_cost + sah_ct * pa) {
          nd.leaf = true; nd.a = w.lo; nd.b = w.hi - 1;
          nodes[w.node] = nd;
          continue;
        }
      }
    }
    int mid;
    if (best_axis >= 0) {
      const float cmin = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
      const float scale = (float)nbins / ext;
      auto it = std::partition(prims.begin() + w.lo, prims.begin() + w.hi, [&](const Prim& p) {
        int b = std::min(nbins - 1, (int)((p.c[best_axis] - cmin) * scale));
        return b <= best_split;
      });
      mid = (int)(it - prims.begin());
      if (mid == w.lo || mid == w.hi) mid = w.lo + count / 2;
    } else {
      mid = w.lo + count / 2;
    }
    nd.leaf = false;
    nd.a = (int32_t)nodes.size();
    nd.b = nd.a + 1;
    nodes[w.node] = nd;
    nodes.push_back(TreeNode{});
    nodes.push_back(TreeNode{});
    stack.push_back({ nd.b, mid, w.hi, w.depth + 1 });
    stack.push_back({ nd.a, w.lo, mid, w.depth + 1 });
  }
}

void stitch_tree(const std::vector<TreeNode>& upper, const std::vector<std::vector<TreeNode>>& sub, std::vector<TreeNode>& out)
{
  out.clear();
  out.push_back(TreeNode{});
  struct Item { int32_t upper_idx, final_idx; };
  std::vector<Item> stack{ { 0, 0 } };
  while (!stack.empty()) {
    const Item it = stack.back();
    stack.pop_back();
    const TreeNode& un = upper[(size_t)it.upper_idx];
    if (!un.leaf && un.a < 0) {
      // a subtree built on its own: its root takes the id the parent reserved, its descendants follow as one block
      const std::vector<TreeNode>& local = sub[(size_t)(-1 - un.a)];
      const int32_t base = (int32_t)out.size();
      auto remap = [base](TreeNode n) { if (!n.leaf) { n.a = base + (n.a - 1); n.b = base + (n.b - 1); } return n; };
      out[(size_t)it.final_idx] = remap(local[0]);
      for (size_t j = 1; j < local.size(); ++j) out.push_back(remap(local[j]));
    } else if (un.leaf) {
      out[(size_t)it.final_idx] = un;
    } else {
      TreeNode nd = un;
      nd.a = (int32_t)out.size();
      nd.b = nd.a + 1;
      out[(size_t)it.final_idx] = nd;
      out.push_back(TreeNode{});
      out.push_back(TreeNode{});
      stack.push_back({ un.b, nd.b });
      stack.push_back({ un.a, nd.a });
    }
  }
}

constexpr int kParallelBuildMin = 200000;   // primitives; smaller trees are built by one thread (meshes run in parallel anyway)

void build_tree(std::vector<Prim>& prims, int leaf_size, int nbins, std::vector<TreeNode>& nodes, int& max_depth)
{
  const int n = (int)prims.size();
  const bool force_serial = std::getenv("CRT_BUILD_SERIAL") != nullptr;   // A/B and the determinism test
  if (n < kParallelBuildMin || force_serial) {
    build_range(prims, 0, n, 0, leaf_size, nbins, nodes, max_depth, 0, nullptr);
    return;
  }
  const bool timing = std::getenv("CRT_BUILD_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[build_tree] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t0).count());
    t0 = now;
  };
  std::vector<TreeNode> upper;
  std::vector<BuildTask> tasks;
  build_range(prims, 0, n, 0, leaf_size, nbins, upper, max_depth, std::max(16384, n / 128), &tasks);
  lap("upper levels");
  std::vector<std::vector<TreeNode>> sub(tasks.size());
  std::vector<int> sub_depth(tasks.size(), 0);
#pragma omp parallel for schedule(dynamic, 1)
  for (long t = 0; t < (long)tasks.size(); ++t)
    build_range(prims, tasks[(size_t)t].lo, tasks[(size_t)t].hi, tasks[(size_t)t].depth, leaf_size, nbins, sub[(size_t)t],
                sub_depth[(size_t)t], 0, nullptr);
  for (int d : sub_depth) max_depth = std::max(max_depth, d);
  lap("subtrees");
  stitch_tree(upper, sub, nodes);
  lap("stitch");
}

// BVH_QuadTree collapse (SURVEY A.3 "optional QUAD_BVH"; north_star: "per-mesh quad trees"): every inner
// node adopts its grandchildren, so an inner node has 2..4 children stored contiguously.  Output nodes
// reuse TreeNode with a = first child, b = child count - 1 for inner nodes; leaves are unchanged.
std::vector<TreeNode> collapse_to_quad(const std::vector<TreeNode>& bin, int& depth)
{
  std::vector<TreeNode> quad;
  depth = 0;
  if (bin.empty()) return quad;
  struct Item { int32_t bin_index, quad_index; int depth; };
  std::vector<Item> queue;
  quad.push_back(bin[0]);
  queue.push_back({ 0, 0, 0 });
  for (size_t q = 0; q < queue.size(); ++q) {
    const Item it = queue[q];
    const TreeNode& bn = bin[it.bin_index];
    depth = std::max(depth, it.depth);
    if (bn.leaf) { quad[it.quad_index] = bn; continue; }
    int32_t kids[4];
    int n = 0;
    for (int32_t c : { bn.a, bn.b }) {
      if (bin[c].leaf) kids[n++] = c;
      else { kids[n++] = bin[c].a; kids[n++] = bin[c].b; }
    }
    const int32_t first = (int32_t)quad.size();
    for (int k = 0; k < n; ++k) {
      quad.push_back(bin[kids[k]]);
      queue.push_back({ kids[k], first + k, it.depth + 1 });
    }
    TreeNode inner = bn;
    inner.leaf = false; inner.a = first; inner.b = n - 1;
    quad[it.quad_index] = inner;
  }
  return quad;
}

size_t align16(size_t x) { return (x + 15u) & ~(size_t)15u; }

// Inverse of a row-major 3x4 affine matrix, computed in double.
bool invert_affine(const float* m, float* out)
{
  double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
  double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  if (det == 0.0 || !std::isfinite(det)) return false;
  double r = 1.0 / det;
  double inv[9] = { (e * i - f * h) * r, (c * h - b * i) * r, (b * f - c * e) * r,
                    (f * g - d * i) * r, (a * i - c * g) * r, (c * d - a * f) * r,
                    (d * h - e * g) * r, (b * g - a * h) * r, (a * e - b * d) * r };
  double t[3] = { m[3], m[7], m[11] };
  for (int row = 0; row < 3; ++row) {
    out[4 * row + 0] = (float)inv[3 * row + 0];
    out[4 * row + 1] = (float)inv[3 * row + 1];
    out[4 * row + 2] = (float)inv[3 * row + 2];
    out[4 * row + 3] = (float)-(inv[3 * row + 0] * t[0] + inv[3 * row + 1] * t[1] + inv[3 * row + 2] * t[2]);
  }
  out[12] = 0.0f; out[13] = 0.0f; out[14] = 0.0f; out[15] = 1.0f;
  return true;
}

using MeshTree = BottomTree;

}  // namespace

bool build_blob(const HostScene& scene, std::vector<uint8_t>& blob, std::string& err, int bvh_width)
{
  const bool timing = std::getenv("CRT_BUILD_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[build_blob] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };

  const bool quad = bvh_width == 4;
  const size_t n_mesh = scene.meshes.size();
  bool any_visible = false;
  for (const Instance& in : scene.instances) any_visible = any_visible || in.visible;
  const size_t n_inst = any_visible ? scene.instances.size() : 0;   // nothing displayed = an empty scene
  scene.tree_cache.resize(n_mesh);
  std::vector<MeshTree>& trees = scene.tree_cache;
  std::vector<char> used(n_mesh, 0);
  for (const Instance& in : scene.instances) {
    if (in.mesh >= n_mesh) { err = "instance refers to an unknown mesh"; return false; }
    used[in.mesh] = 1;
  }

  // bottom-level trees, one per referenced mesh: large meshes one after the other (build_tree spreads each over the
  // threads), then the small ones in parallel
  auto build_mesh_tree = [&](size_t mi) {
    const Mesh& m = scene.meshes[mi];
    const size_t nt = m.idx.size() / 3;
    std::vector<Prim> prims(nt);
#pragma omp parallel for schedule(static) if (nt >= (size_t)kParallelBuildMin)
    for (long t = 0; t < (long)nt; ++t) {
      Prim& p = prims[(size_t)t];
      Box b;
      for (int k = 0; k < 3; ++k) b.grow_pt(&m.pos[3 * (size_t)m.idx[3 * (size_t)t + k]]);
      for (int k = 0; k < 3; ++k) { p.lo[k] = b.lo[k]; p.hi[k] = b.hi[k]; p.c[k] = 0.5f * (b.lo[k] + b.hi[k]); }
      p.id = (uint32_t)t;
    }
    build_tree(prims, env_int("CRT_LEAF_SIZE", kBottomLeafSize), env_int("CRT_BOTTOM_BINS", kBottomBins), trees[mi].nodes, trees[mi].depth);
    trees[mi].order.resize(nt);
    for (size_t t = 0; t < nt; ++t) trees[mi].order[t] = prims[t].id;
    trees[mi].built = true;
  };
  auto is_large = [&](size_t mi) { return scene.meshes[mi].idx.size() / 3 >= (size_t)kParallelBuildMin; };
  for (size_t mi = 0; mi < n_mesh; ++mi)
    if (used[mi] && !trees[mi].built && is_large(mi)) { build_mesh_tree(mi); scene.trees_built++; }
  bool any_small = false;
  for (size_t mi = 0; mi < n_mesh; ++mi) any_small = any_small || (used[mi] && !trees[mi].built && !is_large(mi));
#pragma omp parallel for schedule(dynamic, 1) if (any_small)
  for (long mi = 0; mi < (long)n_mesh; ++mi) {
    if (!used[mi] || trees[mi].built || is_large((size_t)mi)) continue;
    build_mesh_tree((size_t)mi);
#pragma omp atomic
    scene.trees_built++;
  }

  lap("bottom trees");
  // nodes as they go into the blob: the binary trees, or their 4-wide collapse
  std::vector<std::vector<TreeNode>> emit(n_mesh);
  for (size_t mi = 0; mi < n_mesh; ++mi) {
    if (!used[mi]) continue;
    int qd = 0;
    if (quad) emit[mi] = collapse_to_quad(trees[mi].nodes, qd);
  }
  auto mesh_nodes = [&](size_t mi) -> const std::vector<TreeNode>& { return quad ? emit[mi] : trees[mi].nodes; };

  // offsets
  uint32_t n_verts = 0, n_tris = 0, n_bottom_nodes = 0;
  bool any_uv = false;
  for (size_t mi = 0; mi < n_mesh; ++mi) {
    if (!used[mi]) continue;
    trees[mi].vert_off = n_verts;
    trees[mi].tri_off = n_tris;
    trees[mi].node_off = n_bottom_nodes;   // rebased below once the top tree size is known
    n_verts += (uint32_t)(scene.meshes[mi].pos.size() / 3);
    n_tris += (uint32_t)(scene.meshes[mi].idx.size() / 3);
    n_bottom_nodes += (uint32_t)mesh_nodes(mi).size();
    any_uv = any_uv || scene.meshes[mi].has_uv;
  }

  // top-level tree over instance world boxes.  OCCT transforms the 8 corners of the bottom root box;
  // here the box is the tight one of the transformed vertices (rotated parts get up to ~1.7x smaller
  // boxes => fewer false instance entries), padded by a few ulps because the traversal intersects in
  // object space with a differently rounded transform.  Visiting order only; hits are unaffected.
  std::vector<Prim> iprims(n_inst);
  std::vector<float> inv(16 * n_inst);
  int bad_xf = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (long k = 0; k < (long)n_inst; ++k) {
    const Instance& in = scene.instances[k];
    if (!in.visible) {
      // not displayed: never entered by a ray, so its transform is neither inverted nor checked (identity record)
      float* q = &inv[16 * k];
      for (int e = 0; e < 16; ++e) q[e] = (e % 5 == 0) ? 1.0f : 0.0f;
      iprims[k].id = (uint32_t)k;
      continue;
    }
    if (!invert_affine(in.xf, &inv[16 * k])) {
#pragma omp atomic write
      bad_xf = 1;
      continue;
    }
    const Mesh& m = scene.meshes[in.mesh];
    Box b;
    if (in.box_valid && in.box_mesh == in.mesh && std::memcmp(in.box_xf, in.xf, sizeof in.xf) == 0) {
      std::memcpy(b.lo, in.box_lo, 12); std::memcpy(b.hi, in.box_hi, 12);
    } else {
      const size_t nv = m.pos.size() / 3;
      for (size_t v = 0; v < nv; ++v) {
        const float* p = &m.pos[3 * v];
        float q[3];
        for (int r = 0; r < 3; ++r)
          q[r] = in.xf[4 * r] * p[0] + in.xf[4 * r + 1] * p[1] + in.xf[4 * r + 2] * p[2] + in.xf[4 * r + 3];
        b.grow_pt(q);
      }
      std::memcpy(in.box_lo, b.lo, 12); std::memcpy(in.box_hi, b.hi, 12);
      std::memcpy(in.box_xf, in.xf, sizeof in.xf);
      in.box_mesh = in.mesh; in.box_valid = true;
    }
    Prim& p = iprims[k];
    for (int r = 0; r < 3; ++r) {
      const float pad = 4.0e-6f * (std::max(std::fabs(b.lo[r]), std::fabs(b.hi[r])) + (b.hi[r] - b.lo[r])) + 1.0e-30f;
      p.lo[r] = b.lo[r] - pad; p.hi[r] = b.hi[r] + pad; p.c[r] = 0.5f * (b.lo[r] + b.hi[r]);
    }
    p.id = (uint32_t)k;
  }
  if (bad_xf) { err = "instance transform is singular"; return false; }
  {   // the top-level tree is built over the displayed instances only
    size_t n_vis = 0;
    for (size_t k = 0; k < n_inst; ++k)
      if (scene.instances[k].visible) iprims[n_vis++] = iprims[k];
    iprims.resize(n_vis);
  }
  lap("instance world boxes");
  std::vector<TreeNode> top;
  int top_depth = 0;
  build_tree(iprims, kTopLeafSize, env_int("CRT_TOP_BINS", kTopBins), top, top_depth);
  if (quad) top = collapse_to_quad(top, top_depth);
  const uint32_t n_top = (uint32_t)top.size();
  const uint32_t n_nodes = n_inst ? n_top + n_bottom_nodes : 0;

  BlobHeader hdr{};
  hdr.magic = kBlobMagic; hdr.version = 1;
  hdr.n_nodes = n_nodes; hdr.n_verts = n_inst ? n_verts : 0; hdr.n_tris = n_inst ? n_tris : 0;
  hdr.n_inst = (uint32_t)n_inst; hdr.n_top_nodes = n_inst ? n_top : 0;
  hdr.flags = (any_uv ? 1u : 0u) | (quad ? 2u : 0u);
  if (n_inst) {
    for (int k = 0; k < 3; ++k) { hdr.scene_min[k] = top[0].lo[k]; hdr.scene_max[k] = top[0].hi[k]; }
    float sx = top[0].hi[0] - top[0].lo[0], sy = top[0].hi[1] - top[0].lo[1], sz = top[0].hi[2] - top[0].lo[2];
    // uSceneEpsilon = max(1e-6, 1e-4 * |box size|), SURVEY A.7
    hdr.scene_eps = std::max(1.0e-6f, 1.0e-4f * std::sqrt(sx * sx + sy * sy + sz * sz));
  } else {
    hdr.scene_eps = 1.0e-6f;
  }

  size_t off = sizeof(BlobHeader);
  const size_t o_info = off; off = align16(off + (size_t)16 * hdr.n_nodes);
  const size_t o_min = off;  off = align16(off + (size_t)12 * hdr.n_nodes);
  const size_t o_max = off;  off = align16(off + (size_t)12 * hdr.n_nodes);
  const size_t o_pos = off;  off = align16(off + (size_t)12 * hdr.n_verts);
  const size_t o_nrm = off;  off = align16(off + (size_t)12 * hdr.n_verts);
  const size_t o_uv = off;   off = align16(off + (size_t)8 * hdr.n_verts);
  const size_t o_tri = off;  off = align16(off + (size_t)16 * hdr.n_tris);
  const size_t o_inv = off;  off = align16(off + (size_t)64 * hdr.n_inst);
  const size_t o_meta = off; off = align16(off + (size_t)16 * hdr.n_inst);
  lap("top tree");
  // same meshes, same instance -> mesh map, same tree shapes as the blob the caller still holds?
  std::vector<uint64_t> sig;
  sig.reserve(8 + 4 * n_mesh + n_inst);
  sig.push_back(quad); sig.push_back(n_top); sig.push_back(off); sig.push_back(any_uv);
  for (size_t mi = 0; mi < n_mesh; ++mi)
    if (used[mi]) {
      sig.push_back(mi); sig.push_back(scene.meshes[mi].pos.size()); sig.push_back(scene.meshes[mi].idx.size());
      sig.push_back(mesh_nodes(mi).size());
    }
  sig.push_back(~0ull);
  for (const Instance& in : scene.instances) sig.push_back(((uint64_t)in.visible << 32) | in.mesh);
  BlobHeader old_hdr{};
  if (blob.size() >= sizeof old_hdr) std::memcpy(&old_hdr, blob.data(), sizeof old_hdr);
  const bool patch = n_inst && blob.size() == off && sig == scene.blob_signature && old_hdr.magic == kBlobMagic &&
                     old_hdr.n_nodes == hdr.n_nodes && old_hdr.n_verts == hdr.n_verts && old_hdr.n_tris == hdr.n_tris &&
                     old_hdr.n_inst == hdr.n_inst && old_hdr.n_top_nodes == hdr.n_top_nodes && old_hdr.flags == hdr.flags &&
                     std::getenv("CRT_BLOB_NO_PATCH") == nullptr;
  if (!patch) blob.assign(off, 0);
  scene.blob_signature = n_inst ? sig : std::vector<uint64_t>{};
  std::memcpy(blob.data(), &hdr, sizeof hdr);
  if (!n_inst) return true;

  int32_t* info = reinterpret_cast<int32_t*>(blob.data() + o_info);
  float* nmin = reinterpret_cast<float*>(blob.data() + o_min);
  float* nmax = reinterpret_cast<float*>(blob.data() + o_max);
  float* pos = reinterpret_cast<float*>(blob.data() + o_pos);
  float* nrm = reinterpret_cast<float*>(blob.data() + o_nrm);
  float* uv = reinterpret_cast<float*>(blob.data() + o_uv);
  int32_t* tri = reinterpret_cast<int32_t*>(blob.data() + o_tri);
  float* binv = reinterpret_cast<float*>(blob.data() + o_inv);
  int32_t* meta = reinterpret_cast<int32_t*>(blob.data() + o_meta);

  // top-level nodes: x = 0 inner (y,z children), x = inst+1 leaf (y = bottom root, z = vertex
  // offset, w = triangle offset) -- SURVEY A.3
  for (uint32_t n = 0; n < n_top; ++n) {
    const TreeNode& t = top[n];
    std::memcpy(nmin + 3 * n, t.lo, 12);
    std::memcpy(nmax + 3 * n, t.hi, 12);
    if (t.leaf) {
      if (t.a != t.b) { err = "top-level leaf holds more than one instance (tree deeper than the builder's limit)"; return false; }
      const uint32_t k = iprims[t.a].id;   // leaf size 1
      const MeshTree& mt = trees[scene.instances[k].mesh];
      info[4 * n + 0] = (int32_t)k + 1;
      info[4 * n + 1] = (int32_t)(n_top + mt.node_off);
      info[4 * n + 2] = (int32_t)mt.vert_off;
      info[4 * n + 3] = (int32_t)mt.tri_off;
    } else {
      info[4 * n + 0] = 0; info[4 * n + 1] = t.a; info[4 * n + 2] = t.b; info[4 * n + 3] = 0;
    }
  }
  if (patch) scene.blobs_patched++;
  for (size_t mi = 0; mi < n_mesh && !patch; ++mi) {
    if (!used[mi]) continue;
    const MeshTree& mt = trees[mi];
    const Mesh& m = scene.meshes[mi];
    const uint32_t base = n_top + mt.node_off;
    const std::vector<TreeNode>& mnodes = mesh_nodes(mi);
    for (size_t n = 0; n < mnodes.size(); ++n) {
      const TreeNode& t = mnodes[n];
      std::memcpy(nmin + 3 * (base + n), t.lo, 12);
      std::memcpy(nmax + 3 * (base + n), t.hi, 12);
      int32_t* ni = info + 4 * (base + n);
      if (t.leaf) { ni[0] = -1; ni[1] = t.a; ni[2] = t.b; ni[3] = 0; }
      else        { ni[0] = 0;  ni[1] = t.a; ni[2] = t.b; ni[3] = 0; }
    }
    const size_t nv = m.pos.size() / 3;
    std::memcpy(pos + 3 * (size_t)mt.vert_off, m.pos.data(), 12 * nv);
    std::memcpy(nrm + 3 * (size_t)mt.vert_off, m.nrm.data(), 12 * nv);
    if (m.has_uv) std::memcpy(uv + 2 * (size_t)mt.vert_off, m.uv.data(), 8 * nv);
    for (size_t t = 0; t < mt.order.size(); ++t) {
      const uint32_t src = mt.order[t];
      int32_t* tr = tri + 4 * ((size_t)mt.tri_off + t);
      tr[0] = (int32_t)m.idx[3 * src]; tr[1] = (int32_t)m.idx[3 * src + 1]; tr[2] = (int32_t)m.idx[3 * src + 2];
      tr[3] = (int32_t)src;
    }
  }
  std::memcpy(binv, inv.data(), 64 * n_inst);
  for (size_t k = 0; k < n_inst; ++k) {
    const Instance& in = scene.instances[k];
    meta[4 * k + 0] = (int32_t)in.material;
    meta[4 * k + 1] = (int32_t)in.mesh;
    meta[4 * k + 2] = (int32_t)(n_top + trees[in.mesh].node_off);
    meta[4 * k + 3] = 0;
  }
  lap("blob sections");
  return true;
}

bool parse_blob(const void* data, size_t size, BlobView& v, std::string& err)
{
  if (!data || size < sizeof(BlobHeader)) { err = "blob too small"; return false; }
  std::memcpy(&v.hdr, data, sizeof(BlobHeader));
  if (v.hdr.magic != kBlobMagic || v.hdr.version != 1) { err = "bad blob magic/version"; return false; }
  const uint8_t* base = static_cast<const uint8_t*>(data);
  const BlobHeader& h = v.hdr;
  size_t off = sizeof(BlobHeader);
  v.node_info = reinterpret_cast<const int32_t*>(base + off); off = align16(off + (size_t)16 * h.n_nodes);
  v.node_min = reinterpret_cast<const float*>(base + off);    off = align16(off + (size_t)12 * h.n_nodes);
  v.node_max = reinterpret_cast<const float*>(base + off);    off = align16(off + (size_t)12 * h.n_nodes);
  v.vert_pos = reinterpret_cast<const float*>(base + off);    off = align16(off + (size_t)12 * h.n_verts);
  v.vert_nrm = reinterpret_cast<const float*>(base + off);    off = align16(off + (size_t)12 * h.n_verts);
  v.vert_uv = reinterpret_cast<const float*>(base + off);     off = align16(off + (size_t)8 * h.n_verts);
  v.tris = reinterpret_cast<const int32_t*>(base + off);      off = align16(off + (size_t)16 * h.n_tris);
  v.inst_inv = reinterpret_cast<const float*>(base + off);    off = align16(off + (size_t)64 * h.n_inst);
  v.inst_meta = reinterpret_cast<const int32_t*>(base + off); off = align16(off + (size_t)16 * h.n_inst);
  if (off > size) { err = "blob truncated"; return false; }
  if (h.n_top_nodes > h.n_nodes) { err = "blob node counts inconsistent"; return false; }
  if (h.flags & ~3u) { err = "blob: unknown flags"; return false; }
  return true;
}

namespace {

struct Converter {
  const BlobView& v;
  DeviceLayout& out;
  std::string& err;
  bool ok = true;
  std::unordered_map<int32_t, int32_t> mesh_root_ref;   // bottom root node -> device ref
  int cur_depth_max = 0;
  bool top_only = false;   // build_device_layout_top: bottom trees are on the device already, mesh_root_ref is given

  bool node_ok(int64_t n) const { return n >= 0 && n < (int64_t)v.hdr.n_nodes; }

  void set_box(f4* nd, int child, const float* lo, const float* hi)
  {
    f4& nxy = nd[child];
    nxy.x = lo[0]; nxy.y = hi[0]; nxy.z = lo[1]; nxy.w = hi[1];
    if (child == 0) { nd[2].x = lo[2]; nd[2].y = hi[2]; }
    else            { nd[2].z = lo[2]; nd[2].w = hi[2]; }
  }
  static float bits(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }

  // bottom-level subtree; node indices relative to node_off, triangles relative to tri_off
  int32_t bottom(int32_t node_abs, int32_t node_off, int32_t tri_off, int depth)
  {
    if (!ok) return kRefNone;
    if (!node_ok(node_abs) || depth > 64) { err = "blob: bottom tree malformed"; ok = false; return kRefNone; }
    cur_depth_max = std::max(cur_depth_max, depth);
    const int32_t* info = v.node_info + 4 * (size_t)node_abs;
    if (info[0] < 0) {
      const int64_t first = (int64_t)tri_off + info[1], last = (int64_t)tri_off + info[2];
      if (first < 0 || last < first || last >= (int64_t)v.hdr.n_tris) { err = "blob: leaf range"; ok = false; return kRefNone; }
      out.tri_verts[kTriStride * (size_t)last + 1].w = bits(1);
      return (int32_t)(kRefLeafBit | (uint32_t)first);
    }
    if (info[0] != 0) { err = "blob: top-level leaf inside a bottom tree"; ok = false; return kRefNone; }
    const int32_t self = (int32_t)(out.nodes.size() / 4);
    out.nodes.resize(out.nodes.size() + 4, f4{ 0, 0, 0, 0 });
    const int32_t l = node_off + info[1], r = node_off + info[2];
    if (!node_ok(l) || !node_ok(r)) { err = "blob: child index"; ok = false; return kRefNone; }
    const int32_t rl = bottom(l, node_off, tri_off, depth + 1);
    const int32_t rr = bottom(r, node_off, tri_off, depth + 1);
    f4* nd = &out.nodes[4 * (size_t)self];
    set_box(nd, 0, v.node_min + 3 * (size_t)l, v.node_max + 3 * (size_t)l);
    set_box(nd, 1, v.node_min + 3 * (size_t)r, v.node_max + 3 * (size_t)r);
    nd[3].x = bits(rl); nd[3].y = bits(rr);
    return self;
  }

  // 4-wide layout: one recursive converter for both levels (children contiguous in the blob)
  int32_t quad_node(int32_t node_abs, int32_t node_off, int32_t tri_off, bool top_level, int depth)
  {
    if (!ok) return kRefNone;
    if (!node_ok(node_abs) || depth > 64) { err = "blob: quad tree malformed"; ok = false; return kRefNone; }
    const int32_t* info = v.node_info + 4 * (size_t)node_abs;
    if (top_level) out.max_depth_top = std::max(out.max_depth_top, depth);
    else cur_depth_max = std::max(cur_depth_max, depth);
    if (info[0] > 0) {
      if (!top_level) { err = "blob: top-level leaf inside a bottom tree"; ok = false; return kRefNone; }
      return top_leaf(info);
    }
    if (info[0] < 0) {
      if (top_level) { err = "blob: bottom leaf inside the top tree"; ok = false; return kRefNone; }
      const int64_t first = (int64_t)tri_off + info[1], last = (int64_t)tri_off + info[2];
      if (first < 0 || last < first || last >= (int64_t)v.hdr.n_tris) { err = "blob: leaf range"; ok = false; return kRefNone; }
      out.tri_verts[kTriStride * (size_t)last + 1].w = bits(1);
      return (int32_t)(kRefLeafBit | (uint32_t)first);
    }
    const int k = info[2] + 1;
    if (k < 1 || k > 4) { err = "blob: quad node child count"; ok = false; return kRefNone; }
    const int32_t self = (int32_t)(out.nodes.size() / 8);
    out.nodes.resize(out.nodes.size() + 8, f4{ 0, 0, 0, 0 });
    int32_t refs[4] = { kRefNone, kRefNone, kRefNone, kRefNone };
    float boxes[24] = { 0 };
    for (int c = 0; c < k; ++c) {
      const int32_t ch = node_off + info[1] + c;
      if (!node_ok(ch)) { err = "blob: child index"; ok = false; return kRefNone; }
      refs[c] = quad_node(ch, node_off, tri_off, top_level, depth + 1);
      std::memcpy(boxes + 6 * c, v.node_min + 3 * (size_t)ch, 12);
      std::memcpy(boxes + 6 * c + 3, v.node_max + 3 * (size_t)ch, 12);
    }
    f4* nd = &out.nodes[8 * (size_t)self];
    std::memcpy(nd, boxes, sizeof boxes);
    nd[6] = f4{ bits(refs[0]), bits(refs[1]), bits(refs[2]), bits(refs[3]) };
    return self;
  }

  // instance leaf of the top tree: converts (once per mesh) the bottom tree it points to
  int32_t top_leaf(const int32_t* info)
  {
    const int32_t k = info[0] - 1;
    if (k >= (int32_t)v.hdr.n_inst) { err = "blob: instance index"; ok = false; return kRefNone; }
    int32_t ref;
    auto it = mesh_root_ref.find(info[1]);
    if (it == mesh_root_ref.end()) {
      if (top_only) { err = "top-level patch: unknown bottom tree"; ok = false; return kRefNone; }
      cur_depth_max = 0;
      ref = out.quad ? quad_node(info[1], info[1], info[3], false, 0) : bottom(info[1], info[1], info[3], 0);
      mesh_root_ref[info[1]] = ref;
      out.max_depth_bottom = std::max(out.max_depth_bottom, cur_depth_max);
    } else {
      ref = it->second;
    }
    f4* ir = &out.inst[4 * (size_t)k];
    ir[3].x = bits(ref);
    return (int32_t)(kRefLeafBit | kRefInstBit | (uint32_t)k);
  }

  // Top-level tree in breadth-first order: device nodes 0 .. n_top_inner-1 are the top tree level by
  // level, so "the first K nodes" is the top of the tree (what the kernels may stage in shared memory).
  int32_t top(int32_t root_abs)
  {
    if (!node_ok(root_abs)) { err = "blob: top tree malformed"; ok = false; return kRefNone; }
    if (v.node_info[4 * (size_t)root_abs] > 0) return top_leaf(v.node_info + 4 * (size_t)root_abs);
    std::vector<std::pair<int32_t, int>> order;          // (blob node, depth) of inner nodes, BFS
    std::unordered_map<int32_t, int32_t> index;
    order.emplace_back(root_abs, 0);
    index[root_abs] = 0;
    for (size_t q = 0; q < order.size() && ok; ++q) {
      const int32_t n = order[q].first;
      const int depth = order[q].second;
      out.max_depth_top = std::max(out.max_depth_top, depth + 1);
      const int32_t* info = v.node_info + 4 * (size_t)n;
      for (int c = 1; c <= 2; ++c) {
        const int32_t ch = info[c];
        if (!node_ok(ch) || depth > 64 || order.size() > (size_t)v.hdr.n_nodes) { err = "blob: top tree malformed"; ok = false; break; }
        const int32_t ci = v.node_info[4 * (size_t)ch];
        if (ci == 0) { index[ch] = (int32_t)order.size(); order.emplace_back(ch, depth + 1); }
        else if (ci < 0) { err = "blob: bottom leaf inside the top tree"; ok = false; break; }
      }
    }
    if (!ok) return kRefNone;
    out.n_top_inner = (uint32_t)order.size();
    out.nodes.resize(4 * order.size(), f4{ 0, 0, 0, 0 });
    for (size_t q = 0; q < order.size() && ok; ++q) {
      const int32_t* info = v.node_info + 4 * (size_t)order[q].first;
      int32_t refs[2];
      for (int c = 0; c < 2; ++c) {
        const int32_t ch = info[1 + c];
        const int32_t* ci = v.node_info + 4 * (size_t)ch;
        refs[c] = ci[0] == 0 ? index[ch] : top_leaf(ci);
      }
      f4* nd = &out.nodes[4 * q];       // re-fetch: bottom() may have grown the vector
      set_box(nd, 0, v.node_min + 3 * (size_t)info[1], v.node_max + 3 * (size_t)info[1]);
      set_box(nd, 1, v.node_min + 3 * (size_t)info[2], v.node_max + 3 * (size_t)info[2]);
      nd[3].x = bits(refs[0]); nd[3].y = bits(refs[1]);
    }
    return 0;
  }
};

}  // namespace

bool build_device_layout(const BlobView& v, DeviceLayout& out, std::string& err)
{
  out = DeviceLayout{};
  const BlobHeader& h = v.hdr;
  out.n_tris = h.n_tris;
  out.n_inst = h.n_inst;
  if (h.n_nodes == 0 || h.n_inst == 0) return true;

  // de-indexed triangles: the vertex offset of a triangle is that of the mesh whose
  // triangle range contains it, taken from the top-level leaf records
  std::vector<int32_t> tri_voff(h.n_tris, -1);
  {
    // mesh triangle ranges: sort distinct (tri_off, vert_off) pairs
    std::vector<std::pair<int32_t, int32_t>> ranges;
    for (uint32_t n = 0; n < h.n_top_nodes; ++n) {
      const int32_t* info = v.node_info + 4 * (size_t)n;
      if (info[0] > 0) ranges.emplace_back(info[3], info[2]);
    }
    std::sort(ranges.begin(), ranges.end());
    ranges.erase(std::unique(ranges.begin(), ranges.end()), ranges.end());
    for (size_t r = 0; r < ranges.size(); ++r) {
      const int64_t lo = ranges[r].first;
      const int64_t hi = r + 1 < ranges.size() ? ranges[r + 1].first : (int64_t)h.n_tris;
      if (lo < 0 || hi > (int64_t)h.n_tris) { err = "blob: triangle offsets"; return false; }
      for (int64_t t = lo; t < hi; ++t) tri_voff[(size_t)t] = ranges[r].second;
    }
  }
  out.tri_verts.assign(kTriStride * (size_t)h.n_tris, f4{ 0, 0, 0, 0 });
  out.tri_nrm.assign(3 * (size_t)h.n_tris, f4{ 0, 0, 0, 0 });
  out.tri_uv.assign(6 * (size_t)h.n_tris, 0.0f);
  bool bad_index = false;
#pragma omp parallel for schedule(static) if (h.n_tris >= 100000)
  for (long tt = 0; tt < (long)h.n_tris; ++tt) {
    const size_t t = (size_t)tt;
    const int32_t* tr = v.tris + 4 * t;
    const int32_t vo = tri_voff[t];
    if (vo < 0) continue;   // triangle of an unreferenced range
    for (int k = 0; k < 3; ++k) {
      const int64_t vi = (int64_t)vo + tr[k];
      if (vi < 0 || vi >= (int64_t)h.n_verts) {
#pragma omp atomic write
        bad_index = true;
        break;
      }
      const float* p = v.vert_pos + 3 * (size_t)vi;
      const float* n = v.vert_nrm + 3 * (size_t)vi;
      out.tri_verts[kTriStride * t + k] = f4{ p[0], p[1], p[2], 0.0f };
      out.tri_nrm[3 * t + k] = f4{ n[0], n[1], n[2], 0.0f };
      out.tri_uv[6 * t + 2 * k] = v.vert_uv[2 * (size_t)vi];
      out.tri_uv[6 * t + 2 * k + 1] = v.vert_uv[2 * (size_t)vi + 1];
    }
    out.tri_verts[kTriStride * t].w = Converter::bits(tr[3]);
  }
  if (bad_index) { err = "blob: vertex index"; return false; }
  out.inst.assign(4 * (size_t)h.n_inst, f4{ 0, 0, 0, 0 });
  for (size_t k = 0; k < h.n_inst; ++k) {
    const float* m = v.inst_inv + 16 * k;
    for (int r = 0; r < 3; ++r) out.inst[4 * k + r] = f4{ m[4 * r], m[4 * r + 1], m[4 * r + 2], m[4 * r + 3] };
    out.inst[4 * k + 3] = f4{ Converter::bits(kRefNone), Converter::bits(v.inst_meta[4 * k]), 0.0f, 0.0f };
  }
  out.quad = (h.flags & 2u) != 0;
  Converter c{ v, out, err };
  out.top_root = out.quad ? c.quad_node(0, 0, 0, true, 0) : c.top(0);
  out.mesh_root_ref.assign(c.mesh_root_ref.begin(), c.mesh_root_ref.end());
  return c.ok;
}

// The top-level half of build_device_layout for a blob whose mesh sections did not change since `prev` was built
// (build_blob patched it in place: an object moved, changed its material or -- with the same number of visible
// objects -- its visibility): top-level nodes in breadth-first order and the instance records, nothing else.
bool build_device_layout_top(const BlobView& v, const DeviceLayout& prev, DeviceLayout& out, std::string& err)
{
  out = DeviceLayout{};
  const BlobHeader& h = v.hdr;
  if (prev.quad || (h.flags & 2u) || h.n_nodes == 0 || h.n_inst == 0 || h.n_tris != prev.n_tris) return false;
  out.n_tris = h.n_tris;
  out.n_inst = h.n_inst;
  out.max_depth_bottom = prev.max_depth_bottom;
  out.inst.assign(4 * (size_t)h.n_inst, f4{ 0, 0, 0, 0 });
  for (size_t k = 0; k < h.n_inst; ++k) {
    const float* m = v.inst_inv + 16 * k;
    for (int r = 0; r < 3; ++r) out.inst[4 * k + r] = f4{ m[4 * r], m[4 * r + 1], m[4 * r + 2], m[4 * r + 3] };
    out.inst[4 * k + 3] = f4{ Converter::bits(kRefNone), Converter::bits(v.inst_meta[4 * k]), 0.0f, 0.0f };
  }
  Converter c{ v, out, err };
  c.top_only = true;
  c.mesh_root_ref.insert(prev.mesh_root_ref.begin(), prev.mesh_root_ref.end());
  out.top_root = c.top(0);
  if (!c.ok || out.n_top_inner != prev.n_top_inner) return false;
  out.mesh_root_ref = prev.mesh_root_ref;
  return true;
}

}  // namespace crt
