// Host scene code under AddressSanitizer + UBSan (no GPU, no CUDA): random scenes through
// build_blob -> parse_blob -> build_device_layout, both tree widths, then the paths that reuse state between
// commits (cached bottom trees, cached world boxes, blob patched in place, hidden instances), and a few
// malformed blobs that must be rejected, not read out of bounds.
#include "../../cadrays_b200/csrc/host_scene.hpp"

#include <cstdio>
#include <cstring>
#include <random>

using namespace crt;

static Mesh random_mesh(std::mt19937& g, int n_tris)
{
  std::uniform_real_distribution<float> u(-1.0f, 1.0f);
  Mesh m;
  const int nv = 3 + (int)(g() % (unsigned)(n_tris + 2));
  for (int i = 0; i < nv; ++i)
    for (int k = 0; k < 3; ++k) { m.pos.push_back(u(g) * 3.0f); m.nrm.push_back(k == 2 ? 1.0f : 0.0f); }
  m.uv.assign((size_t)2 * nv, 0.25f);
  m.has_uv = (g() & 1u) != 0;
  for (int t = 0; t < n_tris; ++t)
    for (int k = 0; k < 3; ++k) m.idx.push_back(g() % (unsigned)nv);
  return m;
}

static void random_xf(std::mt19937& g, float xf[12])
{
  std::uniform_real_distribution<float> u(-1.0f, 1.0f);
  const float s = 0.5f + 0.5f * (u(g) + 1.0f);
  const float m[12] = { s, 0.1f * u(g), 0, 5 * u(g), 0, s, 0.1f * u(g), 5 * u(g), 0.1f * u(g), 0, s, 5 * u(g) };
  std::memcpy(xf, m, sizeof m);
}

static bool convert(const std::vector<uint8_t>& blob, std::string& err)
{
  BlobView view;
  if (!parse_blob(blob.data(), blob.size(), view, err)) return false;
  DeviceLayout layout;
  return build_device_layout(view, layout, err);
}

int main()
{
  std::mt19937 g(12345);
  std::string err;
  for (int round = 0; round < 12; ++round) {
    HostScene scene;
    const int n_mesh = 1 + (int)(g() % 5u);
    for (int i = 0; i < n_mesh; ++i) scene.meshes.push_back(random_mesh(g, 1 + (int)(g() % (round == 0 ? 4u : 400u))));
    const int n_inst = 1 + (int)(g() % 20u);
    for (int i = 0; i < n_inst; ++i) {
      Instance in;
      in.mesh = g() % (unsigned)n_mesh; in.material = g() % 7u;
      random_xf(g, in.xf);
      scene.instances.push_back(in);
    }
    std::vector<uint8_t> blob;
    for (int width : { 2, 4, 2 }) {
      if (!build_blob(scene, blob, err, width)) { std::printf("build_blob: %s\n", err.c_str()); return 1; }
      if (!convert(blob, err)) { std::printf("convert (width %d): %s\n", width, err.c_str()); return 2; }
    }
    // instance-only edits: the blob is patched in place and must equal a from-scratch build
    for (int edit = 0; edit < 6; ++edit) {
      Instance& in = scene.instances[g() % (unsigned)n_inst];
      if (edit % 3 == 0) random_xf(g, in.xf);
      else if (edit % 3 == 1) in.material = g() % 9u;
      else in.visible = !in.visible;
      if (!build_blob(scene, blob, err, 2)) { std::printf("rebuild: %s\n", err.c_str()); return 3; }
      HostScene fresh;
      fresh.meshes = scene.meshes;
      fresh.instances = scene.instances;
      for (Instance& f : fresh.instances) f.box_valid = false;
      std::vector<uint8_t> expect;
      if (!build_blob(fresh, expect, err, 2)) { std::printf("fresh: %s\n", err.c_str()); return 4; }
      if (blob != expect) { std::printf("round %d edit %d: patched blob differs from a fresh build\n", round, edit); return 5; }
      if (!convert(blob, err)) { std::printf("convert after edit: %s\n", err.c_str()); return 6; }
    }
    // malformed input: truncated and corrupted blobs are rejected
    if (blob.size() > 64) {
      std::vector<uint8_t> bad(blob.begin(), blob.begin() + (long)(blob.size() / 2));
      BlobView view;
      if (parse_blob(bad.data(), bad.size(), view, err)) { std::printf("truncated blob accepted\n"); return 7; }
      bad = blob;
      for (int k = 0; k < 64; ++k) {
        std::vector<uint8_t> c = blob;
        const size_t at = 64 + g() % (c.size() - 64);
        c[at] ^= (uint8_t)(1u << (g() % 8u));
        std::string e2;
        (void)convert(c, e2);           // may succeed or fail; must not touch memory outside the blob
      }
    }
  }
  // singular transform, unknown mesh: reported, not crashed on
  {
    HostScene scene;
    scene.meshes.push_back(random_mesh(g, 5));
    Instance in; in.mesh = 0; in.material = 0;
    std::memset(in.xf, 0, sizeof in.xf);
    scene.instances.push_back(in);
    std::vector<uint8_t> blob;
    if (build_blob(scene, blob, err, 2)) { std::printf("singular transform accepted\n"); return 8; }
    scene.instances[0].mesh = 9;
    if (build_blob(scene, blob, err, 2)) { std::printf("unknown mesh accepted\n"); return 9; }
  }
  std::printf("host scene sanitize ok\n");
  return 0;
}
