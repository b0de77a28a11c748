// The CPU oracle (oracle/cadrays_oracle.c) under AddressSanitizer + UBSan: a random two-level scene built by the
// host scene code, every material class, lights, environment, textures, both tree widths; closest / any-hit /
// brute-force traces, plain and adaptive renders, display.  Test infrastructure checking test infrastructure:
// lives under tests/ and is the only C++ file that links both the oracle and host_scene.cpp.
#include "../../cadrays_b200/csrc/host_scene.hpp"
#include "../../oracle/cadrays_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using namespace crt;

int main()
{
  std::mt19937 g(777);
  std::uniform_real_distribution<float> u(-1.0f, 1.0f);
  for (int width : { 2, 4 }) {
    HostScene scene;
    for (int m = 0; m < 3; ++m) {
      Mesh mesh;
      const int nu = 6 + 2 * m, nv = 5 + m;            // a bumpy sheet
      for (int j = 0; j <= nv; ++j)
        for (int i = 0; i <= nu; ++i) {
          mesh.pos.push_back((float)i / nu - 0.5f); mesh.pos.push_back((float)j / nv - 0.5f); mesh.pos.push_back(0.05f * u(g));
          mesh.nrm.push_back(0); mesh.nrm.push_back(0); mesh.nrm.push_back(1);
          mesh.uv.push_back((float)i / nu); mesh.uv.push_back((float)j / nv);
        }
      mesh.has_uv = true;
      for (int j = 0; j < nv; ++j)
        for (int i = 0; i < nu; ++i) {
          const uint32_t a = (uint32_t)(j * (nu + 1) + i), b = a + 1, c = a + (uint32_t)nu + 1, d = c + 1;
          const uint32_t t[6] = { a, b, d, a, d, c };
          mesh.idx.insert(mesh.idx.end(), t, t + 6);
        }
      scene.meshes.push_back(mesh);
    }
    for (int k = 0; k < 9; ++k) {
      Instance in;
      in.mesh = (uint32_t)(k % 3); in.material = (uint32_t)k;           // material 8 is out of range: default record
      const float s = 0.6f + 0.1f * (float)(k % 4), c = std::cos(0.3f * k), sn = std::sin(0.3f * k);
      const float m[12] = { s * c, -s * sn, 0, 0.7f * u(g), s * sn, s * c, 0, 0.7f * u(g), 0, 0, s, 0.25f * (float)k - 1.0f };
      std::memcpy(in.xf, m, sizeof m);
      scene.instances.push_back(in);
    }
    std::vector<uint8_t> blob;
    std::string err;
    if (!build_blob(scene, blob, err, width)) { std::printf("build_blob: %s\n", err.c_str()); return 1; }
    orc_scene* s = orc_scene_from_blob(blob.data(), blob.size());
    if (!s) { std::printf("orc_scene_from_blob failed\n"); return 2; }

    crt_bsdf mats[8];
    std::memset(mats, 0, sizeof mats);
    for (int k = 0; k < 8; ++k) {
      crt_bsdf& b = mats[k];
      b.FresnelCoat[0] = -1.0f; b.FresnelBase[0] = -1.0f; b.FresnelBase[2] = 1.0f;
      b.Kd[0] = 0.7f; b.Kd[1] = 0.5f; b.Kd[2] = 0.3f;
    }
    mats[1].Ks[0] = mats[1].Ks[1] = mats[1].Ks[2] = 0.3f; mats[1].Ks[3] = 0.2f; mats[1].FresnelBase[0] = 0.9f; mats[1].FresnelBase[1] = 0.6f;
    mats[2].Kc[0] = mats[2].Kc[1] = mats[2].Kc[2] = 1.0f; mats[2].Kc[3] = 0.1f; mats[2].FresnelCoat[0] = -3.0f; mats[2].FresnelCoat[1] = 1.5f;
    mats[3].Kc[0] = mats[3].Kc[1] = mats[3].Kc[2] = 1.0f; mats[3].Kt[0] = mats[3].Kt[1] = mats[3].Kt[2] = 1.0f;
    mats[3].Kd[0] = mats[3].Kd[1] = mats[3].Kd[2] = 0.0f; mats[3].FresnelCoat[0] = -3.0f; mats[3].FresnelCoat[1] = 1.62f;
    mats[3].Absorption[0] = 0.8f; mats[3].Absorption[1] = 0.9f; mats[3].Absorption[2] = 1.0f; mats[3].Absorption[3] = 2.0f;
    mats[4].Ks[0] = mats[4].Ks[1] = mats[4].Ks[2] = 0.9f; mats[4].Ks[3] = 0.0f; mats[4].FresnelBase[0] = -2.0f; mats[4].FresnelBase[1] = 0.8f; mats[4].FresnelBase[2] = 5.8f;
    mats[5].Le[0] = 2.0f; mats[5].Le[1] = 1.0f; mats[5].Le[2] = 0.5f;
    mats[6].Kd[3] = 1.0f; mats[6].Kt[3] = 2.0f; mats[6].Le[3] = 3.0f;     // texture 0, scaled
    orc_set_materials(s, mats, 8);
    crt_light lights[2];
    std::memset(lights, 0, sizeof lights);
    lights[0].emission[0] = lights[0].emission[1] = lights[0].emission[2] = 8.0f; lights[0].smoothness = 0.2f;
    lights[0].posdir[0] = -0.3f; lights[0].posdir[1] = 0.2f; lights[0].posdir[2] = -1.0f; lights[0].is_point = 0;
    lights[1].emission[0] = 30.0f; lights[1].emission[1] = 20.0f; lights[1].emission[2] = 10.0f; lights[1].smoothness = 0.05f;
    lights[1].posdir[0] = 0.5f; lights[1].posdir[1] = 0.5f; lights[1].posdir[2] = 2.0f; lights[1].is_point = 1;
    orc_set_lights(s, lights, 2);
    std::vector<float> env(3 * 16 * 8);
    for (float& e : env) e = 0.5f + 0.5f * u(g);
    orc_set_envmap_rgb32f(s, env.data(), 16, 8);
    std::vector<uint8_t> tex(4 * 5 * 3);
    for (uint8_t& t : tex) t = (uint8_t)(g() & 255u);
    const uint32_t tex_size[2] = { 5, 3 };
    orc_set_textures(s, tex.data(), tex_size, 1);
    crt_params p;
    std::memset(&p, 0, sizeof p);
    p.max_depth = 6; p.max_radiance = 50.0f; p.white_point = 1.0f; p.env_as_background = 1; p.frame_seed0 = 3; p.russian_roulette = 1;
    p.aperture_radius = 0.02f; p.focal_dist = 2.5f; p.tone_map = 1; p.two_sided = width == 4;
    orc_set_params(s, &p);
    crt_camera cam;
    std::memset(&cam, 0, sizeof cam);
    cam.eye[2] = 3.0f; cam.dir[2] = -1.0f; cam.up[1] = 1.0f; cam.fovy_deg = 50.0f; cam.aspect = 40.0f / 24.0f;
    orc_set_camera(s, &cam);

    const uint32_t n = 4000;
    std::vector<float> org(3 * n), dir(3 * n), tmax(n), t(n), uu(n), vv(n);
    std::vector<int32_t> prim(n), inst(n), prim2(n), inst2(n);
    for (uint32_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) { org[3 * i + k] = 2.0f * u(g); dir[3 * i + k] = u(g); }
      if (i % 97 == 0) dir[3 * i] = dir[3 * i + 1] = dir[3 * i + 2] = 0.0f;          // degenerate rays
      tmax[i] = 0.5f + 2.0f * std::fabs(u(g));
    }
    crt_stats st;
    std::memset(&st, 0, sizeof st);
    orc_trace(s, org.data(), dir.data(), nullptr, n, 0, prim.data(), inst.data(), t.data(), uu.data(), vv.data(), &st);
    orc_trace_brute(s, org.data(), dir.data(), nullptr, n, 0, prim2.data(), inst2.data(), t.data(), uu.data(), vv.data());
    uint32_t differ = 0;
    for (uint32_t i = 0; i < n; ++i) differ += (prim[i] >= 0) != (prim2[i] >= 0);
    if (differ > n / 200) { std::printf("width %d: traversal and brute force disagree on %u rays\n", width, differ); return 3; }
    orc_trace(s, org.data(), dir.data(), tmax.data(), n, 1, prim.data(), inst.data(), t.data(), uu.data(), vv.data(), nullptr);

    const uint32_t w = 40, h = 24;
    std::vector<float> accum(4 * w * h, 0.0f);
    orc_render(s, w, h, 0, 3, accum.data(), 2, &st);
    std::vector<uint8_t> rgb(3 * w * h);
    orc_display(s, accum.data(), w, h, rgb.data());
    for (float a : accum) if (!(a == a)) { std::printf("NaN in the accumulation buffer\n"); return 4; }
    const uint32_t nt = ((w + 31) / 32) * ((h + 31) / 32);
    std::vector<uint32_t> count(nt, 0), terr(nt, 0);
    std::vector<float> even(w * h, 0.0f), accum2(4 * w * h, 0.0f);
    uint32_t wave = 0;
    orc_render_adaptive(s, w, h, 0, 7 * nt, 2 * nt, accum2.data(), count.data(), terr.data(), even.data(), &wave, 2);
    uint64_t total = 0;
    for (uint32_t c : count) total += c;
    if (total != 7ull * nt) { std::printf("adaptive budget not spent exactly\n"); return 5; }
    orc_scene_free(s);
  }
  // a truncated blob is refused
  {
    const uint8_t junk[32] = { 0 };
    if (orc_scene_from_blob(junk, sizeof junk) != nullptr) { std::printf("junk blob accepted\n"); return 6; }
  }
  std::printf("oracle sanitize ok\n");
  return 0;
}
