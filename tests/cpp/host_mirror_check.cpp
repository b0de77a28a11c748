// Compiles the C++ host mirror against the C-ABI and exercises everything that needs no GPU:
// material records, host-only scene assembly, BVH build/export, and the error path of a device call.
#include "cadrays_b200.hpp"
#include <cstdio>

int main()
{
  crt::BSDF paint;
  paint.Kc[0] = paint.Kc[1] = paint.Kc[2] = 1.f;
  paint.Kd[0] = 1.0f; paint.Kd[1] = 0.8f; paint.Kd[2] = 0.2f; paint.Ks[0] = paint.Ks[1] = paint.Ks[2] = 0.3f;
  paint.FresnelCoat = crt::Fresnel::CreateDielectric(1.5f);
  paint.Normalize();
  crt_bsdf rec = paint.Record();
  if (rec.FresnelCoat[0] != -3.f || rec.FresnelCoat[1] != 1.5f) return 2;
  float mx = 0.f;
  for (int k = 0; k < 3; ++k) mx = std::max(mx, paint.Kd[k] + paint.Ks[k]);
  if (mx > 1.0001f || mx < 0.9999f) return 3;

  crt::View view{ crt::View::HostOnly{} };
  const float pos[9] = { 0, 0, 0, 1, 0, 0, 0, 1, 0 };
  const uint32_t idx[3] = { 0, 1, 2 };
  uint32_t mesh = view.AddTriangulation(pos, nullptr, nullptr, 3, idx, 1);
  const float xf[12] = { 1, 0, 0, 5, 0, 1, 0, 0, 0, 0, 1, 0 };
  view.Display(mesh, xf, 0);
  view.Display(mesh, nullptr, 0);
  view.SetMaterials({ paint, crt::BSDF::CreateDiffuse(0.8f, 0.8f, 0.8f) });
  view.Update();
  std::vector<uint8_t> blob = view.ExportBVH();
  std::printf("blob %zu bytes\n", blob.size());
  try { view.Redraw(); return 4; }                    // no device bound: must fail loudly
  catch (const crt::Failure& f) { if (f.code != CRT_ERR_STATE && f.code != CRT_ERR_NO_DEVICE) return 5; }
  {
    std::vector<uint8_t> px;
    crt_camera cam = {};
    cam.dir[1] = 1.f; cam.up[2] = 1.f; cam.fovy_deg = 45.f;
    try { if (view.ToPixMap(px, 32, 16, cam, 1)) return 9; }      // host-only: resize refuses, nothing is rendered
    catch (const crt::Failure& f) { if (f.code != CRT_ERR_NO_DEVICE && f.code != CRT_ERR_STATE) return 10; }
  }
  try { view.Display(99, nullptr, 0); return 6; }
  catch (const crt::Failure& f) { if (f.code != CRT_ERR_INVALID_ARG) return 7; }
  return blob.size() > 64 ? 0 : 8;
}
