/* The C-ABI header must be plain C (the adapter inside TKOpenGl is C++, but cgo / ctypes / JNI hosts bind C):
 * compiled with gcc -std=c99 -pedantic -Wall -Werror and run without a GPU. */
#include "cadrays_b200.h"
#include <stdio.h>
#include <string.h>

int main(void)
{
  crt_params p;
  crt_context* c = NULL;
  uint32_t mesh = 0, inst = 0, tx = 0, ty = 0;
  uint64_t total = 0;
  size_t blob = 0;
  const float pos[9] = { 0, 0, 0, 1, 0, 0, 0, 1, 0 };
  const uint32_t idx[3] = { 0, 1, 2 };
  crt_bsdf white;
  if (crt_abi_version() != CRT_ABI_VERSION) return 1;
  if (crt_params_default(&p) != CRT_OK || p.max_depth != 8 || p.adaptive_sampling != 0) return 2;
  if (sizeof(crt_bsdf) != 128 || sizeof(crt_light) != 32 || sizeof(crt_stats) != 112) return 3;
  if (crt_create_host_only(&c) != CRT_OK) return 4;
  if (crt_mesh_create(c, pos, NULL, NULL, 3, idx, 1, &mesh) != CRT_OK) return 5;
  if (crt_instance_add(c, mesh, NULL, 0, &inst) != CRT_OK) return 6;
  memset(&white, 0, sizeof white);
  white.Kd[0] = white.Kd[1] = white.Kd[2] = 0.8f;
  white.FresnelCoat[0] = -1.0f; white.FresnelBase[0] = -1.0f; white.FresnelBase[2] = 1.0f;
  if (crt_materials_set(c, &white, 1) != CRT_OK) return 7;
  p.adaptive_sampling = 1; p.adaptive_tiles = 128;
  if (crt_params_set(c, &p) != CRT_OK) return 8;
  if (crt_commit(c) != CRT_OK) return 9;
  if (crt_bvh_export(c, NULL, 0, &blob) != CRT_OK || blob <= 64) return 10;
  /* every device entry point refuses loudly on a host-only context: there is no CPU fallback */
  if (crt_render(c, 1, &total) == CRT_OK) return 11;
  if (crt_adaptive_tiles_get(c, NULL, NULL, 0, &tx, &ty) == CRT_OK) return 12;
  if (strlen(crt_last_error()) == 0) return 13;
  if (crt_mesh_create(c, pos, NULL, NULL, 3, idx, 0, &mesh) != CRT_ERR_INVALID_ARG) return 14;
  crt_destroy(c);
  printf("c abi ok, blob %lu bytes\n", (unsigned long)blob);
  return 0;
}
