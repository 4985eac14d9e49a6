/* Plain-C consumer of include/jaxib_b200.h: compiled with gcc (no CUDA, no C++), it dlopen()s the library,
 * resolves every entry point THROUGH THE HEADER'S OWN PROTOTYPES (a drifted signature fails to compile, a missing
 * symbol fails at run time) and exercises the argument-error paths, which need no GPU.
 *
 *   gcc -std=c11 -Wall -Werror -Iinclude tests/c_abi_smoke.c -ldl -o /tmp/c_abi_smoke && /tmp/c_abi_smoke <lib.so>
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "jaxib_b200.h"

#define LOAD(name)                                                      \
  __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name);    \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define EXPECT(cond)                                                            \
  do {                                                                          \
    if (!(cond)) { fprintf(stderr, "%s:%d: %s failed\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s libjaxib_b200.so\n", argv[0]); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(jaxib_version) LOAD(jaxib_last_error) LOAD(jaxib_compiled_arch)
  LOAD(jaxib_plan_create) LOAD(jaxib_plan_destroy) LOAD(jaxib_scratch_bytes)
  LOAD(jaxib_explicit_terms) LOAD(jaxib_explicit_predict)
  LOAD(jaxib_ib_interp) LOAD(jaxib_ib_spread_scratch_bytes) LOAD(jaxib_ib_spread) LOAD(jaxib_ib_force) LOAD(jaxib_ib_marker_force) LOAD(jaxib_ib_spread_bodies)
  LOAD(jaxib_pressure_solve) LOAD(jaxib_project)
  LOAD(jaxib_poisson_rows_fwd) LOAD(jaxib_poisson_cols) LOAD(jaxib_poisson_rows_inv)
  LOAD(jaxib_slab_step) LOAD(jaxib_step) LOAD(jaxib_step_profile)
  LOAD(jaxib_tracer_step) LOAD(jaxib_bilinear_sample) LOAD(jaxib_pair_force) LOAD(jaxib_penalty_perm)

  EXPECT(strstr(p_jaxib_version(), "sm_100a") != NULL);
  EXPECT(p_jaxib_compiled_arch() == 100);

  /* argument errors: status codes and messages, no device work */
  jaxib_plan* plan = NULL;
  EXPECT(p_jaxib_plan_create(NULL, &plan) == JAXIB_ERR_INVALID);
  jaxib_grid g;
  memset(&g, 0, sizeof(g));
  g.nx = 0; g.ny = 8; g.batch = 1; g.dx = g.dy = 1.0;
  EXPECT(p_jaxib_plan_create(&g, &plan) == JAXIB_ERR_INVALID && plan == NULL);
  EXPECT(strstr(p_jaxib_last_error(), "nx") != NULL);
  g.nx = 8; g.ny = 7;  /* odd ny: the packed real FFT needs an even row length */
  EXPECT(p_jaxib_plan_create(&g, &plan) == JAXIB_ERR_UNSUPPORTED && plan == NULL);
  g.ny = 8; g.bc = 7;
  EXPECT(p_jaxib_plan_create(&g, &plan) == JAXIB_ERR_INVALID);
  g.bc = JAXIB_BC_PERIODIC; g.convect = 9;
  EXPECT(p_jaxib_plan_create(&g, &plan) == JAXIB_ERR_INVALID);
  g.convect = JAXIB_CONVECT_UPWIND;
  EXPECT(p_jaxib_scratch_bytes(NULL, 0) == 0);
  EXPECT(p_jaxib_plan_destroy(NULL) == JAXIB_OK);

  float dummy[4] = {0};
  EXPECT(p_jaxib_step(NULL, NULL, NULL, NULL, NULL, NULL, NULL, 1, 1, dummy, dummy, dummy, NULL, 0) == JAXIB_ERR_INVALID);
  EXPECT(strstr(p_jaxib_last_error(), "plan") != NULL);
  EXPECT(p_jaxib_project(NULL, NULL, dummy, dummy, dummy, dummy, dummy, dummy, NULL, 0) == JAXIB_ERR_INVALID);
  EXPECT(p_jaxib_explicit_predict(NULL, &g, dummy, dummy, NULL, NULL, dummy, dummy) == JAXIB_ERR_INVALID); /* aliasing */
  EXPECT(strstr(p_jaxib_last_error(), "alias") != NULL);
  EXPECT(p_jaxib_explicit_terms(NULL, &g, NULL, dummy, NULL, dummy, dummy) == JAXIB_ERR_INVALID);
  EXPECT(p_jaxib_ib_interp(NULL, &g, 1.0, 0.5, dummy, dummy, dummy, -1, dummy, NULL, NULL) == JAXIB_ERR_INVALID);
  EXPECT(p_jaxib_ib_interp(NULL, &g, 1.0, 0.5, dummy, dummy, dummy, 0, dummy, NULL, NULL) == JAXIB_OK); /* n = 0: nothing to do */
  EXPECT(p_jaxib_ib_spread(NULL, &g, 1.0, 0.5, dummy, dummy, dummy, dummy, 3, 1.0, dummy, NULL, 0) == JAXIB_ERR_SCRATCH);
  jaxib_bodies b;
  memset(&b, 0, sizeof(b));
  b.n_bodies = 1; b.n_markers = 4;  /* marker arrays missing */
  EXPECT(p_jaxib_ib_force(NULL, &g, &b, dummy, dummy, dummy, NULL, 0) == JAXIB_ERR_INVALID);
  EXPECT(p_jaxib_poisson_cols(NULL, NULL, dummy, 0, 0, 0) == JAXIB_ERR_INVALID);
  jaxib_slab sl;
  memset(&sl, 0, sizeof(sl));
  unsigned epoch = 0;
  EXPECT(p_jaxib_slab_step(NULL, NULL, NULL, &g, &sl, NULL, NULL, 1, 0, 3, 1, &epoch, dummy, dummy) == JAXIB_ERR_INVALID);
  printf("c_abi_smoke: %s, 26 entry points resolved through the header's prototypes, error paths ok\n", p_jaxib_version());
  dlclose(lib);
  return 0;
}
