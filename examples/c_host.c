/* A plain-C host of libd3roma_b200.so: what a non-Python caller binds.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -Ld3roma_b200 -ld3roma_b200 -Wl,-rpath,$PWD/d3roma_b200 -o c_host
 *
 * Without a GPU it checks the library version and the error protocol (argument validation happens before any CUDA call);
 * with one (./c_host gpu) it runs the fused DDIM update (core/scheduler_ddim.py:423-489) on a small buffer through the
 * CUDA runtime that is linked into the library - the caller needs no CUDA headers: device memory comes from
 * d3_peer_alloc / d3_peer_free, the library's own allocation entry points. */
#include <stdio.h>
#include <string.h>

#include "d3roma_b200.h"

static int expect_invalid(int rc, const char* what) {
  if (rc != D3_ERR_INVALID) {
    fprintf(stderr, "%s: expected D3_ERR_INVALID, got %d\n", what, rc);
    return 1;
  }
  if (d3_last_error() == NULL || strlen(d3_last_error()) == 0) {
    fprintf(stderr, "%s: no error text\n", what);
    return 1;
  }
  return 0;
}

int main(int argc, char** argv) {
  int bad = 0;
  printf("libd3roma_b200 version %d\n", d3_version());
  if (d3_version() < 100) return 1;
  /* argument validation: null pointers and out-of-range sizes are refused with a message, never dereferenced */
  bad += expect_invalid(d3_peer_close(NULL), "d3_peer_close(NULL)");
  bad += expect_invalid(d3_layernorm_f32(NULL, 4, 8, NULL, NULL, 1e-5f, NULL, NULL), "d3_layernorm_f32(NULL...)");
  bad += expect_invalid(d3_ensemble_objective(NULL, 40, 16, NULL, NULL, 1, NULL, NULL, NULL), "d3_ensemble_objective(n = 40)");
  printf("last error text: %s\n", d3_last_error());
  if (bad) return 1;

  if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
    /* x0-prediction DDIM update on 1024 zeros: prev = sqrt_alpha_prev * clip(model_output) + dir_coeff * eps = 0 */
    enum { N = 1024 };
    void* buf = NULL;
    unsigned char handle[64];
    d3_ddim_coeffs c;
    int rc;
    memset(&c, 0, sizeof c);
    c.prediction_type = 1; /* sample */
    c.sqrt_alpha_t = 0.8f;
    c.sqrt_beta_t = 0.6f;
    c.sqrt_alpha_prev = 0.9f;
    c.dir_coeff = 0.43589f;
    c.clip_mode = 1;
    c.clip_range = 1.0f;
    rc = d3_peer_alloc(4 * N * sizeof(float), &buf, handle); /* zero-filled device memory */
    if (rc != D3_OK) {
      fprintf(stderr, "d3_peer_alloc: %s\n", d3_last_error());
      return 2;
    }
    {
      float* f = (float*)buf;
      rc = d3_sched_ddim_step(f, f + N, NULL, &c, N, f + 2 * N, f + 3 * N, NULL, NULL, NULL);
    }
    if (rc != D3_OK) {
      fprintf(stderr, "d3_sched_ddim_step: %s\n", d3_last_error());
      return 2;
    }
    if (d3_peer_free(buf) != D3_OK) return 2;
    printf("d3_sched_ddim_step ran on the device\n");
  }
  return 0;
}
