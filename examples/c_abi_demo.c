/* Minimal C host for librgnn (no Python, no torch): one RGCN layer on a 4-node, 2-type graph through the C ABI of
 * include/rgnn.h.  Build:  gcc -std=c99 -Iinclude examples/c_abi_demo.c -o c_abi_demo \
 *                              -Ltf-gnn-samples_b200/lib -lrgnn -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,tf-gnn-samples_b200/lib
 * (the CUDA runtime is only used here to allocate and copy device buffers; the library takes plain device pointers). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rgnn.h"

/* cudart entry points declared by hand to keep this file plain C99 without the CUDA headers */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(call) do { int rc_ = (call); if (rc_ != RGNN_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rgnn_last_error()); return 1; } } while (0)

int main(void) {
  enum { V = 4, L = 2, D = 8 };
  const int32_t adj0[] = {0, 1, 1, 2, 2, 3};          /* type 0: (src, tgt) rows 0->1, 1->2, 2->3 */
  const int32_t adj1[] = {0, 0, 1, 1, 2, 2, 3, 3};    /* type 1: self loops */
  const int64_t counts[L] = {3, 4};
  float h[V * D], w[L][D * D], cnt[L * V] = {0, 1, 1, 1, 1, 1, 1, 1}, out[V * D];
  for (int i = 0; i < V * D; ++i) h[i] = (float)(i % 7) / 7.0f;
  for (int l = 0; l < L; ++l)
    for (int i = 0; i < D * D; ++i) w[l][i] = (i % (D + 1) == 0) ? 1.0f : 0.0f;   /* identity kernels */

  void *d_adj0, *d_adj1, *d_h, *d_w0, *d_w1, *d_cnt, *d_out, *d_ws;
  cudaMalloc(&d_adj0, sizeof adj0); cudaMalloc(&d_adj1, sizeof adj1); cudaMalloc(&d_h, sizeof h);
  cudaMalloc(&d_w0, sizeof w[0]); cudaMalloc(&d_w1, sizeof w[1]); cudaMalloc(&d_cnt, sizeof cnt); cudaMalloc(&d_out, sizeof out);
  cudaMemcpy(d_adj0, adj0, sizeof adj0, H2D); cudaMemcpy(d_adj1, adj1, sizeof adj1, H2D); cudaMemcpy(d_h, h, sizeof h, H2D);
  cudaMemcpy(d_w0, w[0], sizeof w[0], H2D); cudaMemcpy(d_w1, w[1], sizeof w[1], H2D); cudaMemcpy(d_cnt, cnt, sizeof cnt, H2D);

  const int32_t* adj_table[L] = {(const int32_t*)d_adj0, (const int32_t*)d_adj1};
  const float* w_table[L] = {(const float*)d_w0, (const float*)d_w1};
  rgnn_plan_t* plan = NULL;
  CHECK(rgnn_plan_create(&plan, V, L, adj_table, counts, NULL));
  const size_t ws_bytes = rgnn_workspace_bytes(plan, RGNN_LAYER_RGCN, D, D, 0);
  cudaMalloc(&d_ws, ws_bytes);
  CHECK(rgnn_rgcn_forward(plan, (const float*)d_h, D, D, w_table, (const float*)d_cnt, RGNN_ACT_LINEAR, RGNN_AGG_SUM,
                          /*normalize=*/1, /*use_both_source_and_target=*/0, /*num_timesteps=*/1, (float*)d_out, d_ws, ws_bytes, NULL));
  cudaDeviceSynchronize();
  cudaMemcpy(out, d_out, sizeof out, D2H);
  /* identity kernels + in-degree normalisation: out[v] = h[v] + h[v-1] for v >= 1, out[0] = h[0] */
  int bad = 0;
  for (int v = 0; v < V; ++v)
    for (int j = 0; j < D; ++j) {
      const float want = h[v * D + j] + (v > 0 ? h[(v - 1) * D + j] : 0.0f);
      if (want - out[v * D + j] > 1e-5f || out[v * D + j] - want > 1e-5f) ++bad;
    }
  printf("librgnn ABI %d: %d kernel launches, %d mismatches\n", rgnn_version(), (int)rgnn_launch_count(), bad);
  rgnn_plan_destroy(plan);
  cudaFree(d_ws); cudaFree(d_out); cudaFree(d_cnt); cudaFree(d_w1); cudaFree(d_w0); cudaFree(d_h); cudaFree(d_adj1); cudaFree(d_adj0);
  return bad != 0;
}
