/*
 * examples/c_abi_sample.c -- libnova_b200.so driven from plain C: no Python, no torch.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_sample.c -o c_abi_sample \
 *       -L nova_pointcloud_b200/lib -lnova_b200 -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/nova_pointcloud_b200/lib
 *
 * Builds a depth-2, width-256 head with deterministic pseudo-random weights in the reference's state_dict
 * naming (diffnext/models/diffusion_mlp.py:81-87), samples 4 clouds of 200 xyz tokens with the 25-step
 * flow-match Euler schedule (diffnext/schedulers/scheduling_cfm.py:92-104, shift 1), three times (eager,
 * graph capture, graph replay), checks that the three results are identical and finite, and scores the
 * clouds against each other with the Chamfer primitive.  Exit code 0 = ok.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nova_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != 0) {                                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, nova_last_error());          \
      return 1;                                                                  \
    }                                                                            \
  } while (0)
#define CUDA(call)                                                               \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) { /* uniform in [-1, 1) */
  rng_state = rng_state * 1664525u + 1013904223u;
  return (float)(rng_state >> 8) * (2.0f / 16777216.0f) - 1.0f;
}

enum { DEPTH = 2, D = 256, T = 3, B = 4, N = 200, S = 25, MAXK = 14 + 8 * DEPTH };

static char names_buf[MAXK][64];
static const char* names[MAXK];
static const void* ptrs[MAXK];
static int64_t numels[MAXK];
static int nkeys = 0;

static int add_key(const char* name, int64_t numel, float scale) {
  float* host = (float*)malloc(sizeof(float) * numel);
  for (int64_t i = 0; i < numel; ++i) host[i] = frand() * scale;
  void* dev = NULL;
  if (cudaMalloc(&dev, sizeof(float) * numel) != cudaSuccess) return 1;
  if (cudaMemcpy(dev, host, sizeof(float) * numel, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
  free(host);
  snprintf(names_buf[nkeys], sizeof(names_buf[nkeys]), "%s", name);
  names[nkeys] = names_buf[nkeys];
  ptrs[nkeys] = dev;
  numels[nkeys] = numel;
  ++nkeys;
  return 0;
}

int main(void) {
  if (nova_device_check() != 0) {
    fprintf(stderr, "no sm_100 device: %s\n", nova_last_error());
    return 2;
  }
  const float ws = 1.0f / sqrtf((float)D);
  char key[64];
  if (add_key("patch_embed.proj.weight", (int64_t)D * T, 0.5f) || add_key("patch_embed.proj.bias", D, 0.1f) ||
      add_key("time_cond_embed.timestep_proj.fc1.weight", (int64_t)D * 256, 1.0f / 16) ||
      add_key("time_cond_embed.timestep_proj.fc1.bias", D, 0.1f) ||
      add_key("time_cond_embed.timestep_proj.fc2.weight", (int64_t)D * D, ws) ||
      add_key("time_cond_embed.timestep_proj.fc2.bias", D, 0.1f) ||
      add_key("time_cond_embed.condition_proj.fc1.weight", (int64_t)D * D, ws) ||
      add_key("time_cond_embed.condition_proj.fc1.bias", D, 0.1f) ||
      add_key("time_cond_embed.condition_proj.fc2.weight", (int64_t)D * D, ws) ||
      add_key("time_cond_embed.condition_proj.fc2.bias", D, 0.1f))
    return 1;
  for (int i = 0; i < DEPTH; ++i) {
    const char* leaf[] = {"norm1.proj.weight", "norm1.proj.bias", "proj.fc1.weight", "proj.fc1.bias",
                          "proj.fc2.weight",   "proj.fc2.bias",   "norm2.weight",    "norm2.bias"};
    const int64_t size[] = {3ll * D * D, 3 * D, (int64_t)D * D, D, (int64_t)D * D, D, D, D};
    const float scale[] = {ws, 0.1f, ws, 0.1f, ws, 0.1f, 1.0f, 0.1f};
    for (int k = 0; k < 8; ++k) {
      snprintf(key, sizeof(key), "blocks.%d.%s", i, leaf[k]);
      if (add_key(key, size[k], scale[k])) return 1;
    }
  }
  if (add_key("norm.proj.weight", 2ll * D * D, ws) || add_key("norm.proj.bias", 2 * D, 0.1f) ||
      add_key("head.weight", (int64_t)T * D, ws) || add_key("head.bias", T, 0.1f))
    return 1;

  nova_head_config cfg = {DEPTH, D, D, T, NOVA_F32};
  nova_head_t* head = NULL;
  CHECK(nova_head_create(&cfg, &head));
  CHECK(nova_head_load(head, nkeys, names, ptrs, numels, NOVA_F32, /*channels=*/3, NULL));
  CUDA(cudaDeviceSynchronize());

  /* inputs: noise tokens (B, N, 3) fp32, condition (B, N, D) fp32 */
  const int64_t M = (int64_t)B * N;
  float* h_noise = (float*)malloc(sizeof(float) * M * T);
  float* h_z = (float*)malloc(sizeof(float) * M * D);
  for (int64_t i = 0; i < M * T; ++i) h_noise[i] = frand() * 1.7f;
  for (int64_t i = 0; i < M * D; ++i) h_z[i] = frand() * 1.7f;
  float *d_noise, *d_z, *d_out;
  CUDA(cudaMalloc((void**)&d_noise, sizeof(float) * M * T));
  CUDA(cudaMalloc((void**)&d_z, sizeof(float) * M * D));
  CUDA(cudaMalloc((void**)&d_out, sizeof(float) * M * T));
  CUDA(cudaMemcpy(d_noise, h_noise, sizeof(float) * M * T, cudaMemcpyHostToDevice));
  CUDA(cudaMemcpy(d_z, h_z, sizeof(float) * M * D, cudaMemcpyHostToDevice));

  /* FlowMatchEulerDiscreteScheduler.set_timesteps(25), shift 1 (scheduling_cfm.py:92-104), float32 values */
  float timesteps[S];
  double sigmas[S + 1];
  for (int i = 0; i < S; ++i) {
    const float t = (float)(1000.0 + (1.0 - 1000.0) * (double)i / (S - 1)); /* np.linspace(1000, 1, S) as float32 */
    const float sigma = t / 1000.0f;
    timesteps[i] = sigma * 1000.0f;
    sigmas[i] = (double)sigma;
  }
  sigmas[S] = 0.0;

  const size_t ws_bytes = nova_head_workspace_bytes(head, M, S);
  void* d_ws = NULL;
  CUDA(cudaMalloc(&d_ws, ws_bytes));
  cudaStream_t stream;
  CUDA(cudaStreamCreate(&stream));
  nova_guidance g = {1.0f, 0.0f, 1.0f};
  float* h_out[3];
  for (int rep = 0; rep < 3; ++rep) { /* eager, graph capture + launch, graph replay */
    CHECK(nova_head_sample(head, d_noise, d_z, NULL, B, B, N, N, timesteps, sigmas, S, &g, d_out, d_ws, ws_bytes, stream));
    CUDA(cudaStreamSynchronize(stream));
    h_out[rep] = (float*)malloc(sizeof(float) * M * T);
    CUDA(cudaMemcpy(h_out[rep], d_out, sizeof(float) * M * T, cudaMemcpyDeviceToHost));
  }
  double sum = 0.0;
  for (int64_t i = 0; i < M * T; ++i) {
    if (!isfinite(h_out[0][i])) { fprintf(stderr, "non-finite output at %lld\n", (long long)i); return 1; }
    if (h_out[0][i] != h_out[1][i] || h_out[0][i] != h_out[2][i]) { fprintf(stderr, "graph replay differs at %lld\n", (long long)i); return 1; }
    sum += fabs(h_out[0][i]);
  }

  /* Chamfer primitive: clouds 0,1 against clouds 2,3; a cloud against itself must give zeros */
  float *d_d1, *d_d2;
  CUDA(cudaMalloc((void**)&d_d1, sizeof(float) * 2 * N));
  CUDA(cudaMalloc((void**)&d_d2, sizeof(float) * 2 * N));
  CHECK(nova_chamfer_nn(d_out, d_out + 2 * N * T, 2, N, N, d_d1, d_d2, NULL, NULL, stream));
  float h_d1[2 * N];
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_d1, d_d1, sizeof(h_d1), cudaMemcpyDeviceToHost));
  double cd = 0.0;
  for (int i = 0; i < 2 * N; ++i) cd += h_d1[i];
  /* the two means of chamfer_distance (demo.py:50-53) on the device, in float64: per pair mean(d1) + mean(d2) */
  double* d_cd;
  double h_cd[2];
  static float h_d2[2 * N];
  CUDA(cudaMalloc((void**)&d_cd, sizeof(double) * 2));
  CHECK(nova_chamfer_pair_mean(d_d1, d_d2, 2, N, N, d_cd, stream));
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_cd, d_cd, sizeof(h_cd), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_d2, d_d2, sizeof(h_d2), cudaMemcpyDeviceToHost));
  for (int c = 0; c < 2; ++c) {
    double m1 = 0.0, m2 = 0.0;
    for (int i = 0; i < N; ++i) { m1 += h_d1[c * N + i]; m2 += h_d2[c * N + i]; }
    const double want = m1 / N + m2 / N;
    if (fabs(h_cd[c] - want) > 1e-12 * (1.0 + fabs(want))) { fprintf(stderr, "pair mean %d: %.17g vs %.17g\n", c, h_cd[c], want); return 1; }
  }
  CHECK(nova_chamfer_nn(d_out, d_out, 2, N, N, d_d1, d_d2, NULL, NULL, stream));
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_d1, d_d1, sizeof(h_d1), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 2 * N; ++i)
    if (h_d1[i] != 0.0f) { fprintf(stderr, "self-distance %g at %d\n", h_d1[i], i); return 1; }

  /* neighbourhood ops on the generated clouds: the nearest neighbour of a point in its own cloud is the point
     itself (distance 0, index = position), and the local density is the mean of the next 8 distances */
  float* d_knn;
  int* d_idx;
  float* d_dens;
  CUDA(cudaMalloc((void**)&d_knn, sizeof(float) * 2 * N * 9));
  CUDA(cudaMalloc((void**)&d_idx, sizeof(int) * 2 * N * 9));
  CUDA(cudaMalloc((void**)&d_dens, sizeof(float) * 2 * N));
  CHECK(nova_knn(d_out, d_out, 2, N, N, 9, d_knn, d_idx, stream));
  CHECK(nova_local_density(d_out, 2, N, 8, d_dens, stream));
  static float h_knn[2 * N * 9], h_dens[2 * N];
  static int h_idx[2 * N * 9];
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_knn, d_knn, sizeof(h_knn), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_idx, d_idx, sizeof(h_idx), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_dens, d_dens, sizeof(h_dens), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 2 * N; ++i) {
    float acc = 0.f;
    for (int r = 1; r < 9; ++r) acc += h_knn[i * 9 + r];
    if (h_knn[i * 9] != 0.0f || h_idx[i * 9] != i % N || h_dens[i] != acc / 8.0f) {
      fprintf(stderr, "kNN/density mismatch at %d: d0 %g idx0 %d density %g vs %g\n", i, h_knn[i * 9], h_idx[i * 9],
              h_dens[i], acc / 8.0f);
      return 1;
    }
  }
  if (nova_knn(d_out, d_out, 2, N, N, 33, d_knn, d_idx, stream) == NOVA_OK) {
    fprintf(stderr, "k = 33 was not rejected\n");
    return 1;
  }

  /* earth mover's distance (auction algorithm): clouds 0,1 against clouds 2,3.  The matching is a permutation, its mean
     distance is at least the mean nearest-neighbour distance (Chamfer lower bound), and a cloud matches itself at 0 */
  float* d_emd;
  int* d_assign;
  int* d_status;
  CUDA(cudaMalloc((void**)&d_emd, sizeof(float) * 2));
  CUDA(cudaMalloc((void**)&d_assign, sizeof(int) * 2 * N));
  CUDA(cudaMalloc((void**)&d_status, sizeof(int) * 2));
  CHECK(nova_emd(d_out, d_out + 2 * N * T, 2, N, 1e-5f, 400000, d_emd, d_assign, d_status, stream));
  float h_emd[2];
  static int h_assign[2 * N], h_status[2], seen[N];
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_emd, d_emd, sizeof(h_emd), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_assign, d_assign, sizeof(h_assign), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_status, d_status, sizeof(h_status), cudaMemcpyDeviceToHost));
  for (int c = 0; c < 2; ++c) {
    memset(seen, 0, sizeof(seen));
    for (int i = 0; i < N; ++i) {
      const int j = h_assign[c * N + i];
      if (j < 0 || j >= N || seen[j]++) { fprintf(stderr, "EMD assignment of pair %d is not a permutation at %d\n", c, i); return 1; }
    }
    if (h_status[c] <= 0) { fprintf(stderr, "EMD pair %d did not converge (status %d)\n", c, h_status[c]); return 1; }
  }
  if ((h_emd[0] + h_emd[1]) * N < cd - 1e-3) { fprintf(stderr, "EMD %g below the Chamfer bound %g\n", h_emd[0] + h_emd[1], cd / N); return 1; }
  CHECK(nova_emd(d_out, d_out, 2, N, 1e-5f, 400000, d_emd, NULL, NULL, stream));
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_emd, d_emd, sizeof(h_emd), cudaMemcpyDeviceToHost));
  if (h_emd[0] != 0.0f || h_emd[1] != 0.0f) { fprintf(stderr, "EMD of a cloud with itself: %g %g\n", h_emd[0], h_emd[1]); return 1; }

  /* one training step of the head (what autograd does for the reference between get_losses and loss.backward()):
     forward with per-row timesteps, then the gradients of two parameters and of z.  d loss / d head.bias is the column
     sum of dv, which the host can check exactly enough */
  const size_t tr_bytes = nova_head_train_bytes(head, M);
  void* d_tr = NULL;
  CUDA(cudaMalloc(&d_tr, tr_bytes));
  float *d_t, *d_v, *d_dv, *d_gbias, *d_gfc1, *d_dz;
  float* h_t = (float*)malloc(sizeof(float) * M);
  for (int64_t i = 0; i < M; ++i) h_t[i] = 500.0f * (frand() + 1.0f);
  CUDA(cudaMalloc((void**)&d_t, sizeof(float) * M));
  CUDA(cudaMalloc((void**)&d_v, sizeof(float) * M * T));
  CUDA(cudaMalloc((void**)&d_dv, sizeof(float) * M * T));
  CUDA(cudaMalloc((void**)&d_gbias, sizeof(float) * T));
  CUDA(cudaMalloc((void**)&d_gfc1, sizeof(float) * D * D));
  CUDA(cudaMalloc((void**)&d_dz, sizeof(float) * M * D));
  CUDA(cudaMemcpy(d_t, h_t, sizeof(float) * M, cudaMemcpyHostToDevice));
  CHECK(nova_head_train_forward(head, d_noise, d_t, d_z, M, d_v, d_tr, tr_bytes, stream));
  float* h_v = (float*)malloc(sizeof(float) * M * T);
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_v, d_v, sizeof(float) * M * T, cudaMemcpyDeviceToHost));
  double colsum[T] = {0.0, 0.0, 0.0};
  for (int64_t i = 0; i < M * T; ++i) {  /* loss = sum(v^2) / (2 M): dv = v / M */
    h_v[i] /= (float)M;
    colsum[i % T] += h_v[i];
  }
  CUDA(cudaMemcpy(d_dv, h_v, sizeof(float) * M * T, cudaMemcpyHostToDevice));
  const char* gnames[2] = {"head.bias", "blocks.0.proj.fc1.weight"};
  float* gptrs[2] = {d_gbias, d_gfc1};
  CHECK(nova_head_backward(head, d_dv, d_noise, d_z, M, 2, gnames, gptrs, d_dz, d_tr, tr_bytes, stream));
  float h_gbias[T];
  static float h_gfc1[D * D];
  CUDA(cudaStreamSynchronize(stream));
  CUDA(cudaMemcpy(h_gbias, d_gbias, sizeof(h_gbias), cudaMemcpyDeviceToHost));
  CUDA(cudaMemcpy(h_gfc1, d_gfc1, sizeof(h_gfc1), cudaMemcpyDeviceToHost));
  double gnorm = 0.0;
  for (int i = 0; i < D * D; ++i) {
    if (!isfinite(h_gfc1[i])) { fprintf(stderr, "non-finite weight gradient at %d\n", i); return 1; }
    gnorm += fabs(h_gfc1[i]);
  }
  for (int k = 0; k < T; ++k)
    if (fabs(h_gbias[k] - colsum[k]) > 1e-4 * (fabs(colsum[k]) + 1e-3)) {
      fprintf(stderr, "head.bias gradient %g vs column sum %g\n", h_gbias[k], colsum[k]);
      return 1;
    }
  if (gnorm == 0.0) { fprintf(stderr, "zero fc1 weight gradient\n"); return 1; }

  /* error path: a too-small workspace must be refused with a message, not crash */
  if (nova_head_sample(head, d_noise, d_z, NULL, B, B, N, N, timesteps, sigmas, S, &g, d_out, d_ws, 1024, stream) != NOVA_ERR_WORKSPACE) {
    fprintf(stderr, "small workspace was not rejected\n");
    return 1;
  }
  printf("c_abi_sample ok: mean |x| %.6f, mean NN distance %.6f, launches %lld, last refusal: %s\n", sum / (M * T),
         cd / (2 * N), (long long)nova_launch_count(), nova_last_error());
  CHECK(nova_head_destroy(head));
  return 0;
}
