/*
 * A plain C caller of the drop-in boundary (include/difflinker_b200.h): the reverse-diffusion sampler of
 * EDM.sample_chain (reference src/edm.py:126-235) without Python and without a noise tensor.
 *
 *   c_sampler <job.bin> <out.bin>
 *
 * job.bin (little endian, written by difflinker_b200/export_job.py from a DDPM and a batch) holds the dl_config, the
 * weights under the reference's state_dict names, the normalised inputs and masks of one batch, the per-step
 * coefficient table of the noise schedule and a Philox (seed, offset) pair; the noise of the T+2 draws is generated
 * inside the kernels (dl_sample_chain_rng) in the order the reference's torch.randn calls would have produced it on
 * this GPU. out.bin: int32 status, uint64 philox offset consumed, the (keep_frames, B, N, 3+F) chain, B NaN flags.
 *
 * Build: gcc -std=c99 -O2 examples/c_sampler.c -Iinclude -I/usr/local/cuda/include -Ldifflinker_b200 -ldifflinker_b200 \
 *            -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/difflinker_b200 -o c_sampler
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "difflinker_b200.h"

static void die(const char* what) {
  fprintf(stderr, "c_sampler: %s (%s)\n", what, dl_last_error());
  exit(2);
}
static void rd(FILE* f, void* p, size_t n) {
  if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "c_sampler: short read\n"); exit(2); }
}
static void* rd_alloc(FILE* f, size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "c_sampler: out of memory\n"); exit(2); }
  rd(f, p, n);
  return p;
}
static void* to_device(const void* host, size_t n) {
  void* d = NULL;
  if (cudaMalloc(&d, n ? n : 1) != cudaSuccess || cudaMemcpy(d, host, n, cudaMemcpyHostToDevice) != cudaSuccess) {
    fprintf(stderr, "c_sampler: device copy of %zu bytes failed\n", n);
    exit(2);
  }
  return d;
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s job.bin out.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  char magic[8];
  rd(f, magic, 8);
  if (memcmp(magic, "DLJOB1\0\0", 8) != 0) { fprintf(stderr, "c_sampler: not a job file\n"); return 2; }

  dl_config cfg;
  rd(f, &cfg, sizeof cfg);                       /* 11 int32 + 2 float, no padding (checked by the exporter) */
  dl_engine* e = NULL;
  if (dl_create(&cfg, &e) < 0) die("dl_create");

  int32_t n_weights;
  rd(f, &n_weights, 4);
  for (int32_t i = 0; i < n_weights; ++i) {
    int32_t len; int64_t numel; char name[256];
    rd(f, &len, 4);
    if (len <= 0 || len >= (int32_t)sizeof name) { fprintf(stderr, "c_sampler: bad weight name\n"); return 2; }
    rd(f, name, (size_t)len);
    name[len] = 0;
    rd(f, &numel, 8);
    float* w = (float*)rd_alloc(f, (size_t)numel * 4);
    if (dl_set_weight(e, name, w, numel) < 0) die(name);
    free(w);
  }
  if (dl_finalize_weights(e) < 0) die("dl_finalize_weights");

  int32_t dims[6];                               /* B, N, T, keep_frames, xd = 3 + F, C */
  uint64_t rng[2];                               /* philox seed, offset */
  float norm[3];
  rd(f, dims, sizeof dims);
  rd(f, rng, sizeof rng);
  rd(f, norm, sizeof norm);
  const int32_t B = dims[0], N = dims[1], T = dims[2], keep = dims[3], xd = dims[4], C = dims[5];
  const size_t n = (size_t)B * N;
  dl_step_coef* coef = (dl_step_coef*)rd_alloc(f, (size_t)(T + 1) * sizeof(dl_step_coef));   /* host table, as the ABI asks */
  float* xh = (float*)rd_alloc(f, n * xd * 4);
  int8_t* node_mask = (int8_t*)rd_alloc(f, n);
  float* fragment_mask = (float*)rd_alloc(f, n * 4);
  float* linker_mask = (float*)rd_alloc(f, n * 4);
  int32_t has_em, has_ctx;
  rd(f, &has_em, 4);
  int8_t* edge_mask = has_em ? (int8_t*)rd_alloc(f, n * N) : NULL;
  rd(f, &has_ctx, 4);
  float* context = has_ctx ? (float*)rd_alloc(f, n * C * 4) : NULL;
  fclose(f);

  if (cudaSetDevice(cfg.device) != cudaSuccess) { fprintf(stderr, "c_sampler: no CUDA device %d\n", cfg.device); return 2; }
  float* d_xh = (float*)to_device(xh, n * xd * 4);
  int8_t* d_nm = (int8_t*)to_device(node_mask, n);
  float* d_fm = (float*)to_device(fragment_mask, n * 4);
  float* d_lm = (float*)to_device(linker_mask, n * 4);
  int8_t* d_em = has_em ? (int8_t*)to_device(edge_mask, n * N) : NULL;
  float* d_ctx = has_ctx ? (float*)to_device(context, n * C * 4) : NULL;
  float* d_chain = NULL;
  int32_t* d_flags = NULL;
  const size_t chain_bytes = (size_t)keep * n * xd * 4;
  if (cudaMalloc((void**)&d_chain, chain_bytes) != cudaSuccess || cudaMalloc((void**)&d_flags, (size_t)B * 4) != cudaSuccess) {
    fprintf(stderr, "c_sampler: cudaMalloc failed\n");
    return 2;
  }
  cudaStream_t stream;
  if (cudaStreamCreate(&stream) != cudaSuccess) { fprintf(stderr, "c_sampler: cudaStreamCreate failed\n"); return 2; }

  uint64_t consumed = 0;
  const dl_status st = dl_sample_chain_rng(e, DL_SAMPLER_LINKER, B, N, T, keep, d_xh, d_nm, d_fm, d_lm, d_em, d_ctx, rng[0], rng[1],
                                           &consumed, coef, norm, d_chain, d_flags, stream);
  if (st < 0) die("dl_sample_chain_rng");
  if (cudaStreamSynchronize(stream) != cudaSuccess) { fprintf(stderr, "c_sampler: the sampler's stream failed\n"); return 2; }

  float* chain = (float*)malloc(chain_bytes);
  int32_t* flags = (int32_t*)malloc((size_t)B * 4);
  cudaMemcpy(chain, d_chain, chain_bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(flags, d_flags, (size_t)B * 4, cudaMemcpyDeviceToHost);
  FILE* o = fopen(argv[2], "wb");
  if (!o) { perror(argv[2]); return 2; }
  const int32_t status = (int32_t)st;
  fwrite(&status, 4, 1, o);
  fwrite(&consumed, 8, 1, o);
  fwrite(chain, 1, chain_bytes, o);
  fwrite(flags, 4, (size_t)B, o);
  fclose(o);
  printf("c_sampler: %d molecules x %d atoms, T=%d: %.2f ms on the device, %lld kernels, philox offset +%llu\n", B, N, T,
         dl_last_elapsed_ms(e), (long long)dl_launch_count(e), (unsigned long long)consumed);
  dl_destroy(e);
  return 0;
}
