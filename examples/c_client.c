/* Minimal plain-C client of the C ABI (include/rag_b200.h): shows that the boundary needs nothing but a C compiler and
 * the shared library.  Build and run from the repo root:
 *
 *   gcc -std=c99 -Wall -Iinclude examples/c_client.c -o /tmp/c_client \
 *       -Lautomative-rag_b200/lib -lrag_b200 -Wl,-rpath,$PWD/automative-rag_b200/lib && /tmp/c_client
 *
 * Without a B200 it reports the scan's shared-memory plan (host-side arithmetic) and the error rs_create returns —
 * there is no CPU fallback.  With one it runs a 4096 x 64 fp16 top-5 search on device 0 through rs_dense_topk_host
 * with a corpus of its own (rows e_i scaled by i), so the answer is known in closed form.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rag_b200.h"

/* the CUDA runtime, for the one device buffer this example owns */
extern int cudaMalloc(void** p, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaFree(void* p);

static unsigned short f32_to_f16(float f) { /* exact for the small integers used here */
  unsigned int u;
  memcpy(&u, &f, 4);
  if (f == 0.0f) return 0;
  {
    int e = (int)((u >> 23) & 0xFF) - 127 + 15;
    return (unsigned short)(((u >> 16) & 0x8000u) | ((unsigned)e << 10) | ((u >> 13) & 0x3FFu));
  }
}

int main(void) {
  int64_t plan[7];
  rs_handle* h = NULL;
  int rc;
  printf("ABI version %d\n", rs_abi_version());
  if (rs_scan_plan(1024, 10, plan) == RS_OK)
    printf("scan plan d=1024 k=10: %lld rows/tile, %lld consumer warps, %lld ring slots, buffer %lld keys, %lld B smem\n",
           (long long)plan[0], (long long)plan[1], (long long)plan[2], (long long)plan[3], (long long)plan[6]);
  rc = rs_create(0, &h);
  if (rc != RS_OK) {
    printf("rs_create: status %d: %s\n", rc, rs_last_error(NULL));
    return rc == RS_ERR_NO_DEVICE ? 0 : 1;
  }
  {
    enum { N = 4096, D = 64, K = 5 };
    unsigned short* corpus = (unsigned short*)calloc((size_t)N * D, 2);
    unsigned short query[D];
    float scores[K];
    int64_t ids[K];
    void* dcorpus = NULL;
    int i;
    for (i = 0; i < N; ++i) corpus[(size_t)i * D + i % D] = f32_to_f16((float)(i / D + 1)); /* row i = (i/D + 1) e_(i%D) */
    for (i = 0; i < D; ++i) query[i] = f32_to_f16(i == 7 ? 1.0f : 0.0f);                     /* q = e_7 */
    if (cudaMalloc(&dcorpus, (size_t)N * D * 2) != 0 || cudaMemcpy(dcorpus, corpus, (size_t)N * D * 2, 1) != 0) return 1;
    rc = rs_dense_topk_host(h, dcorpus, N, D, RS_F16, NULL, RS_METRIC_IP, query, 1, NULL, NULL, 0, K, 0, scores, ids, NULL);
    if (rc != RS_OK) {
      printf("rs_dense_topk_host: %s\n", rs_last_error(h));
      return 1;
    }
    for (i = 0; i < K; ++i) printf("  #%d id %lld score %.1f\n", i, (long long)ids[i], scores[i]);
    /* rows with column 7 set are 7, 71, 135, ...; the largest values sit at the largest ids: 4039 (64), 3975 (63), ... */
    rc = (ids[0] == 4039 && scores[0] == 64.0f && ids[1] == 3975 && scores[1] == 63.0f) ? 0 : 1;
    printf(rc == 0 ? "ok\n" : "MISMATCH\n");
    cudaFree(dcorpus);
    free(corpus);
  }
  rs_destroy(h);
  return rc;
}
