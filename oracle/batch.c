/*
 * batch.c -- pthread batch runner used as the CPU baseline (test infrastructure;
 * see oracle.h).  One contiguous chunk range per thread, wall-clock timed.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <dlfcn.h>
#include <pthread.h>
#include <time.h>

/* ORACLE_LIBLZ4: the CPU decoder the reference itself links for its LZ4 known-answer tests
 * (reference examples/lz4_cpu_decompression.cu:143-147: LZ4_decompress_safe), taken from the box's own
 * liblz4.so.1 at run time.  Timed next to the port so the LZ4 baseline is the real library. */
typedef int (*lz4_safe_fn)(const char*, char*, int, int);
static lz4_safe_fn g_lz4_safe;
int oracle_have_liblz4(void)
{
  if (!g_lz4_safe) {
    void* h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) g_lz4_safe = (lz4_safe_fn)dlsym(h, "LZ4_decompress_safe");
  }
  return g_lz4_safe != 0;
}

long oracle_cascaded_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) __attribute__((weak));
long oracle_bitcomp_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) __attribute__((weak));
long oracle_ans_decompress(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) __attribute__((weak));

static pthread_barrier_t g_start;

typedef struct {
  int codec;
  const uint8_t* comp;
  const size_t* off;
  const size_t* len;
  size_t begin, end;
  uint8_t* out;
  size_t stride;
  size_t* out_len;
  int failed;
  struct timespec t_start, t_end;   /* stamped by the worker itself, after the start barrier */
} job_t;

static void* worker(void* p)
{
  job_t* j = (job_t*)p;
  pthread_barrier_wait(&g_start);   /* timing starts when every thread is up (spawn cost excluded) */
  clock_gettime(CLOCK_MONOTONIC, &j->t_start);
  for (size_t i = j->begin; i < j->end; ++i) {
    long r = -1;
    const uint8_t* s = j->comp + j->off[i];
    uint8_t* d = j->out + i * j->stride;
    switch (j->codec) {
    case ORACLE_LZ4: r = oracle_lz4_decompress(s, j->len[i], d, j->stride); break;
    case ORACLE_SNAPPY: r = oracle_snappy_decompress(s, j->len[i], d, j->stride); break;
    case ORACLE_CASCADED: if (oracle_cascaded_decompress) r = oracle_cascaded_decompress(s, j->len[i], d, j->stride); break;
    case ORACLE_BITCOMP: if (oracle_bitcomp_decompress) r = oracle_bitcomp_decompress(s, j->len[i], d, j->stride); break;
    case ORACLE_ANS: if (oracle_ans_decompress) r = oracle_ans_decompress(s, j->len[i], d, j->stride); break;
    case ORACLE_LIBLZ4: if (g_lz4_safe) r = g_lz4_safe((const char*)s, (char*)d, (int)j->len[i], (int)j->stride); break;
    default: break;
    }
    if (r < 0) { j->failed = 1; r = 0; }
    if (j->out_len) j->out_len[i] = (size_t)r;
  }
  clock_gettime(CLOCK_MONOTONIC, &j->t_end);
  return 0;
}

double oracle_batch_decompress(int codec, const uint8_t* comp, const size_t* comp_off,
                               const size_t* comp_len, size_t count, uint8_t* out,
                               size_t out_stride, size_t* out_len, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (codec == ORACLE_LIBLZ4 && !oracle_have_liblz4()) return -1.0;
  pthread_t th[256];
  job_t jobs[256];
  struct timespec t0, t1;
  pthread_barrier_init(&g_start, 0, (unsigned)nthreads + 1);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (job_t){codec, comp, comp_off, comp_len, count * t / nthreads, count * (t + 1) / nthreads,
                      out, out_stride, out_len, 0, {0, 0}, {0, 0}};
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  pthread_barrier_wait(&g_start);
  int failed = 0;
  for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], 0); failed |= jobs[t].failed; }
  /* elapsed = last worker finish - first worker start (the caller thread may be descheduled) */
  t0 = jobs[0].t_start; t1 = jobs[0].t_end;
  for (int t = 1; t < nthreads; ++t) {
    if (jobs[t].t_start.tv_sec < t0.tv_sec || (jobs[t].t_start.tv_sec == t0.tv_sec && jobs[t].t_start.tv_nsec < t0.tv_nsec)) t0 = jobs[t].t_start;
    if (jobs[t].t_end.tv_sec > t1.tv_sec || (jobs[t].t_end.tv_sec == t1.tv_sec && jobs[t].t_end.tv_nsec > t1.tv_nsec)) t1 = jobs[t].t_end;
  }
  pthread_barrier_destroy(&g_start);
  double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return failed ? -s : s;
}
