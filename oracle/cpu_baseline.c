/*
 * cpu_baseline.c -- ORACLE (test infrastructure).  Multi-threaded driver of the f32 RX chain used
 * ONLY by bench.py's cpu_baseline / --impl reference legs: one independent channel buffer per
 * thread (mirrors GNU Radio's thread-per-block parallelism without its ring buffers, which
 * favours the CPU).
 */
#define _GNU_SOURCE
#include "amps_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

typedef struct { const float *iq; size_t n; uint32_t fcw; const float *h2; int nh2; float *d; int nb; int reps; } job;

static void *run(void *p) {
    job *j = (job *)p;
    orc_burst *b = (orc_burst *)malloc(sizeof(orc_burst) * 64);
    j->nb = 0;
    for (int rep = 0; rep < j->reps; rep++) {
        orc_rx_chain_f32(j->iq, j->n, j->fcw, j->h2, j->nh2, NULL, j->d);
        int nb = orc_rx_detect(j->d, j->n / 50, b, 64);
        for (int i = 0; i < nb; i++) { orc_recc_result r; orc_recc_decode(b[i].symbols, &r); }
        j->nb += nb;
    }
    free(b);
    return NULL;
}

/* Runs `threads` independent channels, each `reps` times over the same n-sample buffer; returns wall seconds,
 * total bursts decoded via *nbursts.  Samples processed = threads * reps * n. */
double orc_cpu_baseline_run(const float *iq, size_t n, uint32_t fcw, const float *h2, int nh2, int threads, int reps, int *nbursts) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job *jobs = (job *)malloc(sizeof(job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job){iq, n, fcw, h2, nh2, (float *)malloc(sizeof(float) * (n / 50 + 1)), 0, reps};
    }
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, run, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    int nb = 0;
    for (int t = 0; t < threads; t++) { nb += jobs[t].nb; free(jobs[t].d); }
    if (nbursts) *nbursts = nb;
    free(th); free(jobs);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
