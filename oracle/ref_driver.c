/*
 * ref_driver.c -- CPU-baseline driver around the REAL reference (oracle/_ref/libliquid_ref*.so).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY.  This is our code; it calls nothing but the reference's
 * public API (include/liquid.h:5754-5793 firpfbch2_crcf_*, :5667-5728 firpfbch_crcf_*,
 * :1985-2074 firfilt_crcf_*).  One reference object per thread, each on its own independent
 * synthetic stream (same counter-based generator as the GPU bench, seed + thread id), all
 * threads released together; reports the slowest thread's wall time.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "liquid.h"

static unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static void synth(unsigned long long seed, unsigned long long first, unsigned long long n, float *x)
{
    for (unsigned long long i = 0; i < n; i++) {
        unsigned long long k = seed + 2 * (first + i);
        x[2 * i]     = (float)(mix64(k) >> 40) * (1.0f / 16777216.0f) - 0.5f;
        x[2 * i + 1] = (float)(mix64(k + 1) >> 40) * (1.0f / 16777216.0f) - 0.5f;
    }
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

enum { KIND_PFBCH2_ANALYSIS = 0, KIND_PFBCH_ANALYSIS = 1, KIND_FIRFILT = 2 };

struct job {
    int kind;
    unsigned int M, m;
    float As;
    unsigned long long n;       /* input samples per thread */
    unsigned int id;
    pthread_barrier_t *go;
    double secs;
};

static void *worker(void *p)
{
    struct job *j = p;
    const unsigned int blk_frames = 256;
    unsigned int hop = (j->kind == KIND_PFBCH2_ANALYSIS) ? j->M / 2 : (j->kind == KIND_PFBCH_ANALYSIS ? j->M : 4096);
    size_t blk = (size_t)blk_frames * hop;
    float complex *x = malloc(blk * sizeof *x);
    float complex *y = malloc((size_t)blk_frames * (j->kind == KIND_FIRFILT ? hop : j->M) * sizeof *y);
    synth(0x5EED0000ULL + j->id, 0, blk, (float *)x);

    firpfbch2_crcf q2 = NULL; firpfbch_crcf q1 = NULL; firfilt_crcf qf = NULL;
    if (j->kind == KIND_PFBCH2_ANALYSIS) q2 = firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, j->M, j->m, j->As);
    if (j->kind == KIND_PFBCH_ANALYSIS)  q1 = firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, j->M, j->m, j->As);
    if (j->kind == KIND_FIRFILT)         qf = firfilt_crcf_create_kaiser(j->M, 0.25f, j->As, 0.0f);  /* M = taps */

    pthread_barrier_wait(j->go);
    double t0 = now_s();
    unsigned long long done = 0;
    while (done < j->n) {
        for (unsigned int f = 0; f < blk_frames && done < j->n; f++, done += hop) {
            if (q2)      firpfbch2_crcf_execute(q2, x + (size_t)f * hop, y + (size_t)f * j->M);
            else if (q1) firpfbch_crcf_analyzer_execute(q1, x + (size_t)f * hop, y + (size_t)f * j->M);
            else         firfilt_crcf_execute_block(qf, x + (size_t)f * hop, hop, y + (size_t)f * hop);
        }
    }
    j->secs = now_s() - t0;

    if (q2) firpfbch2_crcf_destroy(q2);
    if (q1) firpfbch_crcf_destroy(q1);
    if (qf) firfilt_crcf_destroy(qf);
    free(x); free(y);
    return NULL;
}

/* returns wall seconds of the slowest thread; each thread processes n_samples input samples */
double ref_bench(int kind, unsigned int M, unsigned int m, float As,
                 unsigned long long n_samples, unsigned int n_threads)
{
    pthread_t *th = malloc(n_threads * sizeof *th);
    struct job *jobs = malloc(n_threads * sizeof *jobs);
    pthread_barrier_t go;
    pthread_barrier_init(&go, NULL, n_threads);
    for (unsigned int i = 0; i < n_threads; i++) {
        jobs[i].kind = kind; jobs[i].M = M; jobs[i].m = m; jobs[i].As = As;
        jobs[i].n = n_samples; jobs[i].id = i; jobs[i].go = &go; jobs[i].secs = 0;
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    double worst = 0;
    for (unsigned int i = 0; i < n_threads; i++) {
        pthread_join(th[i], NULL);
        if (jobs[i].secs > worst) worst = jobs[i].secs;
    }
    pthread_barrier_destroy(&go);
    free(th); free(jobs);
    return worst;
}

/* stream helper for parity tests: run n_frames of firpfbch2 analysis from reset state */
void ref_pfbch2_analysis(unsigned int M, unsigned int m, float As,
                         const float complex *x, unsigned int n_frames, float complex *y)
{
    firpfbch2_crcf q = firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, m, As);
    for (unsigned int t = 0; t < n_frames; t++)
        firpfbch2_crcf_execute(q, (float complex *)x + (size_t)t * (M / 2), y + (size_t)t * M);
    firpfbch2_crcf_destroy(q);
}
