/* One rank of a time-sharded firpfbch2_crcf analysis driven through the C multi-GPU driver of libliquid_cuda
 * (include/liquid_cuda.h, "multi-GPU driver"): plain C99, no Python, no MPI -- the unique id travels through a file.
 *
 *   sharded_ranks <rank> <world> <id-file> <M> <m> <frames-per-rank>
 *
 * Every rank holds frames [rank*F, (rank+1)*F) of ONE synthetic stream (liquid_cuda_synth_input_dev), runs
 * firpfbch2_crcf_execute_sharded_dev (device pointers) and firpfbch2_crcf_execute_sharded (host pointers), and checks
 * both against the same frames of a single-GPU run over the whole stream up to the end of its segment.
 * Reference semantics: src/multichannel/src/firpfbch2.c:244-282 (k x execute == execute_block == the sharded run). */
#define _DEFAULT_SOURCE   /* usleep */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "liquid_cuda.h"

static double max_abs_diff(const float *a, const float *b, size_t n, double *scale)
{
    double d = 0.0, s = 0.0;
    for (size_t i = 0; i < n; i++) {
        double e = fabs((double)a[i] - (double)b[i]);
        if (e > d) d = e;
        if (fabs((double)b[i]) > s) s = fabs((double)b[i]);
    }
    *scale = s;
    return d;
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s rank world id-file M m frames-per-rank\n", argv[0]);
        return 2;
    }
    const int rank = atoi(argv[1]), world = atoi(argv[2]);
    const char *idfile = argv[3];
    const unsigned int M = (unsigned int)atoi(argv[4]), m = (unsigned int)atoi(argv[5]);
    const unsigned long long F = strtoull(argv[6], NULL, 10);
    const unsigned long long seed = 0x5EED0000ull;
    liquid_cuda_set_device(rank % liquid_cuda_device_count());

    /* rendezvous: rank 0 writes the id, the others wait for the file */
    const int idb = liquid_cuda_comm_unique_id_bytes();
    char *id = (char *)malloc((size_t)idb);
    if (rank == 0) {
        liquid_cuda_comm_get_unique_id(id);
        char tmp[4096];
        snprintf(tmp, sizeof tmp, "%s.tmp", idfile);
        FILE *f = fopen(tmp, "wb");
        fwrite(id, 1, (size_t)idb, f);
        fclose(f);
        rename(tmp, idfile);
    } else {
        FILE *f = NULL;
        for (int tries = 0; tries < 600 && !(f = fopen(idfile, "rb")); tries++) usleep(100000);
        if (!f || fread(id, 1, (size_t)idb, f) != (size_t)idb) {
            fprintf(stderr, "rank %d: no unique id\n", rank);
            return 2;
        }
        fclose(f);
    }
    liquid_cuda_comm comm = liquid_cuda_comm_create(rank, world, id);

    firpfbch2_crcf q = firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, m, 60.0f);
    const unsigned long long H = firpfbch2_crcf_get_state_len(q), M2 = M / 2, n = F * M2;
    float *d_buf = (float *)liquid_cuda_malloc(8 * (H + n));            /* head-room | segment */
    float *d_y = (float *)liquid_cuda_malloc(8 * F * M);
    /* head-room: zeros (rank 0 = stream start; the others are overwritten by the exchange) */
    float *zeros = (float *)calloc(2 * H, sizeof(float));
    liquid_cuda_memcpy_h2d(d_buf, zeros, 8 * H);
    liquid_cuda_synth_input_dev(d_buf + 2 * H, seed, (unsigned long long)rank * n, n, NULL);
    liquid_cuda_sync();

    /* (1) device-pointer form, twice (the second pass re-uses the head-room) */
    for (int pass = 0; pass < 2; pass++)
        firpfbch2_crcf_execute_sharded_dev(q, comm, d_buf, (unsigned long long)rank * F, F, d_y, NULL);
    liquid_cuda_sync();
    float *y_dev = (float *)malloc(8 * F * M);
    liquid_cuda_memcpy_d2h(y_dev, d_y, 8 * F * M);

    /* (2) host-pointer form */
    float *x_host = (float *)malloc(8 * n), *y_host = (float *)malloc(8 * F * M);
    liquid_cuda_memcpy_d2h(x_host, d_buf + 2 * H, 8 * n);
    firpfbch2_crcf_execute_sharded(q, comm, (liquid_float_complex *)x_host, (unsigned long long)rank * F, (unsigned int)F,
                                   (liquid_float_complex *)y_host);

    /* single-GPU check: the stream from its start to the end of this rank's segment through one object */
    firpfbch2_crcf qs = firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, m, 60.0f);
    const unsigned long long n_all = (unsigned long long)(rank + 1) * n;
    float *d_xa = (float *)liquid_cuda_malloc(8 * n_all), *d_ya = (float *)liquid_cuda_malloc(16 * n_all);
    liquid_cuda_synth_input_dev(d_xa, seed, 0, n_all, NULL);
    firpfbch2_crcf_execute_block_dev(qs, d_xa, (unsigned long long)(rank + 1) * F, d_ya, NULL);
    liquid_cuda_sync();
    float *y_ref = (float *)malloc(8 * F * M);
    liquid_cuda_memcpy_d2h(y_ref, d_ya + 2 * (unsigned long long)rank * F * M, 8 * F * M);

    double scale = 0.0;
    const double e_dev = max_abs_diff(y_dev, y_ref, 2 * F * M, &scale);
    const double e_host = max_abs_diff(y_host, y_ref, 2 * F * M, &scale);
    const int ok = e_dev <= 2e-6 * scale && e_host <= 2e-6 * scale && scale > 0.0;
    printf("rank %d of %d: M=%u m=%u frames %llu: max|sharded_dev - single| = %.3e, max|sharded_host - single| = %.3e (scale %.3e) %s\n",
           rank, world, M, m, F, e_dev, e_host, scale, ok ? "OK" : "MISMATCH");

    firpfbch2_crcf_destroy(q);
    firpfbch2_crcf_destroy(qs);
    liquid_cuda_comm_destroy(comm);
    liquid_cuda_free(d_buf);
    liquid_cuda_free(d_y);
    liquid_cuda_free(d_xa);
    liquid_cuda_free(d_ya);
    free(id); free(zeros); free(y_dev); free(x_host); free(y_host); free(y_ref);
    return ok ? 0 : 1;
}
