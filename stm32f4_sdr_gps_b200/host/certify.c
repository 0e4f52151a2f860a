/*
 * certify.c - proof by exhaustion that the device-resident tracking loop and a host-resident one compute the
 * same loop-filter inputs on THIS machine.
 *
 * The only operations of the per-millisecond step that are not plain IEEE-754 arithmetic are the arctangents
 * of the Costas and frequency discriminators (Firmware/project_main/GPS/tracking.c:180-183, 232-233).  They
 * are evaluated on the prompt sums IP, QP - two integers in [-8184, 8184] (gps_misc.c:140-141) - so their
 * whole input domain has 16369^2 points.  The device values (gpsb_l0_loop_math: fdlibm atanf/atan2f restated in
 * core/gpsb_loop_core.h, CUDA's double atan2) are compared with the host libm's for every one of them.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

#include "host_internal.h"

#define CERT_Q      16369          /* values of QP per IP row */
#define CERT_ROWS   64             /* IP rows per work item: 64 x 16369 x 4 B = 4 MB per transfer */

typedef struct cert_job {
    gpsb_ctx* ctx;
    volatile int32_t* next_row;    /* shared cursor over ip = -8184 .. 8184 */
    pthread_mutex_t* lock;
    int64_t bad;
    int rc;
} cert_job;

static float host_costas(int ip, int qp)
{
    /* same expressions, same libm entry points as the host-resident loop (core/gpsb_loop_core.h, host build) */
    return lc_costas_err((int16_t)ip, (int16_t)qp);
}

static void* cert_worker(void* arg)
{
    cert_job* job = (cert_job*)arg;
    float* dev = (float*)malloc((size_t)CERT_ROWS * CERT_Q * sizeof(float));
    if (!dev) { job->rc = GPSB_ERR_NOMEM; return NULL; }
    for (;;) {
        pthread_mutex_lock(job->lock);
        int32_t lo = *job->next_row;
        *job->next_row = lo + CERT_ROWS;
        pthread_mutex_unlock(job->lock);
        if (lo > 8184) break;
        uint32_t rows = (uint32_t)(lo + CERT_ROWS > 8185 ? 8185 - lo : CERT_ROWS);
        for (int kind = 0; kind < 2; kind++) {
            int rc = gpsb_l0_loop_math(job->ctx, kind, lo, rows, dev);
            if (rc != GPSB_OK) { job->rc = rc; free(dev); return NULL; }
            for (uint32_t r = 0; r < rows; r++) {
                const int ip = lo + (int)r;
                for (int qp = -8184; qp <= 8184; qp++) {
                    float want = kind == 0 ? host_costas(ip, qp) : lc_fll_angle((int16_t)ip, (int16_t)qp);
                    float got = dev[(size_t)r * CERT_Q + (size_t)(qp + 8184)];
                    if (lc_float_bits(want) != lc_float_bits(got)) job->bad++;
                }
            }
        }
    }
    free(dev);
    return NULL;
}

int64_t gpsb_host_certify_loop_math(gpsb_ctx* ctx, uint32_t n_threads)
{
    if (!ctx) return GPSB_ERR_ARG;
    if (n_threads == 0) {
        long cpus = sysconf(_SC_NPROCESSORS_ONLN);
        n_threads = cpus > 0 ? (uint32_t)cpus : 1;
    }
    if (n_threads > 64) n_threads = 64;
    volatile int32_t next_row = -8184;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    pthread_t tid[64];
    cert_job job[64];
    uint32_t started = 0;
    for (uint32_t w = 0; w < n_threads; w++) {
        job[w] = (cert_job){ctx, &next_row, &lock, 0, GPSB_OK};
        if (pthread_create(&tid[w], NULL, cert_worker, &job[w]) != 0) break;
        started++;
    }
    if (started == 0) {                                  /* no thread could be started: do it here */
        job[0] = (cert_job){ctx, &next_row, &lock, 0, GPSB_OK};
        cert_worker(&job[0]);
        return job[0].rc != GPSB_OK ? hx_note(job[0].rc) : job[0].bad;
    }
    int64_t bad = 0;
    int rc = GPSB_OK;
    for (uint32_t w = 0; w < started; w++) {
        pthread_join(tid[w], NULL);
        bad += job[w].bad;
        if (job[w].rc != GPSB_OK) rc = job[w].rc;
    }
    return rc != GPSB_OK ? hx_note(rc) : bad;
}
