/*
 * rx.c - batched receiver: every channel's correlation for one millisecond (or one whole cold-start
 * sweep) goes to the GPU in a single launch; the per-channel votes and loop filters then run on the
 * host exactly as in acq.c / track.c / nav.c.
 *
 * Each channel owns its own gpsb_aux, so the result for a channel equals what the reference computes
 * when that channel is the only active one (the reference's shared scratch, acquisition.c:28-33,
 * tracking.c:33-34, nav_data.c:29,48-51, only works one channel at a time).
 */
#include <pthread.h>
#include <sched.h>
#include <time.h>
#include <stdlib.h>
#include <unistd.h>

#include "../../include/gpsb_flat_state.h"
#include "host_internal.h"

struct gpsb_rx {
    gpsb_ctx* ctx;
    gps_ch_t* ch;
    uint32_t n_ch;
    gpsb_aux* aux;
    gpsb_plan* plan;
    gpsb_epl_req* epl_rq;
    int16_t* epl_out;
    uint32_t* epl_owner;
    gpsb_search_req* s_rq;
    gpsb_search_res* s_res;
    uint32_t* s_owner;
    uint32_t threads;            /* 0 = automatic */
    int loop_site;               /* GPSB_LOOP_AUTO / _HOST / _DEVICE */
    gpsb_loop_result* loop_res;
    uint32_t ring_ms;            /* capacity of the context's signal ring */
    uint64_t device_ms, host_ms; /* channel-milliseconds run by k_track_run / by the per-millisecond host path */
    uint8_t* idx;                /* slot index of each channel in the millisecond being processed (gpsb_rx_track_ms) */
};

/* Slot index of channel i at millisecond ms: (ms + slot phase) % 4, or LC_IDLE_INDEX inside an idle gap of the
 * slot-phase walk (core/gpsb_loop_core.h, lc_walk_*) - then the channel is not served, like a channel the MCU's
 * 17-ms schedule does not reach (main.c:134-155). */
static uint8_t rx_index(gpsb_rx* rx, uint32_t i, uint32_t ms) { return lc_walk_index(&rx->aux[i], ms); }
/* after the loop filters of an E/P/L millisecond: at the end of a slot the walk policy may move the slots */
static void rx_slot_end(gpsb_rx* rx, uint32_t i, uint32_t ms, uint8_t index)
{
    if (index == GPSB_SLOT_LEN - 1 && rx->ch[i].tracking_data.state == GPS_TRACKING_RUN)
        lc_walk_policy(&rx->ch[i], &rx->aux[i], ms);
}

typedef struct track_job {
    gpsb_rx* rx;
    uint32_t worker, n_workers, ms0, n_ms;
    int16_t* iq_log;
    int8_t* nav_log;
    int rc;
} track_job;

int gpsb_rx_create(gpsb_rx** out, gpsb_ctx* ctx, gps_ch_t* channels, uint32_t n_ch)
{
    if (!out || !ctx || !channels || n_ch == 0) return GPSB_ERR_ARG;
    gpsb_rx* rx = (gpsb_rx*)calloc(1, sizeof *rx);
    if (!rx) return GPSB_ERR_NOMEM;
    rx->ctx = ctx;
    rx->ring_ms = gpsb_ring_ms(ctx);
    rx->ch = channels;
    rx->n_ch = n_ch;
    rx->aux = (gpsb_aux*)calloc(n_ch, sizeof(gpsb_aux));
    rx->plan = (gpsb_plan*)calloc(n_ch, sizeof(gpsb_plan));
    rx->epl_rq = (gpsb_epl_req*)calloc(n_ch, sizeof(gpsb_epl_req));
    rx->epl_out = (int16_t*)calloc((size_t)n_ch * 6, sizeof(int16_t));
    rx->epl_owner = (uint32_t*)calloc(n_ch, sizeof(uint32_t));
    rx->s_rq = (gpsb_search_req*)calloc(n_ch, sizeof(gpsb_search_req));
    rx->s_res = (gpsb_search_res*)calloc(n_ch, sizeof(gpsb_search_res));
    rx->s_owner = (uint32_t*)calloc(n_ch, sizeof(uint32_t));
    rx->loop_res = (gpsb_loop_result*)calloc(n_ch, sizeof(gpsb_loop_result));
    rx->idx = (uint8_t*)calloc(n_ch, 1);
    if (!rx->idx || !rx->loop_res || !rx->aux || !rx->plan || !rx->epl_rq || !rx->epl_out || !rx->epl_owner || !rx->s_rq || !rx->s_res ||
        !rx->s_owner) {
        gpsb_rx_destroy(rx);
        return GPSB_ERR_NOMEM;
    }
    for (uint32_t i = 0; i < n_ch; i++) {
        rx->aux[i].last_nav_bit = -1;
        if (channels[i].prn >= 1) {
            gps_generate_prn(channels[i].prn_code, channels[i].prn);
            int rc = gpsb_set_code(ctx, channels[i].prn, channels[i].prn_code);
            if (rc != GPSB_OK) {
                gpsb_rx_destroy(rx);
                return rc;
            }
        }
    }
    *out = rx;
    return GPSB_OK;
}

void gpsb_rx_destroy(gpsb_rx* rx)
{
    if (!rx) return;
    free(rx->aux);
    free(rx->plan);
    free(rx->epl_rq);
    free(rx->epl_out);
    free(rx->epl_owner);
    free(rx->s_rq);
    free(rx->s_res);
    free(rx->s_owner);
    free(rx->loop_res);
    free(rx->idx);
    free(rx);
}

/* Gather the planned cells, run them (one launch per kind), return the counts. */
static int run_plans(gpsb_rx* rx, uint32_t* n_epl_out, uint32_t* n_s_out)
{
    uint32_t n_epl = 0, n_s = 0;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        gpsb_plan* p = &rx->plan[i];
        if (p->want == GPSB_WANT_EPL) {
            rx->epl_rq[n_epl] = p->epl;
            rx->epl_owner[n_epl++] = i;
        } else if (p->want == GPSB_WANT_SEARCH) {
            if (p->search.start < p->search.stop) {
                rx->s_rq[n_s] = p->search;
                rx->s_owner[n_s++] = i;
            }
        }
    }
    int rc = GPSB_OK;
    if (n_epl) rc = gpsb_track_epl(rx->ctx, n_epl, rx->epl_rq, rx->epl_out);
    if (rc == GPSB_OK && n_s) rc = gpsb_search(rx->ctx, n_s, rx->s_rq, rx->s_res);
    *n_epl_out = n_epl;
    *n_s_out = n_s;
    return hx_note(rc);
}

static const gpsb_search_res k_empty_window = {0, 0, 0, 0};

int gpsb_rx_track_ms(gpsb_rx* rx, uint32_t ms)
{
    if (!rx) return GPSB_ERR_ARG;
    gpsb_host_set_packet_cnt(ms);
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        rx->idx[i] = rx_index(rx, i, ms);
        hx_trk_plan(&rx->ch[i], &rx->aux[i], ms, rx->idx[i], &rx->plan[i]);
    }
    uint32_t n_epl, n_s;
    int rc = run_plans(rx, &n_epl, &n_s);
    if (rc != GPSB_OK) return rc;
    for (uint32_t k = 0; k < n_epl; k++) {
        uint32_t i = rx->epl_owner[k];
        hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], rx->idx[i], rx->epl_out + 6u * k);
        rx_slot_end(rx, i, ms, rx->idx[i]);
    }
    uint32_t k = 0;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->plan[i].want != GPSB_WANT_SEARCH) continue;
        const gpsb_search_res* r = &k_empty_window;
        if (k < n_s && rx->s_owner[k] == i) r = &rx->s_res[k++];
        hx_trk_finish_search(&rx->ch[i], &rx->aux[i], rx->idx[i], r);
    }
    rx->host_ms += n_epl + n_s;
    return GPSB_OK;
}

/* One worker of a threaded run: drives its own channels (i % n_workers == worker) through all n_ms milliseconds
 * at its own pace - channels are independent, so nothing forces them through the millisecond in lockstep.  Per
 * ms it posts every channel's cell to that channel's session slot first and only then collects, so the PCIe
 * round trips of its channels overlap. */
static void* track_worker(void* arg)
{
    track_job* job = (track_job*)arg;
    gpsb_rx* rx = job->rx;
    uint32_t seq[128];
    job->rc = GPSB_OK;
    for (uint32_t m = 0; m < job->n_ms && job->rc == GPSB_OK; m++) {
        const uint32_t ms = job->ms0 + m;
        gpsb_host_set_packet_cnt(ms);
        for (uint32_t i = job->worker; i < rx->n_ch; i += job->n_workers) {
            gpsb_plan* p = &rx->plan[i];
            rx->aux[i].last_nav_bit = -1;
            rx->idx[i] = rx_index(rx, i, ms);
            hx_trk_plan(&rx->ch[i], &rx->aux[i], ms, rx->idx[i], p);
            int rc = gpsb_session_post(rx->ctx, i, p->want == GPSB_WANT_EPL ? &p->epl : NULL, &seq[i]);
            if (rc != GPSB_OK) job->rc = hx_note(rc);
        }
        for (uint32_t i = job->worker; i < rx->n_ch && job->rc == GPSB_OK; i += job->n_workers) {
            gpsb_plan* p = &rx->plan[i];
            const uint8_t index = rx->idx[i];
            int16_t iq[6] = {0, 0, 0, 0, 0, 0};
            if (p->want == GPSB_WANT_EPL) {
                int rc = gpsb_session_wait(rx->ctx, i, seq[i], iq);
                if (rc != GPSB_OK) { job->rc = hx_note(rc); break; }
                hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], index, iq);
                rx_slot_end(rx, i, ms, index);
            } else if (p->want == GPSB_WANT_SEARCH) {
                gpsb_search_res res = {0, 0, 0, 0};
                if (p->search.start < p->search.stop) {
                    int rc = gpsb_search(rx->ctx, 1, &p->search, &res);
                    if (rc != GPSB_OK) { job->rc = hx_note(rc); break; }
                }
                hx_trk_finish_search(&rx->ch[i], &rx->aux[i], index, &res);
            }
            if (job->iq_log) memcpy(job->iq_log + ((size_t)m * rx->n_ch + i) * 6, iq, 12);
            if (job->nav_log)
                job->nav_log[(size_t)m * rx->n_ch + i] = p->want == GPSB_WANT_EPL ? rx->aux[i].last_nav_bit : -1;
        }
    }
    return NULL;
}

void gpsb_rx_set_threads(gpsb_rx* rx, uint32_t n) { if (rx) rx->threads = n; }

static uint32_t pick_workers(const gpsb_rx* rx)
{
    uint32_t n = rx->threads;
    if (n == 0) {
        long cpus = sysconf(_SC_NPROCESSORS_ONLN);
        n = cpus > 2 ? (uint32_t)(cpus - 1) : 1;
        if (n > 16) n = 16;
    }
    if (n > rx->n_ch) n = rx->n_ch;
    return n ? n : 1;
}

void gpsb_rx_set_loop_site(gpsb_rx* rx, int site) { if (rx) rx->loop_site = site; }

void gpsb_rx_loop_stats(const gpsb_rx* rx, uint64_t* device_ms, uint64_t* host_ms)
{
    if (device_ms) *device_ms = rx ? rx->device_ms : 0;
    if (host_ms) *host_ms = rx ? rx->host_ms : 0;
}

static int is_tracking(const gps_ch_t* ch)
{
    return ch->tracking_data.state == GPS_TRACKING_RUN || ch->tracking_data.state == GPS_PRE_TRACK_DONE;
}

/* One millisecond of ONE channel on the per-millisecond path (plan on the host, cell on the GPU, finish on the
 * host): what the device-resident loop hands back, and channels that are still in pre-track. */
static int host_step(gpsb_rx* rx, uint32_t i, uint32_t ms, int16_t* iq_row, int8_t* nav_row)
{
    const uint8_t index = rx_index(rx, i, ms);
    gpsb_plan* p = &rx->plan[i];
    int16_t iq[6] = {0, 0, 0, 0, 0, 0};
    gpsb_host_set_packet_cnt(ms);
    rx->aux[i].last_nav_bit = -1;
    const gps_tracking_t before = rx->ch[i].tracking_data;     /* put back if the GPU step fails (see host/track.c) */
    hx_trk_plan(&rx->ch[i], &rx->aux[i], ms, index, p);
    if (p->want == GPSB_WANT_EPL) {
        int rc = gpsb_track_epl(rx->ctx, 1, &p->epl, iq);
        if (rc != GPSB_OK) { rx->ch[i].tracking_data = before; return hx_note(rc); }
        hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], index, iq);
        rx_slot_end(rx, i, ms, index);
    } else if (p->want == GPSB_WANT_SEARCH) {
        gpsb_search_res res = {0, 0, 0, 0};
        if (p->search.start < p->search.stop) {
            int rc = gpsb_search(rx->ctx, 1, &p->search, &res);
            if (rc != GPSB_OK) { rx->ch[i].tracking_data = before; return hx_note(rc); }
        }
        hx_trk_finish_search(&rx->ch[i], &rx->aux[i], index, &res);
    }
    if (iq_row) memcpy(iq_row + 6u * i, iq, 12);
    if (nav_row) nav_row[i] = p->want == GPSB_WANT_EPL ? rx->aux[i].last_nav_bit : -1;
    rx->host_ms++;
    return GPSB_OK;
}

/* Channel i over [ms, end): device-resident loop while the channel is in a tracking state, single host steps for
 * what the loop hands back (a degenerate DLL millisecond) and for channels not tracking yet. */
static int run_channel_span(gpsb_rx* rx, uint32_t i, uint32_t ms0, uint32_t ms, uint32_t end, int16_t* iq_log,
                            int8_t* nav_log)
{
    const uint32_t n_ch = rx->n_ch;
    /* a channel that has not been handed to tracking (still in acquisition, or given up) has nothing to do here */
    if (rx->ch[i].tracking_data.state == GPS_TRACKNG_IDLE) return GPSB_OK;
    while (ms < end) {
        int16_t* iq_row = iq_log ? iq_log + (size_t)(ms - ms0) * n_ch * 6 : NULL;
        int8_t* nav_row = nav_log ? nav_log + (size_t)(ms - ms0) * n_ch : NULL;
        if (!is_tracking(&rx->ch[i])) {
            int rc = host_step(rx, i, ms, iq_row, nav_row);
            if (rc != GPSB_OK) return rc;
            ms++;
            continue;
        }
        const uint32_t n = end - ms;
        int16_t* iq_tmp = iq_log ? (int16_t*)malloc((size_t)n * 12) : NULL;
        int8_t* nav_tmp = nav_log ? (int8_t*)malloc(n) : NULL;
        if ((iq_log && !iq_tmp) || (nav_log && !nav_tmp)) {
            free(iq_tmp);
            free(nav_tmp);
            return hx_note(GPSB_ERR_NOMEM);
        }
        gpsb_loop_result r;
        int rc = gpsb_track_loop(rx->ctx, 1, &rx->ch[i], (uint32_t)sizeof(gps_ch_t), &rx->aux[i], (uint32_t)sizeof(gpsb_aux),
                                 ms, n, iq_tmp, nav_tmp, &r);
        if (rc == GPSB_OK) {
            const uint32_t rows = r.done_ms + (r.stop == LC_STOP_DLL_NAN ? 1u : 0u);
            for (uint32_t k = 0; k < rows && k < n; k++) {
                if (iq_log) memcpy(iq_log + ((size_t)(ms - ms0 + k) * n_ch + i) * 6, iq_tmp + 6u * k, 12);
                if (nav_log) nav_log[(size_t)(ms - ms0 + k) * n_ch + i] = nav_tmp[k];
            }
        }
        free(iq_tmp);
        free(nav_tmp);
        if (rc != GPSB_OK) return hx_note(rc);
        lc_resolve_snr(&rx->ch[i], &rx->aux[i]);
        rx->device_ms += r.done_ms;
        ms += r.done_ms;
        if (r.stop == LC_STOP_DLL_NAN && ms < end) {          /* sums delivered, filters not run: finish on the host */
            gpsb_host_set_packet_cnt(ms);
            rx->aux[i].last_nav_bit = -1;
            const uint8_t index = rx_index(rx, i, ms);
            hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], index, r.iq);
            rx_slot_end(rx, i, ms, index);
            if (nav_log) nav_log[(size_t)(ms - ms0) * n_ch + i] = rx->aux[i].last_nav_bit;
            rx->host_ms++;
            ms++;
        }
    }
    return GPSB_OK;
}

/* After a k_track_run launch over all channels: SNR values, statistics, and whatever the loop handed back (a
 * degenerate DLL millisecond, a starved streaming run) finished channel by channel on the per-millisecond path. */
static int finish_device_run(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log)
{
    const uint32_t n_ch = rx->n_ch;
    for (uint32_t i = 0; i < n_ch; i++) {
        const gpsb_loop_result* r = &rx->loop_res[i];
        lc_resolve_snr(&rx->ch[i], &rx->aux[i]);
        rx->device_ms += r->done_ms;
        uint32_t ms = ms0 + r->done_ms;
        if (r->stop == LC_STOP_DLL_NAN && r->done_ms < n_ms) {
            gpsb_host_set_packet_cnt(ms);
            rx->aux[i].last_nav_bit = -1;
            const uint8_t index = rx_index(rx, i, ms);
            hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], index, r->iq);
            rx_slot_end(rx, i, ms, index);
            if (nav_log) nav_log[(size_t)(ms - ms0) * n_ch + i] = rx->aux[i].last_nav_bit;
            rx->host_ms++;
            ms++;
        }
        rx->plan[i].stage = (int)(ms - ms0);          /* scratch: first row of channel i the loop did not deliver */
    }
    /* rows the loop did not deliver are blank until somebody fills them: one pass over the logs, row by row */
    uint32_t first_blank = n_ms;
    for (uint32_t i = 0; i < n_ch; i++)
        if ((uint32_t)rx->plan[i].stage < first_blank) first_blank = (uint32_t)rx->plan[i].stage;
    for (uint32_t k = first_blank; k < n_ms; k++)
        for (uint32_t i = 0; i < n_ch; i++)
            if (k >= (uint32_t)rx->plan[i].stage) {
                if (iq_log) memset(iq_log + ((size_t)k * n_ch + i) * 6, 0, 12);
                if (nav_log) nav_log[(size_t)k * n_ch + i] = -1;
            }
    for (uint32_t i = 0; i < n_ch; i++) {
        const uint32_t ms = ms0 + (uint32_t)rx->plan[i].stage;
        if (ms < ms0 + n_ms) {
            int rc = run_channel_span(rx, i, ms0, ms, ms0 + n_ms, iq_log, nav_log);
            if (rc != GPSB_OK) return rc;
        }
    }
    return GPSB_OK;
}

/* The whole run with the loops on the device: one k_track_run launch for every channel that is tracking, then
 * whatever is left (channels handed back early, channels still in pre-track) channel by channel. */
static int in_pre_track(const gps_ch_t* ch)
{
    return ch->tracking_data.state == GPS_NEED_PRE_TRACK || ch->tracking_data.state == GPS_PRE_TRACK_RUN;
}

static int track_run_device(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log)
{
    const uint32_t n_ch = rx->n_ch;
    /* Channels still in pre-track (tracking.c:398-450: 7 correlations per millisecond for >= 84 ms) go into the same
     * call: gpsb_track_loop runs k_pretrack_run ahead of k_track_run on the same stream and every channel enters the
     * tracking loop at the millisecond behind its own pre-track - one host round trip for the whole span. */
    uint32_t n_pre = 0;
    for (uint32_t i = 0; i < n_ch; i++) n_pre += in_pre_track(&rx->ch[i]) ? 1u : 0u;
    uint32_t n_trk = 0;
    for (uint32_t i = 0; i < n_ch; i++) n_trk += is_tracking(&rx->ch[i]) ? 1u : 0u;
    if (iq_log && n_trk != n_ch) memset(iq_log, 0, (size_t)n_ms * n_ch * 12);
    if (nav_log && n_trk != n_ch) memset(nav_log, 0xFF, (size_t)n_ms * n_ch);
    if (n_trk + n_pre > 0) {
        /* one launch for every channel; a channel that is neither tracking nor in pre-track is refused by the loop
         * (LC_STOP_STATE, nothing done): idle channels have nothing to do, anything else is taken channel by channel */
        int rc = gpsb_track_loop(rx->ctx, n_ch, rx->ch, (uint32_t)sizeof(gps_ch_t), rx->aux, (uint32_t)sizeof(gpsb_aux), ms0,
                                 n_ms, iq_log, nav_log, rx->loop_res);
        if (rc != GPSB_OK) return hx_note(rc);
        rc = finish_device_run(rx, ms0, n_ms, iq_log, nav_log);
        if (rc != GPSB_OK) return rc;
    } else {
        for (uint32_t i = 0; i < n_ch; i++) {
            int rc = run_channel_span(rx, i, ms0, ms0, ms0 + n_ms, iq_log, nav_log);
            if (rc != GPSB_OK) return rc;
        }
    }
    gpsb_host_set_packet_cnt(ms0 + n_ms - 1);
    return GPSB_OK;
}

/* Streaming form of gpsb_rx_track_run: the samples are still in HOST memory.  The first chunk is uploaded, the
 * device-resident loop is launched for the whole run, and the remaining chunks are DMA-ed into the ring while the
 * loop is already tracking (include/gpsb.h, "streaming ingest"): the upload disappears behind the run instead of
 * preceding it.  Falls back to upload-then-run when some channel is not tracking yet or the loops are on the host. */
static int track_stream_impl(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed, uint32_t chunk_ms,
                             int16_t* iq_log, int8_t* nav_log, int iq2);

int gpsb_rx_track_stream(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed, uint32_t chunk_ms,
                         int16_t* iq_log, int8_t* nav_log)
{
    return track_stream_impl(rx, ms0, n_ms, packed, chunk_ms, iq_log, nav_log, 0);
}

/* The same run fed with the MAX2769-native 2-bit I / 2-bit Q container (one byte per sample, n_ms * 16368 bytes, bit 0 =
 * I sign): every chunk is copied and packed to the 1-bit ring format on the copy stream (k_pack_iq2) behind the loop. */
int gpsb_rx_track_stream_iq2(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* samples, uint32_t chunk_ms,
                             int16_t* iq_log, int8_t* nav_log)
{
    return track_stream_impl(rx, ms0, n_ms, samples, chunk_ms, iq_log, nav_log, 1);
}

#define RX_STREAM_PATIENCE_MS 100u

static int push_chunk(gpsb_rx* rx, uint32_t ms, uint32_t n, const uint8_t* base, uint32_t at, int iq2)
{
    if (iq2) return gpsb_stream_push_iq2(rx->ctx, ms, n, base + (size_t)at * GPSB_MS_SAMPLES);
    return gpsb_stream_push(rx->ctx, ms, n, base + (size_t)at * GPSB_MS_BYTES);
}
static int upload_span(gpsb_rx* rx, uint32_t ms, uint32_t n, const uint8_t* base, uint32_t at, int iq2)
{
    if (iq2) return gpsb_upload_signal_iq2(rx->ctx, ms, n, base + (size_t)at * GPSB_MS_SAMPLES);
    return gpsb_upload_signal(rx->ctx, ms, n, base + (size_t)at * GPSB_MS_BYTES);
}

static int track_stream_impl(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed, uint32_t chunk_ms,
                             int16_t* iq_log, int8_t* nav_log, int iq2)
{
    if (!rx || !packed) return GPSB_ERR_ARG;
    if (n_ms == 0) return GPSB_OK;
    if (chunk_ms == 0) chunk_ms = 64;
    const uint32_t n_ch = rx->n_ch;
    const uint32_t ring_ms = rx->ring_ms;
    if (chunk_ms > ring_ms / 2) chunk_ms = ring_ms / 2 ? ring_ms / 2 : 1;   /* the producer must be able to run ahead */
    uint32_t n_trk = 0;
    for (uint32_t i = 0; i < n_ch; i++) n_trk += is_tracking(&rx->ch[i]) ? 1u : 0u;
    if (n_trk != n_ch || n_ch > 256 || rx->loop_site == GPSB_LOOP_HOST || gpsb_session_slots(rx->ctx) != 0 ||
        (n_ms > ring_ms && ring_ms < 192)) {        /* progress is reported every 64 ms: a smaller ring cannot be refilled behind the loop */
        for (uint32_t at = 0; at < n_ms;) {         /* upload a ring-full, run it, repeat */
            const uint32_t n = n_ms - at < ring_ms ? n_ms - at : ring_ms;
            int rc = upload_span(rx, ms0 + at, n, packed, at, iq2);
            if (rc != GPSB_OK) return hx_note(rc);
            rc = gpsb_rx_track_run(rx, ms0 + at, n, iq_log ? iq_log + (size_t)at * n_ch * 6 : NULL,
                                   nav_log ? nav_log + (size_t)at * n_ch : NULL);
            if (rc != GPSB_OK) return rc;
            at += n;
        }
        return GPSB_OK;
    }
    int rc = gpsb_stream_reset(rx->ctx, ms0);
    /* The samples are in host memory: this call is its own producer and pushes without delay, so a loop that sees no
     * frame for RX_STREAM_PATIENCE_MS is not waiting for a slow source - the copy engine is not running beside the kernel
     * at all (a profiler that makes launches synchronous, CUDA_LAUNCH_BLOCKING).  The loop then ends with its
     * milliseconds complete and the run is finished upload-then-run below, in ~0.1 s instead of the 2-s default. */
    const uint32_t patience_before = gpsb_stream_timeout_ms(rx->ctx);
    if (patience_before > RX_STREAM_PATIENCE_MS) gpsb_stream_set_timeout_ms(rx->ctx, RX_STREAM_PATIENCE_MS);
    /* a short first chunk, so that the loop starts at once; the chunks behind it are as long as asked for */
    const uint32_t first_ms = chunk_ms < 16 ? chunk_ms : 16;
    uint32_t sent = n_ms < first_ms ? n_ms : first_ms;
    if (rc == GPSB_OK) rc = push_chunk(rx, ms0, sent, packed, 0, iq2);
    if (rc != GPSB_OK) { gpsb_stream_set_timeout_ms(rx->ctx, patience_before); return hx_note(rc); }
    rc = gpsb_track_loop_begin(rx->ctx, n_ch, rx->ch, (uint32_t)sizeof(gps_ch_t), rx->aux, (uint32_t)sizeof(gpsb_aux), ms0, n_ms,
                               iq_log, nav_log, rx->loop_res, GPSB_LOOP_STREAMING);
    if (rc != GPSB_OK) { gpsb_stream_wait(rx->ctx); gpsb_stream_set_timeout_ms(rx->ctx, patience_before); return hx_note(rc); }
    int rc_push = GPSB_OK;
    int checked_overlap = 0, no_overlap = 0;
    while (sent < n_ms && rc_push == GPSB_OK && !no_overlap) {
        const uint32_t n = n_ms - sent < chunk_ms ? n_ms - sent : chunk_ms;
        /* a run longer than the ring: chunk [sent, sent+n) replaces frames sent-ring .. - wait until every channel is past them */
        while (sent + n > ring_ms && (int32_t)(gpsb_stream_progress(rx->ctx, n_ch) - (ms0 + sent + n - ring_ms)) < 0 &&
               gpsb_stream_loop_running(rx->ctx))
            sched_yield();
        rc_push = push_chunk(rx, ms0 + sent, n, packed, sent, iq2);
        sent += n;
        if (!checked_overlap && rc_push == GPSB_OK) {
            /* Streaming needs the copy engine to run WHILE the loop kernel does.  Where something serialises the two (a
             * profiler replaying kernels, CUDA_LAUNCH_BLOCKING) the first chunk pushed behind the launch never lands,
             * the loop would wait out its starvation time-out (2 s by default) and the call would take seconds instead of
             * a millisecond.  Looked at once: if that chunk is still queued after 30 ms and the loop is still running,
             * the loop is told to stop waiting (it ends with its milliseconds complete) and the rest of the run is done
             * upload-then-run, below. */
            checked_overlap = 1;
            struct timespec t0, t1;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            while (gpsb_stream_copies_pending(rx->ctx) && gpsb_stream_loop_running(rx->ctx)) {
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if ((t1.tv_sec - t0.tv_sec) * 1000000000L + (t1.tv_nsec - t0.tv_nsec) > 30000000L) {
                    no_overlap = 1;
                    gpsb_stream_abort(rx->ctx);
                    break;
                }
                sched_yield();
            }
        }
    }
    rc = gpsb_track_loop_end(rx->ctx);        /* a failed push starves the loop, which then ends by its time-out */
    int rc_wait = gpsb_stream_wait(rx->ctx);
    gpsb_stream_set_timeout_ms(rx->ctx, patience_before);
    if (rc == GPSB_OK && rc_push == GPSB_OK && rx->loop_res[0].stop == LC_STOP_STARVED) no_overlap = 1;   /* starved by nobody but us */
    if (rc_push != GPSB_OK) return hx_note(rc_push);
    if (rc != GPSB_OK) return hx_note(rc);
    if (rc_wait != GPSB_OK) return hx_note(rc_wait);
    if (no_overlap) {
        /* copies and the loop kernel do not overlap here: every channel stopped where the frames ended (same millisecond,
         * same watermark); the rest of the run goes upload-then-run, all channels per launch */
        uint32_t at = rx->loop_res[0].done_ms;
        int together = 1;
        for (uint32_t i = 0; i < n_ch; i++)
            together &= rx->loop_res[i].done_ms == at && rx->loop_res[i].stop == LC_STOP_STARVED;
        if (together && at < n_ms) {
            for (uint32_t i = 0; i < n_ch; i++) {
                rx->device_ms += at;
                lc_resolve_snr(&rx->ch[i], &rx->aux[i]);
            }
            while (at < n_ms) {
                const uint32_t n = n_ms - at < ring_ms ? n_ms - at : ring_ms;
                rc = upload_span(rx, ms0 + at, n, packed, at, iq2);
                if (rc != GPSB_OK) return hx_note(rc);
                rc = track_run_device(rx, ms0 + at, n, iq_log ? iq_log + (size_t)at * n_ch * 6 : NULL,
                                      nav_log ? nav_log + (size_t)at * n_ch : NULL);
                if (rc != GPSB_OK) return rc;
                at += n;
            }
            gpsb_host_set_packet_cnt(ms0 + n_ms - 1);
            return GPSB_OK;
        }
    }
    if (n_ms <= ring_ms && !no_overlap) {
        rc = finish_device_run(rx, ms0, n_ms, iq_log, nav_log);   /* the ring now holds the whole run */
        if (rc != GPSB_OK) return rc;
    } else {
        /* The ring holds only the tail of the run.  A channel the loop handed back early (degenerate DLL millisecond,
         * starved producer) is finished from the host copy, one ring-full at a time. */
        for (uint32_t i = 0; i < n_ch; i++) {
            rx->device_ms += rx->loop_res[i].done_ms;
            lc_resolve_snr(&rx->ch[i], &rx->aux[i]);
        }
        for (uint32_t i = 0; i < n_ch; i++) {
            gpsb_loop_result* r = &rx->loop_res[i];
            if (r->done_ms >= n_ms) continue;
            uint32_t at = r->done_ms;                              /* first millisecond still to do for this channel */
            while (at < n_ms) {
                const uint32_t n = n_ms - at < ring_ms ? n_ms - at : ring_ms;
                rc = upload_span(rx, ms0 + at, n, packed, at, iq2);
                if (rc != GPSB_OK) return hx_note(rc);
                if (at == r->done_ms && r->stop == LC_STOP_DLL_NAN) {
                    gpsb_host_set_packet_cnt(ms0 + at);
                    rx->aux[i].last_nav_bit = -1;
                    const uint8_t index = rx_index(rx, i, ms0 + at);
                    hx_trk_finish_epl(&rx->ch[i], &rx->aux[i], index, r->iq);
                    rx_slot_end(rx, i, ms0 + at, index);
                    if (nav_log) nav_log[(size_t)at * n_ch + i] = rx->aux[i].last_nav_bit;
                    rx->host_ms++;
                    if (n == 1) { at++; continue; }
                    rc = run_channel_span(rx, i, ms0, ms0 + at + 1, ms0 + at + n, iq_log, nav_log);
                } else {
                    rc = run_channel_span(rx, i, ms0, ms0 + at, ms0 + at + n, iq_log, nav_log);
                }
                if (rc != GPSB_OK) return rc;
                at += n;
            }
        }
    }
    gpsb_host_set_packet_cnt(ms0 + n_ms - 1);
    return GPSB_OK;
}

int gpsb_rx_track_run(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log)
{
    if (!rx) return GPSB_ERR_ARG;
    if (n_ms == 0) return GPSB_OK;
    /* loops on the device: one launch for the whole run, no per-millisecond round trip */
    if (rx->loop_site != GPSB_LOOP_HOST && (rx->loop_site == GPSB_LOOP_DEVICE || n_ms > 1) &&
        gpsb_session_slots(rx->ctx) == 0)
        return track_run_device(rx, ms0, n_ms, iq_log, nav_log);
    if (!rx) return GPSB_ERR_ARG;
    /* a run of many milliseconds is served by resident CTAs (one per channel): no launch per ms */
    const int own_session = n_ms > 8 && rx->n_ch <= 128 && gpsb_session_slots(rx->ctx) == 0 &&
                            gpsb_session_begin(rx->ctx, rx->n_ch) == GPSB_OK;
    const uint32_t n_workers = own_session ? pick_workers(rx) : 1;
    if (own_session && n_workers > 1) {
        pthread_t tid[16];
        track_job job[16];
        for (uint32_t w = 0; w < n_workers; w++) {
            job[w] = (track_job){rx, w, n_workers, ms0, n_ms, iq_log, nav_log, GPSB_OK};
            if (pthread_create(&tid[w], NULL, track_worker, &job[w]) != 0) {
                job[w].rc = GPSB_ERR_NOMEM;
                track_worker(&job[w]);               /* could not start a thread: do its share here */
                tid[w] = 0;
            }
        }
        int rc = GPSB_OK;
        for (uint32_t w = 0; w < n_workers; w++) {
            if (tid[w]) pthread_join(tid[w], NULL);
            if (job[w].rc != GPSB_OK) rc = job[w].rc;
        }
        int rc_end = gpsb_session_end(rx->ctx);
        gpsb_host_set_packet_cnt(ms0 + n_ms - 1);
        return hx_note(rc != GPSB_OK ? rc : rc_end);
    }
    for (uint32_t m = 0; m < n_ms; m++) {
        for (uint32_t i = 0; i < rx->n_ch; i++) rx->aux[i].last_nav_bit = -1;
        int rc = gpsb_rx_track_ms(rx, ms0 + m);
        if (rc != GPSB_OK) {
            if (own_session) gpsb_session_end(rx->ctx);
            return rc;
        }
        if (iq_log) {
            int16_t* row = iq_log + (size_t)m * rx->n_ch * 6;
            memset(row, 0, (size_t)rx->n_ch * 12);
            uint32_t k = 0;
            for (uint32_t i = 0; i < rx->n_ch; i++)
                if (rx->plan[i].want == GPSB_WANT_EPL) memcpy(row + 6u * i, rx->epl_out + 6u * k++, 12);
        }
        if (nav_log)
            for (uint32_t i = 0; i < rx->n_ch; i++)
                nav_log[(size_t)m * rx->n_ch + i] = rx->plan[i].want == GPSB_WANT_EPL ? rx->aux[i].last_nav_bit : -1;
    }
    if (own_session) return hx_note(gpsb_session_end(rx->ctx));
    return GPSB_OK;
}

int gpsb_rx_acquire_ms(gpsb_rx* rx, uint32_t ms)
{
    if (!rx) return GPSB_ERR_ARG;
    gpsb_host_set_packet_cnt(ms);
    for (uint32_t i = 0; i < rx->n_ch; i++) hx_acq_plan(&rx->ch[i], &rx->aux[i], ms, &rx->plan[i]);
    uint32_t n_epl, n_s;
    int rc = run_plans(rx, &n_epl, &n_s);
    if (rc != GPSB_OK) return rc;
    uint32_t k = 0;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->plan[i].want != GPSB_WANT_SEARCH) continue;
        const gpsb_search_res* r = &k_empty_window;
        if (k < n_s && rx->s_owner[k] == i) r = &rx->s_res[k++];
        hx_acq_finish(&rx->ch[i], &rx->aux[i], &rx->plan[i], r);
    }
    return GPSB_OK;
}

/* Start the Doppler search of one channel on its private vote buffers (acquisition.c:68-87). */
static void start_doppler_search(gps_ch_t* ch, gpsb_aux* aux)
{
    gps_acq_t* a = &ch->acq_data;
    if (a->state != GPS_ACQ_NEED_FREQ_SEARCH) return;
    if (a->given_freq_offset_hz != 0) {
        a->found_freq_offset_hz = a->given_freq_offset_hz;
        a->state = GPS_ACQ_FREQ_SEARCH_DONE;
        return;
    }
    memset(aux->freq_hist, 0, sizeof aux->freq_hist);
    memset(aux->bin_phases, 0, sizeof aux->bin_phases);
    aux->bin_count = 0;
    a->freq_index = 0;
    a->state = GPS_ACQ_FREQ_SEARCH_RUN;
}

int gpsb_rx_cold_sweep(gpsb_rx* rx, int32_t first_bin_hz, int32_t bin_step_hz, uint32_t n_bins, uint32_t ms0,
                       uint32_t n_ms, uint8_t* votes, uint16_t* phases)
{
    if (!rx || n_bins == 0 || n_bins > GPSB_MAX_BINS || n_ms == 0 || n_ms > GPSB_FREQ_POINTS_MAX - 1)
        return hx_note(GPSB_ERR_ARG);
    uint32_t* slots = (uint32_t*)malloc(sizeof(uint32_t) * rx->n_ch);
    uint32_t* who = (uint32_t*)malloc(sizeof(uint32_t) * rx->n_ch);
    uint32_t* step32 = (uint32_t*)malloc(sizeof(uint32_t) * n_bins);
    if (!slots || !who || !step32) {
        free(slots); free(who); free(step32);
        return hx_note(GPSB_ERR_NOMEM);
    }
    uint32_t n_sv = 0;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->ch[i].prn < 1) continue;
        start_doppler_search(&rx->ch[i], &rx->aux[i]);
        if (rx->ch[i].acq_data.state != GPS_ACQ_FREQ_SEARCH_RUN) continue;
        slots[n_sv] = rx->ch[i].prn;
        who[n_sv++] = i;
    }
    for (uint32_t b = 0; b < n_bins; b++)        /* int -> float at the call site, acquisition.c:288 */
        step32[b] = hx_nco_step32((float)(IF_FREQ_HZ + (int16_t)(first_bin_hz + (int32_t)b * bin_step_hz)));
    int rc = GPSB_OK;
    gpsb_search_res* cells = NULL;
    if (n_sv) {
        cells = (gpsb_search_res*)malloc(sizeof(gpsb_search_res) * (size_t)n_sv * n_bins * n_ms);
        /* under a communicator (gpsb_comm_init) the cell groups are sharded over the ranks and all-gathered: every rank
         * runs the same votes on the same grid */
        if (!cells) rc = GPSB_ERR_NOMEM;
        else if (gpsb_comm_size(rx->ctx) > 1) rc = gpsb_sweep_gather(rx->ctx, slots, n_sv, step32, n_bins, ms0, n_ms, 0, cells);
        else rc = gpsb_sweep(rx->ctx, slots, n_sv, step32, n_bins, ms0, n_ms, 0, cells);
    }
    if (rc == GPSB_OK) {
        gpsb_host_set_packet_cnt(ms0 + n_ms - 1);
        for (uint32_t s = 0; s < n_sv; s++) {
            gps_ch_t* ch = &rx->ch[who[s]];
            gpsb_aux* aux = &rx->aux[who[s]];
            for (uint32_t b = 0; b < n_bins; b++) {
                uint16_t grp[GPSB_FREQ_POINTS_MAX];
                for (uint32_t m = 0; m < n_ms; m++) grp[m] = cells[((size_t)s * n_bins + b) * n_ms + m].phase;
                uint16_t at = 0;
                uint8_t chain = hx_chain_vote(grp, (uint8_t)n_ms, &at);
                if (votes) votes[(size_t)who[s] * n_bins + b] = chain;
                if (phases) phases[(size_t)who[s] * n_bins + b] = at;
                if (ch->acq_data.state != GPS_ACQ_FREQ_SEARCH_RUN) continue;   /* decided at an earlier bin */
                if (chain >= 2) aux->freq_hist[b] += chain;
                ch->acq_data.freq_index = (uint8_t)b;
                hx_freq_hist_decide(ch, aux->freq_hist, n_bins, first_bin_hz, bin_step_hz);
                /* The reference wipes its vote buffers - the Doppler histogram included - after EVERY bin
                 * (acquisition_buffers_reset, acquisition.c:60-65, called at :303): the "histogram" only ever holds the
                 * bin under test, so a satellite is accepted by the first bin whose ten snapshots agree in a chain of
                 * three.  Reproduced as is. */
                memset(aux->freq_hist, 0, sizeof aux->freq_hist);
                ch->acq_data.freq_index = (uint8_t)(b + 1u >= n_bins ? 0u : b + 1u);   /* next bin, acquisition.c:305-310 */
            }
        }
    }
    free(cells); free(slots); free(who); free(step32);
    return hx_note(rc);
}

/* ---------------------------------------------------------------------------- cold start, N satellites */
/* Code-phase rounds of every channel over the snapshots [ms, ms + n): within a round a channel's window and carrier
 * are fixed, so the cells of the coming snapshots are independent until each vote - they are computed AHEAD in one
 * launch and consumed in order.  A channel whose next cell differs from the one computed ahead (its round ended: new
 * window) has its cells for the REST of the span computed in a second, small launch - the others keep theirs - and the
 * span goes on.  busy_mask: bit s set = a channel in acquisition state s still has work to do; the run ends after the
 * first snapshot that leaves no channel in such a state.  *consumed = snapshots processed. */
static int same_cell(const gpsb_search_req* a, const gpsb_search_req* b)
{
    return a->step32 == b->step32 && a->start == b->start && a->stop == b->stop && a->sv_slot == b->sv_slot &&
           a->off_bits == b->off_bits;
}

/* cells of `cnt` channels (who[]) for the snapshots k0 .. n-1 of the span, one launch; results into ahead[channel][snapshot] */
static int ahead_launch(gpsb_rx* rx, const gpsb_search_req* base, const uint32_t* who, uint32_t cnt, uint32_t ms, uint32_t k0,
                        uint32_t n, gpsb_search_req* rq, gpsb_search_res* rs, gpsb_search_res* ahead, uint32_t* launches)
{
    const uint32_t span = n - k0;
    for (uint32_t j = 0; j < cnt; j++)
        for (uint32_t k = 0; k < span; k++) {
            rq[(size_t)k * cnt + j] = base[who[j]];
            rq[(size_t)k * cnt + j].ms_index = ms + k0 + k;
        }
    int rc = gpsb_search(rx->ctx, cnt * span, rq, rs);
    if (launches) (*launches)++;
    for (uint32_t j = 0; j < cnt && rc == GPSB_OK; j++)
        for (uint32_t k = 0; k < span; k++) ahead[(size_t)who[j] * n + k0 + k] = rs[(size_t)k * cnt + j];
    return rc;
}

static int acquire_ahead(gpsb_rx* rx, uint32_t ms, uint32_t n, uint32_t busy_mask, uint32_t* consumed, uint32_t* launches)
{
    const uint32_t n_ch = rx->n_ch;
    *consumed = 0;
    if (n == 0) return GPSB_OK;
    gpsb_search_req* base = (gpsb_search_req*)calloc(n_ch, sizeof *base);     /* the cell computed ahead for channel i */
    uint8_t* have = (uint8_t*)calloc(n_ch, 1);
    uint32_t* who = (uint32_t*)calloc(n_ch, sizeof *who);
    gpsb_search_res* ahead = (gpsb_search_res*)malloc(sizeof *ahead * (size_t)n_ch * n);    /* [channel][snapshot] */
    gpsb_search_req* rq = (gpsb_search_req*)malloc(sizeof *rq * (size_t)n_ch * n);
    gpsb_search_res* rs = (gpsb_search_res*)malloc(sizeof *rs * (size_t)n_ch * n);
    int rc = (base && have && who && ahead && rq && rs) ? GPSB_OK : GPSB_ERR_NOMEM;
    /* what every channel would correlate at snapshot `ms`, looked at without touching the channel */
    uint32_t n_want = 0;
    for (uint32_t i = 0; i < n_ch && rc == GPSB_OK; i++) {
        const gps_acq_t* a = &rx->ch[i].acq_data;
        if (rx->ch[i].prn < 1) continue;
        if (a->state != GPS_ACQ_CODE_PHASE_SEARCH1 && a->state != GPS_ACQ_CODE_PHASE_SEARCH2 &&
            a->state != GPS_ACQ_CODE_PHASE_SEARCH3)
            continue;
        gps_ch_t probe = rx->ch[i];
        gpsb_aux probe_aux = rx->aux[i];
        gpsb_plan p;
        hx_acq_plan(&probe, &probe_aux, ms, &p);
        if (p.want != GPSB_WANT_SEARCH || p.search.start >= p.search.stop) continue;
        base[i] = p.search;
        have[i] = 1;
        who[n_want++] = i;
    }
    if (rc == GPSB_OK && n_want) rc = ahead_launch(rx, base, who, n_want, ms, 0u, n, rq, rs, ahead, launches);
    for (uint32_t k = 0; k < n && rc == GPSB_OK; k++) {
        gpsb_host_set_packet_cnt(ms + k);
        uint32_t n_miss = 0;
        for (uint32_t i = 0; i < n_ch; i++) {
            gpsb_plan* p = &rx->plan[i];
            p->want = GPSB_WANT_NOTHING;
            /* only channels in the code rounds are served: one whose Doppler vote did not pass would go on with its
             * Doppler search (acquisition.c:141-146), which never ends on a satellite that is not there */
            if (rx->ch[i].acq_data.state < GPS_ACQ_CODE_PHASE_SEARCH1) continue;
            hx_acq_plan(&rx->ch[i], &rx->aux[i], ms + k, p);
            if (p->want != GPSB_WANT_SEARCH) continue;
            if (p->search.start >= p->search.stop) hx_acq_finish(&rx->ch[i], &rx->aux[i], p, &k_empty_window);
            else if (have[i] && same_cell(&p->search, &base[i])) hx_acq_finish(&rx->ch[i], &rx->aux[i], p, &ahead[(size_t)i * n + k]);
            else who[n_miss++] = i;                              /* not computed ahead: a new window from here on */
        }
        if (n_miss) {
            for (uint32_t j = 0; j < n_miss; j++) {
                base[who[j]] = rx->plan[who[j]].search;
                have[who[j]] = 1;
            }
            rc = ahead_launch(rx, base, who, n_miss, ms, k, n, rq, rs, ahead, launches);
            for (uint32_t j = 0; j < n_miss && rc == GPSB_OK; j++) {
                const uint32_t i = who[j];
                hx_acq_finish(&rx->ch[i], &rx->aux[i], &rx->plan[i], &ahead[(size_t)i * n + k]);
            }
        }
        *consumed = k + 1;
        int busy = 0;
        for (uint32_t i = 0; i < n_ch; i++)
            if (rx->ch[i].prn >= 1 && ((busy_mask >> (uint32_t)rx->ch[i].acq_data.state) & 1u)) busy = 1;
        if (!busy) break;
    }
    free(base); free(have); free(who); free(ahead); free(rq); free(rs);
    return hx_note(rc);
}

/* The code-phase rounds of every served channel on the device: one launch for the channels that keep the round alive
 * (each until its state has left busy_mask, at most n snapshots), a second one - only if there are any - for served
 * channels that do not (they see exactly the snapshots the round lasted).  Same outcome as acquire_ahead() window by
 * window: channels never share vote buffers in the code rounds, so each one's snapshots are its own affair.
 * *consumed = snapshots the round lasted. */
static int rounds_on_device(gpsb_rx* rx, uint32_t ms, uint32_t n, uint32_t busy_mask, uint32_t* consumed, uint32_t* launches)
{
    const uint32_t n_ch = rx->n_ch;
    *consumed = 0;
    if (n == 0) return GPSB_OK;
    uint8_t* mode = (uint8_t*)calloc(n_ch, 1);
    uint32_t* used = (uint32_t*)calloc(n_ch, sizeof *used);
    if (!mode || !used) { free(mode); free(used); return hx_note(GPSB_ERR_NOMEM); }
    uint32_t n_busy = 0, n_idle = 0;
    for (uint32_t i = 0; i < n_ch; i++) {
        const uint32_t st = (uint32_t)rx->ch[i].acq_data.state;
        if (rx->ch[i].prn < 1 || st < GPS_ACQ_CODE_PHASE_SEARCH1 || st == GPS_ACQ_DONE) continue;
        if ((busy_mask >> st) & 1u) { mode[i] = 1; n_busy++; }
        else if (st != GPS_ACQ_CODE_PHASE_SEARCH2_DONE) n_idle++;       /* SEARCH2_DONE waits for round 3: nothing happens to it */
    }
    int rc = GPSB_OK;
    uint32_t lasted = 0;
    if (n_busy) {
        rc = gpsb_code_rounds(rx->ctx, n_ch, rx->ch, (uint32_t)sizeof(gps_ch_t), rx->aux, (uint32_t)sizeof(gpsb_aux), ms, n,
                              busy_mask, mode, used);
        if (launches) (*launches)++;
        for (uint32_t i = 0; i < n_ch; i++) lasted = used[i] > lasted ? used[i] : lasted;
    }
    if (rc == GPSB_OK && n_idle && lasted) {
        for (uint32_t i = 0; i < n_ch; i++) {
            const uint32_t st = (uint32_t)rx->ch[i].acq_data.state;
            const int served = rx->ch[i].prn >= 1 && st >= GPS_ACQ_CODE_PHASE_SEARCH1 && st != GPS_ACQ_DONE &&
                               st != GPS_ACQ_CODE_PHASE_SEARCH2_DONE;
            mode[i] = (uint8_t)((served && !mode[i]) ? 2 : 0);          /* mode[i] was 1: that channel has had its snapshots */
        }
        rc = gpsb_code_rounds(rx->ctx, n_ch, rx->ch, (uint32_t)sizeof(gps_ch_t), rx->aux, (uint32_t)sizeof(gpsb_aux), ms, lasted,
                              busy_mask, mode, used);
        if (launches) (*launches)++;
    }
    if (lasted) gpsb_host_set_packet_cnt(ms + lasted - 1);
    *consumed = lasted;
    free(mode); free(used);
    return hx_note(rc);
}

int gpsb_rx_cold_start(gpsb_rx* rx, uint32_t ms0, const gpsb_cold_start_opts* opts, gpsb_cold_start_report* rep)
{
    if (!rx) return hx_note(GPSB_ERR_ARG);
    gpsb_cold_start_opts o;
    memset(&o, 0, sizeof o);
    if (opts) o = *opts;
    if (o.n_bins == 0) {                       /* the reference's own grid, config.h:41-44 */
        o.first_bin_hz = -ACQ_SEARCH_FREQ_HZ;
        o.bin_step_hz = ACQ_SEARCH_STEP_HZ;
        o.n_bins = ACQ_COUNT;
    }
    if (o.sweep_ms == 0) o.sweep_ms = 10;      /* ACQ_SINGLE_FREQ_LENGTH, acquisition.c:18 */
    if (o.sweeps == 0) o.sweeps = 1;
    if (o.round_timeout_ms == 0) o.round_timeout_ms = 400;
    if (o.window_ms == 0) o.window_ms = 16;
    gpsb_cold_start_report r;
    memset(&r, 0, sizeof r);
    r.ms_sweep0 = ms0;
    const uint64_t launches0 = gpsb_launch_count(rx->ctx);
    for (uint32_t i = 0; i < rx->n_ch; i++)
        if (rx->ch[i].prn >= 1 && rx->ch[i].acq_data.state == GPS_ACQ_NEED_FREQ_SEARCH) r.n_searched++;
    /* 1. Doppler: every (channel, bin, snapshot) cell in one launch, the reference's votes bin by bin */
    /*    A satellite the first sweep leaves undecided gets further sweeps, each on the next sweep_ms snapshots, as the
     *    reference's search does when it wraps round (acquisition.c:305-310). */
    int rc = GPSB_OK;
    uint32_t t = ms0;
    for (uint32_t k = 0; k < o.sweeps; k++) {
        uint32_t undecided = 0;
        for (uint32_t i = 0; i < rx->n_ch; i++)
            if (rx->ch[i].prn >= 1 && (rx->ch[i].acq_data.state == GPS_ACQ_NEED_FREQ_SEARCH ||
                                       rx->ch[i].acq_data.state == GPS_ACQ_FREQ_SEARCH_RUN)) undecided++;
        if (!undecided) break;
        rc = gpsb_rx_cold_sweep(rx, o.first_bin_hz, o.bin_step_hz, o.n_bins, t, o.sweep_ms, NULL, NULL);
        if (rc != GPSB_OK) return rc;
        t += o.sweep_ms;
        r.n_sweeps++;
    }
    /* 2. code phase, rounds 1 and 2 (acquisition.c:89-104, 150-170): all channels with a Doppler side by side */
    r.ms_code0 = t;
    gpsb_host_set_packet_cnt(t);
    const uint32_t serve_world = o.serve_world ? o.serve_world : 1u;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->ch[i].prn < 1 || rx->ch[i].acq_data.state != GPS_ACQ_FREQ_SEARCH_DONE) continue;
        r.n_doppler_found++;
        if (i % serve_world != o.serve_rank % serve_world) continue;        /* another rank's satellite from here on */
        r.n_served++;
        acquisition_start_code_search_channel(&rx->ch[i]);
    }
    const uint32_t busy12 = (1u << GPS_ACQ_CODE_PHASE_SEARCH1) | (1u << GPS_ACQ_CODE_PHASE_SEARCH1_DONE) |
                            (1u << GPS_ACQ_CODE_PHASE_SEARCH2);
    uint32_t launches = 0;
    /* Look-ahead windows by default: a code round is throughput work (a 2046-phase window is 68 us on one SM), and the
     * cells of many coming snapshots run side by side on all SMs; a window consumed whole is followed by one twice as
     * long, so a channel that never settles costs a handful of launches for its whole time-out.  On request the rounds
     * run device-resident instead (k_code_rounds_run: same records, same schedule, one launch per round). */
    const int on_device = o.code_rounds == 1 && gpsb_session_slots(rx->ctx) == 0;
    const uint32_t window_max = o.window_max_ms ? o.window_max_ms : o.round_timeout_ms;
    uint32_t window = o.window_ms;
    r.ms_code12_last = t ? t - 1 : 0;
    while (r.n_served && t - r.ms_code0 < o.round_timeout_ms) {
        uint32_t left = o.round_timeout_ms - (t - r.ms_code0), done = 0;
        const uint32_t ahead = left < window ? left : window;
        rc = on_device ? rounds_on_device(rx, t, left, busy12, &done, &launches)
                       : acquire_ahead(rx, t, ahead, busy12, &done, &launches);
        if (done == ahead && window < window_max) window = 2 * window < window_max ? 2 * window : window_max;
        if (rc != GPSB_OK) return rc;
        t += done;
        r.ms_code12_last = t - 1;
        int busy = 0;
        for (uint32_t i = 0; i < rx->n_ch; i++)
            if (rx->ch[i].prn >= 1 && ((busy12 >> (uint32_t)rx->ch[i].acq_data.state) & 1u)) busy = 1;
        if (!busy || done == 0) break;
    }
    /* 3. round 3 for every channel that finished round 2, started together (gps_master.c:113-118) */
    r.ms_code3_first = t;
    gpsb_host_set_packet_cnt(t);
    uint32_t n_round3 = 0;
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->ch[i].prn < 1 || rx->ch[i].acq_data.state != GPS_ACQ_CODE_PHASE_SEARCH2_DONE) continue;
        hx_acq_start_code_search3(&rx->ch[i], &rx->aux[i]);
        n_round3++;
    }
    const uint32_t busy3 = (1u << GPS_ACQ_CODE_PHASE_SEARCH3) | (1u << GPS_ACQ_CODE_PHASE_SEARCH3_DONE);
    window = o.window_ms;
    r.ms_last = t ? t - 1 : 0;
    while (n_round3 && t - r.ms_code3_first < o.round_timeout_ms) {
        uint32_t left = o.round_timeout_ms - (t - r.ms_code3_first), done = 0;
        const uint32_t ahead = left < window ? left : window;
        rc = on_device ? rounds_on_device(rx, t, left, busy3, &done, &launches)
                       : acquire_ahead(rx, t, ahead, busy3, &done, &launches);
        if (done == ahead && window < window_max) window = 2 * window < window_max ? 2 * window : window_max;
        if (rc != GPSB_OK) return rc;
        t += done;
        r.ms_last = t - 1;
        int busy = 0;
        for (uint32_t i = 0; i < rx->n_ch; i++)
            if (rx->ch[i].prn >= 1 && ((busy3 >> (uint32_t)rx->ch[i].acq_data.state) & 1u)) busy = 1;
        if (!busy || done == 0) break;
    }
    /* 4. acquired channels go on to tracking (gps_master.c:121-129) */
    for (uint32_t i = 0; i < rx->n_ch; i++) {
        if (rx->ch[i].prn < 1 || rx->ch[i].acq_data.state != GPS_ACQ_DONE) continue;
        r.n_acquired++;
        if (rx->ch[i].tracking_data.state == GPS_TRACKNG_IDLE) rx->ch[i].tracking_data.state = GPS_NEED_PRE_TRACK;
    }
    r.ms_next = t;
    r.launches = (uint32_t)(gpsb_launch_count(rx->ctx) - launches0);
    if (rep) *rep = r;
    return GPSB_OK;
}

/* ---------------------------------------------------------------------------- slot-phase walk */
int gpsb_rx_set_slot_walk(gpsb_rx* rx, int enable, uint32_t period_ms)
{
    if (!rx || period_ms > 65535u) return hx_note(GPSB_ERR_ARG);
    for (uint32_t i = 0; i < rx->n_ch; i++) gpsb_host_aux_walk(&rx->aux[i], enable ? 1u : 0u, period_ms);
    return GPSB_OK;
}

int gpsb_rx_channel_sync(const gpsb_rx* rx, uint32_t i, gpsb_sync_status* out)
{
    if (!rx || !out || i >= rx->n_ch) return hx_note(GPSB_ERR_ARG);
    const gps_nav_data_t* n = &rx->ch[i].nav_data;
    const gpsb_aux* a = &rx->aux[i];
    memset(out, 0, sizeof *out);
    out->tracking = rx->ch[i].tracking_data.state == GPS_TRACKING_RUN;
    out->bit_period_found = n->period_sync_ok_flag;
    out->bit_edge_refined = n->accurate_swap_ok;
    out->polarity_found = n->polarity_found;
    out->slot_phase = a->slot_phase;
    out->walk_enabled = a->walk_enable;
    out->walk_pending = a->skip_len != 0;
    out->walks = a->walks;
    out->subframes = n->subframe_cnt;
    out->words_ok = n->word_cnt_test;
    return GPSB_OK;
}

void gpsb_host_aux_walk(void* aux_record, uint32_t enable, uint32_t period_ms)
{
    gpsb_aux* a = (gpsb_aux*)aux_record;
    if (!a) return;
    a->walk_enable = (uint8_t)(enable != 0);
    a->walk_period_ms = (uint16_t)period_ms;
}

void gpsb_host_aux_walk_state(const void* aux_record, uint32_t out[6])
{
    const gpsb_aux* a = (const gpsb_aux*)aux_record;
    out[0] = a->slot_phase; out[1] = a->walk_enable; out[2] = a->skip_ms;
    out[3] = a->skip_len;   out[4] = a->walks;       out[5] = a->phase_since_ms;
}

/* ---------------------------------------------------------------------------- split-phase API */
int gpsb_host_plan_acq(gps_ch_t* ch, uint32_t frame_ms, gpsb_plan* plan)
{
    if (!ch || !plan) return GPSB_ERR_ARG;
    hx_acq_plan(ch, &g_shared_aux, frame_ms, plan);
    return GPSB_OK;
}

int gpsb_host_finish_acq(gps_ch_t* ch, const gpsb_plan* plan, const gpsb_search_res* res)
{
    if (!ch || !plan) return GPSB_ERR_ARG;
    hx_acq_finish(ch, &g_shared_aux, plan, res ? res : &k_empty_window);
    return GPSB_OK;
}

int gpsb_host_plan_track(gps_ch_t* ch, uint32_t frame_ms, uint8_t index, gpsb_plan* plan)
{
    if (!ch || !plan) return GPSB_ERR_ARG;
    hx_trk_plan(ch, &g_shared_aux, frame_ms, index, plan);
    return GPSB_OK;
}

int gpsb_host_finish_track(gps_ch_t* ch, uint8_t index, const gpsb_plan* plan, const gpsb_search_res* res,
                           const int16_t* iq6)
{
    if (!ch || !plan) return GPSB_ERR_ARG;
    if (plan->want == GPSB_WANT_SEARCH) hx_trk_finish_search(ch, &g_shared_aux, index, res ? res : &k_empty_window);
    else if (plan->want == GPSB_WANT_EPL && iq6) hx_trk_finish_epl(ch, &g_shared_aux, index, iq6);
    return GPSB_OK;
}

int gpsb_host_last_nav_bit(void) { return g_shared_aux.last_nav_bit; }

/* ---------------------------------------------------------------------------- flat snapshots */
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

uint32_t gpsb_host_sizeof_channel(void) { return (uint32_t)sizeof(gps_ch_t); }

#define COPY_OUT(dst, src) o->dst = (uint32_t)(src)
void gpsb_host_snapshot(const gps_ch_t* ch, struct gpsb_flat_state* o)
{
    const gps_acq_t* a = &ch->acq_data;
    const gps_tracking_t* t = &ch->tracking_data;
    const gps_nav_data_t* n = &ch->nav_data;
    memset(o, 0, sizeof *o);
    o->prn = ch->prn;
    COPY_OUT(acq_state, a->state);                   COPY_OUT(freq_index, a->freq_index);
    o->found_freq_offset_hz = a->found_freq_offset_hz; o->given_freq_offset_hz = a->given_freq_offset_hz;
    COPY_OUT(found_code_phase, a->found_code_phase);
    COPY_OUT(acq_code_search_start, a->code_search_start);
    COPY_OUT(acq_code_search_stop, a->code_search_stop);
    COPY_OUT(code_hist_step, a->code_hist_step);     COPY_OUT(acq_start_timestamp, a->start_timestamp);
    o->hist_ratio_bits = f2u(a->hist_ratio);
    memcpy(o->code_phase_histogram, a->code_phase_histogram, GPSB_FLAT_HIST_SIZE);

    COPY_OUT(trk_state, t->state);
    COPY_OUT(trk_code_search_start, t->code_search_start);
    COPY_OUT(trk_code_search_stop, t->code_search_stop);
    o->if_freq_offset_hz_bits = f2u(t->if_freq_offset_hz);
    COPY_OUT(if_freq_accum, t->if_freq_accum);       COPY_OUT(pre_track_count, t->pre_track_count);
    COPY_OUT(prev_track_timestamp, t->prev_track_timestamp);
    o->code_phase_fine_bits = f2u(t->code_phase_fine);
    o->old_code_phase_fine_bits = f2u(t->old_code_phase_fine);
    COPY_OUT(code_phase_swap_flag, t->code_phase_swap_flag);
    o->dll_code_err_bits = f2u(t->dll_code_err);     o->pll_code_err_bits = f2u(t->pll_code_err);
    o->fll_old_i = t->fll_old_i;                     o->fll_old_q = t->fll_old_q;
    o->fll_err_bits = f2u(t->fll_err);
    COPY_OUT(pll_bad_state_cnt, t->pll_bad_state_cnt);
    COPY_OUT(pll_bad_state_master_cnt, t->pll_bad_state_master_cnt);
    COPY_OUT(i_part_summ, t->i_part_summ);           COPY_OUT(q_part_summ, t->q_part_summ);
    COPY_OUT(snr_summ_cnt, t->snr_summ_cnt);         o->snr_value_bits = f2u(t->snr_value);
    COPY_OUT(filt_start_time_ms, t->filt_start_time_ms);
    COPY_OUT(code_filt_cnt, t->code_filt_cnt);
    o->code_phase_fine_filt_bits = f2u(t->code_phase_fine_filt);
    memcpy(o->pre_track_phases, t->pre_track_phases, sizeof o->pre_track_phases);
    memcpy(o->pll_check_buf, t->pll_check_buf, sizeof o->pll_check_buf);

    COPY_OUT(period_sync_ok_flag, n->period_sync_ok_flag); COPY_OUT(right_period_cnt, n->right_period_cnt);
    COPY_OUT(old_swap_time, n->old_swap_time);       COPY_OUT(old_reminder, n->old_reminder);
    COPY_OUT(accurate_swap_time, n->accurate_swap_time); COPY_OUT(accurate_swap_ok, n->accurate_swap_ok);
    COPY_OUT(last_bit_pos_cnt, n->last_bit_pos_cnt); COPY_OUT(last_bit_neg_cnt, n->last_bit_neg_cnt);
    COPY_OUT(inv_polarity_flag, n->inv_polarity_flag); COPY_OUT(polarity_found, n->polarity_found);
    COPY_OUT(inv_preabmle_cnt, n->inv_preabmle_cnt); COPY_OUT(word_cnt, n->word_cnt);
    COPY_OUT(word_bit_cnt, n->word_bit_cnt);         COPY_OUT(old_D29, n->old_D29);
    COPY_OUT(old_D30, n->old_D30);
    COPY_OUT(word_detection_timestamp, n->word_detection_timestamp);
    COPY_OUT(word_cnt_test, n->word_cnt_test);       COPY_OUT(last_subframe_time, n->last_subframe_time);
    COPY_OUT(first_subframe_time, n->first_subframe_time); COPY_OUT(subframe_cnt, n->subframe_cnt);
    COPY_OUT(new_subframe_flag, n->new_subframe_flag);
    memcpy(o->word_buf, n->word_buf, GPSB_FLAT_WORD_BITS);
    memcpy(o->subframe_data, n->subframe_data, GPSB_FLAT_SUBFRAME_BYTES);
}

void gpsb_host_restore(gps_ch_t* ch, const struct gpsb_flat_state* s)
{
    gps_acq_t* a = &ch->acq_data;
    gps_tracking_t* t = &ch->tracking_data;
    gps_nav_data_t* n = &ch->nav_data;
    a->state = (gps_acq_state_t)s->acq_state;        a->freq_index = (uint8_t)s->freq_index;
    a->found_freq_offset_hz = (int16_t)s->found_freq_offset_hz;
    a->given_freq_offset_hz = (int16_t)s->given_freq_offset_hz;
    a->found_code_phase = (uint16_t)s->found_code_phase;
    a->code_search_start = (uint16_t)s->acq_code_search_start;
    a->code_search_stop = (uint16_t)s->acq_code_search_stop;
    a->code_hist_step = (uint16_t)s->code_hist_step; a->start_timestamp = s->acq_start_timestamp;
    a->hist_ratio = u2f(s->hist_ratio_bits);
    memcpy(a->code_phase_histogram, s->code_phase_histogram, GPSB_FLAT_HIST_SIZE);

    t->state = (gps_tracking_state_t)s->trk_state;
    t->code_search_start = (uint16_t)s->trk_code_search_start;
    t->code_search_stop = (uint16_t)s->trk_code_search_stop;
    t->if_freq_offset_hz = u2f(s->if_freq_offset_hz_bits);
    t->if_freq_accum = s->if_freq_accum;             t->pre_track_count = (uint8_t)s->pre_track_count;
    t->prev_track_timestamp = s->prev_track_timestamp;
    t->code_phase_fine = u2f(s->code_phase_fine_bits);
    t->old_code_phase_fine = u2f(s->old_code_phase_fine_bits);
    t->code_phase_swap_flag = (uint8_t)s->code_phase_swap_flag;
    t->dll_code_err = u2f(s->dll_code_err_bits);     t->pll_code_err = u2f(s->pll_code_err_bits);
    t->fll_old_i = (int16_t)s->fll_old_i;            t->fll_old_q = (int16_t)s->fll_old_q;
    t->fll_err = u2f(s->fll_err_bits);
    t->pll_bad_state_cnt = (uint8_t)s->pll_bad_state_cnt;
    t->pll_bad_state_master_cnt = (uint16_t)s->pll_bad_state_master_cnt;
    t->i_part_summ = s->i_part_summ;                 t->q_part_summ = s->q_part_summ;
    t->snr_summ_cnt = (uint16_t)s->snr_summ_cnt;     t->snr_value = u2f(s->snr_value_bits);
    t->filt_start_time_ms = s->filt_start_time_ms;   t->code_filt_cnt = (uint16_t)s->code_filt_cnt;
    t->code_phase_fine_filt = u2f(s->code_phase_fine_filt_bits);
    memcpy(t->pre_track_phases, s->pre_track_phases, sizeof t->pre_track_phases);
    memcpy(t->pll_check_buf, s->pll_check_buf, sizeof t->pll_check_buf);

    n->period_sync_ok_flag = (uint8_t)s->period_sync_ok_flag; n->right_period_cnt = (uint8_t)s->right_period_cnt;
    n->old_swap_time = s->old_swap_time;             n->old_reminder = (uint8_t)s->old_reminder;
    n->accurate_swap_time = (uint8_t)s->accurate_swap_time; n->accurate_swap_ok = (uint8_t)s->accurate_swap_ok;
    n->last_bit_pos_cnt = (uint8_t)s->last_bit_pos_cnt; n->last_bit_neg_cnt = (uint8_t)s->last_bit_neg_cnt;
    n->inv_polarity_flag = (uint8_t)s->inv_polarity_flag; n->polarity_found = (uint8_t)s->polarity_found;
    n->inv_preabmle_cnt = (uint8_t)s->inv_preabmle_cnt; n->word_cnt = (uint8_t)s->word_cnt;
    n->word_bit_cnt = (uint8_t)s->word_bit_cnt;      n->old_D29 = (uint8_t)s->old_D29;
    n->old_D30 = (uint8_t)s->old_D30;
    n->word_detection_timestamp = s->word_detection_timestamp;
    n->word_cnt_test = s->word_cnt_test;             n->last_subframe_time = s->last_subframe_time;
    n->first_subframe_time = s->first_subframe_time; n->subframe_cnt = (uint16_t)s->subframe_cnt;
    n->new_subframe_flag = (uint8_t)s->new_subframe_flag;
    memcpy(n->word_buf, s->word_buf, GPSB_FLAT_WORD_BITS);
    memcpy(n->subframe_data, s->subframe_data, GPSB_FLAT_SUBFRAME_BYTES);
}

/* Channel array helpers for language bindings that cannot lay out gps_ch_t themselves. */
gps_ch_t* gpsb_host_channels_alloc(uint32_t n) { return (gps_ch_t*)calloc(n, sizeof(gps_ch_t)); }
void gpsb_host_channels_free(gps_ch_t* p) { free(p); }
gps_ch_t* gpsb_host_channel_at(gps_ch_t* base, uint32_t i) { return base + i; }
void gpsb_host_channel_init(gps_ch_t* ch, uint32_t prn, int32_t given_freq_offset_hz)
{
    memset(ch, 0, sizeof *ch);
    ch->prn = (uint8_t)prn;
    ch->acq_data.given_freq_offset_hz = (int16_t)given_freq_offset_hz;
    if (prn >= 1) gps_generate_prn(ch->prn_code, (int)prn);
}
const uint8_t* gpsb_host_channel_code(const gps_ch_t* ch) { return ch->prn_code; }
