/*
 * track.c - per-channel tracking: pre-track refinement of the code phase, then the 1-kHz
 * early/prompt/late loop (2nd-order DLL on normalised early-late power, Costas PLL on slot index 0,
 * FLL on slot indices 1..3, false-lock kicker, SNR estimate).
 *
 * Behaviour follows Firmware/project_main/GPS/tracking.c (cited per function).  Split into PLAN (what
 * the channel needs correlated this millisecond) and FINISH (loop filters on the six sums) so that all
 * channels of a millisecond share one k_epl launch.
 *
 * Float arithmetic: written so that gcc (-O2 -fno-fast-math -ffp-contract=off, x86-64 SSE) performs
 * the same operations in the same precision as the reference compiled with the same flags.  In
 * particular tracking.c ends up using <math.h>'s DOUBLE M_PI on a glibc host (its include order lets
 * math.h redefine the float macro of gps_misc.h:17), so the comparisons and reflections against
 * pi are carried out in double here too.
 */
#include <math.h>
#include <stdlib.h>

#include "host_internal.h"

#define PRE_TRACK_ZONE         30                                   /* tracking.c:17 */
#define PRE_TRACK_PER_MS       (PRE_TRACK_ZONE / GPSB_SLOT_LEN)     /* 7 offsets per ms, tracking.c:20 */

/* ---------------------------------------------------------------------------- plan */
/* tracking.c:52-72: load the acquisition result into a +-15 half-chip pre-track window */
static void arm_pre_track(gps_ch_t* ch)
{
    gps_tracking_t* t = &ch->tracking_data;
    uint16_t lo = (uint16_t)(ch->acq_data.found_code_phase - PRE_TRACK_ZONE / 2);
    uint16_t hi = (uint16_t)(ch->acq_data.found_code_phase + PRE_TRACK_ZONE / 2);
    if (lo > GPSB_HALF_CHIPS) lo = 0;
    if (hi > GPSB_HALF_CHIPS) hi = GPSB_HALF_CHIPS;
    t->code_search_start = lo;
    t->code_search_stop = hi;
    t->if_freq_offset_hz = (float)ch->acq_data.found_freq_offset_hz;
    t->pre_track_count = 0;
    memset(t->pre_track_phases, 0, sizeof t->pre_track_phases);
    t->state = GPS_PRE_TRACK_RUN;
}

void hx_trk_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, uint8_t index, gpsb_plan* plan)
{
    (void)aux;
    gps_tracking_t* t = &ch->tracking_data;
    plan->want = GPSB_WANT_NOTHING;
    plan->stage = 0;

    if (t->state == GPS_NEED_PRE_TRACK) arm_pre_track(ch);

    if (t->state == GPS_PRE_TRACK_RUN) {                          /* tracking.c:398-426 */
        if (index >= GPSB_SLOT_LEN) return;
        unsigned first = (uint16_t)(t->code_search_start + index * PRE_TRACK_PER_MS);
        unsigned last = (uint16_t)(first + PRE_TRACK_PER_MS);
        if (last > GPSB_HALF_CHIPS) last = GPSB_HALF_CHIPS;
        plan->want = GPSB_WANT_SEARCH;
        plan->stage = 3;
        plan->search.sv_slot = ch->prn;
        plan->search.ms_index = frame_ms;
        plan->search.acc0 = 0;
        plan->search.step32 = hx_nco_step32((float)IF_FREQ_HZ + t->if_freq_offset_hz);
        plan->search.off_bits = 0;
        plan->search.start = (uint16_t)first;
        plan->search.stop = (uint16_t)last;
        plan->search.flags = 0;
        return;                                                    /* a run that completes stays DONE this ms */
    }
    if (t->state == GPS_PRE_TRACK_DONE) t->state = GPS_TRACKING_RUN;
    if (t->state != GPS_TRACKING_RUN) return;

    /* tracking.c:92-123 (core/gpsb_loop_core.h, shared with the device-resident loop) */
    if (index >= GPSB_SLOT_LEN) return;                            /* dummy slot */
    plan->want = GPSB_WANT_EPL;
    plan->stage = 4;
    lc_trk_plan_run(ch, hx_now_ms(), frame_ms, &plan->epl);
}

/* ---------------------------------------------------------------------------- pre-track finish */
static int cmp_u16(const void* x, const void* y) { return (int)*(const uint16_t*)x - (int)*(const uint16_t*)y; }

/* tracking.c:459-499: most frequent phase among the collected slot winners (longest run of equal
 * values after sorting; a phase of 0 means "nothing found"). */
static void settle_pre_track(gps_ch_t* ch, uint8_t n)
{
    gps_tracking_t* t = &ch->tracking_data;
    qsort(t->pre_track_phases, n, sizeof(uint16_t), cmp_u16);
    uint8_t run = 0;
    uint16_t best_run = 0, winner = 0;
    for (uint8_t i = 1; i < n; i++) {
        uint16_t gap = (uint16_t)(t->pre_track_phases[i] - t->pre_track_phases[i - 1]);
        if (abs(gap) < 1) {
            run++;
        } else {
            if (run > best_run) { best_run = run; winner = t->pre_track_phases[i - 1]; }
            run = 0;
        }
    }
    if (run > best_run) { best_run = run; winner = t->pre_track_phases[n - 1]; }
    if (winner) {
        t->code_phase_fine = (float)(winner * GPSB_FINE_PER_HALFCHIP);
        t->state = GPS_PRE_TRACK_DONE;
    }
}

/* tracking.c:417-449.  res is the window's (max, first argmax): scanning the window with a strict
 * '>' against the running best is the same as comparing the window maximum once. */
void hx_trk_finish_search(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const gpsb_search_res* res)
{
    gps_tracking_t* t = &ch->tracking_data;
    if (res && (int16_t)res->max > (int16_t)aux->pre_best_value) {
        aux->pre_best_value = res->max;
        aux->pre_best_phase = res->phase;
    }
    if (index != GPSB_SLOT_LEN - 1) return;
    t->pre_track_phases[t->pre_track_count] = aux->pre_best_phase;   /* note: phase is NOT reset per slot */
    t->pre_track_count++;
    if (t->pre_track_count > PRE_TRACK_POINTS_MAX_CNT - 10) settle_pre_track(ch, t->pre_track_count);
    if (t->pre_track_count >= PRE_TRACK_POINTS_MAX_CNT) {
        t->pre_track_count = 0;
        memset(t->pre_track_phases, 0, sizeof t->pre_track_phases);
    }
    aux->pre_best_value = 0;
}

/* ---------------------------------------------------------------------------- loop filters */
/* tracking.c:140-169: DLL, PLL, FLL and false-lock kicker, then nav bits and the SNR estimate.  The code is
 * in core/gpsb_loop_core.h because the device-resident loop runs the very same source. */
void hx_trk_finish_epl(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const int16_t iq[6])
{
    lc_finish_loops(ch, aux, index, iq);
    lc_finish_tail(ch, aux, index, iq[2], iq[3], hx_now_ms());
}

/* ---------------------------------------------------------------------------- reference-named entry point */
/* tracking.c:50-87: one channel, one millisecond, synchronous round trip through the GPU. */
void gps_tracking_process(gps_ch_t* channel, uint8_t* data, uint8_t index)
{
    if (!channel) return;
    gpsb_ctx* ctx = gpsb_host_context();
    gpsb_plan plan;
    /* Planning already moves the channel on (time stamp, carrier NCO, pre-track arming): should the GPU step that
     * follows fail, the tracking record is put back, so that a failed millisecond counts as one the channel did not see
     * (the reference cannot fail here; the status is in gpsb_host_last_status()). */
    const gps_tracking_t before = channel->tracking_data;
    /* the frame goes to ring slot (ms counter mod ring); the plan only needs the number */
    hx_trk_plan(channel, &g_shared_aux, hx_now_ms(), index, &plan);
    if (plan.want == GPSB_WANT_NOTHING) return;
    uint32_t frame;
    if (hx_stage_frame(data, &frame) != GPSB_OK) { channel->tracking_data = before; return; }
    if (plan.want == GPSB_WANT_SEARCH) {
        gpsb_search_res res;
        memset(&res, 0, sizeof res);
        if (plan.search.start < plan.search.stop && hx_note(gpsb_search(ctx, 1, &plan.search, &res)) != GPSB_OK) {
            channel->tracking_data = before;
            return;
        }
        hx_trk_finish_search(channel, &g_shared_aux, index, &res);
    } else {
        int16_t iq[6];
        if (hx_note(gpsb_track_epl(ctx, 1, &plan.epl, iq)) != GPSB_OK) { channel->tracking_data = before; return; }
        hx_trk_finish_epl(channel, &g_shared_aux, index, iq);
    }
}
