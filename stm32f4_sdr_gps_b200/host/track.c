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

/* the pre-track arithmetic (tracking.c:52-72, 398-499) is in core/gpsb_loop_core.h (lc_pre_*): one source for this
 * library and for the device-resident pre-track loop k_pretrack_run */

void hx_trk_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, uint8_t index, gpsb_plan* plan)
{
    (void)aux;
    gps_tracking_t* t = &ch->tracking_data;
    plan->want = GPSB_WANT_NOTHING;
    plan->stage = 0;

    if (t->state == GPS_NEED_PRE_TRACK) lc_pre_arm(ch);

    if (t->state == GPS_PRE_TRACK_RUN) {                          /* tracking.c:398-426 */
        if (index >= GPSB_SLOT_LEN) return;
        uint16_t first, last;
        lc_pre_window(t, index, &first, &last);
        plan->want = GPSB_WANT_SEARCH;
        plan->stage = 3;
        plan->search.sv_slot = ch->prn;
        plan->search.ms_index = frame_ms;
        plan->search.acc0 = 0;
        plan->search.step32 = hx_nco_step32((float)IF_FREQ_HZ + t->if_freq_offset_hz);
        plan->search.off_bits = 0;
        plan->search.start = first;
        plan->search.stop = last;
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
/* tracking.c:417-449, 459-499 (core/gpsb_loop_core.h, lc_pre_finish / lc_pre_settle) */
void hx_trk_finish_search(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const gpsb_search_res* res)
{
    lc_pre_finish(ch, aux, index, res ? res->max : 0, res ? res->phase : 0);
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
