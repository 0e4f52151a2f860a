/*
 * acq.c - acquisition state machine: Doppler-bin search with a 10-snapshot chain vote, then three
 * rounds of code-phase narrowing with a 32-bin histogram vote.
 *
 * Behaviour follows Firmware/project_main/GPS/acquisition.c (cited per function).  The work is split
 * into a PLAN half (state transitions + the one search cell the channel needs this snapshot) and a
 * FINISH half (the vote on the cell's result) so that the cells of many channels can share one GPU
 * launch; run back to back they are exactly one acquisition_process_channel() call.
 */
#include <stdlib.h>

#include "host_internal.h"
#include "../core/gpsb_acq_core.h"

/* the code-phase rounds live in core/gpsb_acq_core.h (ac_*): one source for this file and for k_code_rounds_run */
#define SNAPSHOTS_PER_BIN        10          /* ACQ_SINGLE_FREQ_LENGTH, acquisition.c:18 */

static void clear_vote_buffers(gpsb_aux* aux) { ac_clear_vote_buffers(aux); }

/* acquisition.c:68-87 */
static void start_channel(gps_ch_t* ch, gpsb_aux* aux)
{
    gps_acq_t* a = &ch->acq_data;
    if (a->state != GPS_ACQ_NEED_FREQ_SEARCH) return;
    if (a->given_freq_offset_hz != 0) {            /* Doppler supplied by the user: nothing to search */
        a->found_freq_offset_hz = a->given_freq_offset_hz;
        a->state = GPS_ACQ_FREQ_SEARCH_DONE;
        return;
    }
    clear_vote_buffers(aux);
    a->freq_index = 0;
    a->state = GPS_ACQ_FREQ_SEARCH_RUN;
}

void acquisition_start_channel(gps_ch_t* channel) { if (channel) start_channel(channel, &g_shared_aux); }

/* acquisition.c:89-104: first narrowing round covers every half chip with 64-wide histogram bins */
void acquisition_start_code_search_channel(gps_ch_t* channel)
{
    if (channel) ac_start_code_search(channel, hx_now_ms());
}

/* acquisition.c:106-130.  NOTE: shares (and clears) the one set of vote buffers, like the reference. */
void acquisition_start_code_search3_channel(gps_ch_t* channel)
{
    if (channel) ac_start_code_search3(channel, &g_shared_aux, hx_now_ms());
}

/* the same start on a channel's own vote buffers (batched receiver) */
void hx_acq_start_code_search3(gps_ch_t* ch, gpsb_aux* aux) { ac_start_code_search3(ch, aux, hx_now_ms()); }

uint32_t* acquisition_get_hist(void) { return g_shared_aux.freq_hist; }

/* ---------------------------------------------------------------------------- plan */
static void fill_search(gpsb_plan* plan, const gps_ch_t* ch, uint32_t frame_ms, int carrier_hz, unsigned start,
                        unsigned stop, int stage)
{
    plan->want = GPSB_WANT_SEARCH;
    plan->stage = stage;
    plan->search.sv_slot = ch->prn;
    plan->search.ms_index = frame_ms;
    plan->search.acc0 = 0;                                   /* stateless mixer, gps_misc.c:223 */
    plan->search.step32 = hx_nco_step32((float)carrier_hz);  /* int -> float at the call, acquisition.c:288 */
    plan->search.off_bits = 0;
    plan->search.start = (uint16_t)start;
    plan->search.stop = (uint16_t)stop;
    plan->search.flags = 0;
}

/* acquisition.c:134-192 up to (not including) the correlation itself */
void hx_acq_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, gpsb_plan* plan)
{
    gps_acq_t* a = &ch->acq_data;
    plan->want = GPSB_WANT_NOTHING;
    plan->stage = 0;
    if (ch->prn < 1 || a->state == GPS_ACQ_DONE) return;

    if (a->state == GPS_ACQ_FREQ_SEARCH_RUN) {
        int offset_hz = (int16_t)(-ACQ_SEARCH_FREQ_HZ + a->freq_index * ACQ_SEARCH_STEP_HZ);   /* :285 */
        fill_search(plan, ch, frame_ms, IF_FREQ_HZ + offset_hz, 0, GPSB_HALF_CHIPS, 1);
        return;
    }
    if (ac_code_plan(ch, aux, hx_now_ms()))                   /* the code rounds: core/gpsb_acq_core.h */
        fill_search(plan, ch, frame_ms, IF_FREQ_HZ + a->found_freq_offset_hz, a->code_search_start,
                    a->code_search_stop, 2);
}

/* ---------------------------------------------------------------------------- votes */

/* Longest run of neighbouring code phases among the n best phases of one Doppler bin
 * (acquisition.c:322-352).  Neighbours closer than 15 half chips extend a run; a run only counts if it
 * contained at least one pair closer than 3; the last run is accepted without that condition, as in the
 * reference.  phases[] is sorted in place.  *chain_phase (optional) = a member of the winning run. */
uint8_t hx_chain_vote(uint16_t* phases, uint8_t n, uint16_t* chain_phase)
{
    /* ascending order, as the reference's qsort (acquisition.c:326) leaves it; at most 25 values: insertion sort */
    for (uint8_t i = 1; i < n; i++) {
        const uint16_t v = phases[i];
        int k = (int)i - 1;
        for (; k >= 0 && phases[k] > v; k--) phases[k + 1] = phases[k];
        phases[k + 1] = v;
    }
    uint8_t run = 0, tight = 0;
    uint16_t best = 0, best_phase = 0;
    for (uint8_t i = 1; i < n; i++) {
        int gap = abs((int16_t)phases[i] - (int16_t)phases[i - 1]);
        if (gap < 3) tight = 1;
        if (gap < 15) {
            run++;
        } else {
            if (run > best && tight) { best = run; best_phase = phases[i - 1]; }
            run = 0;
            tight = 0;
        }
    }
    if (run > best) { best = run; best_phase = n ? phases[n - 1] : 0; }
    if (chain_phase) *chain_phase = best_phase;
    return (uint8_t)best;
}

/* Decision on the Doppler histogram (acquisition.c:365-416): accept a lone bin with >= 3 votes, or the
 * largest bin when it beats every other non-empty bin by more than 1.7x. */
void hx_freq_hist_decide(gps_ch_t* ch, const uint32_t* hist, uint32_t n_bins, int32_t first_bin_hz, int32_t step_hz)
{
    gps_acq_t* a = &ch->acq_data;
    uint8_t occupied = 0, top_bin = 0, top = 0;
    for (uint32_t b = 0; b < n_bins; b++) {
        if (hist[b] > 0) occupied++;
        if (hist[b] > top) { top = (uint8_t)hist[b]; top_bin = (uint8_t)b; }
    }
    if (occupied == 1 && top >= 3) {
        a->state = GPS_ACQ_FREQ_SEARCH_DONE;
        a->found_freq_offset_hz = (int16_t)(first_bin_hz + top_bin * step_hz);
        a->hist_ratio = 10.0f;
    } else if (occupied > 1) {
        float worst = 10.0;
        for (uint32_t b = 0; b < n_bins; b++) {
            if (hist[b] > 0 && b != top_bin) {
                float r = (float)top / (float)hist[b];
                if (r < worst) worst = r;
            }
        }
        if (worst > 1.7f) {
            a->hist_ratio = worst;
            a->state = GPS_ACQ_FREQ_SEARCH_DONE;
            a->found_freq_offset_hz = (int16_t)(first_bin_hz + top_bin * step_hz);
        }
    }
}

/* ---------------------------------------------------------------------------- finish */
/* acquisition.c:296-311 */
static void finish_freq_cell(gps_ch_t* ch, gpsb_aux* aux, const gpsb_search_res* res)
{
    gps_acq_t* a = &ch->acq_data;
    aux->bin_phases[aux->bin_count++] = res->phase;
    if (aux->bin_count < SNAPSHOTS_PER_BIN) return;

    uint8_t chain = hx_chain_vote(aux->bin_phases, aux->bin_count, NULL);
    if (chain >= 2) aux->freq_hist[a->freq_index] += chain;
    hx_freq_hist_decide(ch, aux->freq_hist, ACQ_COUNT, -ACQ_SEARCH_FREQ_HZ, ACQ_SEARCH_STEP_HZ);
    clear_vote_buffers(aux);
    if (++a->freq_index >= ACQ_COUNT) a->freq_index = 0;
}

/* acquisition.c:211-275 */
static void finish_code_window(gps_ch_t* ch, const gpsb_search_res* res) { ac_finish_code_window(ch, res->phase, hx_now_ms()); }

void hx_acq_finish(gps_ch_t* ch, gpsb_aux* aux, const gpsb_plan* plan, const gpsb_search_res* res)
{
    if (plan->want != GPSB_WANT_SEARCH) return;
    if (plan->stage == 1) finish_freq_cell(ch, aux, res);
    else if (plan->stage == 2) finish_code_window(ch, res);
}

/* ---------------------------------------------------------------------------- reference-named entry points */
void acquisition_process_channel(gps_ch_t* channel, uint8_t* data)
{
    if (!channel) return;
    gpsb_ctx* ctx = gpsb_host_context();
    uint32_t frame;
    gpsb_plan plan;
    hx_acq_plan(channel, &g_shared_aux, 0, &plan);
    if (plan.want != GPSB_WANT_SEARCH) return;
    if (hx_stage_frame(data, &frame) != GPSB_OK) return;
    plan.search.ms_index = frame;
    gpsb_search_res res;
    if (plan.search.start >= plan.search.stop) {
        memset(&res, 0, sizeof res);                        /* empty window: max 0, phase 0 (gps_misc.c:161-181) */
    } else if (hx_note(gpsb_search(ctx, 1, &plan.search, &res)) != GPSB_OK) {
        return;
    }
    hx_acq_finish(channel, &g_shared_aux, &plan, &res);
}

/* acquisition.c:51-57 - all channels on one snapshot; the cells go to the GPU in ONE launch when the
 * outcome cannot depend on the order (channels never share vote buffers while in code search, and at
 * most one channel is in the Doppler search at a time, gps_master.c:93-103). */
void acquisition_process(gps_ch_t* channel, uint8_t* data)
{
    if (!channel) return;
    uint32_t n = gpsb_host_sat_cnt();
    for (uint32_t i = 0; i < n; i++) acquisition_process_channel(&channel[i], data);
}
