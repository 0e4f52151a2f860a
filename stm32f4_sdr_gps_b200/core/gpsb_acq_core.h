/*
 * gpsb_acq_core.h - the code-phase narrowing rounds of the acquisition state machine as ONE source compiled twice:
 *
 *   - by gcc into libgpsb_host.so (host/acq.c: hx_acq_plan / hx_acq_finish, one search cell per snapshot and channel);
 *   - by nvcc into libgpsb_cuda.so (k_code_rounds_run): a channel's CTA runs the rounds over a whole span of
 *     snapshots in one launch - plan, window search, vote - with no host round trip in between.
 *
 * Behaviour follows Firmware/project_main/GPS/acquisition.c of iliasam/STM32F4_SDR_GPS (cited per function).  Integer
 * arithmetic plus two fp32 divisions and comparisons that gcc (-ffp-contract=off) and nvcc (--fmad=false, IEEE divide)
 * perform identically.  `now` is the reference's millisecond counter (signal_capture_get_packet_cnt) at the snapshot.
 */
#ifndef GPSB_ACQ_CORE_H
#define GPSB_ACQ_CORE_H

#include "gpsb_loop_core.h"

#define AC_CODE_SEARCH_TIMEOUT_MS   120000u     /* acquisition.c:13 */
#define AC_CODE_SEARCH2_WIDTH       500         /* acquisition.c:15 */
#define AC_CODE_SEARCH3_WIDTH       60          /* acquisition.c:16 */

/* acquisition_buffers_reset, acquisition.c:60-65 */
LC_FN void ac_clear_vote_buffers(gpsb_aux* aux)
{
    for (unsigned i = 0; i < sizeof aux->freq_hist / sizeof aux->freq_hist[0]; i++) aux->freq_hist[i] = 0;
    for (unsigned i = 0; i < sizeof aux->bin_phases / sizeof aux->bin_phases[0]; i++) aux->bin_phases[i] = 0;
    aux->bin_count = 0;
}

/* Window of `width` half chips centred on the phase found so far; uint16 wrap-around below zero and
 * overshoot past 2046 are both clamped the way the reference does (acquisition.c:112-118, 156-164). */
LC_FN void ac_centre_window(gps_acq_t* a, unsigned width)
{
    uint16_t lo = (uint16_t)(a->found_code_phase - width / 2);
    uint16_t hi = (uint16_t)(a->found_code_phase + width / 2);
    if (lo > LC_HALF_CHIPS) lo = 0;
    if (hi > LC_HALF_CHIPS) hi = LC_HALF_CHIPS;
    a->code_search_start = lo;
    a->code_search_stop = hi;
    a->code_hist_step = (uint16_t)(width / ACQ_PHASE1_HIST_SIZE + 1);
}

LC_FN void ac_begin_narrow_round(gps_ch_t* ch, gpsb_aux* aux, unsigned width, gps_acq_state_t next, uint32_t now)
{
    gps_acq_t* a = &ch->acq_data;
    for (unsigned i = 0; i < ACQ_PHASE1_HIST_SIZE; i++) a->code_phase_histogram[i] = 0;
    ac_centre_window(a, width);
    ac_clear_vote_buffers(aux);
    a->start_timestamp = now;
    a->state = next;
}

/* acquisition.c:89-104: first narrowing round covers every half chip with 64-wide histogram bins */
LC_FN void ac_start_code_search(gps_ch_t* ch, uint32_t now)
{
    gps_acq_t* a = &ch->acq_data;
    if (a->state != GPS_ACQ_FREQ_SEARCH_DONE) return;
    for (unsigned i = 0; i < ACQ_PHASE1_HIST_SIZE; i++) a->code_phase_histogram[i] = 0;
    a->code_search_start = 0;
    a->code_search_stop = LC_HALF_CHIPS;
    a->code_hist_step = ACQ_PHASE1_HIST_STEP;
    a->start_timestamp = now;
    a->state = GPS_ACQ_CODE_PHASE_SEARCH1;
}

/* acquisition.c:106-130 on a channel's own vote buffers */
LC_FN void ac_start_code_search3(gps_ch_t* ch, gpsb_aux* aux, uint32_t now)
{
    if (ch->acq_data.state != GPS_ACQ_CODE_PHASE_SEARCH2_DONE) return;
    ac_begin_narrow_round(ch, aux, AC_CODE_SEARCH3_WIDTH, GPS_ACQ_CODE_PHASE_SEARCH3, now);
}

/* The code-round part of acquisition.c:134-192 up to (not including) the correlation: state transitions that happen on a
 * snapshot, then 1 when the channel wants its window [code_search_start, code_search_stop) searched on this snapshot at
 * IF + found_freq_offset_hz. */
LC_FN int ac_code_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t now)
{
    gps_acq_t* a = &ch->acq_data;
    if (a->state == GPS_ACQ_CODE_PHASE_SEARCH1_DONE) {        /* arm round 2, work starts next snapshot */
        ac_begin_narrow_round(ch, aux, AC_CODE_SEARCH2_WIDTH, GPS_ACQ_CODE_PHASE_SEARCH2, now);
        return 0;
    }
    if (a->state == GPS_ACQ_CODE_PHASE_SEARCH3_DONE) a->state = GPS_ACQ_DONE;                   /* :176-180 */
    return a->state == GPS_ACQ_CODE_PHASE_SEARCH1 || a->state == GPS_ACQ_CODE_PHASE_SEARCH2 ||
           a->state == GPS_ACQ_CODE_PHASE_SEARCH3;
}

/* acquisition.c:211-275: the vote on one window's best phase */
LC_FN void ac_finish_code_window(gps_ch_t* ch, uint16_t best, uint32_t now)
{
    gps_acq_t* a = &ch->acq_data;
    if (best < a->code_search_start || best >= a->code_search_stop) return;

    if (now - a->start_timestamp > AC_CODE_SEARCH_TIMEOUT_MS) {   /* stale votes: start the round over */
        for (unsigned i = 0; i < ACQ_PHASE1_HIST_SIZE; i++) a->code_phase_histogram[i] = 0;
        a->start_timestamp = now;
    }
    uint8_t cell = (uint8_t)((best - a->code_search_start) / a->code_hist_step);
    if (cell < ACQ_PHASE1_HIST_SIZE) a->code_phase_histogram[cell]++;

    uint16_t used = (uint16_t)((a->code_search_stop + 2 - a->code_search_start) / a->code_hist_step);
    uint8_t top = 0, top_cell = 0, occupied = 0;
    for (uint8_t i = 0; i < used; i++) {          /* reference reads past 32 cells only if used > 32: it is not */
        uint8_t v = a->code_phase_histogram[i];
        if (v > top) { top = v; top_cell = i; }
        if (v > 0) occupied++;
    }
    if (top < 2) return;

    uint32_t sum = 0;
    uint8_t cnt = 0;
    for (uint8_t i = 0; i < ACQ_PHASE1_HIST_SIZE; i++)
        if (a->code_phase_histogram[i] > 0) { sum += a->code_phase_histogram[i]; cnt++; }
    float mean = (float)sum / (float)cnt;
    if (mean < 0.01f) return;
    float ratio = (float)top / mean;
    if (occupied == 1 && top > 3) ratio = 10.0f;
    if (ratio > 3.2f) {
        a->found_code_phase = (uint16_t)(a->code_search_start + top_cell * a->code_hist_step);
        if (a->state == GPS_ACQ_CODE_PHASE_SEARCH1) a->state = GPS_ACQ_CODE_PHASE_SEARCH1_DONE;
        else if (a->state == GPS_ACQ_CODE_PHASE_SEARCH2) a->state = GPS_ACQ_CODE_PHASE_SEARCH2_DONE;
        else if (a->state == GPS_ACQ_CODE_PHASE_SEARCH3) a->state = GPS_ACQ_CODE_PHASE_SEARCH3_DONE;
    }
}

/* One snapshot of one channel in the code rounds, the window's search result supplied by `search` semantics of the
 * caller: returns 1 and the window when a search is wanted; the caller then calls ac_finish_code_window with the best
 * phase (0 for an empty window, gps_misc.c:161-181). */

#endif /* GPSB_ACQ_CORE_H */
