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
#define FALSE_LOCK_LIMIT       80                                   /* tracking.c:14 */
#define SNR_WINDOW             200                                  /* tracking.c:26 */
#define LOOP_DT_S              0.001f

static const double kPi = 3.14159265358979323846;

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

/* tracking.c:115-130: byte offsets of the three arms and the sub-byte replica shift */
static void arm_offsets(float code_phase_fine, gpsb_epl_req* rq)
{
    int16_t fine = (int16_t)code_phase_fine;
    uint16_t prompt = (uint16_t)(fine / GPSB_FINE_PER_HALFCHIP);
    uint16_t early = (uint16_t)(prompt - 1);
    uint16_t late = (uint16_t)(prompt + 1);
    if (early >= GPSB_HALF_CHIPS) early = GPSB_HALF_CHIPS - 1;
    if (late >= GPSB_HALF_CHIPS) late = 0;
    /* A code phase just above 16368 (the DLL's "16368 - x" wrap of a negative x, tracking.c:353-358) gives
     * a prompt offset of exactly 2046; the reference's pointer arithmetic then runs its second loop over
     * the whole buffer from byte 0 (gps_misc.c:57,73-81), i.e. it computes offset 0. */
    if (prompt >= GPSB_HALF_CHIPS) prompt = (uint16_t)(prompt - GPSB_HALF_CHIPS);
    rq->off_bits = (uint16_t)(fine & (GPSB_FINE_PER_HALFCHIP - 1));
    rq->off_e = early;
    rq->off_p = prompt;
    rq->off_l = late;
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

    /* tracking.c:92-123 */
    if (index >= GPSB_SLOT_LEN) return;                            /* dummy slot */
    uint32_t now = hx_now_ms();
    uint32_t gap = now - t->prev_track_timestamp;
    t->prev_track_timestamp = now;
    if (gap > 50) gap = 1;                                         /* first step after start-up */
    if (gap != 1) gps_rewind_if_phase(t, (uint8_t)(gap - 1));      /* ms this channel did not see */

    plan->want = GPSB_WANT_EPL;
    plan->stage = 4;
    plan->epl.sv_slot = ch->prn;
    plan->epl.ms_index = frame_ms;
    arm_offsets(t->code_phase_fine, &plan->epl);
    plan->epl.acc0 = t->if_freq_accum;
    plan->epl.step32 = hx_nco_step32((float)IF_FREQ_HZ + t->if_freq_offset_hz);
    t->if_freq_accum += 511u * plan->epl.step32;                   /* what the mixer leaves behind, gps_misc.c:261-273 */
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
/* tracking.c:333-393 */
static void dll_update(gps_tracking_t* t, int16_t ie, int16_t qe, int16_t il, int16_t ql)
{
    int32_t early = (int32_t)ie * ie + (int32_t)qe * qe;
    int32_t late = (int32_t)il * il + (int32_t)ql * ql;
    float err = (float)(early - late) / (float)(early + late);
    err = -err;
    t->code_phase_fine += (TRACKING_DLL1_C1 * (err - t->dll_code_err) + TRACKING_DLL1_C2 * LOOP_DT_S * err);

    uint8_t wrapped = 0;
    if (t->code_phase_fine < 0.0f) {
        t->code_phase_fine = (float)GPSB_FINE_RANGE - t->code_phase_fine;    /* sic: minus a negative */
        wrapped = 1;
    } else if (t->code_phase_fine > (float)GPSB_FINE_RANGE) {
        t->code_phase_fine = t->code_phase_fine - (float)GPSB_FINE_RANGE;
        wrapped = 1;
    }
#if (ENABLE_CODE_FILTER)
    if (wrapped) {
        t->code_phase_fine_filt = -1.0f;                  /* averaging across a wrap is meaningless: stop */
    } else if (t->code_phase_fine_filt >= 0.0f) {
        t->code_phase_fine_filt += t->code_phase_fine;
        t->code_filt_cnt++;
    }
#endif
    t->dll_code_err = err;
}

/* Fold an angle difference back into [-pi/2, pi/2] the way the reference does (reflection, in double). */
static float fold_half_pi(float x)
{
    if (x > kPi / 2) x = (float)(kPi - x);
    if (x < -kPi / 2) x = (float)(-kPi - x);
    return x;
}

/* tracking.c:175-209 */
static void pll_update(gps_ch_t* ch, uint8_t index, int16_t ip, int16_t qp)
{
    gps_tracking_t* t = &ch->tracking_data;
    float err;
    if (ip > 0) err = (float)(atan2f((float)qp, (float)ip) / kPi);
    else err = (float)(atan2((float)-qp, (float)-ip) / kPi);       /* double atan2 on this branch */
    if (index != 0) return;

    float delta = fold_half_pi(err - t->pll_code_err);
    if (ch->nav_data.period_sync_ok_flag)
        t->if_freq_offset_hz -= TRACKING_PLL2_C1 * delta + (TRACKING_PLL2_C2 * LOOP_DT_S * err);
    else
        t->if_freq_offset_hz -= TRACKING_PLL1_C1 * delta + (TRACKING_PLL1_C2 * LOOP_DT_S * err);
    t->pll_code_err = err;
}

/* tracking.c:261-327: two or more sign flips of IP inside one 4-ms slot cannot be data; count them and,
 * after a long bad streak, jump the carrier to a random frequency at least 200 Hz away. */
static void lock_check(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t ip)
{
    gps_tracking_t* t = &ch->tracking_data;
    if (index >= GPSB_SLOT_LEN) return;
    t->pll_check_buf[index] = ip;
    if (index < GPSB_SLOT_LEN - 1) return;

    uint8_t flips = 0;
    uint8_t prev = t->pll_check_buf[0] > 0;
    for (uint8_t i = 1; i < GPSB_SLOT_LEN; i++) {
        uint8_t cur = t->pll_check_buf[i] > 0;
        if (cur != prev) flips++;
        prev = cur;
    }
    if (flips > 1) {
        if (++t->pll_bad_state_cnt > 10) t->pll_bad_state_cnt = 10;
    } else if (t->pll_bad_state_cnt > 0) {
        t->pll_bad_state_cnt--;
    }
    if (t->pll_bad_state_cnt > 9) t->pll_bad_state_master_cnt++;
    else if (t->pll_bad_state_cnt == 0) t->pll_bad_state_master_cnt = 0;

    if (t->pll_bad_state_master_cnt > FALSE_LOCK_LIMIT) {
        t->pll_bad_state_master_cnt = 0;
        t->pll_bad_state_cnt = 0;
        int16_t candidate, away;
        do {
            uint16_t r = (uint16_t)(hx_rand(aux) % ACQ_SEARCH_STEP_HZ);
            candidate = (int16_t)(ch->acq_data.found_freq_offset_hz - r + (ACQ_SEARCH_STEP_HZ / 2));
            away = (int16_t)((int16_t)t->if_freq_offset_hz - candidate);
        } while (abs(away) < 200);
        t->if_freq_offset_hz = (float)candidate;
    }
}

/* tracking.c:214-256 */
static void fll_update(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t ip, int16_t qp)
{
    gps_tracking_t* t = &ch->tracking_data;
    lock_check(ch, aux, index, ip);
    if (index == 0) {                                 /* first ms of a slot: previous sample is from another time */
        t->fll_old_i = ip;
        t->fll_old_q = qp;
        return;
    }
    int16_t ip0 = t->fll_old_i, qp0 = t->fll_old_q;
    float now = (ip == 0) ? (float)(kPi / 2) : atanf((float)qp / (float)ip);
    float before = (ip0 == 0) ? (float)(kPi / 2) : atanf((float)qp0 / (float)ip0);
    float rot = fold_half_pi(now - before);
    float rot_change = fold_half_pi(rot - t->fll_err);
    float step_hz = TRACKING_FLL1_C1 * LOOP_DT_S * rot_change + (TRACKING_FLL1_C2 * LOOP_DT_S * rot);
    t->if_freq_offset_hz -= step_hz;
    t->fll_old_i = ip;
    t->fll_old_q = qp;
    t->fll_err = rot;
}

/* tracking.c:140-169 */
void hx_trk_finish_epl(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const int16_t iq[6])
{
    gps_tracking_t* t = &ch->tracking_data;
    const int16_t ie = iq[0], qe = iq[1], ip = iq[2], qp = iq[3], il = iq[4], ql = iq[5];
    dll_update(t, ie, qe, il, ql);
    pll_update(ch, index, ip, qp);
    fll_update(ch, aux, index, ip, qp);
    hx_nav_new_code(ch, aux, index, ip);

    t->i_part_summ += (uint32_t)abs(ip);
    t->q_part_summ += (uint32_t)abs(qp);
    t->snr_summ_cnt++;
    if (t->snr_summ_cnt > SNR_WINDOW) {
        if (t->q_part_summ == 0) {
            t->snr_value = 1.0f;
            return;                                   /* sums are left running, like the reference */
        }
        float ratio = (float)t->i_part_summ / (float)t->q_part_summ;
        t->snr_value = 10.0f * log10f(ratio);
        t->snr_summ_cnt = 0;
        t->i_part_summ = 0;
        t->q_part_summ = 0;
    }
}

/* ---------------------------------------------------------------------------- reference-named entry point */
/* tracking.c:50-87: one channel, one millisecond, synchronous round trip through the GPU. */
void gps_tracking_process(gps_ch_t* channel, uint8_t* data, uint8_t index)
{
    if (!channel) return;
    gpsb_ctx* ctx = gpsb_host_context();
    gpsb_plan plan;
    /* the frame goes to ring slot (ms counter mod ring); the plan only needs the number */
    hx_trk_plan(channel, &g_shared_aux, hx_now_ms(), index, &plan);
    if (plan.want == GPSB_WANT_NOTHING) return;
    uint32_t frame;
    if (hx_stage_frame(data, &frame) != GPSB_OK) return;
    if (plan.want == GPSB_WANT_SEARCH) {
        gpsb_search_res res;
        memset(&res, 0, sizeof res);
        if (plan.search.start < plan.search.stop && hx_note(gpsb_search(ctx, 1, &plan.search, &res)) != GPSB_OK) return;
        hx_trk_finish_search(channel, &g_shared_aux, index, &res);
    } else {
        int16_t iq[6];
        if (hx_note(gpsb_track_epl(ctx, 1, &plan.epl, iq)) != GPSB_OK) return;
        hx_trk_finish_epl(channel, &g_shared_aux, index, iq);
    }
}
