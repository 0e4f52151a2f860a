/*
 * ref_shim.c - glue compiled INTO oracle/_ref/libgpsref.so next to the unmodified reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product
 * library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.
 *
 * What it provides:
 *   - the one external symbol the reference hot path needs from the MCU side:
 *     signal_capture_get_packet_cnt() (Firmware/project_main/signal_capture.h:15), backed by a
 *     settable counter;
 *   - allocation + flat snapshot/restore of gps_ch_t (gps_misc.h:184-193) so Python never has to
 *     know the reference struct layout;
 *   - C driver loops (tracking pass, acquisition cells) so the CPU baseline is timed without
 *     per-call ctypes overhead, with the E/P/L sums and nav bits logged on the way.
 *
 * No reference source text is reproduced here; the reference is compiled from /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gps_misc.h"
#include "acquisition.h"
#include "tracking.h"
#include "nav_data.h"
#include "common_ram.h"
#include "config.h"

#include "../include/gpsb_flat_state.h"

/* reference functions that are global but not declared in its headers */
void acquisition_process_channel(gps_ch_t* channel, uint8_t* data);
void gps_generate_prn(uint8_t* dest, int prn);
uint16_t* sim_generate_data(void);
void sim_add_noise(uint16_t* buff_p, uint8_t noise_level);

double ref_now_s(void);

/* ------------------------------------------------------------------ gps_master.c link stubs
 * gps_master.c (compiled unmodified for its acquisition/tracking sequencing, gps_master.c:68-156)
 * also references the terminal UI and the key handler; neither is on the hot path, so they are inert here
 * (the position solver, RTK/solving.c, is compiled in as it lies for row N4). */
#include "gps_master.h"
#include "rtk_common.h"
uint8_t key_up_presed = 0;
void print_state_handling(uint32_t time_ms) { (void)time_ms; }
void print_state_update_acquisition(gps_ch_t* channels, uint32_t time_ms) { (void)channels; (void)time_ms; }
void print_state_update_tracking(gps_ch_t* channels, uint32_t time_ms) { (void)channels; (void)time_ms; }
uint32_t get_dwt_value(void) { return 0; }            /* cycle counter the solver times itself with (delay_us_timer.h) */

/* ------------------------------------------------------------------ ms counter seam */
static uint32_t g_packet_cnt = 0;
uint32_t signal_capture_get_packet_cnt(void) { return g_packet_cnt; }
void ref_set_packet_cnt(uint32_t v) { g_packet_cnt = v; }

/* ------------------------------------------------------------------ channel storage */
uint32_t ref_sizeof_channel(void) { return (uint32_t)sizeof(gps_ch_t); }
uint32_t ref_sat_cnt(void) { return GPS_SAT_CNT; }

gps_ch_t* ref_channels_alloc(uint32_t n) { return (gps_ch_t*)calloc(n, sizeof(gps_ch_t)); }
void ref_channels_free(gps_ch_t* p) { free(p); }
gps_ch_t* ref_channel_at(gps_ch_t* base, uint32_t i) { return base + i; }

void ref_channel_init(gps_ch_t* ch, uint32_t prn, int32_t given_freq_offset_hz)
{
    memset(ch, 0, sizeof(*ch));
    ch->prn = (uint8_t)prn;
    ch->acq_data.given_freq_offset_hz = (int16_t)given_freq_offset_hz;
    gps_channell_prepare(ch);
}

const uint8_t* ref_channel_prn_code(const gps_ch_t* ch) { return ch->prn_code; }

uint16_t* ref_tmp_prn_data(void) { return tmp_prn_data; }
uint16_t* ref_tmp_data_i(void) { return tmp_data_i; }
uint16_t* ref_tmp_data_q(void) { return tmp_data_q; }

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

void ref_channel_snapshot(const gps_ch_t* ch, gpsb_flat_state* o)
{
    const gps_acq_t* a = &ch->acq_data;
    const gps_tracking_t* t = &ch->tracking_data;
    const gps_nav_data_t* n = &ch->nav_data;
    memset(o, 0, sizeof(*o));
    o->prn = ch->prn;

    o->acq_state = (uint32_t)a->state;
    o->freq_index = a->freq_index;
    o->found_freq_offset_hz = a->found_freq_offset_hz;
    o->given_freq_offset_hz = a->given_freq_offset_hz;
    o->found_code_phase = a->found_code_phase;
    o->acq_code_search_start = a->code_search_start;
    o->acq_code_search_stop = a->code_search_stop;
    o->code_hist_step = a->code_hist_step;
    o->acq_start_timestamp = a->start_timestamp;
    o->hist_ratio_bits = f2u(a->hist_ratio);
    memcpy(o->code_phase_histogram, a->code_phase_histogram, GPSB_FLAT_HIST_SIZE);

    o->trk_state = (uint32_t)t->state;
    o->trk_code_search_start = t->code_search_start;
    o->trk_code_search_stop = t->code_search_stop;
    o->if_freq_offset_hz_bits = f2u(t->if_freq_offset_hz);
    o->if_freq_accum = t->if_freq_accum;
    o->pre_track_count = t->pre_track_count;
    o->prev_track_timestamp = t->prev_track_timestamp;
    o->code_phase_fine_bits = f2u(t->code_phase_fine);
    o->old_code_phase_fine_bits = f2u(t->old_code_phase_fine);
    o->code_phase_swap_flag = t->code_phase_swap_flag;
    o->dll_code_err_bits = f2u(t->dll_code_err);
    o->pll_code_err_bits = f2u(t->pll_code_err);
    o->fll_old_i = t->fll_old_i;
    o->fll_old_q = t->fll_old_q;
    o->fll_err_bits = f2u(t->fll_err);
    o->pll_bad_state_cnt = t->pll_bad_state_cnt;
    o->pll_bad_state_master_cnt = t->pll_bad_state_master_cnt;
    o->i_part_summ = t->i_part_summ;
    o->q_part_summ = t->q_part_summ;
    o->snr_summ_cnt = t->snr_summ_cnt;
    o->snr_value_bits = f2u(t->snr_value);
    o->filt_start_time_ms = t->filt_start_time_ms;
    o->code_filt_cnt = t->code_filt_cnt;
    o->code_phase_fine_filt_bits = f2u(t->code_phase_fine_filt);
    memcpy(o->pre_track_phases, t->pre_track_phases, sizeof(o->pre_track_phases));
    memcpy(o->pll_check_buf, t->pll_check_buf, sizeof(o->pll_check_buf));

    o->period_sync_ok_flag = n->period_sync_ok_flag;
    o->right_period_cnt = n->right_period_cnt;
    o->old_swap_time = n->old_swap_time;
    o->old_reminder = n->old_reminder;
    o->accurate_swap_time = n->accurate_swap_time;
    o->accurate_swap_ok = n->accurate_swap_ok;
    o->last_bit_pos_cnt = n->last_bit_pos_cnt;
    o->last_bit_neg_cnt = n->last_bit_neg_cnt;
    o->inv_polarity_flag = n->inv_polarity_flag;
    o->polarity_found = n->polarity_found;
    o->inv_preabmle_cnt = n->inv_preabmle_cnt;
    o->word_cnt = n->word_cnt;
    o->word_bit_cnt = n->word_bit_cnt;
    o->old_D29 = n->old_D29;
    o->old_D30 = n->old_D30;
    o->word_detection_timestamp = n->word_detection_timestamp;
    o->word_cnt_test = n->word_cnt_test;
    o->last_subframe_time = n->last_subframe_time;
    o->first_subframe_time = n->first_subframe_time;
    o->subframe_cnt = n->subframe_cnt;
    o->new_subframe_flag = n->new_subframe_flag;
    memcpy(o->word_buf, n->word_buf, GPSB_FLAT_WORD_BITS);
    memcpy(o->subframe_data, n->subframe_data, GPSB_FLAT_SUBFRAME_BYTES);
}

/* Inverse of the snapshot (prn / prn_code / eph / obs are left untouched). */
void ref_channel_restore(gps_ch_t* ch, const gpsb_flat_state* s)
{
    gps_acq_t* a = &ch->acq_data;
    gps_tracking_t* t = &ch->tracking_data;
    gps_nav_data_t* n = &ch->nav_data;

    a->state = (gps_acq_state_t)s->acq_state;
    a->freq_index = (uint8_t)s->freq_index;
    a->found_freq_offset_hz = (int16_t)s->found_freq_offset_hz;
    a->given_freq_offset_hz = (int16_t)s->given_freq_offset_hz;
    a->found_code_phase = (uint16_t)s->found_code_phase;
    a->code_search_start = (uint16_t)s->acq_code_search_start;
    a->code_search_stop = (uint16_t)s->acq_code_search_stop;
    a->code_hist_step = (uint16_t)s->code_hist_step;
    a->start_timestamp = s->acq_start_timestamp;
    a->hist_ratio = u2f(s->hist_ratio_bits);
    memcpy(a->code_phase_histogram, s->code_phase_histogram, GPSB_FLAT_HIST_SIZE);

    t->state = (gps_tracking_state_t)s->trk_state;
    t->code_search_start = (uint16_t)s->trk_code_search_start;
    t->code_search_stop = (uint16_t)s->trk_code_search_stop;
    t->if_freq_offset_hz = u2f(s->if_freq_offset_hz_bits);
    t->if_freq_accum = s->if_freq_accum;
    t->pre_track_count = (uint8_t)s->pre_track_count;
    t->prev_track_timestamp = s->prev_track_timestamp;
    t->code_phase_fine = u2f(s->code_phase_fine_bits);
    t->old_code_phase_fine = u2f(s->old_code_phase_fine_bits);
    t->code_phase_swap_flag = (uint8_t)s->code_phase_swap_flag;
    t->dll_code_err = u2f(s->dll_code_err_bits);
    t->pll_code_err = u2f(s->pll_code_err_bits);
    t->fll_old_i = (int16_t)s->fll_old_i;
    t->fll_old_q = (int16_t)s->fll_old_q;
    t->fll_err = u2f(s->fll_err_bits);
    t->pll_bad_state_cnt = (uint8_t)s->pll_bad_state_cnt;
    t->pll_bad_state_master_cnt = (uint16_t)s->pll_bad_state_master_cnt;
    t->i_part_summ = s->i_part_summ;
    t->q_part_summ = s->q_part_summ;
    t->snr_summ_cnt = (uint16_t)s->snr_summ_cnt;
    t->snr_value = u2f(s->snr_value_bits);
    t->filt_start_time_ms = s->filt_start_time_ms;
    t->code_filt_cnt = (uint16_t)s->code_filt_cnt;
    t->code_phase_fine_filt = u2f(s->code_phase_fine_filt_bits);
    memcpy(t->pre_track_phases, s->pre_track_phases, sizeof(s->pre_track_phases));
    memcpy(t->pll_check_buf, s->pll_check_buf, sizeof(s->pll_check_buf));

    n->period_sync_ok_flag = (uint8_t)s->period_sync_ok_flag;
    n->right_period_cnt = (uint8_t)s->right_period_cnt;
    n->old_swap_time = s->old_swap_time;
    n->old_reminder = (uint8_t)s->old_reminder;
    n->accurate_swap_time = (uint8_t)s->accurate_swap_time;
    n->accurate_swap_ok = (uint8_t)s->accurate_swap_ok;
    n->last_bit_pos_cnt = (uint8_t)s->last_bit_pos_cnt;
    n->last_bit_neg_cnt = (uint8_t)s->last_bit_neg_cnt;
    n->inv_polarity_flag = (uint8_t)s->inv_polarity_flag;
    n->polarity_found = (uint8_t)s->polarity_found;
    n->inv_preabmle_cnt = (uint8_t)s->inv_preabmle_cnt;
    n->word_cnt = (uint8_t)s->word_cnt;
    n->word_bit_cnt = (uint8_t)s->word_bit_cnt;
    n->old_D29 = (uint8_t)s->old_D29;
    n->old_D30 = (uint8_t)s->old_D30;
    n->word_detection_timestamp = s->word_detection_timestamp;
    n->word_cnt_test = s->word_cnt_test;
    n->last_subframe_time = s->last_subframe_time;
    n->first_subframe_time = s->first_subframe_time;
    n->subframe_cnt = (uint16_t)s->subframe_cnt;
    n->new_subframe_flag = (uint8_t)s->new_subframe_flag;
    memcpy(n->word_buf, s->word_buf, GPSB_FLAT_WORD_BITS);
    memcpy(n->subframe_data, s->subframe_data, GPSB_FLAT_SUBFRAME_BYTES);
}

/* ------------------------------------------------------------------ simulator fixture */
/* The single-sat project's 1-ms generator (SS/GPS/simulator.c:88) + its noise injector (:40).
 * srand(seed) first so that noise>0 fixtures are reproducible with this libc. */
void ref_sim_buffer(uint8_t* out2046, uint32_t noise_level, uint32_t seed)
{
    uint16_t* p = sim_generate_data();
    if (noise_level) {
        srand(seed);
        sim_add_noise(p, (uint8_t)noise_level);
    }
    memcpy(out2046, p, 2046);
}

/* ------------------------------------------------------------------ acquisition cell drivers */
/* One Doppler-bin x 1 ms cell exactly as acquisition_freq_search() evaluates it
 * (acquisition.c:282-294) but with bits / window exposed, returning the triple the reference
 * computes. */
uint16_t ref_search_cell(gps_ch_t* ch, const uint8_t* signal, int32_t freq_offset_hz,
                         uint32_t offset_bits, uint32_t start, uint32_t stop,
                         uint16_t* avr, uint16_t* phase)
{
    gps_generate_prn_data2(ch, tmp_prn_data, (uint16_t)offset_bits);
    gps_shift_to_zero_freq((uint8_t*)signal, (uint8_t*)tmp_data_i, (uint8_t*)tmp_data_q,
                           IF_FREQ_HZ + freq_offset_hz);
    return correlation_search(tmp_prn_data, tmp_data_i, tmp_data_q,
                              (uint16_t)start, (uint16_t)stop, avr, phase);
}

/* Same with an explicit float carrier frequency (pre-track path, tracking.c:403-407). */
uint16_t ref_search_cell_f(gps_ch_t* ch, const uint8_t* signal, float freq_hz,
                           uint32_t offset_bits, uint32_t start, uint32_t stop,
                           uint16_t* avr, uint16_t* phase)
{
    gps_generate_prn_data2(ch, tmp_prn_data, (uint16_t)offset_bits);
    gps_shift_to_zero_freq((uint8_t*)signal, (uint8_t*)tmp_data_i, (uint8_t*)tmp_data_q, freq_hz);
    return correlation_search(tmp_prn_data, tmp_data_i, tmp_data_q,
                              (uint16_t)start, (uint16_t)stop, avr, phase);
}

/* All (I,Q) pairs of one cell: out[2*k] = I, out[2*k+1] = Q for offset start+k. */
void ref_iq_cell(gps_ch_t* ch, const uint8_t* signal, float freq_hz, uint32_t offset_bits,
                 uint32_t start, uint32_t stop, int16_t* out)
{
    gps_generate_prn_data2(ch, tmp_prn_data, (uint16_t)offset_bits);
    gps_shift_to_zero_freq((uint8_t*)signal, (uint8_t*)tmp_data_i, (uint8_t*)tmp_data_q, freq_hz);
    for (uint32_t off = start; off < stop; off++)
        gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, (uint16_t)off,
                           &out[2 * (off - start)], &out[2 * (off - start) + 1]);
}

/* A full sweep of cells: n_sv x n_bins x n_ms, triple per cell in (sv, bin, ms) order.
 * out[3*c + {0,1,2}] = {max, phase, avr}.  This is the CPU baseline of the cold-acquisition metric. */
void ref_sweep_cells(gps_ch_t* chans, uint32_t n_sv, const uint8_t* signal, uint32_t n_ms,
                     int32_t first_bin_hz, int32_t bin_step_hz, uint32_t n_bins,
                     uint32_t offset_bits, uint16_t* out)
{
    for (uint32_t s = 0; s < n_sv; s++)
        for (uint32_t b = 0; b < n_bins; b++)
            for (uint32_t m = 0; m < n_ms; m++) {
                uint16_t avr = 0, phase = 0;
                uint16_t mx = ref_search_cell(&chans[s], signal + 2046u * m,
                                              first_bin_hz + (int32_t)b * bin_step_hz, offset_bits,
                                              0, 2 * PRN_LENGTH, &avr, &phase);
                uint16_t* o = out + 3u * ((s * n_bins + b) * n_ms + m);
                o[0] = mx; o[1] = phase; o[2] = avr;
            }
}

/* ------------------------------------------------------------------ tracking pass driver */
/* gps_misc.c is compiled as-is; tracking.c calls gps_correlation_iq() three times per step in the
 * order E, P, L (tracking.c:136-138).  To log those sums without touching the reference we read them
 * back by re-running the same three correlations on the scratch buffers the step just used:
 * tmp_prn_data / tmp_data_i / tmp_data_q still hold the replica and the mixed data of this ms when
 * gps_tracking_process() returns (common_ram.c:3-5), and the offsets are a pure function of the
 * code_phase_fine value the step started from (tracking.c:115-130). */
static void epl_offsets(float code_phase_fine, uint16_t* e, uint16_t* p, uint16_t* l)
{
    int16_t fine = (int16_t)code_phase_fine;
    uint16_t op = (uint16_t)(fine / 8);
    uint16_t oe = (uint16_t)(op - 1);
    uint16_t ol = (uint16_t)(op + 1);
    if (oe >= 2 * PRN_LENGTH) oe = 2 * PRN_LENGTH - 1;
    if (ol >= 2 * PRN_LENGTH) ol = 0;
    *e = oe; *p = op; *l = ol;
}

/*
 * Run n_ms consecutive tracking steps of ONE channel ("every SV every ms" schedule of
 * index = ms % 4, ms counter = ms_first + k).
 *   iq_log    : NULL or int16[n_ms][6] = IE,QE,IP,QP,IL,QL (zeros for steps that were not TRACKING_RUN)
 *   nav_log   : NULL or int8[n_ms]     = -1 no bit this ms, else the 20-ms bit handed to
 *               gps_nav_data_words_detection() (nav_data.c:239)
 *   state_log : NULL or float[n_ms][2] = code_phase_fine, if_freq_offset_hz after the step
 */
void ref_track_run(gps_ch_t* ch, const uint8_t* signal, uint32_t ms_first, uint32_t n_ms,
                   int16_t* iq_log, int8_t* nav_log, float* state_log)
{
    for (uint32_t k = 0; k < n_ms; k++) {
        uint32_t ms = ms_first + k;
        uint8_t index = (uint8_t)(ms % TRACKING_CH_LENGTH);
        g_packet_cnt = ms;

        /* observer state taken BEFORE the step */
        float fine_before = ch->tracking_data.code_phase_fine;
        gps_nav_data_t nb = ch->nav_data;
        int will_track = (ch->tracking_data.state == GPS_TRACKING_RUN) ||
                         (ch->tracking_data.state == GPS_PRE_TRACK_DONE);

        gps_tracking_process(ch, (uint8_t*)(signal + 2046u * k), index);

        if (iq_log) {
            int16_t* o = iq_log + 6u * k;
            memset(o, 0, 12);
            if (will_track) {
                uint16_t oe, op, ol;
                epl_offsets(fine_before, &oe, &op, &ol);
                gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, oe, &o[0], &o[1]);
                gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, op, &o[2], &o[3]);
                gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, ol, &o[4], &o[5]);
            }
        }
        if (nav_log) {
            int8_t bit = -1;
            if (will_track && nb.period_sync_ok_flag == 1) {
                /* the emit condition of gps_nav_data_bits_extraction (nav_data.c:226-239) */
                uint32_t diff = ms - nb.old_swap_time;
                uint8_t rem = (uint8_t)(diff % 20);
                if (rem < nb.old_reminder)
                    bit = (nb.last_bit_pos_cnt > nb.last_bit_neg_cnt) ? 1 : 0;
            }
            nav_log[k] = bit;
        }
        if (state_log) {
            state_log[2 * k] = ch->tracking_data.code_phase_fine;
            state_log[2 * k + 1] = ch->tracking_data.if_freq_offset_hz;
        }
    }
}

/*
 * The same run on a schedule with idle gaps: the checker's own restatement of the slot-phase walk of this library's
 * batched paths (include/gpsb_host.h, gpsb_rx_set_slot_walk), written from its description, driving the UNMODIFIED
 * reference.  A channel is called with slot index (ms + slot_phase) % 4; inside an idle gap it is called with the
 * reference's dummy index 0xFF (main.c:146-147), which tracking.c:96 ignores.  After a complete slot of a channel that
 * has no refined bit edge (nav_data.accurate_swap_ok == 0):
 *   three or more on-grid edges since the last gap at slot positions 3 or 1, and fewer than a fifth of all on-grid
 *   edges at position 2                                                   -> idle 1 ms (most at 3, ties too) or 3 ms;
 *   no on-grid edge at all for period_ms (default 400) at this slot phase -> idle 2 ms;
 *   the first slot end only starts that clock;
 * the gap begins 5 ms after the slot end at which it was decided, and the millisecond behind a gap is slot index 0.
 * Edge positions are observed from outside: sign of the prompt sum XOR the polarity flag before the call, one flip in
 * the slot, "on grid" = (edge - old_swap_time before the call) % 20 in {0, 1, 19} (nav_data.c:87-113).
 *   idx_log : NULL or uint8[n_ms] = slot index used (0xFF = idle)
 */
typedef struct ref_walk {
    uint32_t enable, period_ms;
    uint32_t slot_phase, gap_first, gap_len, phase_since, gaps_taken, edge_pos;
    uint32_t edges_at[4];             /* on-grid edges seen per slot position since the last gap was decided */
    uint32_t armed;
    uint32_t slot_first_ms;
    uint8_t sign[4];
    int16_t ip[4];
    uint32_t slot_fill;               /* entries of sign[] / ip[] that belong to the slot in progress */
} ref_walk;
uint32_t ref_sizeof_walk(void) { return (uint32_t)sizeof(ref_walk); }

void gps_nav_data_analyse_new_code(gps_ch_t* channel, uint8_t index, int16_t new_i);

void ref_track_run_walk(gps_ch_t* ch, const uint8_t* signal, uint32_t ms_first, uint32_t n_ms, ref_walk* w,
                        int16_t* iq_log, int8_t* nav_log, uint8_t* idx_log)
{
    /* The reference keeps the samples of the slot in progress in function statics shared by all channels
     * (nav_data.c:29,48-51).  A run that resumes inside a slot after ANOTHER channel was run puts this channel's
     * samples back first, through the reference's own function on a scratch channel that does nothing else. */
    if (w->slot_fill) {
        static gps_ch_t scratch;
        for (uint32_t i = 0; i < w->slot_fill && i < TRACKING_CH_LENGTH - 1; i++) {
            memset(&scratch.nav_data, 0, sizeof scratch.nav_data);
            scratch.nav_data.inv_polarity_flag = (uint8_t)(w->sign[i] ^ (w->ip[i] > 0 ? 1 : 0));
            g_packet_cnt = w->slot_first_ms + i;
            gps_nav_data_analyse_new_code(&scratch, (uint8_t)i, w->ip[i]);
        }
    }
    for (uint32_t k = 0; k < n_ms; k++) {
        const uint32_t ms = ms_first + k;
        uint8_t* frame = (uint8_t*)(signal + 2046u * (size_t)k);
        g_packet_cnt = ms;
        if (iq_log) memset(iq_log + 6u * k, 0, 12);
        if (nav_log) nav_log[k] = -1;
        if (w->gap_len && ms >= w->gap_first && ms < w->gap_first + w->gap_len) {      /* not served this millisecond */
            if (idx_log) idx_log[k] = 0xFF;
            gps_tracking_process(ch, frame, 0xFF);
            if (ms == w->gap_first + w->gap_len - 1) {
                w->slot_phase = (4u - ((ms + 1u) % 4u)) % 4u;
                w->phase_since = ms + 1u;
            }
            continue;
        }
        const uint8_t index = (uint8_t)((ms + w->slot_phase) % TRACKING_CH_LENGTH);
        if (idx_log) idx_log[k] = index;

        const float fine_before = ch->tracking_data.code_phase_fine;
        const gps_nav_data_t nb = ch->nav_data;
        const int will_track = (ch->tracking_data.state == GPS_TRACKING_RUN) ||
                               (ch->tracking_data.state == GPS_PRE_TRACK_DONE);
        gps_tracking_process(ch, frame, index);
        if (!will_track) continue;

        int16_t o[6];
        uint16_t oe, op, ol;
        epl_offsets(fine_before, &oe, &op, &ol);
        gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, oe, &o[0], &o[1]);
        gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, op, &o[2], &o[3]);
        gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, ol, &o[4], &o[5]);
        if (iq_log) memcpy(iq_log + 6u * k, o, 12);
        if (nav_log && nb.period_sync_ok_flag == 1) {
            uint8_t rem = (uint8_t)((ms - nb.old_swap_time) % 20);
            if (rem < nb.old_reminder) nav_log[k] = (nb.last_bit_pos_cnt > nb.last_bit_neg_cnt) ? 1 : 0;
        }

        /* the observer and the policy */
        w->sign[index] = (uint8_t)((o[2] > 0 ? 1 : 0) ^ (nb.inv_polarity_flag ? 1 : 0));
        w->ip[index] = o[2];
        if (index == 0) w->slot_first_ms = ms;
        w->slot_fill = (index + 1u) % TRACKING_CH_LENGTH;
        if (index != TRACKING_CH_LENGTH - 1) continue;
        uint32_t changes = 0, where = 0;
        for (uint32_t i = 1; i < TRACKING_CH_LENGTH; i++)
            if (w->sign[i] != w->sign[i - 1]) { changes++; where = i; }
        if (changes == 1) {
            const uint32_t rem = (w->slot_first_ms + where - nb.old_swap_time) % 20u;
            if (rem == 0 || rem == 1 || rem == 19) {
                w->edge_pos = where;
                if (w->edges_at[where] < 255) w->edges_at[where]++;
            }
        }
        if (w->gap_len) {
            if (ms < w->gap_first + w->gap_len) continue;          /* decided, not taken yet */
            w->gap_len = 0;
        }
        if (!w->enable || ch->nav_data.accurate_swap_ok) continue;
        uint32_t gap = 0;
        if (!w->armed) {
            w->armed = 1;
            w->phase_since = ms;
            continue;
        }
        const uint32_t side = w->edges_at[1] + w->edges_at[3];
        if (side >= 3 && 4u * w->edges_at[2] < side) gap = (w->edges_at[3] >= w->edges_at[1]) ? 1 : 3;
        else if (side + w->edges_at[2] == 0 && ms - w->phase_since >= (w->period_ms ? w->period_ms : 400u)) gap = 2;
        if (gap) {
            w->gap_first = ms + 5u;
            w->gap_len = gap;
            w->edge_pos = 0;
            memset(w->edges_at, 0, sizeof w->edges_at);
            w->gaps_taken++;
        }
    }
}

/* The three fused calls of one tracking step on explicit parameters (no loop filters), used to pin
 * the fused E/P/L cell against the reference primitives for arbitrary NCO state. */
void ref_epl_cell(gps_ch_t* ch, const uint8_t* signal, float if_freq_offset_hz, uint32_t accum_in,
                  float code_phase_fine, int16_t out6[6], uint32_t* accum_out)
{
    gps_tracking_t trk;
    memset(&trk, 0, sizeof(trk));
    trk.if_freq_offset_hz = if_freq_offset_hz;
    trk.if_freq_accum = accum_in;
    int16_t fine = (int16_t)code_phase_fine;
    gps_generate_prn_data2(ch, tmp_prn_data, (uint16_t)(fine & 7));
    gps_shift_to_zero_freq_track(&trk, (uint8_t*)signal, (uint8_t*)tmp_data_i, (uint8_t*)tmp_data_q);
    uint16_t oe, op, ol;
    epl_offsets(code_phase_fine, &oe, &op, &ol);
    gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, oe, &out6[0], &out6[1]);
    gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, op, &out6[2], &out6[3]);
    gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, ol, &out6[4], &out6[5]);
    if (accum_out) *accum_out = trk.if_freq_accum;
}

/* ------------------------------------------------------------------ wall-clock helper */
double ref_now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------ CPU-baseline batch drivers */
/* Open-loop E/P/L cells on explicit per-cell state, in one C loop (no per-call ctypes overhead).
 * Cell i uses channel sv[i], millisecond ms[i] of `signal` (2046 bytes per ms). Returns seconds. */
double ref_epl_batch(gps_ch_t* chans, uint32_t n, const uint32_t* sv, const uint32_t* ms,
                     const uint8_t* signal, const float* if_freq_offset_hz, const uint32_t* accum_in,
                     const float* code_phase_fine, int16_t* out6)
{
    double t0 = ref_now_s();
    for (uint32_t i = 0; i < n; i++)
        ref_epl_cell(&chans[sv[i]], signal + 2046u * (size_t)ms[i], if_freq_offset_hz[i], accum_in[i],
                     code_phase_fine[i], out6 + 6u * (size_t)i, 0);
    return ref_now_s() - t0;
}

/* Closed-loop tracking of one channel over n_ms, no logging: the reference's own per-ms step
 * (tracking.c:50) including loop filters and nav-bit extraction. Returns seconds. */
double ref_track_time(gps_ch_t* ch, const uint8_t* signal, uint32_t ms_first, uint32_t n_ms)
{
    double t0 = ref_now_s();
    for (uint32_t k = 0; k < n_ms; k++) {
        g_packet_cnt = ms_first + k;
        gps_tracking_process(ch, (uint8_t*)(signal + 2046u * (size_t)k), (uint8_t)((ms_first + k) % TRACKING_CH_LENGTH));
    }
    return ref_now_s() - t0;
}

/* Timed sweep (cold-acquisition cells). Returns seconds. */
double ref_sweep_time(gps_ch_t* chans, uint32_t n_sv, const uint8_t* signal, uint32_t n_ms,
                      int32_t first_bin_hz, int32_t bin_step_hz, uint32_t n_bins,
                      uint32_t offset_bits, uint16_t* out)
{
    double t0 = ref_now_s();
    ref_sweep_cells(chans, n_sv, signal, n_ms, first_bin_hz, bin_step_hz, n_bins, offset_bits, out);
    return ref_now_s() - t0;
}

/* SURVEY.md section 8(d) config 1 repeated over a recording: per millisecond the reference's replica generator,
 * stateless mixer and ONE gps_correlation_iq at a fixed offset (the recipe of project_single_sat/main.c:59-68 with
 * the prompt correlation instead of the search).  out2[2*m], out2[2*m+1] = I, Q.  Returns seconds. */
double ref_prompt_time(gps_ch_t* ch, const uint8_t* signal, uint32_t n_ms, float freq_hz, uint32_t offset,
                       uint32_t offset_bits, int16_t* out2)
{
    double t0 = ref_now_s();
    for (uint32_t m = 0; m < n_ms; m++) {
        gps_generate_prn_data2(ch, tmp_prn_data, (uint16_t)offset_bits);
        gps_shift_to_zero_freq((uint8_t*)(signal + 2046u * (size_t)m), (uint8_t*)tmp_data_i, (uint8_t*)tmp_data_q, freq_hz);
        gps_correlation_iq(tmp_prn_data, tmp_data_i, tmp_data_q, (uint16_t)offset, &out2[2 * m], &out2[2 * m + 1]);
    }
    return ref_now_s() - t0;
}

/* ------------------------------------------------------------------ ephemeris container, flat image */
static uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
void ref_channel_eph(const gps_ch_t* ch, gpsb_flat_eph* o)
{
    const sdreph_t* d = &ch->eph_data;
    const eph_t* e = &d->eph;
    memset(o, 0, sizeof *o);
    o->sat = e->sat; o->iode = e->iode; o->iodc = e->iodc; o->sva = e->sva; o->svh = e->svh; o->week = e->week;
    o->code = e->code; o->flag = e->flag;
    o->toe_time = (int64_t)e->toe.time; o->toc_time = (int64_t)e->toc.time; o->ttr_time = (int64_t)e->ttr.time;
    o->toe_sec_bits = d2u(e->toe.sec); o->toc_sec_bits = d2u(e->toc.sec); o->ttr_sec_bits = d2u(e->ttr.sec);
    o->A = d2u(e->A); o->e = d2u(e->e); o->i0 = d2u(e->i0); o->OMG0 = d2u(e->OMG0); o->omg = d2u(e->omg);
    o->M0 = d2u(e->M0); o->deln = d2u(e->deln); o->OMGd = d2u(e->OMGd); o->idot = d2u(e->idot);
    o->crc = d2u(e->crc); o->crs = d2u(e->crs); o->cuc = d2u(e->cuc); o->cus = d2u(e->cus); o->cic = d2u(e->cic);
    o->cis = d2u(e->cis); o->toes = d2u(e->toes); o->fit = d2u(e->fit); o->f0 = d2u(e->f0); o->f1 = d2u(e->f1);
    o->f2 = d2u(e->f2);
    for (int i = 0; i < 4; i++) o->tgd[i] = d2u(e->tgd[i]);
    o->ctype = d->ctype; o->week_gpst = d->week_gpst; o->cnt = d->cnt; o->cntth = d->cntth; o->update = d->update;
    o->prn = d->prn; o->week_gst = d->week_gst; o->sub_cnt = d->sub_cnt; o->received_mask = d->received_mask;
    o->received_mask_proc = d->received_mask_proc; o->tow_gpst = d2u(d->tow_gpst);
}

/* gps_nav_data_decode_subframe() on a given 300-bit subframe image (nav_data_decode.c:33) */
uint8_t gps_nav_data_decode_subframe(gps_ch_t* channel);
uint32_t ref_decode_subframe(gps_ch_t* ch, const uint8_t image[38])
{
    memcpy(ch->nav_data.subframe_data, image, 38);
    return gps_nav_data_decode_subframe(ch);
}
/* the word assembler fed bit by bit (nav_data.c:257), ms counter advancing 20 per bit from ms0 */
void gps_nav_data_words_detection(gps_ch_t* channel, uint8_t new_bit);
void ref_feed_nav_bits(gps_ch_t* ch, const uint8_t* bits, uint32_t n, uint32_t ms0)
{
    for (uint32_t i = 0; i < n; i++) {
        g_packet_cnt = ms0 + 20u * i;
        gps_nav_data_words_detection(ch, bits[i]);
    }
}

/* ------------------------------------------------------------------ observations (gps_master.c:159-327) */
void gps_master_nav_handling(gps_ch_t* channels);
void ref_nav_handling(gps_ch_t* chans, uint32_t now_ms)
{
    g_packet_cnt = now_ms;
    gps_master_nav_handling(chans);
}
void ref_channel_obs(const gps_ch_t* ch, uint64_t out2[2])
{
    out2[0] = d2u(ch->obs_data.pseudorange_m);
    out2[1] = d2u(ch->obs_data.tow_s);
}
void ref_channel_set_tow(gps_ch_t* ch, double tow_gpst) { ch->eph_data.tow_gpst = tow_gpst; }


/* ------------------------------------------------------------------ position fix (RTK/solving.c, row N4) */
#include "solving.h"
extern sol_t gps_sol;
extern double final_pos[3];
extern double azel[];
static obsd_t g_ref_obsd[GPS_SAT_CNT];
static double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

void ref_channel_set_eph(gps_ch_t* ch, const gpsb_flat_eph* in)
{
    sdreph_t* d = &ch->eph_data;
    eph_t* e = &d->eph;
    e->sat = in->sat; e->iode = in->iode; e->iodc = in->iodc; e->sva = in->sva; e->svh = in->svh; e->week = in->week;
    e->code = in->code; e->flag = in->flag;
    e->toe.time = (time_t)in->toe_time; e->toc.time = (time_t)in->toc_time; e->ttr.time = (time_t)in->ttr_time;
    e->toe.sec = u2d(in->toe_sec_bits); e->toc.sec = u2d(in->toc_sec_bits); e->ttr.sec = u2d(in->ttr_sec_bits);
    e->A = u2d(in->A); e->e = u2d(in->e); e->i0 = u2d(in->i0); e->OMG0 = u2d(in->OMG0); e->omg = u2d(in->omg);
    e->M0 = u2d(in->M0); e->deln = u2d(in->deln); e->OMGd = u2d(in->OMGd); e->idot = u2d(in->idot);
    e->crc = u2d(in->crc); e->crs = u2d(in->crs); e->cuc = u2d(in->cuc); e->cus = u2d(in->cus); e->cic = u2d(in->cic);
    e->cis = u2d(in->cis); e->toes = u2d(in->toes); e->fit = u2d(in->fit); e->f0 = u2d(in->f0); e->f1 = u2d(in->f1);
    e->f2 = u2d(in->f2);
    for (int i = 0; i < 4; i++) e->tgd[i] = u2d(in->tgd[i]);
    d->ctype = in->ctype; d->week_gpst = in->week_gpst; d->cnt = in->cnt; d->cntth = in->cntth; d->update = in->update;
    d->prn = in->prn; d->week_gst = in->week_gst; d->sub_cnt = (uint16_t)in->sub_cnt;
    d->received_mask = (uint8_t)in->received_mask; d->received_mask_proc = (uint8_t)in->received_mask_proc;
    d->tow_gpst = u2d(in->tow_gpst);
}
void ref_channel_set_obs(gps_ch_t* ch, double pseudorange_m, double tow_s)
{
    ch->obs_data.pseudorange_m = pseudorange_m;
    ch->obs_data.tow_s = tow_s;
}

void ref_fix_init(gps_ch_t* chans) { gps_pos_solve_init(chans); }

/* sdrobs2obsd + gps_pos_solve until solving_is_busy() drops (at most max_calls); returns the number of calls made */
uint32_t ref_fix_run(gps_ch_t* chans, uint32_t max_calls)
{
    uint32_t calls = 0;
    sdrobs2obsd(chans, GPS_SAT_CNT, g_ref_obsd);
    do { gps_pos_solve(g_ref_obsd); calls++; } while (solving_is_busy() && calls < max_calls);
    return calls;
}

/* exactly n calls of gps_pos_solve on the channels' current observations (a solve left in flight) */
void ref_fix_steps(gps_ch_t* chans, uint32_t n)
{
    sdrobs2obsd(chans, GPS_SAT_CNT, g_ref_obsd);
    for (uint32_t k = 0; k < n; k++) gps_pos_solve(g_ref_obsd);
}

/* the reference's one-shot pntpos (solving.c:153) into gps_sol, then the conversion gps_pos_solve would do */
extern nav_t nav_data;
int ref_fix_once(gps_ch_t* chans)
{
    sdrobs2obsd(chans, GPS_SAT_CNT, g_ref_obsd);
    int ok = pntpos(g_ref_obsd, GPS_SAT_CNT, &nav_data, &gps_sol);
    if (ok) {
        ecef2pos(gps_sol.rr, final_pos);
        final_pos[0] = final_pos[0] * R2D;
        final_pos[1] = final_pos[1] * R2D;
    }
    return ok;
}

/* the solver's global outputs back to zero (its function statics cannot be reached) */
void ref_fix_clear(void)
{
    memset(&gps_sol, 0, sizeof gps_sol);
    memset(final_pos, 0, sizeof final_pos);
    memset(azel, 0, 2 * MAXSAT * sizeof(double));
}

void ref_fix_set_start(const double rr3[3]) { for (int i = 0; i < 3; i++) gps_sol.rr[i] = rr3[i]; }

void ref_fix_state(gpsb_flat_fix* o)
{
    memset(o, 0, sizeof *o);
    o->stat = gps_sol.stat; o->ns = gps_sol.ns; o->type = gps_sol.type; o->busy = solving_is_busy();
    o->time_time = (int64_t)gps_sol.time.time;
    o->time_sec_bits = d2u(gps_sol.time.sec);
    for (int k = 0; k < 6; k++) { o->rr[k] = d2u(gps_sol.rr[k]); o->qr[k] = f2u(gps_sol.qr[k]); }
    o->dtr0 = d2u(gps_sol.dtr[0]);
    for (int k = 0; k < 3; k++) o->final_pos[k] = d2u(final_pos[k]);
    for (int k = 0; k < 2 * MAXSAT; k++) o->azel[k] = d2u(azel[k]);
}

void ref_obsd(gps_ch_t* chans, uint8_t* out, uint32_t bytes)
{
    sdrobs2obsd(chans, GPS_SAT_CNT, g_ref_obsd);
    memcpy(out, g_ref_obsd, bytes < sizeof g_ref_obsd ? bytes : sizeof g_ref_obsd);
}
uint32_t ref_sizeof_obsd(void) { return (uint32_t)sizeof(obsd_t); }
uint32_t ref_sizeof_sol(void) { return (uint32_t)sizeof(sol_t); }

/* ------------------------------------------------------------------ RTCM frames (obs_publish.c via ref_rtcm_unit.c, RTK/rtcm3e.c) */
#include "obs_publish.h"
static uint8_t g_uart_frame[1200];
static uint32_t g_uart_bytes = 0, g_uart_calls = 0;
uint8_t uart_prim_dma_send_data(uint8_t* data, uint16_t size)      /* the UART the frames go out on (uart_comm.h:16) */
{
    g_uart_bytes = size;
    g_uart_calls++;
    memcpy(g_uart_frame, data, size < sizeof g_uart_frame ? size : sizeof g_uart_frame);
    return 0;
}
uint8_t uart_prim_is_busy(void) { return 0; }

static uint32_t take_frame(uint8_t* out, uint32_t cap)
{
    const uint32_t n = g_uart_bytes < cap ? g_uart_bytes : cap;
    memcpy(out, g_uart_frame, n);
    return g_uart_bytes;
}
/* sdrobs2obsd + sendrtcmobs (message 1075); returns the frame length */
uint32_t ref_rtcm_obs(gps_ch_t* chans, uint8_t* out, uint32_t cap)
{
    g_uart_bytes = 0;
    sdrobs2obsd(chans, GPS_SAT_CNT, g_ref_obsd);
    sendrtcmobs(g_ref_obsd, GPS_SAT_CNT);
    return take_frame(out, cap);
}
/* the same on caller-made observation records (n <= GPS_SAT_CNT) */
uint32_t ref_rtcm_obs_records(const uint8_t* records, uint32_t n, uint8_t* out, uint32_t cap)
{
    g_uart_bytes = 0;
    memcpy(g_ref_obsd, records, n * sizeof(obsd_t));
    sendrtcmobs(g_ref_obsd, (int)n);
    return take_frame(out, cap);
}
/* sendrtcmnav (message 1019) */
uint32_t ref_rtcm_nav(gps_ch_t* ch, uint8_t* out, uint32_t cap)
{
    g_uart_bytes = 0;
    sendrtcmnav(ch);
    return take_frame(out, cap);
}
