/*
 * gpsb_flat_state.h - fixed-layout snapshot of one receiver channel.
 *
 * The reference keeps all per-satellite state in gps_ch_t
 * (Firmware/project_main/GPS/gps_misc.h:43-133,184-193).  Parity tests need to compare that state
 * between the reference build (oracle/_ref) and this library without depending on either side's
 * struct padding, so both sides export the same flat, explicitly-sized record.  Floats are carried
 * as their IEEE-754 bit patterns so the comparison is bit-exact.
 *
 * Field names follow the reference struct members they mirror.
 */
#ifndef GPSB_FLAT_STATE_H
#define GPSB_FLAT_STATE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSB_FLAT_HIST_SIZE       32   /* ACQ_PHASE1_HIST_SIZE, config.h:48          */
#define GPSB_FLAT_PRETRACK_POINTS 30   /* PRE_TRACK_POINTS_MAX_CNT, config.h:50       */
#define GPSB_FLAT_SLOT_LEN        4    /* TRACKING_CH_LENGTH, config.h:56             */
#define GPSB_FLAT_WORD_BITS       30   /* GPS_NAV_WORD_LENGTH, gps_misc.h:11          */
#define GPSB_FLAT_SUBFRAME_BYTES  38   /* GPS_NAV_SUBFRAME_LENGTH_BYTES, gps_misc.h:12 */

typedef struct gpsb_flat_state {
    /* ---- gps_ch_t ---- */
    uint32_t prn;

    /* ---- gps_acq_t (gps_misc.h:43-60) ---- */
    uint32_t acq_state;
    uint32_t freq_index;
    int32_t  found_freq_offset_hz;
    int32_t  given_freq_offset_hz;
    uint32_t found_code_phase;
    uint32_t acq_code_search_start;
    uint32_t acq_code_search_stop;
    uint32_t code_hist_step;
    uint32_t acq_start_timestamp;
    uint32_t hist_ratio_bits;
    uint8_t  code_phase_histogram[GPSB_FLAT_HIST_SIZE];

    /* ---- gps_tracking_t (gps_misc.h:62-99) ---- */
    uint32_t trk_state;
    uint32_t trk_code_search_start;
    uint32_t trk_code_search_stop;
    uint32_t if_freq_offset_hz_bits;
    uint32_t if_freq_accum;
    uint32_t pre_track_count;
    uint32_t prev_track_timestamp;
    uint32_t code_phase_fine_bits;
    uint32_t old_code_phase_fine_bits;
    uint32_t code_phase_swap_flag;
    uint32_t dll_code_err_bits;
    uint32_t pll_code_err_bits;
    int32_t  fll_old_i;
    int32_t  fll_old_q;
    uint32_t fll_err_bits;
    uint32_t pll_bad_state_cnt;
    uint32_t pll_bad_state_master_cnt;
    uint32_t i_part_summ;
    uint32_t q_part_summ;
    uint32_t snr_summ_cnt;
    uint32_t snr_value_bits;
    uint32_t filt_start_time_ms;
    uint32_t code_filt_cnt;
    uint32_t code_phase_fine_filt_bits;
    uint16_t pre_track_phases[GPSB_FLAT_PRETRACK_POINTS];
    int16_t  pll_check_buf[GPSB_FLAT_SLOT_LEN];

    /* ---- gps_nav_data_t (gps_misc.h:101-133) ---- */
    uint32_t period_sync_ok_flag;
    uint32_t right_period_cnt;
    uint32_t old_swap_time;
    uint32_t old_reminder;
    uint32_t accurate_swap_time;
    uint32_t accurate_swap_ok;
    uint32_t last_bit_pos_cnt;
    uint32_t last_bit_neg_cnt;
    uint32_t inv_polarity_flag;
    uint32_t polarity_found;
    uint32_t inv_preabmle_cnt;
    uint32_t word_cnt;
    uint32_t word_bit_cnt;
    uint32_t old_D29;
    uint32_t old_D30;
    uint32_t word_detection_timestamp;
    uint32_t word_cnt_test;
    uint32_t last_subframe_time;
    uint32_t first_subframe_time;
    uint32_t subframe_cnt;
    uint32_t new_subframe_flag;
    uint8_t  word_buf[GPSB_FLAT_WORD_BITS];
    uint8_t  subframe_data[GPSB_FLAT_SUBFRAME_BYTES];
} gpsb_flat_state;

/* Layout-independent image of the channel's ephemeris container (sdreph_t / eph_t, PM/GPS/gps_misc.h:135-182), filled by
 * gps_nav_data_decode_subframe (PM/GPS/nav_data_decode.c:33): doubles as their bit patterns. */
typedef struct gpsb_flat_eph {
    int32_t  sat, iode, iodc, sva, svh, week, code, flag;
    int64_t  toe_time, toc_time, ttr_time;
    uint64_t toe_sec_bits, toc_sec_bits, ttr_sec_bits;
    uint64_t A, e, i0, OMG0, omg, M0, deln, OMGd, idot;          /* bit patterns of the doubles */
    uint64_t crc, crs, cuc, cus, cic, cis;
    uint64_t toes, fit, f0, f1, f2;
    uint64_t tgd[4];
    int32_t  ctype;
    int32_t  week_gpst, cnt, cntth, update, prn, week_gst;
    uint32_t sub_cnt, received_mask, received_mask_proc;
    uint64_t tow_gpst;
} gpsb_flat_eph;

/* Position solver outputs (RTK/solving.c: gps_sol, final_pos, azel) as bit patterns. */
#define GPSB_FLAT_FIX_SATS 32
typedef struct gpsb_flat_fix {
    int32_t  stat, ns, type, busy;
    int64_t  time_time;
    uint64_t time_sec_bits;
    uint64_t rr[6];
    uint32_t qr[6];
    uint64_t dtr0;
    uint64_t final_pos[3];                                       /* latitude, longitude (deg), height (m) */
    uint64_t azel[2 * GPSB_FLAT_FIX_SATS];                       /* the reference holds 2 x 4 */
} gpsb_flat_fix;

#ifdef __cplusplus
}
#endif

#endif /* GPSB_FLAT_STATE_H */
