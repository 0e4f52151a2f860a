/* host_internal.h - private declarations of libgpsb_host.so (see include/gpsb_host.h). */
#ifndef GPSB_HOST_INTERNAL_H
#define GPSB_HOST_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gpsb_host.h"

#define GPSB_HALF_CHIPS        (2 * PRN_LENGTH)          /* 2046 code phases, acquisition.c:294 */
#define GPSB_FINE_PER_HALFCHIP 8                          /* GPS_FINE_RATIO, tracking.c:23 */
#define GPSB_FINE_RANGE        (GPSB_HALF_CHIPS * GPSB_FINE_PER_HALFCHIP)   /* 16368 */
#define GPSB_SLOT_LEN          TRACKING_CH_LENGTH
#define GPSB_FREQ_POINTS_MAX   25                         /* FREQ_SEARCH_POINTS_MAX_CNT, acquisition.c:12 */
#define GPSB_MAX_BINS          64

/* Cross-call scratch that the reference keeps in file-scope variables.  One shared instance backs
 * the reference-named API; the batched receiver owns one per channel. */
typedef struct gpsb_aux {
    uint32_t freq_hist[GPSB_MAX_BINS];          /* acq_freq_histogram, acquisition.c:28 (ACQ_COUNT used) */
    uint16_t bin_phases[GPSB_FREQ_POINTS_MAX];  /* acq_single_freq_phases, acquisition.c:32 */
    uint8_t  bin_count;                         /* acq_single_freq_count, acquisition.c:33 */
    uint16_t pre_best_value;                    /* pre_track_best_corr_value, tracking.c:33 */
    uint16_t pre_best_phase;                    /* pre_track_best_corr_phase, tracking.c:34 */
    int16_t  slot_ip[GPSB_SLOT_LEN];            /* raw_ip_values, nav_data.c:48 */
    uint8_t  slot_bits[GPSB_SLOT_LEN];          /* tmp_nav_data, nav_data.c:51 */
    uint32_t slot_start_ticks;                  /* gps_channel_tmp_start_time_ticks, nav_data.c:29 */
    int8_t   last_nav_bit;                      /* observer: bit handed to the word assembler this ms, or -1 */
    /* private rand() stream of a batched channel: same generator and default seed as libc rand(), so the
     * channel draws what it would draw as the only channel of a process (tracking.c:316) */
    int      rnd_ready;
    struct random_data rnd;
    char     rnd_state[128];
} gpsb_aux;

extern gpsb_aux g_shared_aux;

uint32_t hx_now_ms(void);
int hx_rand(gpsb_aux* aux);
uint32_t hx_nco_step(float freq_hz);
uint32_t hx_nco_step32(float freq_hz);

/* acquisition (acq.c) */
void hx_acq_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, gpsb_plan* plan);
void hx_acq_finish(gps_ch_t* ch, gpsb_aux* aux, const gpsb_plan* plan, const gpsb_search_res* res);
uint8_t hx_chain_vote(uint16_t* phases, uint8_t n, uint16_t* chain_phase);
void hx_freq_hist_decide(gps_ch_t* ch, const uint32_t* hist, uint32_t n_bins, int32_t first_bin_hz, int32_t step_hz);

/* tracking (track.c) */
void hx_trk_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, uint8_t index, gpsb_plan* plan);
void hx_trk_finish_search(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const gpsb_search_res* res);
void hx_trk_finish_epl(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const int16_t iq[6]);

/* nav bits (nav.c) */
void hx_nav_new_code(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t new_i);
void hx_nav_word_bit(gps_ch_t* ch, uint8_t new_bit);

/* binding (bind.c) */
int hx_note(int status);
int hx_stage_frame(const uint8_t* data, uint32_t* frame_ms);

#endif
