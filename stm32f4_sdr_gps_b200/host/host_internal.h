/* host_internal.h - private declarations of libgpsb_host.so (see include/gpsb_host.h). */
#ifndef GPSB_HOST_INTERNAL_H
#define GPSB_HOST_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gpsb_host.h"
#include "../core/gpsb_loop_core.h"

#define GPSB_HALF_CHIPS        LC_HALF_CHIPS             /* 2046 code phases, acquisition.c:294 */
#define GPSB_FINE_PER_HALFCHIP LC_FINE_PER_HALFCHIP      /* GPS_FINE_RATIO, tracking.c:23 */
#define GPSB_FINE_RANGE        LC_FINE_RANGE             /* 16368 */
#define GPSB_SLOT_LEN          LC_SLOT_LEN
#define GPSB_FREQ_POINTS_MAX   LC_FREQ_POINTS_MAX        /* FREQ_SEARCH_POINTS_MAX_CNT, acquisition.c:12 */
#define GPSB_MAX_BINS          LC_MAX_BINS

/* gpsb_aux (the cross-call scratch the reference keeps in file-scope variables) is defined in the core. */
extern gpsb_aux g_shared_aux;

uint32_t hx_now_ms(void);
int hx_rand(gpsb_aux* aux);
uint32_t hx_nco_step(float freq_hz);
uint32_t hx_nco_step32(float freq_hz);

/* acquisition (acq.c) */
void hx_acq_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, gpsb_plan* plan);
void hx_acq_finish(gps_ch_t* ch, gpsb_aux* aux, const gpsb_plan* plan, const gpsb_search_res* res);
void hx_acq_start_code_search3(gps_ch_t* ch, gpsb_aux* aux);
uint8_t hx_chain_vote(uint16_t* phases, uint8_t n, uint16_t* chain_phase);
void hx_freq_hist_decide(gps_ch_t* ch, const uint32_t* hist, uint32_t n_bins, int32_t first_bin_hz, int32_t step_hz);

/* tracking (track.c) */
void hx_trk_plan(gps_ch_t* ch, gpsb_aux* aux, uint32_t frame_ms, uint8_t index, gpsb_plan* plan);
void hx_trk_finish_search(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const gpsb_search_res* res);
void hx_trk_finish_epl(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const int16_t iq[6]);

/* nav bits (nav.c) */
void hx_nav_new_code(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t new_i);
void hx_nav_word_bit(gps_ch_t* ch, uint8_t new_bit);

/* observations shared by the position solver and the RTCM publisher (fix.c), gps_master.c:41 */
extern obsd_t hx_obsd[];

/* binding (bind.c) */
int hx_note(int status);
int hx_stage_frame(const uint8_t* data, uint32_t* frame_ms);

#endif
