/*
 * master.c - sequencing of acquisition and tracking over the receiver's channels.
 *
 * Behaviour follows the sequencing half of Firmware/project_main/GPS/gps_master.c:68-156 and the
 * helpers at :453-510.  The other half of that file (pseudorange assembly, position solver, RTCM,
 * terminal output, the key handler) is outside the correlator hot path and is
 * not part of this library; gps_master_handling() therefore stops after the state sequencing.
 */
#include "host_internal.h"

static uint8_t g_need_acq = 1;       /* gps_common_need_acq, gps_master.c:44 */
static uint8_t g_first_call = 1;     /* gps_start_flag, gps_master.c:46 */

void gps_master_handling(gps_ch_t* ch, uint8_t index)
{
    (void)index;
    if (!ch) return;
    const uint32_t n = gpsb_host_sat_cnt();
    if (g_first_call) {
        g_first_call = 0;
        acquisition_start_channel(&ch[0]);
    }

    uint8_t doppler_pending = 0, round2_done = 0;
    g_need_acq = 0;
    for (uint32_t i = 0; i < n; i++) {
        gps_acq_state_t s = ch[i].acq_data.state;
        if (s != GPS_ACQ_DONE) g_need_acq = 1;
        if (s < GPS_ACQ_FREQ_SEARCH_DONE) doppler_pending = 1;
        if (s == GPS_ACQ_CODE_PHASE_SEARCH2_DONE) round2_done++;
    }

    /* Doppler searches run one channel after another (they share the vote buffers) */
    if (g_need_acq) {
        for (uint32_t i = 0; i + 1 < n; i++) {
            if (ch[i].acq_data.state == GPS_ACQ_FREQ_SEARCH_DONE &&
                ch[i + 1].acq_data.state == GPS_ACQ_NEED_FREQ_SEARCH) {
                acquisition_start_channel(&ch[i + 1]);
                return;
            }
        }
    }
    /* once every Doppler is known the code searches of all channels run side by side; round 3 starts
     * for everyone at the same snapshot */
    if (!doppler_pending && g_need_acq) {
        for (uint32_t i = 0; i < n; i++) {
            if (ch[i].acq_data.state == GPS_ACQ_FREQ_SEARCH_DONE) acquisition_start_code_search_channel(&ch[i]);
            if (round2_done == n) acquisition_start_code_search3_channel(&ch[i]);
        }
    }
    if (!g_need_acq) {
        for (uint32_t i = 0; i < n; i++)
            if (ch[i].tracking_data.state == GPS_TRACKNG_IDLE) ch[i].tracking_data.state = GPS_NEED_PRE_TRACK;
    }
}

uint8_t gps_master_need_acq(void) { return g_need_acq; }

uint8_t gps_master_need_freq_search(gps_ch_t* ch)                       /* gps_master.c:453-463 */
{
    uint8_t pending = 0;
    for (uint32_t i = 0; ch && i < gpsb_host_sat_cnt(); i++)
        if (ch[i].acq_data.state < GPS_ACQ_FREQ_SEARCH_DONE) pending = 1;
    return pending;
}

uint8_t gps_master_is_code_search3(gps_ch_t* ch)                        /* gps_master.c:466-476 */
{
    uint32_t past_round2 = 0, n = gpsb_host_sat_cnt();
    for (uint32_t i = 0; ch && i < n; i++)
        if (ch[i].acq_data.state > GPS_ACQ_CODE_PHASE_SEARCH2) past_round2++;
    return past_round2 == n;
}

/* gps_master.c:490-510: back to the start of the code search, keeping a Doppler that proved itself */
void gps_master_reset_to_aqc_start(gps_ch_t* ch)
{
    if (!ch) return;
    const uint32_t n = gpsb_host_sat_cnt();
    for (uint32_t i = 0; i < n; i++)
        if (ch[i].acq_data.state < GPS_ACQ_FREQ_SEARCH_DONE) return;
    for (uint32_t i = 0; i < n; i++) {
        if (ch[i].nav_data.word_cnt_test > 1)
            ch[i].acq_data.found_freq_offset_hz = (int16_t)ch[i].tracking_data.if_freq_offset_hz;
        ch[i].acq_data.state = GPS_ACQ_FREQ_SEARCH_DONE;
        memset(&ch[i].tracking_data, 0, sizeof ch[i].tracking_data);
        memset(&ch[i].nav_data, 0, sizeof ch[i].nav_data);
    }
}

/* Restart the sequencing (tests and re-runs inside one process). */
void gpsb_host_master_reset(void)
{
    g_need_acq = 1;
    g_first_call = 1;
}
