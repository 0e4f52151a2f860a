/*
 * master.c - sequencing of acquisition and tracking over the receiver's channels.
 *
 * Behaviour follows Firmware/project_main/GPS/gps_master.c: the sequencing half (:68-156, helpers :453-510) and the
 * observation half - subframe-time bookkeeping, code-phase filter and pseudorange / time-of-week assembly
 * (gps_master_nav_handling, :159-388; row N3 of SURVEY.md section 8(f)).  The position fix and the RTCM frames that follow
 * the observations in that file live in fix.c and rtcm.c (row N4); its terminal output and key handler are not part
 * of this library.
 */
#include <math.h>

#include "host_internal.h"

#define HX_SUBFRAME_MS      6000                    /* SUBFRAME_DURATION_MS, gps_master.c:34 */
#define HX_OFFSET_TIME_MS   (68.802)                /* GPS_OFFSET_TIME_MS, gps_master.c:31: nominal time of flight */
#define HX_M_PER_MS         (299792458.0 / PRN_SPEED_HZ)   /* CLIGHT_NORM, gps_master.c:33 */

static uint8_t g_need_acq = 1;       /* gps_common_need_acq, gps_master.c:44 */
static uint8_t g_first_call = 1;     /* gps_start_flag, gps_master.c:46 */

void gps_master_handling(gps_ch_t* ch, uint8_t index)
{
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
    if (index == 0xFF) gps_master_nav_handling(ch);          /* the idle slot of the 17-ms schedule, gps_master.c:144-146 */
}

#if (ENABLE_CODE_FILTER)
/* gps_master.c:378-388 */
static void code_filter_restart(gps_ch_t* ch, uint32_t n, uint32_t now)
{
    for (uint32_t i = 0; i < n; i++) {
        ch[i].tracking_data.code_phase_fine_filt = 0.0f;
        ch[i].tracking_data.code_filt_cnt = 0;
        ch[i].tracking_data.filt_start_time_ms = now;
    }
}

/* gps_master.c:334-376: once every channel has gathered more than CODE_FILTER_LENGTH code-phase samples without a
 * wrap (the DLL marks a wrap with a negative sum) inside one second, turn the sums into means.  Returns the span of
 * the averaging window in ms, 0 when there is nothing to use yet. */
static uint16_t code_filter_close(gps_ch_t* ch, uint32_t n, uint32_t now)
{
    uint32_t ready = 0, wrapped = 0;
    for (uint32_t i = 0; i < n; i++) ready += ch[i].tracking_data.code_filt_cnt > CODE_FILTER_LENGTH;
    if (ready < n) return 0;
    for (uint32_t i = 0; i < n; i++) wrapped += ch[i].tracking_data.code_phase_fine_filt < -0.5f;
    const uint32_t span = now - ch[0].tracking_data.filt_start_time_ms;
    if (wrapped || span > 1000) {
        code_filter_restart(ch, n, now);
        return 0;
    }
    for (uint32_t i = 0; i < n; i++)
        ch[i].tracking_data.code_phase_fine_filt = ch[i].tracking_data.code_phase_fine_filt / ch[i].tracking_data.code_filt_cnt;
    return (uint16_t)span;
}
#endif

/* gps_master.c:289-327.  since_ref_ms arrives as uint32_t exactly like the reference's parameter: a negative value
 * (possible after half the filter window is taken off) wraps, and the time of week inherits that - reproduced. */
static void assemble_observations(gps_ch_t* ch, uint32_t n, uint32_t since_ref_ms, uint32_t epoch_ms, uint32_t ref)
{
    for (uint32_t i = 0; i < n; i++) {
        const int32_t whole_ms = (int32_t)(ch[i].nav_data.last_subframe_time - epoch_ms);
#if (ENABLE_CODE_FILTER)
        const float fine = ch[i].tracking_data.code_phase_fine_filt;
#else
        const float fine = ch[i].tracking_data.code_phase_fine;
#endif
        double flight_ms = (double)whole_ms + fine / ((double)PRN_LENGTH * 16.0f);
        if (ch[i].tracking_data.code_phase_swap_flag == 1)          /* the code epoch wrapped but the subframe stamp has not yet */
            flight_ms = flight_ms - (ch[i].tracking_data.if_freq_offset_hz < 0.0f ? -1.0 : 1.0);
        ch[i].obs_data.pseudorange_m = (HX_OFFSET_TIME_MS + flight_ms) * HX_M_PER_MS;
        ch[i].obs_data.tow_s = ch[ref].eph_data.tow_gpst +
                               ((float)(since_ref_ms + (uint8_t)i * TRACKING_CH_LENGTH) / PRN_SPEED_HZ);
    }
}

/* gps_master.c:159-287 */
void gps_master_nav_handling(gps_ch_t* ch)
{
    if (!ch) return;
    const uint32_t n = gpsb_host_sat_cnt();
    uint32_t stamped = 0, unlocked = 0, ref = 0;
    uint32_t t_min = 0xFFFFFFFFu, t_max = 0;
    uint16_t c_max = 0;
    for (uint32_t i = 0; i < n; i++) {
        const gps_nav_data_t* nv = &ch[i].nav_data;
        stamped += nv->last_subframe_time != 0;
        unlocked += nv->first_subframe_time == 0;
        if (nv->last_subframe_time < t_min) { t_min = nv->last_subframe_time; ref = i; }   /* earliest = nearest satellite */
        if (nv->last_subframe_time > t_max) t_max = nv->last_subframe_time;
        if (nv->subframe_cnt > c_max) c_max = nv->subframe_cnt;
    }
    if (t_min == 0) return;
    if (t_max - t_min > 100) return;             /* this epoch's subframes have not all arrived yet */

    if (stamped == n && unlocked == n) {         /* once: the zero moment of every channel */
        for (uint32_t i = 0; i < n; i++) {
            ch[i].nav_data.first_subframe_time = ch[i].nav_data.last_subframe_time;
            ch[i].nav_data.subframe_cnt = 0;
        }
    }
    if (ch[0].nav_data.first_subframe_time == 0) return;

    /* NB: the count was taken before the zero moment cleared it, like the reference */
    const uint32_t epoch_ms = ch[ref].nav_data.first_subframe_time + (uint32_t)c_max * HX_SUBFRAME_MS;

    for (uint32_t i = 0; i < n; i++) {           /* a code-epoch wrap shows as a jump of more than half the range */
        gps_tracking_t* t = &ch[i].tracking_data;
        if (t->code_phase_swap_flag && ch[i].nav_data.new_subframe_flag) {
            ch[i].nav_data.new_subframe_flag = 0;
            t->code_phase_swap_flag = 0;
        }
        const float jump = (float)fabs(t->old_code_phase_fine - t->code_phase_fine);
        if (jump > ((float)PRN_LENGTH * 16.0f / 2.0f)) t->code_phase_swap_flag = 1;
        t->old_code_phase_fine = t->code_phase_fine;
    }

    const uint32_t now = signal_capture_get_packet_cnt();
    int32_t since_ref_ms = (int32_t)now - (int32_t)ch[ref].nav_data.last_subframe_time;
    if (since_ref_ms < 0) since_ref_ms = since_ref_ms % HX_SUBFRAME_MS;
    int usable = 1;
#if (ENABLE_CODE_FILTER)
    const uint16_t window_ms = code_filter_close(ch, n, now);
    if (window_ms < 1) usable = 0;
    since_ref_ms = since_ref_ms - window_ms / 2;          /* the mean belongs to the middle of the window */
#endif
    if (usable) {
        assemble_observations(ch, n, (uint32_t)since_ref_ms, epoch_ms, ref);
#if (ENABLE_CODE_FILTER)
        code_filter_restart(ch, n, now);
#endif
    }
    if (gpsb_host_rtcm_enabled()) gps_master_transmit_obs(ch);   /* gps_master.c:279-281; rtcm.c */
    gps_master_calculate_pos(ch);                          /* gps_master.c:283-285; fix.c */
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
    gpsb_host_fix_reset();
}
