/*
 * gpsb_host.h - host-side C mirror of the reference receiver API for the acquisition / tracking hot
 * path (libgpsb_host.so), running its correlations on the B200 through include/gpsb.h.
 *
 * Reference = iliasam/STM32F4_SDR_GPS, Firmware/project_main ("PM/").  The reference has no plugin
 * interface; its seam is a set of plain C headers.  This library exports
 *
 *   (1) the SAME function names and signatures as PM/GPS/acquisition.h:7-12, PM/GPS/tracking.h:6,
 *       PM/GPS/gps_master.h:7-15 and PM/GPS/gps_misc.h:195-216, operating on the SAME channel record
 *       (gps_ch_t, PM/GPS/gps_misc.h:184-193), so that a build of the reference firmware logic can link
 *       against it instead of acquisition.c / tracking.c / gps_misc.c / nav_data.c / gps_master.c;
 *   (2) batched forms (gpsb_rx_*) that evaluate ALL channels of a millisecond - or a whole cold-start
 *       sweep - in one GPU launch; these are what a B200 deployment calls.
 *
 *   (3) the steps that follow the path in the reference: subframe decode, observation
 *       assembly, position fix and RTCM frames under the reference's names (PM/GPS/nav_data_decode.h,
 *       RTK/solving.h, obs_publish.h) - plain host C, bit-exact against the compiled reference.
 *
 * What runs where: every XOR/popcount correlation runs on the GPU (no CPU correlator exists in this
 * library; without a CUDA device the calls fail and report through gpsb_host_last_status()).  The
 * loop filters, votes and nav-bit logic are scalar float/integer code that feeds the next step's NCO
 * words: ONE source (core/gpsb_loop_core.h) compiled for the host - used by the reference-named per-call
 * entry points, exactly as the reference does - and for the device, where gpsb_rx_track_run keeps whole
 * runs resident in one kernel launch; both are bit-identical to the reference.
 *
 * The reference keeps several pieces of cross-call state in file-scope globals that only work under
 * its one-channel-at-a-time schedule (PM/GPS/acquisition.c:28-33, tracking.c:33-34, nav_data.c:29,
 * 48-51).  The reference-named entry points below keep ONE such shared set, like the reference; the
 * gpsb_rx_* batched entry points keep one set PER CHANNEL, which equals the reference run with a single
 * active channel.
 */
#ifndef GPSB_HOST_H
#define GPSB_HOST_H

#include <stdint.h>
#include <time.h>

#include "gpsb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Receiver constants (PM/config.h).  Names are kept so reference code compiles against this header.
 * ---------------------------------------------------------------------------------------------- */
#ifndef _CONFIG_H
#define IF_FREQ_HZ                 ((int)4092000)              /* config.h:23 */
#define SPI_BAUDRATE_HZ            ((int)16368000)             /* config.h:24 */
#define PRN_SPEED_HZ               1000                        /* config.h:25 */
#define BITS_IN_PRN                (SPI_BAUDRATE_HZ / PRN_SPEED_HZ)
#define PRN_SPI_WORDS_CNT          (BITS_IN_PRN / 16)
#define PRN_LENGTH                 1023
#define ENABLE_CODE_FILTER         1                           /* config.h:36 */
#define CODE_FILTER_LENGTH         100
#define ACQ_SEARCH_FREQ_HZ         (7000)                      /* config.h:41 */
#define ACQ_SEARCH_STEP_HZ         (500)
#define ACQ_COUNT                  (ACQ_SEARCH_FREQ_HZ * 2 / ACQ_SEARCH_STEP_HZ + 1)
#define ACQ_PHASE1_HIST_STEP       (64)
#define ACQ_PHASE1_HIST_SIZE       ((PRN_LENGTH + 1) * 2 / ACQ_PHASE1_HIST_STEP)
#define PRE_TRACK_POINTS_MAX_CNT   30
#define IF_NCO_STEP_HZ             (0.003810972f)              /* config.h:53 */
#define TRACKING_CH_LENGTH         4
#define GPS_SAT_CNT                4                           /* default; see gpsb_host_set_sat_cnt */
#define TRACKING_DLL1_C1           (1.0f)
#define TRACKING_DLL1_C2           (300.0f)
#define TRACKING_PLL1_C1           (4.0f)
#define TRACKING_PLL1_C2           (3000.0f)
#define TRACKING_PLL2_C1           (8.0f)
#define TRACKING_PLL2_C2           (5000.0f)
#define TRACKING_FLL1_C1           (200.0f)
#define TRACKING_FLL1_C2           (2000.0f)
#endif

/* ------------------------------------------------------------------------------------------------
 * Channel record.  Field order, names and types reproduce PM/GPS/gps_misc.h:20-193 because the
 * record IS the interface: callers own it, fill prn / given_freq_offset_hz, and read results out of it.
 * If the reference header was included first its definitions are used as they are.
 * ---------------------------------------------------------------------------------------------- */
#ifndef _GPS_MISC_H
#define GPS_NAV_WORD_LENGTH            30
#define GPS_NAV_SUBFRAME_LENGTH_BYTES  38

typedef enum {
    GPS_ACQ_NEED_FREQ_SEARCH = 0, GPS_ACQ_FREQ_SEARCH_RUN, GPS_ACQ_FREQ_SEARCH_DONE,
    GPS_ACQ_CODE_PHASE_SEARCH1, GPS_ACQ_CODE_PHASE_SEARCH1_DONE,
    GPS_ACQ_CODE_PHASE_SEARCH2, GPS_ACQ_CODE_PHASE_SEARCH2_DONE,
    GPS_ACQ_CODE_PHASE_SEARCH3, GPS_ACQ_CODE_PHASE_SEARCH3_DONE,
    GPS_ACQ_DONE,
} gps_acq_state_t;

typedef enum {
    GPS_TRACKNG_IDLE, GPS_NEED_PRE_TRACK, GPS_PRE_TRACK_RUN, GPS_PRE_TRACK_DONE, GPS_TRACKING_RUN,
} gps_tracking_state_t;

typedef struct {
    uint8_t  freq_index;              /* Doppler bin under test, 0 == -ACQ_SEARCH_FREQ_HZ         */
    int16_t  found_freq_offset_hz;
    int16_t  given_freq_offset_hz;    /* non-zero: skip the Doppler search                        */
    uint16_t found_code_phase;        /* half chips, 0..2046                                      */
    uint16_t code_search_start;
    uint16_t code_search_stop;
    uint16_t code_hist_step;
    gps_acq_state_t state;
    uint8_t  code_phase_histogram[ACQ_PHASE1_HIST_SIZE];
    uint32_t start_timestamp;
    float    hist_ratio;
} gps_acq_t;

typedef struct {
    uint16_t code_search_start;
    uint16_t code_search_stop;
    float    if_freq_offset_hz;
    uint32_t if_freq_accum;
    uint16_t pre_track_phases[PRE_TRACK_POINTS_MAX_CNT];
    uint8_t  pre_track_count;
    uint32_t prev_track_timestamp;
    float    code_phase_fine;         /* samples (1/16 chip), 0..16368                            */
    float    old_code_phase_fine;
    uint8_t  code_phase_swap_flag;
    float    dll_code_err;
    float    pll_code_err;
    int16_t  fll_old_i;
    int16_t  fll_old_q;
    float    fll_err;
    int16_t  pll_check_buf[TRACKING_CH_LENGTH];
    uint8_t  pll_bad_state_cnt;
    uint16_t pll_bad_state_master_cnt;
    uint32_t i_part_summ;
    uint32_t q_part_summ;
    uint16_t snr_summ_cnt;
    float    snr_value;
#if (ENABLE_CODE_FILTER)
    uint32_t filt_start_time_ms;
    uint16_t code_filt_cnt;
    float    code_phase_fine_filt;
#endif
    gps_tracking_state_t state;
} gps_tracking_t;

typedef struct {
    uint8_t  period_sync_ok_flag;
    uint8_t  right_period_cnt;
    uint32_t old_swap_time;
    uint8_t  old_reminder;
    uint8_t  accurate_swap_time;
    uint8_t  accurate_swap_ok;
    uint8_t  last_bit_pos_cnt;
    uint8_t  last_bit_neg_cnt;
    uint8_t  inv_polarity_flag;
    uint8_t  polarity_found;
    uint8_t  inv_preabmle_cnt;
    uint8_t  word_buf[GPS_NAV_WORD_LENGTH];
    uint8_t  word_cnt;
    uint8_t  word_bit_cnt;
    uint8_t  old_D29;
    uint8_t  old_D30;
    uint32_t word_detection_timestamp;
    uint32_t word_cnt_test;
    uint32_t last_subframe_time;
    uint32_t first_subframe_time;
    uint16_t subframe_cnt;
    uint8_t  new_subframe_flag;
    uint8_t  subframe_data[GPS_NAV_SUBFRAME_LENGTH_BYTES];
} gps_nav_data_t;

typedef struct { double pseudorange_m; double tow_s; } gps_obs_data_t;

/* RTKLIB-derived ephemeris containers (gps_misc.h:143-182): filled by the subframe decode, read by the position fix
 * and the RTCM encoder; the correlator path never touches them. */
typedef struct { time_t time; double sec; } gtime_t;
typedef struct {
    int sat, iode, iodc, sva, svh, week, code, flag;
    gtime_t toe, toc, ttr;
    double A, e, i0, OMG0, omg, M0, deln, OMGd, idot;
    double crc, crs, cuc, cus, cic, cis;
    double toes, fit, f0, f1, f2;
    double tgd[4];
} eph_t;
typedef struct {
    eph_t eph;
    int ctype;
    double tow_gpst;
    int week_gpst, cnt, cntth, update, prn, week_gst;
    uint16_t sub_cnt;
    uint8_t received_mask, received_mask_proc;
} sdreph_t;

typedef struct {
    gps_acq_t      acq_data;
    gps_tracking_t tracking_data;
    gps_nav_data_t nav_data;
    gps_obs_data_t obs_data;
    sdreph_t       eph_data;
    uint8_t        prn;
    uint8_t        prn_code[PRN_LENGTH];
} gps_ch_t;
#endif /* _GPS_MISC_H */

/* Records of the position solver (row N4): PM/GPS/RTK/rtk_common.h:50-59 and solving.h:17-32, same layout. */
#define GPSB_FIX_MAX_SATS 32
#ifndef _RTCM_COMMON_H
typedef struct {
    gtime_t time;                    /* receiver sampling time (GPST) */
    unsigned char sat, rcv;
    unsigned char SNR[1], LLI[1], code[1];
    double L[1];                     /* carrier phase (cycles) - not produced */
    double P[1];                     /* pseudorange (m) */
    float  D[1];                     /* Doppler (Hz) */
} obsd_t;
#endif
#ifndef _GPS_SOLVING_H
#define SOLQ_NONE   0
#define SOLQ_SINGLE 5
typedef struct {
    gtime_t time;                    /* GPST of the fix */
    double rr[6];                    /* ECEF position (m) / velocity (always 0) */
    float  qr[6];                    /* position covariance xx yy zz xy yz zx (m^2) */
    double dtr[6];                   /* receiver clock bias (s) in [0] */
    unsigned char type, stat, ns;    /* stat: SOLQ_NONE / SOLQ_SINGLE */
    float age, ratio;
} sol_t;
#endif

/* ------------------------------------------------------------------------------------------------
 * Binding to a GPU context and to the millisecond clock.
 * ---------------------------------------------------------------------------------------------- */
/* All reference-named calls below run on this context (one per process for the drop-in API).
 * The context must have been created with max_sv >= 211: satellite slot == PRN number. */
int  gpsb_host_attach(gpsb_ctx* ctx);
gpsb_ctx* gpsb_host_context(void);
/* Status of the most recent GPU call made on behalf of a void reference-named function
 * (the reference API has no error channel, PM/GPS/acquisition.c:136, tracking.c:96). */
int  gpsb_host_last_status(void);
/* Number of receiver channels the array forms iterate over (GPS_SAT_CNT, PM/config.h:59). */
void gpsb_host_set_sat_cnt(uint32_t n);
uint32_t gpsb_host_sat_cnt(void);
/* Replaces the SPI/DMA millisecond counter, PM/signal_capture.c:57-82.  The library provides
 * signal_capture_get_packet_cnt() (signal_capture.h:15) as a WEAK symbol backed by this value, so an
 * application that has its own capture layer overrides it by simply defining the function. */
void gpsb_host_set_packet_cnt(uint32_t ms);
uint32_t signal_capture_get_packet_cnt(void);
/* Deterministic replacement of rand() in the false-lock reseed (PM/GPS/tracking.c:316): NULL = libc. */
void gpsb_host_set_rand(int (*fn)(void));

/* ------------------------------------------------------------------------------------------------
 * (1) Reference-named API.
 * ---------------------------------------------------------------------------------------------- */
/* PM/GPS/gps_misc.h:195-196 */
void gps_fill_summ_table(void);                 /* no table is needed on the GPU; kept for link parity */
void gps_channell_prepare(gps_ch_t* channel);   /* fills prn_code[] and loads device slot `prn`      */

/* PM/GPS/acquisition.h:7-12 */
void acquisition_process(gps_ch_t* channel, uint8_t* data);
uint32_t* acquisition_get_hist(void);
void acquisition_start_channel(gps_ch_t* channel);
void acquisition_start_code_search_channel(gps_ch_t* channel);
void acquisition_start_code_search3_channel(gps_ch_t* channel);
void acquisition_process_channel(gps_ch_t* channel, uint8_t* data);   /* acquisition.c:134 */

/* PM/GPS/tracking.h:6 */
void gps_tracking_process(gps_ch_t* channel, uint8_t* data, uint8_t index);

/* PM/GPS/nav_data.h:7 and the non-static word assembler, nav_data.c:257 */
void gps_nav_data_analyse_new_code(gps_ch_t* channel, uint8_t index, int16_t new_i);
void gps_nav_data_words_detection(gps_ch_t* channel, uint8_t new_bit);

/* PM/GPS/nav_data_decode.h: ephemeris / clock fields of a completed subframe (channel->nav_data.subframe_data) into
 * channel->eph_data; returns the subframe id.  Called by the word assembler whenever a subframe completes. */
uint8_t gps_nav_data_decode_subframe(gps_ch_t* channel);

/* PM/GPS/gps_master.h:7-15 (sequencing and, in the idle slot index == 0xFF, the observations, the RTCM frames when enabled
 * and the position fix; no UART or keys here) */
void    gps_master_handling(gps_ch_t* channels, uint8_t index);
/* PM/GPS/gps_master.c:159: subframe-time bookkeeping, code-phase filter, pseudorange and time of week of every channel
 * into channels[i].obs_data (the reference calls it from gps_master_handling's idle slot). */
void    gps_master_nav_handling(gps_ch_t* channels);
uint8_t gps_master_need_acq(void);
uint8_t gps_master_need_freq_search(gps_ch_t* channels);
uint8_t gps_master_is_code_search3(gps_ch_t* channels);
void    gps_master_reset_to_aqc_start(gps_ch_t* channels);

/* PM/GPS/gps_misc.h:198-216 - the DSP primitives with their exact reference signatures, each one a
 * round trip through the GPU (level 0 of include/gpsb.h). */
int16_t  gps_correlation8(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t offset);
void     gps_correlation_iq(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t offset,
                            int16_t* res_i, int16_t* res_q);
uint16_t correlation_search(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t start_shift,
                            uint16_t stop_shift, uint16_t* aver_val, uint16_t* phase);
void     gps_shift_to_zero_freq(uint8_t* signal_data, uint8_t* data_i, uint8_t* data_q, float freq_hz);
void     gps_shift_to_zero_freq_track(gps_tracking_t* trk_channel, uint8_t* signal_data, uint8_t* data_i,
                                      uint8_t* data_q);
void     gps_generate_prn_data2(gps_ch_t* channel, uint16_t* data, uint16_t offset_bits);
void     gps_rewind_if_phase(gps_tracking_t* trk_channel, uint8_t steps);
void     gps_generate_prn(uint8_t* dest, int prn);                       /* gps_misc.c:317 */

/* ------------------------------------------------------------------------------------------------
 * (2) Batched receiver: many channels per launch, signal resident in the context's HBM ring.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gpsb_rx gpsb_rx;

/* channels: caller-owned array of n_ch records (prn set, gps_channell_prepare not required).
 * The receiver borrows the array until gpsb_rx_destroy. */
int  gpsb_rx_create(gpsb_rx** out, gpsb_ctx* ctx, gps_ch_t* channels, uint32_t n_ch);
void gpsb_rx_destroy(gpsb_rx* rx);

/* One millisecond of tracking for every channel (gps_tracking_process for each, PM/main.c:155, in the
 * "every satellite every millisecond" schedule: slot index = ms % 4, or as moved by gpsb_rx_set_slot_walk).  Frame `ms` must already be in
 * the ring (gpsb_upload_signal).  Exactly one k_epl launch (+ one search launch while any channel is
 * still in pre-track). */
int gpsb_rx_track_ms(gpsb_rx* rx, uint32_t ms);
/* n_ms consecutive milliseconds starting at ms0; optional logs, each may be NULL:
 *   iq_log  [n_ms][n_ch][6]  IE,QE,IP,QP,IL,QL (zeros for channels not in GPS_TRACKING_RUN that ms)
 *   nav_log [n_ms][n_ch]     -1, or the 20-ms data bit handed to the word assembler that ms        */
int gpsb_rx_track_run(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, int16_t* iq_log, int8_t* nav_log);
/* Same run with the samples still in HOST memory (n_ms * 2046 bytes at packed, ideally pinned): uploads chunk_ms
 * milliseconds (0 = 64), launches the device-resident loop and streams the rest into the HBM ring while the loop is
 * already running - the B200 form of the reference's capture double buffer, PM/signal_capture.c:57-123. */
int gpsb_rx_track_stream(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* packed, uint32_t chunk_ms,
                         int16_t* iq_log, int8_t* nav_log);
/* Same, fed with the MAX2769-native 2-bit I / 2-bit Q container (one byte per sample, n_ms * 16368 bytes). */
int gpsb_rx_track_stream_iq2(gpsb_rx* rx, uint32_t ms0, uint32_t n_ms, const uint8_t* samples, uint32_t chunk_ms,
                             int16_t* iq_log, int8_t* nav_log);
/* The same run fed from a recording on disk (row N1, host/ingest.c).  The sample of ms0 starts at first_byte.  Default
 * container: the MCU's memory image of the SPI stream (signal_capture.c:9-16), 2046 bytes per millisecond, LSB first -
 * mapped and streamed as it lies.  GPSB_FILE_MSB_FIRST: first sample of each byte in bit 7 (bytes are bit-reversed on
 * the way in).  GPSB_FILE_IQ2: 2-bit I / 2-bit Q, one byte per sample.  GPSB_ERR_ARG when the file is missing or
 * shorter than the run.  gpsb_file_ms: whole milliseconds available from first_byte on (negative status on error). */
#define GPSB_FILE_MSB_FIRST 1u
#define GPSB_FILE_IQ2       2u
int gpsb_rx_track_file(gpsb_rx* rx, const char* path, uint64_t first_byte, uint32_t ms0, uint32_t n_ms, uint32_t flags,
                       int16_t* iq_log, int8_t* nav_log);
int64_t gpsb_file_ms(const char* path, uint64_t first_byte, uint32_t flags);
/* Where gpsb_rx_track_run keeps the loop filters.  AUTO (default) and DEVICE: the whole run is one launch of the
 * device-resident loop k_track_run (include/gpsb.h, gpsb_track_loop) for every channel that is tracking; its
 * float discriminators are the fdlibm atanf/atan2f glibc ships and CUDA's double atan2, checked against the host
 * libm over their whole (finite) input domain by gpsb_host_certify_loop_math().  HOST: one GPU round trip per
 * millisecond with the filters on this machine's libm - for hosts whose libm the certificate rejects. */
enum { GPSB_LOOP_AUTO = 0, GPSB_LOOP_HOST = 1, GPSB_LOOP_DEVICE = 2 };
void gpsb_rx_set_loop_site(gpsb_rx* rx, int site);
/* channel-milliseconds processed so far by the device-resident loop and by the per-millisecond host path */
void gpsb_rx_loop_stats(const gpsb_rx* rx, uint64_t* device_ms, uint64_t* host_ms);
/* Compares the device's Costas and FLL discriminators with THIS host's libm (atan2f, atan2, atanf as called by
 * PM/GPS/tracking.c:180-183,232-233) for every (IP, QP) pair in [-8184, 8184]^2 - 2 x 268 M values - using
 * n_threads host threads (0 = all).  Returns the number of differing bit patterns (0 = the device-resident loop
 * is bit-exact with a host-resident one on this machine), or a negative gpsb_status. */
int64_t gpsb_host_certify_loop_math(gpsb_ctx* ctx, uint32_t n_threads);
/* Host threads used by gpsb_rx_track_run for the per-channel loop filters (each drives its own channels'
 * session slots): 0 = automatic (online CPUs - 1, at most 16, at most one per channel), 1 = single thread. */
void gpsb_rx_set_threads(gpsb_rx* rx, uint32_t n);

/* Slot-phase walk.  The reference's bit synchroniser sees a data-bit edge only inside a 4-ms channel slot and refines
 * only an edge at slot position 2 (PM/GPS/nav_data.c:87-138); subframe time stamps and pseudoranges need that refined
 * edge (nav_data.c:356-360).  The MCU serves a channel 4 of every 17 ms (PM/main.c:134-155), so its slots walk over
 * every edge alignment.  The batched paths process every millisecond - by default with slot index = ms % 4, an
 * alignment that never moves.  With the walk enabled a channel that tracks but has no refined edge leaves 1..3
 * milliseconds out between two slots (the MCU's own means: an unserved channel, PM/GPS/tracking.c:102-113) until its
 * edges show at slot position 2: every satellite then delivers subframe stamps.  Results equal the unmodified
 * reference called on the same (millisecond, slot index) schedule; idle milliseconds log zero sums and no nav bit.
 * period_ms: patience at one slot phase without any bit edge seen, 0 = 400.  Default: off (index = ms % 4). */
int gpsb_rx_set_slot_walk(gpsb_rx* rx, int enable, uint32_t period_ms);
/* Where a channel stands on its way to a time stamp (so a caller can tell "not yet" from "never"). */
typedef struct gpsb_sync_status {
    uint8_t  tracking;            /* in GPS_TRACKING_RUN */
    uint8_t  bit_period_found;    /* nav_data.period_sync_ok_flag */
    uint8_t  bit_edge_refined;    /* nav_data.accurate_swap_ok: subframe stamps can be made */
    uint8_t  polarity_found;
    uint8_t  slot_phase;          /* slot index of millisecond ms is (ms + slot_phase) % 4 */
    uint8_t  walk_enabled, walk_pending;
    uint8_t  reserved;
    uint16_t walks;               /* idle gaps taken so far */
    uint16_t subframes;           /* nav_data.subframe_cnt */
    uint32_t words_ok;            /* nav_data.word_cnt_test */
} gpsb_sync_status;
int gpsb_rx_channel_sync(const gpsb_rx* rx, uint32_t i, gpsb_sync_status* out);
/* The same switch / observer on one raw per-channel scratch record (callers of gpsb_track_loop that own their records).
 * out: slot_phase, walk_enable, skip_ms, skip_len, walks, phase_since_ms. */
void gpsb_host_aux_walk(void* aux_record, uint32_t enable, uint32_t period_ms);
void gpsb_host_aux_walk_state(const void* aux_record, uint32_t out[6]);

/* One acquisition snapshot for every channel (acquisition_process, PM/main.c:166) in one launch. */
int gpsb_rx_acquire_ms(gpsb_rx* rx, uint32_t ms);

/* Cold start, all channels at once: every (channel, Doppler bin, ms) cell of a sweep
 * (n_bins bins from first_bin_hz in steps of bin_step_hz, ms0 .. ms0+n_ms-1, n_ms <= 24) in one launch,
 * followed per channel by the reference's 10-cell chain vote and histogram decision
 * (PM/GPS/acquisition.c:322-416) applied bin by bin.  On return each channel that passed the vote is in
 * GPS_ACQ_FREQ_SEARCH_DONE with found_freq_offset_hz set; votes[ch*n_bins+b] (may be NULL) receives
 * the chain length of every bin and phases[ch*n_bins+b] the code phase of the longest chain. */
int gpsb_rx_cold_sweep(gpsb_rx* rx, int32_t first_bin_hz, int32_t bin_step_hz, uint32_t n_bins,
                       uint32_t ms0, uint32_t n_ms, uint8_t* votes, uint16_t* phases);

/* Cold start of every channel, any number of satellites - the acquisition half of gps_master_handling
 * (PM/GPS/gps_master.c:68-129) without its one-channel-at-a-time Doppler search, which never ends on a satellite that is
 * not in the sky (PM/GPS/acquisition.c:306-310):
 *   1. Doppler: gpsb_rx_cold_sweep over snapshots ms0 .. ms0+sweep_ms-1 for the channels in GPS_ACQ_NEED_FREQ_SEARCH;
 *      channels it leaves undecided get up to `sweeps` sweeps in all, each on the next sweep_ms snapshots, as the
 *      reference's search does when it wraps round (acquisition.c:305-310);
 *   2. code-phase rounds 1 and 2 (acquisition.c:89-104, 150-170, 196-275) from the next snapshot (ms_code0) on for every channel
 *      whose vote passed (the others are not served any more), every such channel on EVERY snapshot, until none is left in rounds 1 / 2 or round_timeout_ms
 *      snapshots have gone by (a Doppler vote that fired on noise does not get through the rounds; it is left behind);
 *   3. round 3 (acquisition.c:106-130) started together at the next snapshot for the channels that finished round 2
 *      (gps_master.c:113-118), until none is left in it or the time-out;
 *   4. channels in GPS_ACQ_DONE get GPS_NEED_PRE_TRACK (gps_master.c:121-129): gpsb_rx_track_run takes them from there.
 * The cells of up to window_ms coming snapshots are computed ahead in one launch (they are independent until each
 * channel's vote) and consumed in order.  Several GPUs: with a communicator on the context (gpsb_comm_init, include/
 * gpsb.h) the sweep's cell groups are sharded over the ranks and all-gathered, every rank runs the same votes, and
 * serve_rank / serve_world deal the satellites found round-robin over the ranks for everything that follows (no further
 * exchange: channels are independent).  Per channel the result equals the unmodified reference called on the same
 * snapshots - acquisition_start_channel, its 10 cells per bin and sweep, acquisition_start_code_search_channel at ms_code0,
 * acquisition_process_channel for ms_code0 .. ms_code12_last, acquisition_start_code_search3_channel at ms_code3_first,
 * acquisition_process_channel for ms_code3_first .. ms_last - which is how the tests check it.  All frames
 * ms0 .. ms_last must be in the ring.  opts == NULL or zero fields: the reference's Doppler grid (-7000 .. +7000 Hz in
 * steps of 500, PM/config.h:41-44), 10 snapshots per bin, time-out 400, 16 snapshots ahead at first. */
typedef struct gpsb_cold_start_opts {
    int32_t  first_bin_hz, bin_step_hz;
    uint32_t n_bins, sweep_ms, round_timeout_ms, window_ms;
    uint32_t sweeps;                    /* Doppler sweeps at most (each on the next sweep_ms snapshots); 0 = 1 */
    uint32_t serve_rank, serve_world;   /* multi-GPU: after the sweep this process goes on with channels i % serve_world ==
                                           serve_rank only (0, 0 = all); the sweep itself is sharded by gpsb_sweep_gather
                                           when the context has a communicator */
    uint32_t window_max_ms;             /* a window that was consumed whole (no channel changed its window) is followed by one
                                           twice as long, up to this many snapshots; 0 = up to the round time-out */
    uint32_t code_rounds;               /* 0: look-ahead windows, search cells of many snapshots side by side on all SMs
                                           (default); 1: k_code_rounds_run (gpsb_code_rounds, include/gpsb.h) - one launch per
                                           round, each channel's snapshots strictly one after the other on one SM: fewer
                                           launches, but a channel that never settles holds the launch for its whole time-out */
} gpsb_cold_start_opts;
typedef struct gpsb_cold_start_report {
    uint32_t ms_sweep0, ms_code0, ms_code12_last, ms_code3_first, ms_last, ms_next;   /* the snapshot schedule */
    uint32_t n_sweeps, n_searched, n_doppler_found, n_served, n_acquired;
    uint32_t launches;                                                                 /* kernel launches spent */
} gpsb_cold_start_report;
int gpsb_rx_cold_start(gpsb_rx* rx, uint32_t ms0, const gpsb_cold_start_opts* opts, gpsb_cold_start_report* rep);

/* ------------------------------------------------------------------------------------------------
 * (3) Split-phase form of the two per-channel steps.  PLAN performs the state transitions of
 * acquisition_process_channel() / gps_tracking_process() up to the point where a correlation is needed
 * and describes that correlation; FINISH consumes its result (votes, loop filters, nav bits).  Running
 * PLAN, the described cell on the GPU, then FINISH equals one reference call; callers that schedule
 * channels themselves use this to batch cells (gpsb_rx_* is built on it).  Uses the shared scratch.
 * ---------------------------------------------------------------------------------------------- */
typedef enum { GPSB_WANT_NOTHING = 0, GPSB_WANT_SEARCH, GPSB_WANT_EPL } gpsb_want;
typedef struct gpsb_plan {
    gpsb_want want;
    gpsb_search_req search;     /* valid when want == GPSB_WANT_SEARCH (start >= stop: empty window) */
    gpsb_epl_req epl;           /* valid when want == GPSB_WANT_EPL                                  */
    int stage;                  /* 1 Doppler cell, 2 code-search window, 3 pre-track window, 4 E/P/L */
} gpsb_plan;
int gpsb_host_plan_acq(gps_ch_t* ch, uint32_t frame_ms, gpsb_plan* plan);
int gpsb_host_finish_acq(gps_ch_t* ch, const gpsb_plan* plan, const gpsb_search_res* res);
int gpsb_host_plan_track(gps_ch_t* ch, uint32_t frame_ms, uint8_t index, gpsb_plan* plan);
int gpsb_host_finish_track(gps_ch_t* ch, uint8_t index, const gpsb_plan* plan, const gpsb_search_res* res,
                           const int16_t* iq6);
/* The 20-ms data bit handed to the word assembler by the most recent finish (shared scratch), or -1. */
int gpsb_host_last_nav_bit(void);
void gpsb_host_master_reset(void);

/* Channel-array helpers for bindings that cannot lay out gps_ch_t themselves (ctypes, cgo, ...). */
gps_ch_t* gpsb_host_channels_alloc(uint32_t n);
void gpsb_host_channels_free(gps_ch_t* p);
gps_ch_t* gpsb_host_channel_at(gps_ch_t* base, uint32_t i);
void gpsb_host_channel_init(gps_ch_t* ch, uint32_t prn, int32_t given_freq_offset_hz);
const uint8_t* gpsb_host_channel_code(const gps_ch_t* ch);

/* Replay helper: hand n data bits to the word assembler, the millisecond counter advancing 20 per bit from ms0. */
void gpsb_host_feed_nav_bits(gps_ch_t* ch, const uint8_t* bits, uint32_t n, uint32_t ms0);
struct gpsb_flat_eph;
void gpsb_host_channel_eph(const gps_ch_t* ch, struct gpsb_flat_eph* out);

void gpsb_host_channel_obs(const gps_ch_t* ch, uint64_t out2[2]);     /* bit patterns of obs_data.pseudorange_m, tow_s */
void gpsb_host_channel_set_tow(gps_ch_t* ch, double tow_gpst);

/* ------------------------------------------------------------------------------------------------
 * Position fix: PM/GPS/RTK/solving.h:34-40, rtk_common.h:103-110 and
 * gps_master.c:394.  Host C only - a fix is a few thousand double operations twice a second; bit-exact against the
 * compiled reference (tests/test_fix.py).  gps_master_nav_handling ends with gps_master_calculate_pos, as in the
 * reference, once gps_pos_solve_init has registered the channels.
 * ------------------------------------------------------------------------------------------------ */
extern sol_t  gps_sol;                               /* solving.c:49 */
extern double final_pos[3];                          /* solving.c:51: latitude, longitude (deg), height (m) */
extern double azel[2 * GPSB_FIX_MAX_SATS];           /* solving.c:52: azimuth / elevation per satellite (deg) */
void    gps_pos_solve_init(gps_ch_t* channels);      /* registers the ephemerides of the first sat_cnt channels: the
                                                        solver keeps POINTERS into them (like the reference); NULL unregisters */
void    gps_pos_solve(obsd_t* obs_p);                /* one sub-millisecond slice per call, like the reference */
uint8_t solving_is_busy(void);
void    gps_master_calculate_pos(gps_ch_t* channels);
void    ecef2pos(const double* r, double* pos);
void    sdrobs2obsd(gps_ch_t* channels, int ns, obsd_t* out);
double  timediff(gtime_t t1, gtime_t t2);
gtime_t timeadd(gtime_t t, double sec);
gtime_t gpst2time(int week, double sec);
double  time2gpst(gtime_t t, int* week);
/* The same fix without the slicing (the reference's pntpos, solving.c:153): 1 = fix in *sol (and pos_deg, may be
 * NULL), 0 = none.  gpsb_host_fix_channels runs it on the channels' current observations into gps_sol / final_pos. */
int  gpsb_host_fix_once(const obsd_t* obs, int n, sol_t* sol, double pos_deg[3]);
int  gpsb_host_fix_channels(gps_ch_t* channels, uint32_t n);
void gpsb_host_fix_set_iono(const double coeff[8]);  /* Klobuchar a0..a3, b0..b3; NULL / zeros = the reference's defaults */
void gpsb_host_fix_set_start(const double ecef_m[3]);   /* first iterate; default = previous fix (0,0,0 at start-up) */
void gpsb_host_fix_reset(void);
struct gpsb_flat_fix;
void gpsb_host_fix_state(struct gpsb_flat_fix* out);
void gpsb_host_channel_set_eph(gps_ch_t* ch, const struct gpsb_flat_eph* in);   /* assisted start: inject an ephemeris */
void gpsb_host_channel_set_obs(gps_ch_t* ch, double pseudorange_m, double tow_s);
uint32_t gpsb_host_sizeof_obsd(void);               /* 48 and 152, checked against the compiled reference */
uint32_t gpsb_host_sizeof_sol(void);

/* ------------------------------------------------------------------------------------------------
 * RTCM 3 output (row N4, second half): PM/GPS/obs_publish.h:8-9, RTK/rtk_common.h:100, gps_master.c:431.
 * Message 1019 (GPS ephemeris) and 1075 (GPS MSM5 observations), byte-exact against the compiled reference
 * (tests/test_rtcm.py).  Compiled out in the reference's shipped configuration (config.h:30); here a run-time
 * switch, off by default.  Frames go to the sink the application registers (the reference's UART).
 * ------------------------------------------------------------------------------------------------ */
void gpsb_host_set_rtcm_sink(void (*send)(const uint8_t* frame, uint32_t bytes), int (*busy)(void));
void gpsb_host_enable_rtcm(int on);                  /* gps_master_nav_handling then also calls gps_master_transmit_obs */
int  gpsb_host_rtcm_enabled(void);
void sendrtcmobs(obsd_t* obsd, int nsat);
void sendrtcmnav(gps_ch_t* channel);
void gps_master_transmit_obs(gps_ch_t* channels);
void setbitu(unsigned char* buff, int pos, int len, unsigned int data);
/* The same frames into a caller's buffer; return the frame length in bytes, 0 = nothing to send / does not fit. */
int  gpsb_rtcm_encode_eph(const eph_t* eph, int sat, uint8_t* out, uint32_t cap);
int  gpsb_rtcm_encode_obs(const obsd_t* obs, int n_obs, uint8_t* out, uint32_t cap);

/* Flat, layout-independent snapshot of one channel (include/gpsb_flat_state.h) for parity tests. */
struct gpsb_flat_state;
void gpsb_host_snapshot(const gps_ch_t* ch, struct gpsb_flat_state* out);
void gpsb_host_restore(gps_ch_t* ch, const struct gpsb_flat_state* in);
uint32_t gpsb_host_sizeof_channel(void);

#ifdef __cplusplus
}
#endif
#endif /* GPSB_HOST_H */
