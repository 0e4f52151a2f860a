/*
 * bind.c - process-wide binding of the reference-named API to one GPU context, the millisecond clock
 * seam, C/A code generation and the level-0 primitives with their reference signatures.
 *
 * Reference: Firmware/project_main/GPS/gps_misc.c (primitives :98-300, code generator :317-372),
 * Firmware/project_main/signal_capture.h:15 (ms counter).
 */
#include <stdlib.h>

#include "host_internal.h"

static gpsb_ctx* g_ctx = NULL;
static __thread int g_last_status = 0;
static uint32_t g_sat_cnt = GPS_SAT_CNT;
static __thread uint32_t g_packet_cnt = 0;   /* per thread: batched workers run channels at their own pace */
static int (*g_rand)(void) = NULL;
gpsb_aux g_shared_aux;

int gpsb_host_attach(gpsb_ctx* ctx)
{
    g_ctx = ctx;
    memset(&g_shared_aux, 0, sizeof g_shared_aux);
    g_shared_aux.last_nav_bit = -1;
    g_last_status = ctx ? GPSB_OK : GPSB_ERR_ARG;
    return g_last_status;
}

gpsb_ctx* gpsb_host_context(void) { return g_ctx; }
int gpsb_host_last_status(void) { return g_last_status; }
int hx_note(int status) { g_last_status = status; return status; }

void gpsb_host_set_sat_cnt(uint32_t n) { if (n) g_sat_cnt = n; }
uint32_t gpsb_host_sat_cnt(void) { return g_sat_cnt; }

void gpsb_host_set_packet_cnt(uint32_t ms) { g_packet_cnt = ms; }
__attribute__((weak)) uint32_t signal_capture_get_packet_cnt(void) { return g_packet_cnt; }
uint32_t hx_now_ms(void) { return signal_capture_get_packet_cnt(); }

void gpsb_host_set_rand(int (*fn)(void)) { g_rand = fn; }
int hx_rand(gpsb_aux* aux)
{
    if (g_rand) return g_rand();
    if (!aux || aux == &g_shared_aux || aux->process_rand) return rand();   /* reference-named API: the process-wide stream */
    /* batched channel: glibc's generator with its default seed, kept in the channel's own scratch so the
     * channel draws what it would draw as the only channel of a process, on the host or on the device */
    return lc_rand31_next(&aux->rnd);
}

uint32_t hx_nco_step(float freq_hz) { return lc_nco_step(freq_hz); }
uint32_t hx_nco_step32(float freq_hz) { return lc_nco_step32(freq_hz); }

/* Put one millisecond of host samples where the kernels can see it: ring frame (ms counter mod ring). */
int hx_stage_frame(const uint8_t* data, uint32_t* frame_ms)
{
    if (!g_ctx) return hx_note(GPSB_ERR_STATE);
    if (!data) return hx_note(GPSB_ERR_ARG);
    *frame_ms = hx_now_ms();
    return hx_note(gpsb_upload_signal(g_ctx, *frame_ms, 1, data));
}

/* ------------------------------------------------------------------ C/A code (gps_misc.c:317-372) */
/* Two 10-stage shift registers, all ones at start: G1 taps 3,10; G2 taps 2,3,6,8,9,10; the satellite
 * is selected by delaying the G2 output (IS-GPS-200).  chip = G1 xor delayed G2, stored as 0/1. */
static const uint16_t k_g2_delay_chips[210] = {
    5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
    473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862, 863, 950, 947, 948, 950, 67, 103,
    91, 19, 679, 225, 625, 946, 638, 161, 1001, 554, 280, 710, 709, 775, 864, 558, 220, 397, 55,
    898, 759, 367, 299, 1018, 729, 695, 780, 801, 788, 732, 34, 320, 327, 389, 407, 525, 405, 221,
    761, 260, 326, 955, 653, 699, 422, 188, 438, 959, 539, 879, 677, 586, 153, 792, 814, 446, 264,
    1015, 278, 536, 819, 156, 957, 159, 712, 885, 461, 248, 713, 126, 807, 279, 122, 197, 693, 632,
    771, 467, 647, 203, 145, 175, 52, 21, 237, 235, 886, 657, 634, 762, 355, 1012, 176, 603, 130,
    359, 595, 68, 386, 797, 456, 499, 883, 307, 127, 211, 121, 118, 163, 628, 853, 484, 289, 811,
    202, 1021, 463, 568, 904, 670, 230, 911, 684, 309, 644, 932, 12, 314, 891, 212, 185, 675, 503,
    150, 395, 345, 846, 798, 992, 357, 995, 877, 112, 144, 476, 193, 109, 445, 291, 87, 399, 292,
    901, 339, 208, 711, 189, 263, 537, 663, 942, 173, 900, 30, 500, 935, 556, 373, 85, 652, 310};

void gps_generate_prn(uint8_t* dest, int prn)
{
    if (prn < 1 || prn > 210 || !dest) return;
    uint8_t g1[PRN_LENGTH], g2[PRN_LENGTH];
    unsigned r1 = 0x3FFu, r2 = 0x3FFu;
    for (int i = 0; i < PRN_LENGTH; i++) {
        g1[i] = (uint8_t)((r1 >> 9) & 1u);
        g2[i] = (uint8_t)((r2 >> 9) & 1u);
        unsigned fb1 = ((r1 >> 2) ^ (r1 >> 9)) & 1u;
        unsigned fb2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1u;
        r1 = ((r1 << 1) | fb1) & 0x3FFu;
        r2 = ((r2 << 1) | fb2) & 0x3FFu;
    }
    const int lag = k_g2_delay_chips[prn - 1];
    for (int i = 0; i < PRN_LENGTH; i++) dest[i] = g1[i] ^ g2[(i + PRN_LENGTH - lag) % PRN_LENGTH];
}

void gps_fill_summ_table(void) { /* popcount is an instruction on the GPU (gps_misc.c:19-38 not needed) */ }

void gps_channell_prepare(gps_ch_t* channel)
{
    if (!channel || channel->prn < 1) return;           /* gps_misc.c:308 */
    gps_generate_prn(channel->prn_code, channel->prn);
    if (g_ctx) hx_note(gpsb_set_code(g_ctx, channel->prn, channel->prn_code));
}

/* ------------------------------------------------------------------ level-0 primitives */
int16_t gps_correlation8(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t offset)
{
    int16_t r = 0;
    hx_note(g_ctx ? gpsb_l0_correlation8(g_ctx, prn_p, data_i, data_q, offset, &r) : GPSB_ERR_STATE);
    return r;
}

void gps_correlation_iq(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t offset, int16_t* res_i,
                        int16_t* res_q)
{
    hx_note(g_ctx ? gpsb_l0_correlation_iq(g_ctx, prn_p, data_i, data_q, offset, res_i, res_q) : GPSB_ERR_STATE);
}

uint16_t correlation_search(uint16_t* prn_p, uint16_t* data_i, uint16_t* data_q, uint16_t start_shift,
                            uint16_t stop_shift, uint16_t* aver_val, uint16_t* phase)
{
    uint16_t mx = 0;
    hx_note(g_ctx ? gpsb_l0_correlation_search(g_ctx, prn_p, data_i, data_q, start_shift, stop_shift, aver_val,
                                               phase, &mx)
                  : GPSB_ERR_STATE);
    return mx;
}

void gps_shift_to_zero_freq(uint8_t* signal_data, uint8_t* data_i, uint8_t* data_q, float freq_hz)
{
    hx_note(g_ctx ? gpsb_l0_shift_to_zero_freq(g_ctx, signal_data, data_i, data_q, 0u, hx_nco_step32(freq_hz), NULL)
                  : GPSB_ERR_STATE);
}

void gps_shift_to_zero_freq_track(gps_tracking_t* trk, uint8_t* signal_data, uint8_t* data_i, uint8_t* data_q)
{
    if (!trk) return;
    float carrier = (float)IF_FREQ_HZ + trk->if_freq_offset_hz;          /* gps_misc.c:250-251 */
    uint32_t after = trk->if_freq_accum;
    int rc = g_ctx ? gpsb_l0_shift_to_zero_freq(g_ctx, signal_data, data_i, data_q, trk->if_freq_accum,
                                                hx_nco_step32(carrier), &after)
                   : GPSB_ERR_STATE;
    if (hx_note(rc) == GPSB_OK) trk->if_freq_accum = after;
}

void gps_generate_prn_data2(gps_ch_t* channel, uint16_t* data, uint16_t offset_bits)
{
    if (!channel || !data) return;
    int rc = hx_note(g_ctx ? gpsb_l0_generate_prn_data2(g_ctx, channel->prn_code, data, offset_bits) : GPSB_ERR_STATE);
    /* the reference's last 32-bit OR spills the top `bits` samples of chip 1022 into word 1023
     * (gps_misc.c:294; never read by the correlator, never cleared) - keep the caller's buffer identical */
    unsigned b = offset_bits & 15u;
    if (rc == GPSB_OK && b && channel->prn_code[PRN_LENGTH - 1]) data[PRN_SPI_WORDS_CNT] |= (uint16_t)((0xFFFFu << b) >> 16);
}

/* Catch-up of the carrier NCO over skipped milliseconds (gps_misc.c:196-204), pure host arithmetic. */
void gps_rewind_if_phase(gps_tracking_t* trk, uint8_t steps)
{
    if (trk) lc_rewind_if_phase(trk, steps);
}
