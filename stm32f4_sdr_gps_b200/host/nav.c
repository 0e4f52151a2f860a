/*
 * nav.c - navigation-bit stream from the sign of the prompt in-phase sum: 20-ms bit-edge
 * synchronisation from sign flips inside 4-ms slots, majority-vote data bits, preamble / polarity
 * detection, IS-GPS-200 parity and 10-word subframe assembly with a sub-bit subframe time stamp.
 *
 * Behaviour follows Firmware/project_main/GPS/nav_data.c (cited per function).  The reference keeps the
 * per-slot sample buffers in function statics shared by every channel (nav_data.c:48-51); here they
 * live in the gpsb_aux the caller passes (shared for the reference-named API, per channel in the
 * batched receiver).  Ephemeris field decoding (nav_data_decode.c) is outside the hot path: a completed
 * subframe is left in nav_data.subframe_data for whoever wants to decode it.
 */
#include <stdlib.h>

#include "host_internal.h"

#define MS_PER_BIT            20          /* CODES_IN_BIT, nav_data.c:15 */
#define WORDS_PER_SUBFRAME    10          /* nav_data.c:17 */
#define POLARITY_TIMEOUT_MS   12000u      /* two subframes, nav_data.c:22 */

static const uint8_t k_preamble[8] = {1, 0, 0, 0, 1, 0, 1, 1};    /* nav_data.c:26 */

/* 1 when the first eight buffered bits equal the preamble (flip = 0) or its complement (flip = 1) */
static int starts_with_preamble(const gps_nav_data_t* n, uint8_t flip)
{
    for (unsigned i = 0; i < sizeof k_preamble; i++)
        if (n->word_buf[i] != (k_preamble[i] ^ flip)) return 0;
    return 1;
}

/* nav_data.c:409-426: copy the 30 buffered bits to bit positions word_cnt*30.. of the subframe image
 * (bit i at byte i/8, bit i%8) and remember D29/D30 for the next word's parity. */
static void store_word(gps_nav_data_t* n)
{
    unsigned pos = n->word_cnt * GPS_NAV_WORD_LENGTH;
    for (unsigned i = 0; i < GPS_NAV_WORD_LENGTH; i++, pos++) {
        uint8_t mask = (uint8_t)(1u << (pos & 7u));
        if (n->word_buf[i] == 1) n->subframe_data[pos >> 3] |= mask;
        else n->subframe_data[pos >> 3] &= (uint8_t)~mask;
    }
    n->old_D29 = n->word_buf[28];
    n->old_D30 = n->word_buf[29];
}

/* IS-GPS-200 table 20-XIV parity over the buffered word; data bits are first complemented in place by
 * the previous D30 as the reference does (nav_data.c:433-453), which also changes what store_word saves. */
static int parity_ok(gps_nav_data_t* n)
{
    static const uint8_t taps[6][15] = {
        {1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23, 0},
        {2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24, 0},
        {1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 0},
        {2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23, 0},
        {1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24},
        {3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24, 0, 0}};
    const uint8_t seed[6] = {n->old_D29, n->old_D30, n->old_D29, n->old_D30, n->old_D30, n->old_D29};
    uint8_t* w = n->word_buf;                         /* ICD bit d[k] is w[k-1] */
    for (unsigned k = 1; k < 25; k++) w[k - 1] ^= n->old_D30;
    for (unsigned p = 0; p < 6; p++) {
        uint8_t v = seed[p];
        for (unsigned k = 0; k < 15 && taps[p][k]; k++) v ^= w[taps[p][k] - 1];
        if (w[24 + p] != v) return 0;
    }
    return 1;
}

/* nav_data.c:356-380: time stamp (ms counter) of the bit edge that started the subframe just completed */
static void stamp_subframe(gps_nav_data_t* n)
{
    if (!n->accurate_swap_ok) return;
    uint32_t now = hx_now_ms();
    uint32_t edge = (now / MS_PER_BIT) * MS_PER_BIT + n->accurate_swap_time;
    if ((int32_t)(now - edge) < 0) edge -= MS_PER_BIT;      /* the edge estimate was late: use the previous one */
    n->subframe_cnt++;
    n->last_subframe_time = edge;
}

/* nav_data.c:257-352 */
void hx_nav_word_bit(gps_ch_t* ch, uint8_t new_bit)
{
    gps_nav_data_t* n = &ch->nav_data;
    if (n->word_cnt == 0) {                                   /* hunting for a preamble */
        memmove(n->word_buf, n->word_buf + 1, GPS_NAV_WORD_LENGTH - 1);
        n->word_buf[GPS_NAV_WORD_LENGTH - 1] = new_bit;
        if (starts_with_preamble(n, 0)) {
            store_word(n);
            n->word_cnt = 1;
            n->word_bit_cnt = 0;
            n->inv_preabmle_cnt = 0;
        }
        if (n->polarity_found == 0 && n->word_cnt == 0) {    /* 0/180 degree ambiguity of the Costas loop */
            if (starts_with_preamble(n, 1)) n->inv_preabmle_cnt++;
            if (n->inv_preabmle_cnt >= 2) n->inv_polarity_flag = 1;
        }
        if (n->polarity_found) {
            uint32_t now = hx_now_ms();
            if (now - n->word_detection_timestamp > POLARITY_TIMEOUT_MS) {
                n->word_detection_timestamp = now;
                n->polarity_found = 0;
                n->inv_polarity_flag = 0;
            }
        }
        return;
    }
    n->word_buf[n->word_bit_cnt++] = new_bit;                 /* collecting words 2..10 */
    if (n->word_bit_cnt < GPS_NAV_WORD_LENGTH) return;
    if (!parity_ok(n)) {
        n->word_cnt = 0;
        memset(n->word_buf, 0, GPS_NAV_WORD_LENGTH);
        return;
    }
    n->word_cnt_test++;
    store_word(n);
    n->word_cnt++;
    n->word_bit_cnt = 0;
    n->word_detection_timestamp = hx_now_ms();
    n->polarity_found = 1;
    if (n->word_cnt == WORDS_PER_SUBFRAME) {
        ch->eph_data.sub_cnt++;                               /* nav_data_decode.c:47 (field decode not done here) */
        stamp_subframe(n);
        n->word_cnt = 0;
        n->new_subframe_flag = 1;
        memset(n->word_buf, 0, GPS_NAV_WORD_LENGTH);
    }
}

void gps_nav_data_words_detection(gps_ch_t* channel, uint8_t new_bit) { if (channel) hx_nav_word_bit(channel, new_bit); }

/* nav_data.c:223-252: close a data bit when the position inside the 20-ms period wraps */
static void count_ms_into_bit(gps_ch_t* ch, gpsb_aux* aux, uint8_t ms_bit, uint32_t now)
{
    gps_nav_data_t* n = &ch->nav_data;
    uint8_t pos = (uint8_t)((now - n->old_swap_time) % MS_PER_BIT);
    if (pos < n->old_reminder) {
        uint8_t bit = n->last_bit_pos_cnt > n->last_bit_neg_cnt;
        aux->last_nav_bit = (int8_t)bit;
        hx_nav_word_bit(ch, bit);
        n->last_bit_pos_cnt = 0;
        n->last_bit_neg_cnt = 0;
    }
    if (ms_bit) n->last_bit_pos_cnt++;
    else n->last_bit_neg_cnt++;
    n->old_reminder = pos;
}

/* nav_data.c:145-218: decide whether the single sign flip seen at slot position 2 really happened
 * between samples 0/1 or 1/2, from the prompt amplitudes (the circular correlator smears an edge over
 * the millisecond in which it falls). */
static void refine_edge(gps_ch_t* ch, const gpsb_aux* aux)
{
    gps_nav_data_t* n = &ch->nav_data;
    const int16_t* v = aux->slot_ip;
    if (abs(v[1]) > abs(v[0])) return;
    if (v[3] == 0) return;
    float ends = (float)abs(v[0]) / (float)abs(v[3]);
    if (ends > 1.5f || ends < 0.7f) return;

    int16_t chip = (int16_t)((int16_t)ch->tracking_data.code_phase_fine / 16);
    if (chip < 0 || chip > PRN_LENGTH) return;

    uint8_t edge_at = 0;
    if (chip < PRN_LENGTH / 4 || chip > PRN_LENGTH * 3 / 4) {
        if (v[1] == 0) return;
        float head = (float)abs(v[0]) / (float)abs(v[1]);
        if (head > 1.5f || head < 0.7f) return;
        edge_at = (chip < PRN_LENGTH / 4) ? 2 : 1;
    } else {
        uint16_t step_a = (uint16_t)abs(v[0] - v[1]);
        uint16_t step_b = (uint16_t)abs(v[2] - v[3]);
        if (step_a > step_b) {
            if (step_b == 0) return;
            if ((float)step_a / (float)step_b < 2.5f) return;
            edge_at = 1;
        } else {
            if (step_a == 0) return;
            if ((float)step_b / (float)step_a < 2.5f) return;
            edge_at = 2;
        }
    }
    n->accurate_swap_time = (uint8_t)((aux->slot_start_ticks + edge_at) % MS_PER_BIT);
    n->accurate_swap_ok = 1;
}

/* nav_data.c:46-138 */
void hx_nav_new_code(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t new_i)
{
    gps_nav_data_t* n = &ch->nav_data;
    aux->last_nav_bit = -1;
    if (index >= GPSB_SLOT_LEN) return;
    uint8_t ms_bit = (uint8_t)((new_i > 0) ^ (n->inv_polarity_flag != 0));
    aux->slot_bits[index] = ms_bit;
    aux->slot_ip[index] = new_i;
    uint32_t now = hx_now_ms();
    if (index == 0) aux->slot_start_ticks = now;
    if (n->period_sync_ok_flag == 1) count_ms_into_bit(ch, aux, ms_bit, now);
    if (index < GPSB_SLOT_LEN - 1) return;

    /* end of the 4-ms slot: exactly one sign flip is a candidate bit edge */
    uint8_t flips = 0, flip_pos = 0;
    for (uint8_t i = 1; i < GPSB_SLOT_LEN; i++)
        if (aux->slot_bits[i] != aux->slot_bits[i - 1]) { flips++; flip_pos = i; }
    if (flips != 1) return;

    uint32_t edge = aux->slot_start_ticks + flip_pos;
    uint8_t phase = (uint8_t)((edge - n->old_swap_time) % MS_PER_BIT);
    if (phase < 2 || phase == MS_PER_BIT - 1) {               /* a multiple of 20 ms since the last edge */
        if (n->right_period_cnt < 10) n->right_period_cnt++;
        if (n->right_period_cnt > 8) n->period_sync_ok_flag = 1;
    } else {
        if (n->right_period_cnt > 0) n->right_period_cnt--;
        if (n->right_period_cnt < 3) n->period_sync_ok_flag = 0;
    }
    n->old_swap_time = edge;
    if (n->period_sync_ok_flag && flip_pos == 2) refine_edge(ch, aux);
}

void gps_nav_data_analyse_new_code(gps_ch_t* channel, uint8_t index, int16_t new_i)
{
    if (channel) hx_nav_new_code(channel, &g_shared_aux, index, new_i);
}
