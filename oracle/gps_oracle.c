/*
 * gps_oracle.c - see gps_oracle.h.  TEST INFRASTRUCTURE ONLY; never linked into the product.
 *
 * Independent byte/bit-domain restatement of the reference hot path
 * (Firmware/project_main/GPS/gps_misc.c, callers in acquisition.c / tracking.c).
 */
#include "gps_oracle.h"

#include <math.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * C/A code.  gps_misc.c:317-372 emulates the two 10-stage LFSRs with +/-1 arithmetic
 * (-1 == logic 1, product == XOR) and selects the satellite by delaying G2 by a per-PRN number
 * of chips (table :319-341, IS-GPS-200 Table 3-Ia/Ib expressed as delays).  Restated with plain
 * bit registers: chip[i] = G1[i] ^ G2[(i - delay) mod 1023].
 * ------------------------------------------------------------------------------------------ */
static const uint16_t k_g2_delay[210] = {
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

int orc_ca_code(int prn, uint8_t chips[ORC_CHIPS])
{
    if (prn < 1 || prn > 210) return -1;
    uint8_t g1[ORC_CHIPS], g2[ORC_CHIPS];
    unsigned r1 = 0x3FF, r2 = 0x3FF; /* bit k = stage k+1, all ones at start */
    for (int i = 0; i < ORC_CHIPS; i++) {
        g1[i] = (r1 >> 9) & 1;
        g2[i] = (r2 >> 9) & 1;
        unsigned f1 = ((r1 >> 2) ^ (r1 >> 9)) & 1;                                    /* x^3 + x^10 */
        unsigned f2 = ((r2 >> 1) ^ (r2 >> 2) ^ (r2 >> 5) ^ (r2 >> 7) ^ (r2 >> 8) ^ (r2 >> 9)) & 1;
        r1 = ((r1 << 1) | f1) & 0x3FF;
        r2 = ((r2 << 1) | f2) & 0x3FF;
    }
    int d = k_g2_delay[prn - 1];
    for (int i = 0; i < ORC_CHIPS; i++)
        chips[i] = g1[i] ^ g2[(i + ORC_CHIPS - d) % ORC_CHIPS];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Code replica.  gps_misc.c:282-300: clear 2046 bytes, then for chip k OR 0x0000FFFF<<b into the
 * unaligned 32-bit word at 16-bit index k  =>  chip k covers sample positions [16k+b, 16k+b+16).
 * Positions below b stay 0 (no wrap); chip 1022 spills b samples past 16368 (never read).
 * ------------------------------------------------------------------------------------------ */
void orc_replica(const uint8_t chips[ORC_CHIPS], unsigned bits, uint8_t rep[2048])
{
    unsigned b = bits & 15u;
    memset(rep, 0, 2048);
    for (unsigned n = b; n < 16384u; n++) {
        unsigned k = (n - b) >> 4;
        if (k < ORC_CHIPS && chips[k]) rep[n >> 3] |= (uint8_t)(1u << (n & 7));
    }
}

/* ------------------------------------------------------------------------------------------
 * Carrier NCO.  gps_misc.c:219: acc_step = (uint32_t)(freq_hz / IF_NCO_STEP_HZ) in fp32
 * (IF_NCO_STEP_HZ = 0.003810972f, PM/config.h:53); :220-221 per-word advance = acc_step*32 mod 2^32.
 * ------------------------------------------------------------------------------------------ */
uint32_t orc_nco_step(float freq_hz)
{
    volatile float q = freq_hz / 0.003810972f; /* volatile: force a rounded fp32 quotient */
    return (uint32_t)q;
}

uint32_t orc_nco_step32(uint32_t acc_step) { return (uint32_t)((uint64_t)acc_step * 32u); }

/* Quadrant patterns, gps_misc.c:216-217.  0x9999999 has seven nibbles in the reference source:
 * the top nibble is 0 and that is part of the arithmetic to reproduce. */
static const uint32_t k_sin_pat[4] = {0x33333333u, 0x09999999u, 0xCCCCCCCCu, 0x66666666u};
static const uint32_t k_cos_pat[4] = {0x09999999u, 0xCCCCCCCCu, 0x66666666u, 0x33333333u};

static uint32_t ld32(const uint8_t* p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static void st32(uint8_t* p, uint32_t v)
{
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

uint32_t orc_mix(const uint8_t sig[ORC_MS_BYTES], uint32_t acc0, uint32_t step32,
                 uint8_t* data_i, uint8_t* data_q)
{
    uint32_t acc = acc0;
    for (unsigned w = 0; w < 511u; w++) { /* PRN_SPI_WORDS_CNT/2, gps_misc.c:229 */
        uint32_t s = ld32(sig + 4u * w);
        unsigned ph = acc >> 30;
        st32(data_i + 4u * w, k_cos_pat[ph] ^ s);
        st32(data_q + 4u * w, k_sin_pat[ph] ^ s);
        acc += step32;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * XOR / popcount correlator.  gps_misc.c:48-93 in the byte domain:
 * replica byte p is compared with data byte (p + offset) mod 2046.  Even offsets use all 1023
 * replica 16-bit words.  For an odd offset 2k+1 the reference's two loops skip replica words
 * 1022-k and 1022 (the first would straddle the buffer end, the second falls off the shortened
 * second loop, gps_misc.c:81).
 * ------------------------------------------------------------------------------------------ */
static unsigned pop8(unsigned v)
{
    v = v - ((v >> 1) & 0x55u);
    v = (v & 0x33u) + ((v >> 2) & 0x33u);
    return (v + (v >> 4)) & 0x0Fu;
}

void orc_corr_sums(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                   unsigned offset, int* sum_i, int* sum_q)
{
    int si = 0, sq = 0;
    unsigned odd = offset & 1u, k = offset >> 1;
    for (unsigned w = 0; w < 1023u; w++) {
        if (odd && (w == 1022u - k || w == 1022u)) continue;
        for (unsigned h = 0; h < 2u; h++) {
            unsigned p = 2u * w + h;
            unsigned d = p + offset;
            if (d >= ORC_MS_BYTES) d -= ORC_MS_BYTES;
            si += (int)pop8(rep[p] ^ data_i[d]);
            sq += (int)pop8(rep[p] ^ data_q[d]);
        }
    }
    *sum_i = si;
    *sum_q = sq;
}

void orc_correlation_iq(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                        unsigned offset, int16_t* res_i, int16_t* res_q)
{
    int si, sq;
    orc_corr_sums(rep, data_i, data_q, offset, &si, &sq);
    *res_i = (int16_t)(si - ORC_HALF_SUM); /* gps_misc.c:140-141 */
    *res_q = (int16_t)(sq - ORC_HALF_SUM);
}

int16_t orc_correlation8(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                         unsigned offset)
{
    int si, sq;
    orc_corr_sums(rep, data_i, data_q, offset, &si, &sq);
    int a = si - ORC_HALF_SUM, b = sq - ORC_HALF_SUM;
    if (a < 0) a = 0; /* half-wave rectified, gps_misc.c:111-114 */
    if (b < 0) b = 0;
    /* gps_misc.c:116-118: two int->float conversions (round to nearest), one fp32 add, sqrtf, trunc */
    volatile float fa = (float)(a * a);
    volatile float fb = (float)(b * b);
    volatile float s = fa + fb;
    return (int16_t)sqrtf(s);
}

uint16_t orc_correlation_search(const uint8_t* rep, const uint8_t* data_i, const uint8_t* data_q,
                                unsigned start_shift, unsigned stop_shift,
                                uint16_t* aver_val, uint16_t* phase)
{
    int best = 0;
    unsigned best_pos = 0;
    long total = 0;
    for (unsigned off = start_shift; off < stop_shift; off++) {
        int c = orc_correlation8(rep, data_i, data_q, off);
        if (c > best) { /* strict: the lowest offset wins ties, all-zero leaves phase 0 (:170) */
            best = c;
            best_pos = off;
        }
        total += c;
    }
    total /= (ORC_CHIPS * 2); /* always 2046, also for sub-windows (:178) */
    if (total < 0) total = 0;
    *aver_val = (uint16_t)total;
    *phase = (uint16_t)best_pos;
    return (uint16_t)best;
}

uint32_t orc_rewind_if_phase(uint32_t accum, float if_freq_offset_hz, unsigned steps)
{
    volatile float f = (float)ORC_IF_HZ + if_freq_offset_hz; /* gps_misc.c:199 */
    uint32_t acc_step = orc_nco_step(f);
    uint64_t adv = (uint64_t)acc_step * ORC_MS_SAMPLES * (uint64_t)(steps & 0xFFu);
    return accum + (uint32_t)adv;
}

/* ------------------------------------------------------------------------------------------ fused */
uint16_t orc_search_cell(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                         float freq_hz, unsigned bits, unsigned start, unsigned stop,
                         uint16_t* aver_val, uint16_t* phase)
{
    uint8_t rep[2048], di[2048], dq[2048];
    memset(di, 0, sizeof di); /* bytes 2044/2045 of the reference's global scratch are always 0 */
    memset(dq, 0, sizeof dq);
    orc_replica(chips, bits, rep);
    orc_mix(sig, 0u, orc_nco_step32(orc_nco_step(freq_hz)), di, dq);
    return orc_correlation_search(rep, di, dq, start, stop, aver_val, phase);
}

void orc_epl_offsets(float code_phase_fine, unsigned* off_e, unsigned* off_p, unsigned* off_l,
                     unsigned* bits)
{
    int16_t fine = (int16_t)code_phase_fine;  /* tracking.c:115 */
    *bits = (unsigned)(fine & 7);             /* :116 */
    uint16_t p = (uint16_t)(fine / 8);        /* :123, C division truncates toward zero */
    uint16_t e = (uint16_t)(p - 1);
    uint16_t l = (uint16_t)(p + 1);
    if (e >= 2 * ORC_CHIPS) e = 2 * ORC_CHIPS - 1; /* :127-130 */
    if (l >= 2 * ORC_CHIPS) l = 0;
    *off_e = e; *off_p = p; *off_l = l;
}

void orc_epl_explicit(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                      uint32_t acc0, uint32_t step32, unsigned off_e, unsigned off_p,
                      unsigned off_l, unsigned bits, int16_t out6[6])
{
    uint8_t rep[2048], di[2048], dq[2048];
    memset(di, 0, sizeof di);
    memset(dq, 0, sizeof dq);
    orc_replica(chips, bits, rep);
    orc_mix(sig, acc0, step32, di, dq);
    orc_correlation_iq(rep, di, dq, off_e, &out6[0], &out6[1]);
    orc_correlation_iq(rep, di, dq, off_p, &out6[2], &out6[3]);
    orc_correlation_iq(rep, di, dq, off_l, &out6[4], &out6[5]);
}

uint32_t orc_track_epl(const uint8_t chips[ORC_CHIPS], const uint8_t sig[ORC_MS_BYTES],
                       float if_freq_offset_hz, uint32_t accum_in, float code_phase_fine,
                       int16_t out6[6])
{
    unsigned oe, op, ol, bits;
    orc_epl_offsets(code_phase_fine, &oe, &op, &ol, &bits);
    volatile float f = (float)ORC_IF_HZ + if_freq_offset_hz; /* gps_misc.c:250-251 */
    uint32_t step32 = orc_nco_step32(orc_nco_step(f));
    orc_epl_explicit(chips, sig, accum_in, step32, oe, op, ol, bits, out6);
    return accum_in + 511u * step32; /* gps_misc.c:261-273 */
}

/* ------------------------------------------------------------------------------------------ timing */
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_time_epl(const uint8_t* chips_all, unsigned n_sv, const uint8_t* signal, unsigned n_ms,
                    const uint32_t* acc0, const uint32_t* step32, const uint16_t* off_p,
                    const uint8_t* bits, int16_t* out)
{
    double t0 = now_s();
    for (unsigned m = 0; m < n_ms; m++)
        for (unsigned s = 0; s < n_sv; s++) {
            unsigned i = m * n_sv + s;
            unsigned p = off_p[i];
            unsigned e = (p == 0) ? 2045u : p - 1u;
            unsigned l = (p + 1u >= 2046u) ? 0u : p + 1u;
            orc_epl_explicit(chips_all + (size_t)s * ORC_CHIPS, signal + (size_t)m * ORC_MS_BYTES,
                             acc0[i], step32[i], e, p, l, bits[i], out + 6u * i);
        }
    return now_s() - t0;
}

double orc_time_sweep(const uint8_t* chips_all, unsigned n_sv, const uint8_t* signal, unsigned n_ms,
                      int first_bin_hz, int bin_step_hz, unsigned n_bins, unsigned bits,
                      uint16_t* out)
{
    double t0 = now_s();
    for (unsigned s = 0; s < n_sv; s++)
        for (unsigned b = 0; b < n_bins; b++)
            for (unsigned m = 0; m < n_ms; m++) {
                uint16_t avr = 0, ph = 0;
                float f = (float)(ORC_IF_HZ + first_bin_hz + (int)b * bin_step_hz);
                uint16_t mx = orc_search_cell(chips_all + (size_t)s * ORC_CHIPS,
                                              signal + (size_t)m * ORC_MS_BYTES, f, bits, 0, 2046,
                                              &avr, &ph);
                uint16_t* o = out + 3u * ((s * n_bins + b) * n_ms + m);
                o[0] = mx; o[1] = ph; o[2] = avr;
            }
    return now_s() - t0;
}
