/*
 * gpsb_epl_core.h - the early/prompt/late integrate-and-dump of one millisecond evaluated straight from
 * the RAW packed frame (no staged copy of the mixed samples), written once for nvcc (k_track_run) and for a
 * C compiler (the CPU emulation the tests use to check this very source against the reference).
 *
 * Reference arithmetic (Firmware/project_main/GPS/):
 *   replica   gps_generate_prn_data2, gps_misc.c:282-300   chip k covers sample bits [16k+b, 16k+b+16), no wrap
 *   mixer     gps_shift_to_zero_freq_track, :244-274       word w < 511: I = cos[ph] ^ S, Q = sin[ph] ^ S with
 *                                                          ph = (acc0 + w*step32) >> 30; bytes 2044..2045 stay 0
 *   sums      gps_mult_and_summ, :48-93                    data byte = (replica byte + offset) mod 2046; odd
 *                                                          offsets 2k+1 skip the 16-bit replica words 1022-k, 1022
 *
 * A thread owns EC_NW consecutive 32-bit replica words W0.. and, for each arm, needs the 32-bit window of the
 * mixed stream at byte position 4W + off for each of them.  The mixed stream, extended periodically with its
 * period of 2046 BYTES (= 511.5 words, so the second lap is shifted by 16 bits), is "ext":
 *
 *   ext(x) = M(x)                          x <= 510        M(w) = pattern(ph(w)) ^ S[w], M(511) = 0
 *          = M(0) << 16                    x == 511
 *          = M(y) >> 16 | M(y+1) << 16     512 <= x <= 1022, y = x - 512
 *          = 0                             x == 1023
 */
#ifndef GPSB_EPL_CORE_H
#define GPSB_EPL_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define EC_FN static __device__ __forceinline__
#define EC_POPC(x) __popc(x)
#define EC_FSHR(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#define EC_FSHL(lo, hi, s) __funnelshift_l((lo), (hi), (s))
#define EC_UNROLL _Pragma("unroll")
#else
#define EC_UNROLL
#define EC_FN static inline
#define EC_POPC(x) __builtin_popcount(x)
static inline uint32_t ec_fshr(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
static inline uint32_t ec_fshl(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
#define EC_FSHR(lo, hi, s) ec_fshr((lo), (hi), (s))
#define EC_FSHL(lo, hi, s) ec_fshl((lo), (hi), (s))
#endif

#define EC_WORDS      512      /* 32-bit words per ms frame (incl. the 2 pad bytes) */
#define EC_MIX_WORDS  511      /* PRN_SPI_WORDS_CNT/2, gps_misc.c:229 */
#define EC_HALF_SUM   8184     /* BITS_IN_PRN/2, gps_misc.c:140 */

/* Quadrant patterns of the fs/4 carrier, gps_misc.c:216-217.  The reference literal 0x9999999 has seven
 * nibbles, i.e. the top nibble is zero; sin[ph] == cos[(ph+3)&3]. */
EC_FN uint32_t ec_cos_pattern(uint32_t ph)
{
    return (ph & 2u) ? ((ph & 1u) ? 0x33333333u : 0x66666666u)
                     : ((ph & 1u) ? 0xCCCCCCCCu : 0x09999999u);
}

/* mixed I and Q word w of the frame (0 beyond the 511 words the reference mixes) */
EC_FN void ec_mixed(const uint32_t* S, uint32_t acc0, uint32_t step32, int w, uint32_t* mi, uint32_t* mq)
{
    if (w >= EC_MIX_WORDS) {
        *mi = 0u;
        *mq = 0u;
        return;
    }
    const uint32_t ph = (acc0 + (uint32_t)w * step32) >> 30;
    const uint32_t s = S[w];
    *mi = ec_cos_pattern(ph) ^ s;
    *mq = ec_cos_pattern((ph + 3u) & 3u) ^ s;
}

/* word x (0..1023) of the periodically extended mixed streams */
EC_FN void ec_ext(const uint32_t* S, uint32_t acc0, uint32_t step32, int x, uint32_t* ei, uint32_t* eq)
{
    if (x < EC_MIX_WORDS) {
        ec_mixed(S, acc0, step32, x, ei, eq);
    } else if (x == EC_MIX_WORDS) {
        uint32_t a, b;
        ec_mixed(S, acc0, step32, 0, &a, &b);
        *ei = a << 16;
        *eq = b << 16;
    } else if (x < 2 * EC_WORDS - 1) {
        uint32_t ai, aq, bi, bq;
        ec_mixed(S, acc0, step32, x - EC_WORDS, &ai, &aq);
        ec_mixed(S, acc0, step32, x - EC_WORDS + 1, &bi, &bq);
        *ei = (ai >> 16) | (bi << 16);
        *eq = (aq >> 16) | (bq << 16);
    } else {
        *ei = 0u;
        *eq = 0u;
    }
}

/* Replica word W for sub-byte shift b from the chip-expanded table E (E[w] = chips 2w, 2w+1 as 0xFFFF
 * halves; E[511] = chip 1022 in the low half).  No wrap: bits below b stay 0. */
EC_FN uint32_t ec_replica_word(const uint32_t* E, int W, uint32_t b)
{
    const uint32_t hi = E[W];
    const uint32_t lo = (W > 0) ? E[W - 1] : 0u;
    return EC_FSHL(lo, hi, b);
}

/* Mask of replica word W for byte offset off (gps_misc.c:59-89): the upper half of word 511 is replica
 * word 1023, which does not exist; for odd offsets 2k+1 the 16-bit replica words 1022-k and 1022 are skipped. */
EC_FN uint32_t ec_word_mask(int W, uint32_t off)
{
    const uint32_t odd = off & 1u;
    if (W == EC_WORDS - 1) return odd ? 0u : 0x0000FFFFu;
    if (odd) {
        const int u = 1022 - (int)(off >> 1);
        if (W == (u >> 1)) return (u & 1) ? 0x0000FFFFu : 0xFFFF0000u;
    }
    return 0xFFFFFFFFu;
}

/* Mismatch counts of the three arms over replica words W0 .. W0+nw-1, packed as I | Q << 16 (a whole
 * millisecond is at most 16368 per component, so the halves never carry into each other).
 * off[3] = byte offsets early, prompt, late; bits = sub-byte replica shift. */
#ifndef EC_NW_MAX
#define EC_NW_MAX 4
#endif
EC_FN void ec_epl_partial(const uint32_t* S, const uint32_t* E, uint32_t acc0, uint32_t step32,
                          const uint32_t off[3], uint32_t bits, int W0, int nw, uint32_t acc[3])
{
    uint32_t R[EC_NW_MAX], ei[EC_NW_MAX + 1], eq[EC_NW_MAX + 1];
    EC_UNROLL
    for (int j = 0; j < nw; j++) R[j] = ec_replica_word(E, W0 + j, bits);
    int have_x0 = -1;
    EC_UNROLL
    for (int a = 0; a < 3; a++) {
        const int x0 = (int)(off[a] >> 2);
        const uint32_t sh = (off[a] & 3u) * 8u;
        if (x0 != have_x0) {                       /* neighbouring arms usually share their data words */
            EC_UNROLL
            for (int j = 0; j <= nw; j++) ec_ext(S, acc0, step32, x0 + W0 + j, &ei[j], &eq[j]);
            have_x0 = x0;
        }
        uint32_t sum = 0;
        EC_UNROLL
        for (int j = 0; j < nw; j++) {
            const uint32_t m = ec_word_mask(W0 + j, off[a]);
            const uint32_t vi = (EC_FSHR(ei[j], ei[j + 1], sh) ^ R[j]) & m;
            const uint32_t vq = (EC_FSHR(eq[j], eq[j + 1], sh) ^ R[j]) & m;
            sum += (uint32_t)EC_POPC(vi) + ((uint32_t)EC_POPC(vq) << 16);
        }
        acc[a] += sum;
    }
}

/* packed sums of the three arms -> IE,QE,IP,QP,IL,QL (gps_misc.c:140-141: popcount - 8184) */
EC_FN void ec_unpack_sums(const uint32_t packed[3], int16_t iq[6])
{
    for (int a = 0; a < 3; a++) {
        iq[2 * a] = (int16_t)((int)(packed[a] & 0xFFFFu) - EC_HALF_SUM);
        iq[2 * a + 1] = (int16_t)((int)(packed[a] >> 16) - EC_HALF_SUM);
    }
}

#endif /* GPSB_EPL_CORE_H */
