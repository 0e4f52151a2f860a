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
 * The sums are indexed here by DATA word: a thread owns nw consecutive 32-bit words w of the frame, so the
 * carrier pattern of a word is formed once and serves all three arms, and the 2046-byte period only shows on the
 * replica side: data byte d meets replica byte (d - off) mod 2046, i.e. the thread needs the 32-bit window of
 * the periodically extended replica stream RX at byte position (4w - off) mod 2046.  RX depends on the code and
 * the sub-byte shift only (ec_rx_word; 513 words, built once per run for all eight shifts).
 *
 * For an odd offset 2k+1 the reference's two skipped replica words are the ones whose data bytes would be
 * {2045, 0} (word 1022-k, which would straddle the end of the buffer) and {2k-1, 2k} (word 1022): in data space
 * four byte positions to mask.  Data bytes 2044..2045 are zero after mixing (the mixer stops at word 511) but
 * still count against the replica; bytes 2046..2047 of word 511 do not exist.
 *
 * Phase 1 needs only the code offsets (DLL output): T = data word ^ replica window under its byte mask, and -
 * because the 1-bit carrier of a word is one of only four patterns - the mismatch count of T against each of
 * them (4 x POPC, off the serial path).  Phase 2 needs the carrier NCO words (PLL/FLL output): the phase of
 * each word selects its I and Q count; three integer instructions per arm and word.
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
#define EC_PRMT(a, b, sel) __byte_perm((a), (b), (sel))
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
/* PRMT: result byte i = byte (sel >> 4i & 7) of the eight bytes b:a (selectors 0..7 only) */
static inline uint32_t ec_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
#define EC_PRMT(a, b, sel) ec_prmt((a), (b), (sel))
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

/* cos and sin quadrant patterns of phase ph as rotations of 0x33333333 (same values as ec_cos_pattern(ph),
 * ec_cos_pattern((ph+3)&3), without a select chain) */
EC_FN void ec_patterns(uint32_t ph, uint32_t* c, uint32_t* s)
{
    const uint32_t P = 0x33333333u;
    const uint32_t rc = EC_FSHR(P, P, (ph + 1u) & 3u);
    const uint32_t rs = EC_FSHR(P, P, ph & 3u);
    *c = (ph & 3u) == 0u ? (rc & 0x0FFFFFFFu) : rc;
    *s = (ph & 3u) == 1u ? (rs & 0x0FFFFFFFu) : rs;
}

/* Replica word W for sub-byte shift b from the chip-expanded table E (E[w] = chips 2w, 2w+1 as 0xFFFF
 * halves; E[511] = chip 1022 in the low half).  No wrap: bits below b stay 0. */
EC_FN uint32_t ec_replica_word(const uint32_t* E, int W, uint32_t b)
{
    const uint32_t hi = E[W];
    const uint32_t lo = (W > 0) ? E[W - 1] : 0u;
    return EC_FSHL(lo, hi, b);
}

/* Word x (0..512) of the replica buffer extended with its period of 2046 bytes: the 1023 16-bit words the
 * correlator reads (the spill of chip 1022 beyond sample 16368 is dropped), then the buffer again from byte 0. */
#define EC_RX_WORDS 513
EC_FN uint32_t ec_rx_word(const uint32_t* E, int x, uint32_t b)
{
    if (x < EC_WORDS - 1) return ec_replica_word(E, x, b);
    const uint32_t r0 = ec_replica_word(E, 0, b);
    if (x == EC_WORDS - 1) return (ec_replica_word(E, EC_WORDS - 1, b) & 0xFFFFu) | (r0 << 16);
    return (r0 >> 16) | (ec_replica_word(E, 1, b) << 16);
}

#ifndef EC_NW_MAX
#define EC_NW_MAX 4
#endif

/* What a thread keeps between the two phases of a millisecond: for each arm and owned data word the mismatch
 * counts of (raw word ^ replica window) against each of the FOUR quadrant patterns, one count per byte
 * (C[a][j] = n0 | n1 << 8 | n2 << 16 | n3 << 24, each at most 32).  The carrier only decides, per word, which
 * two of the four are the I and the Q count, so no XOR / POPC is left for phase 2. */
typedef struct ec_partial {
    uint32_t C[3][EC_NW_MAX];
} ec_partial;

/* the four pattern counts of one word under a byte mask (the rare path) */
EC_FN uint32_t ec_counts_masked(uint32_t t, uint32_t m, int mixed)
{
    uint32_t c = 0;
    EC_UNROLL
    for (uint32_t ph = 0; ph < 4; ph++) {
        const uint32_t pat = mixed ? ec_cos_pattern(ph) : 0u;
        c |= (uint32_t)EC_POPC((t ^ pat) & m) << (8u * ph);
    }
    return c;
}

/* the four pattern counts of a word in which all 32 bits count, with TWO population counts instead of four.
 * Below the top nibble the four patterns are 0x9999999, 0xCCCCCCC, 0x6666666, 0x3333333: two complementary pairs, so
 * over those 28 bits n0 = 28 - n2 and n1 = 28 - n3.  The top nibble (patterns 0x0, 0xC, 0x6, 0x3 - the reference's
 * 0x9999999 literal has only seven nibbles) contributes one of 16 packed constants, ec_top_nibble_counts(x >> 28). */
EC_FN uint32_t ec_top_nibble_counts(uint32_t v)
{
    uint32_t c = 0;
    EC_UNROLL
    for (uint32_t ph = 0; ph < 4; ph++) c |= (uint32_t)EC_POPC((v ^ (ec_cos_pattern(ph) >> 28)) & 0xFu) << (8u * ph);
    return c;
}
#if defined(__CUDACC__)
/* 16 x 4 bytes; lanes reading the same entry are served by one broadcast, different entries sit in different banks */
#define EC_TOP_LUT(v) ec_top_lut[(v)]
#else
#define EC_TOP_LUT(v) ec_top_nibble_counts(v)
#endif
#define EC_COUNTS_FULL(x, out)                                                                                   \
    do {                                                                                                         \
        const uint32_t x_ = (x);                                                                                 \
        const uint32_t n2_ = (uint32_t)EC_POPC((x_ & 0x0FFFFFFFu) ^ 0x06666666u);                                \
        const uint32_t n3_ = (uint32_t)EC_POPC((x_ & 0x0FFFFFFFu) ^ 0x03333333u);                                \
        /* (28 - n2) | (28 - n3) << 8 | n2 << 16 | n3 << 24, plus the top nibble's four counts */               \
        (out) = n2_ * 0x0000FFFFu + n3_ * 0x00FFFF00u + 0x1C1Cu + EC_TOP_LUT(x_ >> 28);                          \
    } while (0)

/* ---- Phase 1 ------------------------------------------------------------------------------------------
 * Work split of a millisecond over a CTA:
 *   plain threads own data words 1..510 in runs of nw; all 32 bits of such a word count for every arm EXCEPT the
 *                 (at most two per odd arm) bytes 2k-1, 2k - which they count anyway;
 *   edge lanes    twelve lanes of one extra warp put that right and own the two irregular words: per arm one lane
 *                 each for word 0, word 511, and for taking data byte off-2 and data byte off-1 back out
 *                 (a negative contribution).  Keeping the masks out of the plain threads keeps them free of
 *                 branches: a warp that diverges into a masked path is what every other warp would wait for. */

/* Plain thread.  S = raw frame, RX = extended replica for the current sub-byte shift, off[3] = byte offsets early,
 * prompt, late, data words w0 .. w0+nw-1 with 1 <= w0 and w0+nw <= 511. */
#if defined(__CUDACC__)
EC_FN void ec_epl_phase1(const uint32_t* S, const uint32_t* RX, const uint32_t off[3], int w0, int nw, ec_partial* p,
                         const uint32_t* ec_top_lut)
#else
EC_FN void ec_epl_phase1(const uint32_t* S, const uint32_t* RX, const uint32_t off[3], int w0, int nw, ec_partial* p)
#endif
{
    uint32_t t[3][EC_NW_MAX], s[EC_NW_MAX];
    EC_UNROLL
    for (int j = 0; j < EC_NW_MAX; j++)
        if (j < nw) s[j] = S[w0 + j];

    /* The tracking loop asks for neighbouring arms (early = prompt - 1, late = prompt + 1 byte): the replica
     * windows of all arms and owned words then lie in one run of nw*4 + 2 bytes starting at byte
     * 4*w0 - off_late, i.e. in nw + 2 consecutive words of RX - unless that run wraps around the period. */
    const int lo = 4 * w0 - (int)off[2];                        /* first byte of the run (late arm, first word) */
    const int hi = 4 * (w0 + nw - 1) - (int)off[0] + 3;         /* last byte (early arm, last word) */
    const int neighbours = off[0] + 1u == off[1] && off[1] + 1u == off[2];
    if (neighbours && ((lo < 0) == (hi < 0))) {
        const int base = lo + (lo < 0 ? 2 * 1023 : 0);
        const int xr = base >> 2;
        const uint32_t sh = ((uint32_t)base & 3u) * 8u;         /* late arm; prompt is 1 byte, early 2 bytes further */
        uint32_t r[EC_NW_MAX + 2];
        EC_UNROLL
        for (int j = 0; j < EC_NW_MAX + 2; j++)
            if (j < nw + 2) r[j] = RX[xr + j];                  /* xr + nw + 1 <= 512: the run ends at byte <= 2045 + 3 */
        EC_UNROLL
        for (int j = 0; j < EC_NW_MAX; j++) {
            if (j < nw) {
                /* 64-bit window r[j+1]:r[j] shifted by sh + 8*(2 - a) bits, at most 40: two funnel shifts */
                const uint32_t l0 = EC_FSHR(r[j], r[j + 1], sh), l1 = EC_FSHR(r[j + 1], r[j + 2], sh);
                t[2][j] = s[j] ^ l0;
                t[1][j] = s[j] ^ EC_FSHR(l0, l1, 8u);
                t[0][j] = s[j] ^ EC_FSHR(l0, l1, 16u);
            }
        }
    } else {
        EC_UNROLL
        for (int j = 0; j < EC_NW_MAX; j++) {
            if (j < nw) {
                EC_UNROLL
                for (int a = 0; a < 3; a++) {
                    int pos = 4 * (w0 + j) - (int)off[a];
                    pos += (pos < 0) ? 2 * 1023 : 0;
                    t[a][j] = s[j] ^ EC_FSHR(RX[pos >> 2], RX[(pos >> 2) + 1], ((uint32_t)pos & 3u) * 8u);
                }
            }
        }
    }
    EC_UNROLL
    for (int j = 0; j < EC_NW_MAX; j++) {
        if (j < nw) {
            EC_UNROLL
            for (int a = 0; a < 3; a++) EC_COUNTS_FULL(t[a][j], p->C[a][j]);
        }
    }
}

/* Edge lane e = 0..11: role e / 3 (0: word 0, 1: word 511, 2: take back data byte off-2, 3: data byte off-1),
 * arm e % 3.  Returns the packed counts of its one (word, arm) entry, *w the data word it belongs to (for the
 * carrier phase in phase 2) and *negative = 1 when the entry is to be subtracted. */
#define EC_EDGE_LANES 12
/* The entry of one edge lane given its role (0..3) and the byte offset o of ITS arm: the caller reads that one offset
 * (a lane only ever needs its own arm's; indexing an array of three by a run-time arm would put it in local memory). */
EC_FN uint32_t ec_epl_edge_entry(const uint32_t* S, const uint32_t* RX, uint32_t o, int role, int* w, int* negative)
{
    const int odd = (int)(o & 1u);
    int word, active = 1;
    uint32_t m;
    if (role == 0) {
        word = 0;
        m = odd ? 0xFFFFFF00u : 0xFFFFFFFFu;                    /* odd: data byte 0 belongs to the skipped word 1022-k */
    } else if (role == 1) {
        word = EC_WORDS - 1;
        m = odd ? 0x000000FFu : 0x0000FFFFu;                    /* bytes 2046, 2047 do not exist; odd: nor does 2045 */
    } else {
        const int d = (int)o - (role == 2 ? 2 : 1);             /* the data bytes of the skipped replica word 1022 */
        active = odd && o >= 3u;
        word = active ? d >> 2 : 0;
        m = active ? 0xFFu << (8 * (d & 3)) : 0u;
    }
    int pos = 4 * word - (int)o;
    pos += (pos < 0) ? 2 * 1023 : 0;
    const int mixed = word < EC_MIX_WORDS;                      /* word 511 is never mixed: stays 0, gps_misc.c:229 */
    const uint32_t x = (mixed ? S[word] : 0u) ^ EC_FSHR(RX[pos >> 2], RX[(pos >> 2) + 1], ((uint32_t)pos & 3u) * 8u);
    *w = word;
    *negative = role >= 2;
    return active ? ec_counts_masked(x, m, mixed) : 0u;
}
EC_FN uint32_t ec_epl_edge_phase1(const uint32_t* S, const uint32_t* RX, const uint32_t off[3], int e, int* w, int* negative)
{
    return ec_epl_edge_entry(S, RX, off[e % 3], e / 3, w, negative);
}

/* Phase 2: per word the carrier phase ph picks the I count (pattern cos[ph]) and the Q count (pattern
 * sin[ph] = cos[(ph+3)&3]); sums of the three arms over the owned words, packed as I | Q << 16 (a whole
 * millisecond is at most 16368 per component, so the halves never carry into each other). */
EC_FN void ec_epl_phase2(uint32_t acc0, uint32_t step32, int w0, int nw, const ec_partial* p, uint32_t acc[3])
{
    EC_UNROLL
    for (int j = 0; j < EC_NW_MAX; j++) {
        if (j < nw) {
            const uint32_t ph = (acc0 + (uint32_t)(w0 + j) * step32) >> 30;
            /* byte selector: result = count[ph] | count[(ph+3)&3] << 16 */
            const uint32_t sel = 0x4040u | ph | (((ph + 3u) & 3u) << 8);
            EC_UNROLL
            for (int a = 0; a < 3; a++) acc[a] += EC_PRMT(p->C[a][j], 0u, sel);
        }
    }
}

/* Phase 2 of an edge lane: its single entry, signed (packed subtraction borrows across the halves, which the
 * sum over all threads undoes - the totals of both halves are non-negative). */
EC_FN uint32_t ec_epl_edge_phase2(uint32_t acc0, uint32_t step32, int w, int negative, uint32_t counts)
{
    const uint32_t ph = (acc0 + (uint32_t)w * step32) >> 30;
    const uint32_t sel = 0x4040u | ph | (((ph + 3u) & 3u) << 8);
    const uint32_t v = EC_PRMT(counts, 0u, sel);
    return negative ? 0u - v : v;
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
