/*
 * gpsb_loop_core.h - the per-channel, per-millisecond tracking step as ONE source compiled twice:
 *
 *   - by gcc into libgpsb_host.so (host/track.c, host/nav.c): the host-resident loop, libm from the host;
 *   - by nvcc into libgpsb_cuda.so (csrc/gpsb_track_loop.cuh): the device-resident loop k_track_run, where a
 *     control thread of the channel's CTA runs exactly these functions between two integrate-and-dump
 *     phases, so a whole run of milliseconds costs one launch and no host round trip.
 *
 * Behaviour follows Firmware/project_main/GPS/tracking.c and nav_data.c of iliasam/STM32F4_SDR_GPS (cited
 * per function).  Everything is integer or IEEE-754 arithmetic that both compilers perform identically
 * (gcc -ffp-contract=off, nvcc --fmad=false, IEEE divide/sqrt), with three exceptions that are handled
 * explicitly:
 *
 *   atan2f / atanf   the device uses lc_atan2f / lc_atanf below: the fdlibm single-precision algorithm that
 *                    glibc 2.39 ships (sysdeps/ieee754/flt-32/s_atanf.c, e_atan2f.c; plain SSE2, no ifunc
 *                    variant), restated operation for operation.  The loop only ever evaluates them on
 *                    ratios of two integers in [-8184, 8184], a finite domain, and tests compare them with
 *                    the host libm over that whole domain (tests/test_loop_core.py).
 *   atan2 (double)   tracking.c:183 evaluates one branch of the Costas discriminator in double.  The device
 *                    uses CUDA's double atan2; after the division by pi the value is rounded to float, and
 *                    the float result is compared with the host libm over the whole input domain on the GPU
 *                    (tests/test_gpu_loop.py, gpsb_selftest_costas).
 *   log10f, rand     log10f only feeds snr_value, which the loop never reads back: the device records the
 *                    two sums and the host finishes the value after the run (lc_resolve_snr).  rand()
 *                    (tracking.c:316) is replaced on batched channels by the same additive-feedback
 *                    generator and default seed glibc's rand() uses, so its state can travel to the device.
 *
 * An all-zero early+late power (0/0 in the DLL, tracking.c:341) makes x86 and the GPU produce different NaN
 * bit patterns; the device loop does not take that step but hands the millisecond back to the host
 * (LC_STOP_DLL_NAN), as it does for anything else it does not implement (a channel not in GPS_TRACKING_RUN).
 */
#ifndef GPSB_LOOP_CORE_H
#define GPSB_LOOP_CORE_H

#include <stdint.h>

#include "gpsb_host.h"

#if defined(__CUDACC__)
#define LC_FN static __device__ __forceinline__
#define LC_FN_BIG static __device__ __noinline__
#define LC_DEVICE_MATH 1
#else
#define LC_FN static inline
#define LC_FN_BIG static
#if defined(LC_EMULATE_DEVICE)
#define LC_DEVICE_MATH 1
#else
#define LC_DEVICE_MATH 0
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>
#endif

/* The loop-filter functions below touch only a few scalar fields of gps_tracking_t.  A C compiler sees them with
 * the record itself (LC_TRK = gps_tracking_t); nvcc sees them as templates over the record type, so that k_track_run
 * can instantiate them on small structs of the same field names that its control threads keep in REGISTERS (a
 * private copy of the whole 152-byte record ends up in local memory: its arrays defeat scalar replacement). */
#if defined(__CUDACC__)
#define LC_TPL template <class LC_TRK>
#else
#define LC_TPL
#define LC_TRK gps_tracking_t
#endif

#define LC_HALF_CHIPS        (2 * PRN_LENGTH)            /* 2046 code phases, acquisition.c:294 */
#define LC_FINE_PER_HALFCHIP 8                            /* GPS_FINE_RATIO, tracking.c:23 */
#define LC_FINE_RANGE        (LC_HALF_CHIPS * LC_FINE_PER_HALFCHIP)   /* 16368 */
#define LC_SLOT_LEN          TRACKING_CH_LENGTH
#define LC_FREQ_POINTS_MAX   25                           /* FREQ_SEARCH_POINTS_MAX_CNT, acquisition.c:12 */
#define LC_MAX_BINS          64
#define LC_FALSE_LOCK_LIMIT  80                           /* tracking.c:14 */
#define LC_SNR_WINDOW        200                          /* tracking.c:26 */
#define LC_LOOP_DT_S         0.001f
#define LC_MS_PER_BIT        20                           /* CODES_IN_BIT, nav_data.c:15 */
#define LC_WORDS_PER_SUBFRAME 10                          /* nav_data.c:17 */
#define LC_POLARITY_TIMEOUT_MS 12000u                     /* two subframes, nav_data.c:22 */
#define LC_WALK_PERIOD_MS    400u                         /* default patience of the slot-phase walk (this library's) */
#define LC_WALK_CONFIDENCE   3u                           /* on-grid edges seen before their slot position is believed */
#define LC_WALK_LEAD_MS      5u                           /* a walk decided at the end of a slot idles after the NEXT slot */
#define LC_IDLE_INDEX        0xFFu                        /* the reference's "dummy" slot index (main.c:146-147, tracking.c:96) */
#define LC_PI                3.14159265358979323846       /* <math.h>'s double M_PI, see host/track.c */

/* why a device-resident run stopped before its last millisecond */
#define LC_STOP_NONE       0
#define LC_STOP_STATE      1      /* channel is not in GPS_TRACKING_RUN: nothing was done for that ms */
#define LC_STOP_DLL_NAN    2      /* early+late power is zero: sums delivered, filters not run */
#define LC_STOP_STARVED    3      /* streaming run: the producer did not deliver a frame in time; done_ms are complete */
#define LC_STOP_PRE_DONE   4      /* k_pretrack_run: pre-track settled in millisecond done_ms - 1, the channel is in GPS_PRE_TRACK_DONE */

/* glibc's default rand(): TYPE_3 additive feedback x^31 + x^3 + 1 over 32-bit words (stdlib/random_r.c) */
typedef struct gpsb_rand31 {
    int32_t r[31];
    uint8_t f, b, ready;
} gpsb_rand31;

/* Cross-call scratch that the reference keeps in file-scope variables.  One shared instance backs
 * the reference-named API; the batched receiver owns one per channel. */
typedef struct gpsb_aux {
    uint32_t freq_hist[LC_MAX_BINS];            /* acq_freq_histogram, acquisition.c:28 (ACQ_COUNT used) */
    uint16_t bin_phases[LC_FREQ_POINTS_MAX];    /* acq_single_freq_phases, acquisition.c:32 */
    uint8_t  bin_count;                         /* acq_single_freq_count, acquisition.c:33 */
    uint16_t pre_best_value;                    /* pre_track_best_corr_value, tracking.c:33 */
    uint16_t pre_best_phase;                    /* pre_track_best_corr_phase, tracking.c:34 */
    int16_t  slot_ip[LC_SLOT_LEN];              /* raw_ip_values, nav_data.c:48 */
    uint8_t  slot_bits[LC_SLOT_LEN];            /* tmp_nav_data, nav_data.c:51 */
    uint32_t slot_start_ticks;                  /* gps_channel_tmp_start_time_ticks, nav_data.c:29 */
    int8_t   last_nav_bit;                      /* observer: bit handed to the word assembler this ms, or -1 */
    uint8_t  process_rand;                      /* 1: false-lock reseed draws from the process-wide rand() */
    uint8_t  snr_pending;                       /* device run: snr_value awaits log10f of the two sums below */
    uint32_t snr_i, snr_q;
    gpsb_rand31 rnd;                            /* private rand() stream of a batched channel */
    /* Slot-phase walk of the batched paths (lc_walk_*, below).  All zero = slots start at ms % 4 == 0 for ever. */
    uint8_t  slot_phase;                        /* slot index of millisecond ms is (ms + slot_phase) % 4 */
    uint8_t  walk_enable;                       /* 1: lc_walk_policy may move the slots of this channel */
    uint8_t  skip_len;                          /* the channel idles in [skip_ms, skip_ms + skip_len): 0 = no walk decided */
    uint8_t  last_flip_pos;                     /* observer: slot position (1..3) of the last bit edge seen on the 20-ms grid */
    uint8_t  flip_cnt[4];                       /* observer: such edges seen per slot position at this slot phase (saturating) */
    uint8_t  walk_armed;                        /* the policy has seen this channel's first slot end */
    uint16_t walk_period_ms;                    /* patience at one slot phase without any bit edge seen (0 = LC_WALK_PERIOD_MS) */
    uint16_t walks;                             /* statistics: idle gaps taken so far */
    uint32_t skip_ms;                           /* first idle millisecond of the pending (or last) walk */
    uint32_t phase_since_ms;                    /* millisecond at which the current slot phase began */
} gpsb_aux;

/* ------------------------------------------------------------------------------------------ bits */
LC_FN int32_t lc_float_bits(float x)
{
#if defined(__CUDACC__)
    return __float_as_int(x);
#else
    int32_t i;
    memcpy(&i, &x, 4);
    return i;
#endif
}
LC_FN float lc_bits_float(int32_t i)
{
#if defined(__CUDACC__)
    return __int_as_float(i);
#else
    float x;
    memcpy(&x, &i, 4);
    return x;
#endif
}

/* ------------------------------------------------------------------------------------------ rand */
LC_FN void lc_rand31_seed(gpsb_rand31* g, uint32_t seed)
{
    int32_t word = seed ? (int32_t)seed : 1;
    g->r[0] = word;
    for (int i = 1; i < 31; i++) {                   /* 16807 * word mod (2^31 - 1) without overflow */
        int64_t hi = word / 127773, lo = word % 127773;
        int64_t next = 16807 * lo - 2836 * hi;
        if (next < 0) next += 2147483647;
        word = (int32_t)next;
        g->r[i] = word;
    }
    g->f = 3;
    g->b = 0;
    g->ready = 1;
    for (int i = 0; i < 310; i++) {                  /* glibc discards ten rounds of the register */
        g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
        g->f = (uint8_t)(g->f == 30 ? 0 : g->f + 1);
        g->b = (uint8_t)(g->b == 30 ? 0 : g->b + 1);
    }
}
LC_FN int lc_rand31_next(gpsb_rand31* g)
{
    if (!g->ready) lc_rand31_seed(g, 1u);
    uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
    g->r[g->f] = (int32_t)v;
    g->f = (uint8_t)(g->f == 30 ? 0 : g->f + 1);
    g->b = (uint8_t)(g->b == 30 ? 0 : g->b + 1);
    return (int)(v >> 1);
}

/* ------------------------------------------------------------------------------------------ arctangent */
/* fdlibm atanf: argument reduction to one of atan(0.5), atan(1), atan(1.5), atan(inf) and an odd
 * polynomial of degree 23 split into even and odd halves.  Constants are the ones glibc carries. */
LC_FN float lc_atanf(float x)
{
    const float hi0 = lc_bits_float(0x3eed6338), hi1 = lc_bits_float(0x3f490fda);
    const float hi2 = lc_bits_float(0x3f7b985e), hi3 = lc_bits_float(0x3fc90fda);
    const float lo0 = lc_bits_float(0x31ac3769), lo1 = lc_bits_float(0x33222168);
    const float lo2 = lc_bits_float(0x33140fb4), lo3 = lc_bits_float(0x33a22168);
    const float a0 = lc_bits_float(0x3eaaaaab), a1 = lc_bits_float((int32_t)0xbe4ccccdu);
    const float a2 = lc_bits_float(0x3e124925), a3 = lc_bits_float((int32_t)0xbde38e38u);
    const float a4 = lc_bits_float(0x3dba2e6e), a5 = lc_bits_float((int32_t)0xbd9d8795u);
    const float a6 = lc_bits_float(0x3d886b35), a7 = lc_bits_float((int32_t)0xbd6ef16bu);
    const float a8 = lc_bits_float(0x3d4bda59), a9 = lc_bits_float((int32_t)0xbd15a221u);
    const float a10 = lc_bits_float(0x3c8569d7);
    const int32_t hx = lc_float_bits(x);
    const int32_t ix = hx & 0x7fffffff;
    float hi, lo;
    int reduced = 1;
    if (ix >= 0x4c000000) {                           /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x;            /* NaN */
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {                            /* |x| < 0.4375 */
        if (ix < 0x31000000) {                        /* |x| < 2^-29 */
            if (1.0e30f + x > 1.0f) return x;
        }
        reduced = 0;
        hi = lo = 0.0f;
    } else {
        x = lc_bits_float(ix);                        /* fabsf */
        if (ix < 0x3f980000) {                        /* |x| < 1.1875 */
            if (ix < 0x3f300000) {                    /* 7/16 <= |x| < 11/16 */
                hi = hi0; lo = lo0;
                x = ((x + x) - 1.0f) / (2.0f + x);
            } else {                                  /* 11/16 <= |x| < 19/16 */
                hi = hi1; lo = lo1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else if (ix < 0x401c0000) {                 /* |x| < 2.4375 */
            hi = hi2; lo = lo2;
            x = (x - 1.5f) / (1.0f + 1.5f * x);
        } else {                                      /* 2.4375 <= |x| < 2^25 */
            hi = hi3; lo = lo3;
            x = -1.0f / x;
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    const float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (!reduced) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -r : r;
}

/* fdlibm atan2f (glibc e_atan2f.c): quadrant bookkeeping around atanf(|y/x|). */
LC_FN float lc_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f;
    const float pi_o_4 = lc_bits_float(0x3f490fdb), pi_o_2 = lc_bits_float(0x3fc90fdb);
    const float pi = lc_bits_float(0x40490fdb), pi_lo = lc_bits_float((int32_t)0xb3bbbd2eu);
    const int32_t hx = lc_float_bits(x), hy = lc_float_bits(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;               /* NaN */
    if (hx == 0x3f800000) return lc_atanf(y);                           /* x == 1 */
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);                   /* 2*sign(x) + sign(y) */
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            if (m == 0) return pi_o_4 + tiny;
            if (m == 1) return -pi_o_4 - tiny;
            if (m == 2) return 3.0f * pi_o_4 + tiny;
            return -3.0f * pi_o_4 - tiny;
        }
        if (m == 0) return 0.0f;
        if (m == 1) return -0.0f;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;                               /* |y/x| > 2^60 */
    else if (hx < 0 && k < -60) z = 0.0f;                                /* |y|/x < -2^60 */
    else z = lc_atanf(lc_bits_float(lc_float_bits(y / x) & 0x7fffffff));
    if (m == 0) return z;
    if (m == 1) return lc_bits_float(lc_float_bits(z) ^ (int32_t)0x80000000u);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

#if LC_DEVICE_MATH
#define LC_ATANF(x)      lc_atanf(x)
#define LC_ATAN2F(y, x)  lc_atan2f((y), (x))
#else
#define LC_ATANF(x)      atanf(x)
#define LC_ATAN2F(y, x)  atan2f((y), (x))
#endif
#define LC_ATAN2(y, x)   atan2((double)(y), (double)(x))

/* The Costas discriminator of tracking.c:180-183 in units of pi.  The reference promotes the float atan2f
 * to double for the division by (double) M_PI on one branch and calls the double atan2 on the other. */
LC_FN float lc_costas_err(int16_t ip, int16_t qp)
{
#if defined(__CUDACC__) && !defined(GPSB_COSTAS_DIVIDE)
    /* On the device the double DIVISION by pi is a double MULTIPLICATION by 1/pi (a divide is ~100 dependent cycles on
     * the carrier thread's chain at every slot index 0).  The two differ in the last bit of the double now and then,
     * never after the rounding to float on this function's domain: ip, qp are sums in [-8184, 8184], and
     * gpsb_l0_loop_math evaluates all 16369 x 16369 pairs on the GPU against the host's division
     * (tests/test_gpu_loop.py, test_loop_math_certificate_whole_domain) - the certificate that already covers the
     * arctangents. */
    if (ip > 0) return (float)(LC_ATAN2F((float)qp, (float)ip) * 0.31830988618379067154);
    return (float)(LC_ATAN2((float)-qp, (float)-ip) * 0.31830988618379067154);
#else
    if (ip > 0) return (float)(LC_ATAN2F((float)qp, (float)ip) / LC_PI);
    return (float)(LC_ATAN2((float)-qp, (float)-ip) / LC_PI);
#endif
}

/* One arm of the frequency discriminator, tracking.c:232-233 */
LC_FN float lc_fll_angle(int16_t ip, int16_t qp)
{
    return (ip == 0) ? (float)(LC_PI / 2) : LC_ATANF((float)qp / (float)ip);
}

/* ------------------------------------------------------------------------------------------ carrier NCO */
/* NCO word: fp32 divide then truncation (gps_misc.c:219, :199, :250). */
LC_FN uint32_t lc_nco_step(float freq_hz)
{
    float q = freq_hz / IF_NCO_STEP_HZ;
    return (uint32_t)q;
}
LC_FN uint32_t lc_nco_step32(float freq_hz)
{
    uint64_t wide = (uint64_t)lc_nco_step(freq_hz) * 32u;   /* 32 samples per mixed word, gps_misc.c:220 */
    return (uint32_t)wide;
}
/* Catch-up of the carrier NCO over skipped milliseconds (gps_misc.c:196-204): the reference advances by
 * acc_step*16368 per skipped ms although a processed ms advances by 511*32 samples - reproduced as is. */
LC_TPL LC_FN void lc_rewind_if_phase(LC_TRK* trk, uint8_t steps)
{
    uint32_t per_sample = lc_nco_step((float)IF_FREQ_HZ + trk->if_freq_offset_hz);
    uint64_t advance = (uint64_t)per_sample * BITS_IN_PRN * steps;
    trk->if_freq_accum += (uint32_t)advance;
}

/* ------------------------------------------------------------------------------------------ plan */
/* tracking.c:115-130: byte offsets of the three arms and the sub-byte replica shift */
LC_FN void lc_arm_offsets(float code_phase_fine, gpsb_epl_req* rq)
{
    int16_t fine = (int16_t)code_phase_fine;
    uint16_t prompt = (uint16_t)(fine / LC_FINE_PER_HALFCHIP);
    uint16_t early = (uint16_t)(prompt - 1);
    uint16_t late = (uint16_t)(prompt + 1);
    if (early >= LC_HALF_CHIPS) early = LC_HALF_CHIPS - 1;
    if (late >= LC_HALF_CHIPS) late = 0;
    /* A code phase just above 16368 (the DLL's "16368 - x" wrap of a negative x, tracking.c:353-358) gives
     * a prompt offset of exactly 2046; the reference's pointer arithmetic then runs its second loop over
     * the whole buffer from byte 0 (gps_misc.c:57,73-81), i.e. it computes offset 0. */
    if (prompt >= LC_HALF_CHIPS) prompt = (uint16_t)(prompt - LC_HALF_CHIPS);
    rq->off_bits = (uint16_t)(fine & (LC_FINE_PER_HALFCHIP - 1));
    rq->off_e = early;
    rq->off_p = prompt;
    rq->off_l = late;
}

/* tracking.c:92-123 for a channel in GPS_TRACKING_RUN: what to correlate this millisecond.  Two independent
 * halves - the carrier NCO words depend on the PLL/FLL state only, the code offsets on the DLL state only - so
 * the device-resident loop can run them on different threads. */
LC_TPL LC_FN void lc_plan_carrier(LC_TRK* t, uint8_t prn, uint32_t now, uint32_t frame_ms, gpsb_epl_req* rq)
{
    /* the NCO word first: its divide is the long dependent chain of this function, and nothing below changes
     * if_freq_offset_hz */
    rq->step32 = lc_nco_step32((float)IF_FREQ_HZ + t->if_freq_offset_hz);
    uint32_t gap = now - t->prev_track_timestamp;
    t->prev_track_timestamp = now;
    if (gap > 50) gap = 1;                                         /* first step after start-up */
    if (gap != 1) lc_rewind_if_phase(t, (uint8_t)(gap - 1));       /* ms this channel did not see */
    rq->sv_slot = prn;
    rq->ms_index = frame_ms;
    rq->acc0 = t->if_freq_accum;
    t->if_freq_accum += 511u * rq->step32;                         /* what the mixer leaves behind, gps_misc.c:261-273 */
}
LC_TPL LC_FN void lc_plan_code(const LC_TRK* t, gpsb_epl_req* rq) { lc_arm_offsets(t->code_phase_fine, rq); }
LC_FN void lc_trk_plan_run(gps_ch_t* ch, uint32_t now, uint32_t frame_ms, gpsb_epl_req* rq)
{
    lc_plan_carrier(&ch->tracking_data, ch->prn, now, frame_ms, rq);
    lc_plan_code(&ch->tracking_data, rq);
}

/* ------------------------------------------------------------------------------------------ pre-track */
#define LC_PRE_TRACK_ZONE    30                                   /* GPS_PRE_TRACK_ZONE, tracking.c:17 */
#define LC_PRE_TRACK_PER_MS  (LC_PRE_TRACK_ZONE / LC_SLOT_LEN)    /* 7 offsets per ms, tracking.c:20 */

/* tracking.c:52-72: load the acquisition result into a +-15 half-chip pre-track window */
LC_FN void lc_pre_arm(gps_ch_t* ch)
{
    gps_tracking_t* t = &ch->tracking_data;
    uint16_t lo = (uint16_t)(ch->acq_data.found_code_phase - LC_PRE_TRACK_ZONE / 2);
    uint16_t hi = (uint16_t)(ch->acq_data.found_code_phase + LC_PRE_TRACK_ZONE / 2);
    if (lo > LC_HALF_CHIPS) lo = 0;
    if (hi > LC_HALF_CHIPS) hi = LC_HALF_CHIPS;
    t->code_search_start = lo;
    t->code_search_stop = hi;
    t->if_freq_offset_hz = (float)ch->acq_data.found_freq_offset_hz;
    t->pre_track_count = 0;
    for (unsigned i = 0; i < PRE_TRACK_POINTS_MAX_CNT; i++) t->pre_track_phases[i] = 0;
    t->state = GPS_PRE_TRACK_RUN;
}

/* tracking.c:411-415: the seven offsets of slot index `index` (0..3); *last is exclusive */
LC_FN void lc_pre_window(const gps_tracking_t* t, uint8_t index, uint16_t* first, uint16_t* last)
{
    unsigned lo = (uint16_t)(t->code_search_start + index * LC_PRE_TRACK_PER_MS);
    unsigned hi = (uint16_t)(lo + LC_PRE_TRACK_PER_MS);
    if (hi > LC_HALF_CHIPS) hi = LC_HALF_CHIPS;
    *first = (uint16_t)lo;
    *last = (uint16_t)hi;
}

/* tracking.c:459-499: most frequent phase among the collected slot winners (longest run of equal values after
 * sorting; a phase of 0 means "nothing found").  The reference sorts with qsort; any sort of plain integers leaves
 * the same array behind, so an insertion sort (at most 30 values) serves the host and the device alike. */
LC_FN void lc_pre_settle(gps_ch_t* ch, uint8_t n)
{
    gps_tracking_t* t = &ch->tracking_data;
    for (uint8_t i = 1; i < n; i++) {
        const uint16_t v = t->pre_track_phases[i];
        int k = (int)i - 1;
        for (; k >= 0 && t->pre_track_phases[k] > v; k--) t->pre_track_phases[k + 1] = t->pre_track_phases[k];
        t->pre_track_phases[k + 1] = v;
    }
    uint8_t run = 0;
    uint16_t best_run = 0, winner = 0;
    for (uint8_t i = 1; i < n; i++) {
        if (t->pre_track_phases[i] == t->pre_track_phases[i - 1]) {
            run++;
        } else {
            if (run > best_run) { best_run = run; winner = t->pre_track_phases[i - 1]; }
            run = 0;
        }
    }
    if (run > best_run) { best_run = run; winner = t->pre_track_phases[n - 1]; }
    if (winner) {
        t->code_phase_fine = (float)(winner * LC_FINE_PER_HALFCHIP);
        t->state = GPS_PRE_TRACK_DONE;
    }
}

/* tracking.c:417-449.  (max, phase) = the window's maximum and its first position: scanning the window with a strict
 * '>' against the running best is the same as comparing the window maximum once. */
LC_FN void lc_pre_finish(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, uint16_t max, uint16_t phase)
{
    gps_tracking_t* t = &ch->tracking_data;
    if ((int16_t)max > (int16_t)aux->pre_best_value) {
        aux->pre_best_value = max;
        aux->pre_best_phase = phase;
    }
    if (index != LC_SLOT_LEN - 1) return;
    t->pre_track_phases[t->pre_track_count] = aux->pre_best_phase;   /* note: the phase is NOT reset per slot */
    t->pre_track_count++;
    if (t->pre_track_count > PRE_TRACK_POINTS_MAX_CNT - 10) lc_pre_settle(ch, t->pre_track_count);
    if (t->pre_track_count >= PRE_TRACK_POINTS_MAX_CNT) {
        t->pre_track_count = 0;
        for (unsigned i = 0; i < PRE_TRACK_POINTS_MAX_CNT; i++) t->pre_track_phases[i] = 0;
    }
    aux->pre_best_value = 0;
}

/* ------------------------------------------------------------------------------------------ loop filters */
/* 1 when the DLL discriminator of this millisecond would be 0/0 */
LC_FN int lc_dll_is_degenerate(const int16_t iq[6])
{
    return (iq[0] | iq[1] | iq[4] | iq[5]) == 0;          /* a sum of four squares is zero iff all four sums are */
}

/* tracking.c:333-393 */
LC_TPL LC_FN void lc_dll_update(LC_TRK* t, int16_t ie, int16_t qe, int16_t il, int16_t ql)
{
    int32_t early = (int32_t)ie * ie + (int32_t)qe * qe;
    int32_t late = (int32_t)il * il + (int32_t)ql * ql;
    float err = (float)(early - late) / (float)(early + late);
    err = -err;
    t->code_phase_fine += (TRACKING_DLL1_C1 * (err - t->dll_code_err) + TRACKING_DLL1_C2 * LC_LOOP_DT_S * err);

    uint8_t wrapped = 0;
    if (t->code_phase_fine < 0.0f) {
        t->code_phase_fine = (float)LC_FINE_RANGE - t->code_phase_fine;      /* sic: minus a negative */
        wrapped = 1;
    } else if (t->code_phase_fine > (float)LC_FINE_RANGE) {
        t->code_phase_fine = t->code_phase_fine - (float)LC_FINE_RANGE;
        wrapped = 1;
    }
#if (ENABLE_CODE_FILTER)
    if (wrapped) {
        t->code_phase_fine_filt = -1.0f;                  /* averaging across a wrap is meaningless: stop */
    } else if (t->code_phase_fine_filt >= 0.0f) {
        t->code_phase_fine_filt += t->code_phase_fine;
        t->code_filt_cnt++;
    }
#endif
    t->dll_code_err = err;
}

/* Fold an angle difference back into [-pi/2, pi/2] the way the reference does: `if (x > M_PI / 2) x = M_PI - x;
 * if (x < -M_PI / 2) x = -M_PI - x;` with the DOUBLE M_PI (tracking.c:189-192, 238-247).  The comparisons promote
 * the float to double; since (float)(pi/2) = 0x3FC90FDB is the smallest float above the double pi/2, "x > pi/2 in
 * double" is exactly "x >= 0x3FC90FDB in float" - same truth value for every float, one compare instead of a
 * conversion and a double compare.  The reflections themselves stay in double. */
LC_FN_BIG float lc_reflect_half_pi(float x)
{
    const float above_half_pi = lc_bits_float(0x3fc90fdb);
    if (x >= above_half_pi) x = (float)(LC_PI - x);
    if (x <= -above_half_pi) x = (float)(-LC_PI - x);
    return x;
}
LC_FN float lc_fold_half_pi(float x)
{
    /* Inside (-pi/2, pi/2) - nearly always - nothing happens.  The reflections (double arithmetic) sit behind a real
     * call so that a compiler cannot evaluate them speculatively on the serial path of the device-resident loop. */
    const float above_half_pi = lc_bits_float(0x3fc90fdb);
    if (x >= above_half_pi || x <= -above_half_pi) return lc_reflect_half_pi(x);
    return x;
}

/* tracking.c:175-209.  The reference evaluates the discriminator on every millisecond and uses it on slot
 * index 0 only; it is a pure function of (ip, qp), so it is evaluated only where it is used. */
LC_TPL LC_FN void lc_pll_update(LC_TRK* t, int period_sync_ok, uint8_t index, int16_t ip, int16_t qp)
{
    if (index != 0) return;
    float err = lc_costas_err(ip, qp);
    float delta = lc_fold_half_pi(err - t->pll_code_err);
    if (period_sync_ok)
        t->if_freq_offset_hz -= TRACKING_PLL2_C1 * delta + (TRACKING_PLL2_C2 * LC_LOOP_DT_S * err);
    else
        t->if_freq_offset_hz -= TRACKING_PLL1_C1 * delta + (TRACKING_PLL1_C2 * LC_LOOP_DT_S * err);
    t->pll_code_err = err;
}

#if defined(__CUDACC__) || defined(LC_EMULATE_DEVICE)
LC_FN int lc_rand(gpsb_aux* aux) { return lc_rand31_next(&aux->rnd); }
#else
int hx_rand(gpsb_aux* aux);      /* host/bind.c: override hook, process-wide rand(), or the private stream */
LC_FN int lc_rand(gpsb_aux* aux) { return hx_rand(aux); }
#endif

/* tracking.c:261-327: two or more sign flips of IP inside one 4-ms slot cannot be data; count them and,
 * after a long bad streak, jump the carrier to a random frequency at least 200 Hz away. */
LC_TPL LC_FN void lc_lock_check(LC_TRK* t, gpsb_aux* aux, int16_t found_freq_offset_hz, uint8_t index, int16_t ip)
{
    if (index >= LC_SLOT_LEN) return;
    /* pll_check_buf[index] = ip, spelled with constant subscripts (LC_SLOT_LEN is 4) so that a register-resident
     * record never needs an address */
    if (index == 0) t->pll_check_buf[0] = ip;
    else if (index == 1) t->pll_check_buf[1] = ip;
    else if (index == 2) t->pll_check_buf[2] = ip;
    else t->pll_check_buf[3] = ip;
    if (index < LC_SLOT_LEN - 1) return;

    const uint8_t s0 = t->pll_check_buf[0] > 0, s1 = t->pll_check_buf[1] > 0;
    const uint8_t s2 = t->pll_check_buf[2] > 0, s3 = t->pll_check_buf[3] > 0;
    uint8_t flips = (uint8_t)((s0 != s1) + (s1 != s2) + (s2 != s3));
    if (flips > 1) {
        if (++t->pll_bad_state_cnt > 10) t->pll_bad_state_cnt = 10;
    } else if (t->pll_bad_state_cnt > 0) {
        t->pll_bad_state_cnt--;
    }
    if (t->pll_bad_state_cnt > 9) t->pll_bad_state_master_cnt++;
    else if (t->pll_bad_state_cnt == 0) t->pll_bad_state_master_cnt = 0;

    if (t->pll_bad_state_master_cnt > LC_FALSE_LOCK_LIMIT) {
        t->pll_bad_state_master_cnt = 0;
        t->pll_bad_state_cnt = 0;
        int16_t candidate, away;
        do {
            uint16_t r = (uint16_t)(lc_rand(aux) % ACQ_SEARCH_STEP_HZ);
            candidate = (int16_t)(found_freq_offset_hz - r + (ACQ_SEARCH_STEP_HZ / 2));
            away = (int16_t)((int16_t)t->if_freq_offset_hz - candidate);
        } while ((away < 0 ? -away : away) < 200);
        t->if_freq_offset_hz = (float)candidate;
    }
}

/* The angle of the PREVIOUS prompt sample is, from slot index 2 on, the angle computed one millisecond
 * earlier from the same two integers; a caller may hand it back to save one divide + arctangent. */
typedef struct lc_angle_cache {
    int16_t i, q;
    uint8_t valid;
    float angle;
} lc_angle_cache;

/* tracking.c:214-256 */
LC_TPL LC_FN void lc_fll_update(LC_TRK* t, gpsb_aux* aux, int16_t found_freq_offset_hz, uint8_t index, int16_t ip,
                         int16_t qp, lc_angle_cache* cache)
{
    lc_lock_check(t, aux, found_freq_offset_hz, index, ip);
    if (index == 0) {                                 /* first ms of a slot: previous sample is from another time */
        t->fll_old_i = ip;
        t->fll_old_q = qp;
        return;
    }
    float now = lc_fll_angle(ip, qp);
    float before;
    if (cache && cache->valid && cache->i == t->fll_old_i && cache->q == t->fll_old_q) before = cache->angle;
    else before = lc_fll_angle(t->fll_old_i, t->fll_old_q);
    if (cache) {
        cache->i = ip;
        cache->q = qp;
        cache->angle = now;
        cache->valid = 1;
    }
    float rot = lc_fold_half_pi(now - before);
    float rot_change = lc_fold_half_pi(rot - t->fll_err);
    float step_hz = TRACKING_FLL1_C1 * LC_LOOP_DT_S * rot_change + (TRACKING_FLL1_C2 * LC_LOOP_DT_S * rot);
    t->if_freq_offset_hz -= step_hz;
    t->fll_old_i = ip;
    t->fll_old_q = qp;
    t->fll_err = rot;
}

/* First half of tracking.c:140-169: everything the NEXT millisecond's correlation depends on. */
LC_FN void lc_finish_loops(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, const int16_t iq[6])
{
    gps_tracking_t* t = &ch->tracking_data;
    lc_dll_update(t, iq[0], iq[1], iq[4], iq[5]);
    lc_pll_update(t, ch->nav_data.period_sync_ok_flag, index, iq[2], iq[3]);
    lc_fll_update(t, aux, ch->acq_data.found_freq_offset_hz, index, iq[2], iq[3], (lc_angle_cache*)0);
}

/* ------------------------------------------------------------------------------------------ nav bits */
/* 1 when the first eight buffered bits equal the preamble 10001011 (flip = 0) or its complement (flip = 1),
 * nav_data.c:26 */
LC_FN int lc_starts_with_preamble(const gps_nav_data_t* n, uint8_t flip)
{
    const uint32_t preamble = 0xD1u;                       /* bit i = element i of {1,0,0,0,1,0,1,1} */
    for (unsigned i = 0; i < 8; i++)
        if (n->word_buf[i] != (((preamble >> i) & 1u) ^ flip)) return 0;
    return 1;
}

/* nav_data.c:409-426: copy the 30 buffered bits to bit positions word_cnt*30.. of the subframe image
 * (bit i at byte i/8, bit i%8) and remember D29/D30 for the next word's parity. */
LC_FN void lc_store_word(gps_nav_data_t* n)
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
 * the previous D30 as the reference does (nav_data.c:433-453), which also changes what lc_store_word saves.
 * Each parity equation is a 24-bit mask over d1..d24 (bit k-1 = d_k). */
LC_FN int lc_parity_ok(gps_nav_data_t* n)
{
    uint8_t* w = n->word_buf;                         /* ICD bit d[k] is w[k-1] */
    uint32_t d = 0;
    for (unsigned k = 1; k < 25; k++) {
        w[k - 1] ^= n->old_D30;
        d |= (uint32_t)(w[k - 1] & 1u) << (k - 1);
    }
    /* d_k lists of the six equations, table 20-XIV */
    const uint32_t m25 = (1u<<0)|(1u<<1)|(1u<<2)|(1u<<4)|(1u<<5)|(1u<<9)|(1u<<10)|(1u<<11)|(1u<<12)|(1u<<13)|(1u<<16)|(1u<<17)|(1u<<19)|(1u<<22);
    const uint32_t m26 = (1u<<1)|(1u<<2)|(1u<<3)|(1u<<5)|(1u<<6)|(1u<<10)|(1u<<11)|(1u<<12)|(1u<<13)|(1u<<14)|(1u<<17)|(1u<<18)|(1u<<20)|(1u<<23);
    const uint32_t m27 = (1u<<0)|(1u<<2)|(1u<<3)|(1u<<4)|(1u<<6)|(1u<<7)|(1u<<11)|(1u<<12)|(1u<<13)|(1u<<14)|(1u<<15)|(1u<<18)|(1u<<19)|(1u<<21);
    const uint32_t m28 = (1u<<1)|(1u<<3)|(1u<<4)|(1u<<5)|(1u<<7)|(1u<<8)|(1u<<12)|(1u<<13)|(1u<<14)|(1u<<15)|(1u<<16)|(1u<<19)|(1u<<20)|(1u<<22);
    const uint32_t m29 = (1u<<0)|(1u<<2)|(1u<<4)|(1u<<5)|(1u<<6)|(1u<<8)|(1u<<9)|(1u<<13)|(1u<<14)|(1u<<15)|(1u<<16)|(1u<<17)|(1u<<20)|(1u<<21)|(1u<<23);
    const uint32_t m30 = (1u<<2)|(1u<<4)|(1u<<5)|(1u<<7)|(1u<<8)|(1u<<9)|(1u<<10)|(1u<<12)|(1u<<14)|(1u<<18)|(1u<<21)|(1u<<22)|(1u<<23);
    const uint32_t masks[6] = {m25, m26, m27, m28, m29, m30};
    const uint8_t seed[6] = {n->old_D29, n->old_D30, n->old_D29, n->old_D30, n->old_D30, n->old_D29};
    for (unsigned p = 0; p < 6; p++) {
        uint32_t x = d & masks[p];
        x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
        uint8_t v = (uint8_t)(seed[p] ^ (x & 1u));
        if (w[24 + p] != v) return 0;
    }
    return 1;
}

/* nav_data.c:356-380: time stamp (ms counter) of the bit edge that started the subframe just completed */
LC_FN void lc_stamp_subframe(gps_nav_data_t* n, uint32_t now)
{
    if (!n->accurate_swap_ok) return;
    uint32_t edge = (now / LC_MS_PER_BIT) * LC_MS_PER_BIT + n->accurate_swap_time;
    if ((int32_t)(now - edge) < 0) edge -= LC_MS_PER_BIT;   /* the edge estimate was late: use the previous one */
    n->subframe_cnt++;
    n->last_subframe_time = edge;
}

LC_FN void lc_clear_word(gps_nav_data_t* n)
{
    for (unsigned i = 0; i < GPS_NAV_WORD_LENGTH; i++) n->word_buf[i] = 0;
}

/* ------------------------------------------------------------------------------------------ ephemeris fields */
/* nav_data_decode.c:144-181.  The subframe image keeps navigation bit i (0 = first bit of the TLM word) at byte
 * i/8, bit i%8 (nav_data.c:409-426); fields are MSB first.  lc_field reads `len` bits from bit `pos`; a field split
 * over two words is its upper part shifted up by the length of the lower one (getbitu2 / getbits2). */
LC_FN uint32_t lc_field(const uint8_t* sf, int pos, int len)
{
    uint32_t v = 0;
    for (int i = pos; i < pos + len; i++) v = (v << 1) | ((uint32_t)(sf[i >> 3] >> (i & 7)) & 1u);
    return v;
}
LC_FN int32_t lc_field_signed(const uint8_t* sf, int pos, int len)
{
    uint32_t v = lc_field(sf, pos, len);
    if (len > 0 && len < 32 && (v >> (len - 1)) & 1u) v |= ~0u << len;      /* two's complement of `len` bits */
    return (int32_t)v;
}
LC_FN uint32_t lc_field2(const uint8_t* sf, int p1, int l1, int p2, int l2)
{
    return (lc_field(sf, p1, l1) << l2) + lc_field(sf, p2, l2);
}
LC_FN int32_t lc_field2_signed(const uint8_t* sf, int p1, int l1, int p2, int l2)
{
    /* the sign lives in the upper part; a non-negative value is the plain concatenation */
    if (lc_field(sf, p1, 1)) return (int32_t)(((uint32_t)lc_field_signed(sf, p1, l1) << l2) + lc_field(sf, p2, l2));
    return (int32_t)lc_field2(sf, p1, l1, p2, l2);
}

/* RTK/rtklib_common.c:32-43: GPS week + seconds of week -> time stamp (whole seconds since the Unix epoch, fraction) */
LC_FN gtime_t lc_gpst2time(int week, double sec)
{
    gtime_t t;
    if (sec < -1e9 || 1e9 < sec) sec = 0.0;
    t.time = (time_t)315964800 + (time_t)(86400 * 7 * week + (int)sec);    /* the reference adds in int: reproduce */
    t.sec = sec - (int)sec;
    return t;
}

#define LC_GPS_BUILD_WEEK 2290                    /* config.h:73; resolves the 10-bit week number, nav_data_decode.c:184 */
/* Scale factors as the reference spells them (rtk_common.h:9-32): 16-digit decimals, three of which (2^-33, 2^-43,
 * 2^-55) do NOT round to the power of two they stand for but to the double one ulp below it - parity needs those. */
#define LC_2P_5  0.03125
#define LC_2P_19 1.907348632812500E-06
#define LC_2P_29 1.862645149230957E-09
#define LC_2P_31 4.656612873077393E-10
#define LC_2P_33 1.164153218269348E-10
#define LC_2P_43 1.136868377216160E-13
#define LC_2P_55 2.775557561562891E-17
#define LC_2P(n) LC_2P_##n
#define LC_SC2RAD 3.1415926535898                 /* semicircles -> radians, rtk_common.h:45 */

/* nav_data_decode.c:33-141: the ephemeris / clock fields of subframes 1..3 (IS-GPS-200 figure 20-1), transmit time and
 * bookkeeping for 4 and 5.  Fields go straight into the channel's record, like the reference.  Returns the subframe id.
 * Everything is integer -> double conversion and IEEE double multiplication in the reference's order of operations:
 * identical bits from gcc (-ffp-contract=off) and nvcc (--fmad=false). */
LC_FN_BIG uint8_t lc_decode_subframe(gps_ch_t* ch)
{
    const uint8_t* sf = ch->nav_data.subframe_data;
    sdreph_t* d = &ch->eph_data;
    eph_t* e = &d->eph;
    const uint32_t id = lc_field(sf, 49, 3);
    e->sat = ch->prn;
    if (id >= 1 && id <= 5) d->tow_gpst = lc_field(sf, 30, 17) * 6.0;      /* HOW: time of week of the next subframe */
    if (id == 1) {
        const int week = (int)lc_field(sf, 60, 10) + 1024;
        e->code = (int)lc_field(sf, 70, 2);
        e->sva = (int)lc_field(sf, 72, 4);
        e->svh = (int)lc_field(sf, 76, 6);
        e->iodc = (int)lc_field2(sf, 82, 2, 210, 8);
        e->flag = (int)lc_field(sf, 90, 1);
        e->tgd[0] = lc_field_signed(sf, 196, 8) * LC_2P(31);
        const double toc = lc_field(sf, 218, 16) * 16.0;
        e->f2 = lc_field_signed(sf, 240, 8) * LC_2P(55);
        e->f1 = lc_field_signed(sf, 248, 16) * LC_2P(43);
        e->f0 = lc_field_signed(sf, 270, 22) * LC_2P(31);
        e->week = week + (LC_GPS_BUILD_WEEK - week + 512) / 1024 * 1024;
        d->week_gpst = e->week;
        e->ttr = lc_gpst2time(e->week, d->tow_gpst);
        e->toc = lc_gpst2time(e->week, toc);
    } else if (id == 2) {
        e->iode = (int)lc_field(sf, 60, 8);
        e->crs = lc_field_signed(sf, 68, 16) * LC_2P(5);
        e->deln = lc_field_signed(sf, 90, 16) * LC_2P(43) * LC_SC2RAD;
        e->M0 = lc_field2_signed(sf, 106, 8, 120, 24) * LC_2P(31) * LC_SC2RAD;
        e->cuc = lc_field_signed(sf, 150, 16) * LC_2P(29);
        e->e = lc_field2(sf, 166, 8, 180, 24) * LC_2P(33);
        e->cus = lc_field_signed(sf, 210, 16) * LC_2P(29);
        const double root_a = lc_field2(sf, 226, 8, 240, 24) * LC_2P(19);
        e->toes = lc_field(sf, 270, 16) * 16.0;
        e->fit = lc_field(sf, 286, 1);
        e->A = root_a * root_a;
        e->toe = lc_gpst2time(e->week, e->toes);
    } else if (id == 3) {
        e->cic = lc_field_signed(sf, 60, 16) * LC_2P(29);
        e->OMG0 = lc_field2_signed(sf, 76, 8, 90, 24) * LC_2P(31) * LC_SC2RAD;
        e->cis = lc_field_signed(sf, 120, 16) * LC_2P(29);
        e->i0 = lc_field2_signed(sf, 136, 8, 150, 24) * LC_2P(31) * LC_SC2RAD;
        e->crc = lc_field_signed(sf, 180, 16) * LC_2P(5);
        e->omg = lc_field2_signed(sf, 196, 8, 210, 24) * LC_2P(31) * LC_SC2RAD;
        e->OMGd = lc_field_signed(sf, 240, 24) * LC_2P(43) * LC_SC2RAD;
        e->iode = (int)lc_field(sf, 270, 8);
        e->idot = lc_field_signed(sf, 278, 14) * LC_2P(43) * LC_SC2RAD;
    }
    if (id >= 1 && id <= 4) d->cnt++;                         /* subframe 5 does not count, nav_data_decode.c:137-141 */
    if (id >= 1 && id <= 5) {
        d->received_mask |= (uint8_t)(1u << (id - 1));
        d->received_mask_proc |= (uint8_t)(1u << (id - 1));
    }
    d->sub_cnt++;
    return (uint8_t)id;
}

/* nav_data.c:257-352 */
LC_FN_BIG void lc_nav_word_bit(gps_ch_t* ch, uint8_t new_bit, uint32_t now)
{
    gps_nav_data_t* n = &ch->nav_data;
    if (n->word_cnt == 0) {                                   /* hunting for a preamble */
        for (unsigned i = 0; i + 1 < GPS_NAV_WORD_LENGTH; i++) n->word_buf[i] = n->word_buf[i + 1];
        n->word_buf[GPS_NAV_WORD_LENGTH - 1] = new_bit;
        if (lc_starts_with_preamble(n, 0)) {
            lc_store_word(n);
            n->word_cnt = 1;
            n->word_bit_cnt = 0;
            n->inv_preabmle_cnt = 0;
        }
        if (n->polarity_found == 0 && n->word_cnt == 0) {    /* 0/180 degree ambiguity of the Costas loop */
            if (lc_starts_with_preamble(n, 1)) n->inv_preabmle_cnt++;
            if (n->inv_preabmle_cnt >= 2) n->inv_polarity_flag = 1;
        }
        if (n->polarity_found) {
            if (now - n->word_detection_timestamp > LC_POLARITY_TIMEOUT_MS) {
                n->word_detection_timestamp = now;
                n->polarity_found = 0;
                n->inv_polarity_flag = 0;
            }
        }
        return;
    }
    n->word_buf[n->word_bit_cnt++] = new_bit;                 /* collecting words 2..10 */
    if (n->word_bit_cnt < GPS_NAV_WORD_LENGTH) return;
    if (!lc_parity_ok(n)) {
        n->word_cnt = 0;
        lc_clear_word(n);
        return;
    }
    n->word_cnt_test++;
    lc_store_word(n);
    n->word_cnt++;
    n->word_bit_cnt = 0;
    n->word_detection_timestamp = now;
    n->polarity_found = 1;
    if (n->word_cnt == LC_WORDS_PER_SUBFRAME) {
        lc_decode_subframe(ch);                               /* nav_data.c:335 */
        lc_stamp_subframe(n, now);
        n->word_cnt = 0;
        n->new_subframe_flag = 1;
        lc_clear_word(n);
    }
}

/* nav_data.c:223-252: close a data bit when the position inside the 20-ms period wraps */
LC_FN void lc_count_ms_into_bit(gps_ch_t* ch, gpsb_aux* aux, uint8_t ms_bit, uint32_t now)
{
    gps_nav_data_t* n = &ch->nav_data;
    uint8_t pos = (uint8_t)((now - n->old_swap_time) % LC_MS_PER_BIT);
    if (pos < n->old_reminder) {
        uint8_t bit = n->last_bit_pos_cnt > n->last_bit_neg_cnt;
        aux->last_nav_bit = (int8_t)bit;
        lc_nav_word_bit(ch, bit, now);
        n->last_bit_pos_cnt = 0;
        n->last_bit_neg_cnt = 0;
    }
    if (ms_bit) n->last_bit_pos_cnt++;
    else n->last_bit_neg_cnt++;
    n->old_reminder = pos;
}

LC_FN int lc_iabs(int v) { return v < 0 ? -v : v; }

/* nav_data.c:145-218: decide whether the single sign flip seen at slot position 2 really happened
 * between samples 0/1 or 1/2, from the prompt amplitudes (the circular correlator smears an edge over
 * the millisecond in which it falls). */
LC_FN_BIG void lc_refine_edge(gps_ch_t* ch, const gpsb_aux* aux)
{
    gps_nav_data_t* n = &ch->nav_data;
    const int16_t* v = aux->slot_ip;
    if (lc_iabs(v[1]) > lc_iabs(v[0])) return;
    if (v[3] == 0) return;
    float ends = (float)lc_iabs(v[0]) / (float)lc_iabs(v[3]);
    if (ends > 1.5f || ends < 0.7f) return;

    int16_t chip = (int16_t)((int16_t)ch->tracking_data.code_phase_fine / 16);
    if (chip < 0 || chip > PRN_LENGTH) return;

    uint8_t edge_at = 0;
    if (chip < PRN_LENGTH / 4 || chip > PRN_LENGTH * 3 / 4) {
        if (v[1] == 0) return;
        float head = (float)lc_iabs(v[0]) / (float)lc_iabs(v[1]);
        if (head > 1.5f || head < 0.7f) return;
        edge_at = (chip < PRN_LENGTH / 4) ? 2 : 1;
    } else {
        uint16_t step_a = (uint16_t)lc_iabs(v[0] - v[1]);
        uint16_t step_b = (uint16_t)lc_iabs(v[2] - v[3]);
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
    n->accurate_swap_time = (uint8_t)((aux->slot_start_ticks + edge_at) % LC_MS_PER_BIT);
    n->accurate_swap_ok = 1;
}

/* nav_data.c:46-138 without its last step: returns 1 when the bit edge found at the end of this slot still has
 * to be refined by lc_refine_edge (the only part that reads the code phase the DLL produced this millisecond). */
LC_FN int lc_nav_new_code(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t new_i, uint32_t now)
{
    gps_nav_data_t* n = &ch->nav_data;
    aux->last_nav_bit = -1;
    if (index >= LC_SLOT_LEN) return 0;
    uint8_t ms_bit = (uint8_t)((new_i > 0) ^ (n->inv_polarity_flag != 0));
    aux->slot_bits[index] = ms_bit;
    aux->slot_ip[index] = new_i;
    if (index == 0) aux->slot_start_ticks = now;
    if (n->period_sync_ok_flag == 1) lc_count_ms_into_bit(ch, aux, ms_bit, now);
    if (index < LC_SLOT_LEN - 1) return 0;

    /* end of the 4-ms slot: exactly one sign flip is a candidate bit edge */
    uint8_t flips = 0, flip_pos = 0;
    for (uint8_t i = 1; i < LC_SLOT_LEN; i++)
        if (aux->slot_bits[i] != aux->slot_bits[i - 1]) { flips++; flip_pos = i; }
    if (flips != 1) return 0;

    uint32_t edge = aux->slot_start_ticks + flip_pos;
    uint8_t phase = (uint8_t)((edge - n->old_swap_time) % LC_MS_PER_BIT);
    if (phase < 2 || phase == LC_MS_PER_BIT - 1) {            /* a multiple of 20 ms since the last edge */
        if (n->right_period_cnt < 10) n->right_period_cnt++;
        if (n->right_period_cnt > 8) n->period_sync_ok_flag = 1;
        aux->last_flip_pos = flip_pos;                        /* observers for lc_walk_policy; not reference state */
        if (aux->flip_cnt[flip_pos & 3u] < 255u) aux->flip_cnt[flip_pos & 3u]++;
    } else {
        if (n->right_period_cnt > 0) n->right_period_cnt--;
        if (n->right_period_cnt < 3) n->period_sync_ok_flag = 0;
    }
    n->old_swap_time = edge;
    return n->period_sync_ok_flag && flip_pos == 2;
}

/* ------------------------------------------------------------------------------------------ slot-phase walk */
/* The reference's bit synchroniser sees a data-bit edge only INSIDE a 4-ms slot and refines only an edge at slot
 * position 2 (nav_data.c:87-138); subframe time stamps - and with them pseudoranges - need that refinement
 * (nav_data.c:356-360).  On the MCU a channel is served 4 of every 17 ms (main.c:134-155), so its slots start one
 * millisecond later, modulo 4, every cycle and every edge alignment passes the window.  The batched paths of this
 * library process EVERY millisecond (index = ms % 4): data bits last 20 ms = 5 slots, so the alignment never moves
 * and three satellites in four would never deliver a time stamp.
 *
 * The walk restores the MCU's behaviour with the MCU's own means - an idle gap between two slots: a channel that
 * tracks but has no refined edge leaves 1..3 milliseconds unprocessed after a complete slot (exactly what the
 * reference does with a channel it does not serve: no call, and gps_rewind_if_phase catches the carrier NCO up,
 * tracking.c:102-113) and starts its next slot behind the gap.  Every slot stays a whole index 0..3 sequence, so the
 * result equals the unmodified reference called on the same (millisecond, index) schedule - which is how the tests
 * check it (oracle/ref_shim.c: ref_track_run_walk).  Policy (this library's, evaluated at the end of a slot, from what the
 * bit synchroniser has seen at the current slot phase):
 *   - at least LC_WALK_CONFIDENCE edges on the 20-ms grid at slot positions 3 / 1, and fewer than a fifth of all edges
 *     at position 2: idle 1 / 3 ms (majority) - the edges then show at 2.  (An edge in the middle of a millisecond
 *     shows at two neighbouring positions in turn: if one of them is 2 that is good enough and nothing happens; a lone
 *     spurious edge at 2 while the loops pull in does not hold the channel back.)
 *   - no edge on the grid seen for walk_period_ms at this slot phase (they fall on the slot boundary, or the loops have
 *     not pulled in yet): idle 2 ms;
 *   - edge refined (accurate_swap_ok): nothing, for ever.
 * A walk decided at the end of a slot takes effect LC_WALK_LEAD_MS later, behind the next slot, so that the thread
 * that plans the carrier NCO of the device-resident loop knows it a whole slot ahead. */
LC_FN int lc_walk_idle(uint32_t skip_ms, uint32_t skip_len, uint32_t ms) { return (uint32_t)(ms - skip_ms) < skip_len; }
/* slot phase after an idle gap whose last millisecond is `ms`: the next millisecond starts a slot */
LC_FN uint8_t lc_walk_phase_after(uint32_t ms) { return (uint8_t)((0u - (ms + 1u)) & (LC_SLOT_LEN - 1u)); }
/* Slot index of millisecond `ms` for this channel, LC_IDLE_INDEX inside an idle gap; the gap's last millisecond
 * moves the slot phase.  The one place the host paths take their index from. */
LC_FN uint8_t lc_walk_index(gpsb_aux* aux, uint32_t ms)
{
    if (lc_walk_idle(aux->skip_ms, aux->skip_len, ms)) {
        if (ms + 1u - aux->skip_ms == aux->skip_len) {
            aux->slot_phase = lc_walk_phase_after(ms);
            aux->phase_since_ms = ms + 1u;
        }
        return LC_IDLE_INDEX;
    }
    return (uint8_t)((ms + aux->slot_phase) & (LC_SLOT_LEN - 1u));
}
/* End of a slot (index 3) of a channel in GPS_TRACKING_RUN, after the nav-bit logic of that millisecond. */
LC_FN_BIG void lc_walk_policy(const gps_ch_t* ch, gpsb_aux* aux, uint32_t ms)
{
    const gps_nav_data_t* n = &ch->nav_data;
    if (aux->skip_len) {
        if ((int32_t)(ms - (aux->skip_ms + aux->skip_len)) < 0) return;     /* a walk is pending */
        aux->skip_len = 0;                                                  /* taken: forget it */
    }
    if (!aux->walk_enable || n->accurate_swap_ok) return;
    const uint32_t patience = aux->walk_period_ms ? aux->walk_period_ms : LC_WALK_PERIOD_MS;
    uint8_t idle = 0;
    const unsigned off_centre = (unsigned)aux->flip_cnt[1] + aux->flip_cnt[3], centre = aux->flip_cnt[2];
    if (!aux->walk_armed) {                       /* first slot end of this channel's tracking: the clock starts here */
        aux->walk_armed = 1;
        aux->phase_since_ms = ms;
        return;
    }
    if (off_centre >= LC_WALK_CONFIDENCE && centre * 4u < off_centre) idle = aux->flip_cnt[3] >= aux->flip_cnt[1] ? 1 : 3;
    else if (off_centre + centre == 0 && ms - aux->phase_since_ms >= patience) idle = 2;
    if (!idle) return;
    aux->skip_ms = ms + LC_WALK_LEAD_MS;
    aux->skip_len = idle;
    aux->last_flip_pos = 0;
    aux->flip_cnt[1] = aux->flip_cnt[2] = aux->flip_cnt[3] = 0;
    aux->walks++;
}

/* ------------------------------------------------------------------------------------------ tail of the step */
/* tracking.c:154-169: SNR estimate from the prompt sums of 200 ms */
LC_FN void lc_snr_update(gps_ch_t* ch, gpsb_aux* aux, int16_t ip, int16_t qp)
{
    gps_tracking_t* t = &ch->tracking_data;
    t->i_part_summ += (uint32_t)lc_iabs(ip);
    t->q_part_summ += (uint32_t)lc_iabs(qp);
    t->snr_summ_cnt++;
    if (t->snr_summ_cnt > LC_SNR_WINDOW) {
        if (t->q_part_summ == 0) {
            t->snr_value = 1.0f;
            aux->snr_pending = 0;
            return;                                   /* sums are left running, like the reference */
        }
#if LC_DEVICE_MATH
        aux->snr_pending = 1;                         /* log10f on the host after the run: lc_resolve_snr */
        aux->snr_i = t->i_part_summ;
        aux->snr_q = t->q_part_summ;
#else
        float ratio = (float)t->i_part_summ / (float)t->q_part_summ;
        t->snr_value = 10.0f * log10f(ratio);
#endif
        t->snr_summ_cnt = 0;
        t->i_part_summ = 0;
        t->q_part_summ = 0;
    }
}

/* Second half of tracking.c:140-169: nav bits and the SNR estimate; nothing here feeds the next correlation
 * (period_sync_ok_flag is written at slot index 3 and read by lc_pll_update at slot index 0). */
LC_FN void lc_finish_tail(gps_ch_t* ch, gpsb_aux* aux, uint8_t index, int16_t ip, int16_t qp, uint32_t now)
{
    if (lc_nav_new_code(ch, aux, index, ip, now)) lc_refine_edge(ch, aux);
    lc_snr_update(ch, aux, ip, qp);
}

#if !defined(__CUDACC__)
/* Host side of the deferred SNR value of a device-resident run. */
static inline void lc_resolve_snr(gps_ch_t* ch, gpsb_aux* aux)
{
    if (!aux->snr_pending) return;
    float ratio = (float)aux->snr_i / (float)aux->snr_q;
    ch->tracking_data.snr_value = 10.0f * log10f(ratio);
    aux->snr_pending = 0;
}
#endif

#endif /* GPSB_LOOP_CORE_H */
