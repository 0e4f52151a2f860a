/*
 * loop_emu.c - TEST INFRASTRUCTURE: a CPU emulation of the device-resident tracking loop k_track_run
 * (stm32f4_sdr_gps_b200/csrc/gpsb_track_loop.cuh).  It compiles the SAME sources the kernel is made of
 * (core/gpsb_loop_core.h with the device math selected, core/gpsb_epl_core.h) with gcc and runs the kernel's
 * control flow thread by thread, so that the tests can check them against the compiled reference without a
 * GPU.  What it cannot cover - CUDA's double atan2 and nvcc's code generation - is covered by the GPU tests.
 * Never linked into the product libraries.
 */
#define LC_EMULATE_DEVICE 1
#include "../../stm32f4_sdr_gps_b200/core/gpsb_loop_core.h"
#include "../../stm32f4_sdr_gps_b200/core/gpsb_epl_core.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

uint32_t emu_sizeof_aux(void) { return (uint32_t)sizeof(gpsb_aux); }
uint32_t emu_sizeof_channel(void) { return (uint32_t)sizeof(gps_ch_t); }

static void expand_code(const uint8_t* chips, uint32_t* E)
{
    for (int w = 0; w < EC_WORDS; w++) {
        uint32_t lo = chips[2 * w] ? 0x0000FFFFu : 0u;
        uint32_t hi = (2 * w + 1 < PRN_LENGTH && chips[2 * w + 1]) ? 0xFFFF0000u : 0u;
        E[w] = lo | hi;
    }
}

static void build_rx(const uint32_t* E, uint32_t bits, uint32_t* RX)
{
    for (int x = 0; x < EC_RX_WORDS; x++) RX[x] = ec_rx_word(E, x, bits);
}

/* the kernel's two worker phases: plain threads over data words 1..510 (nw consecutive words each, the last
 * thread takes what is left) and the twelve edge lanes */
static void epl_two_phase(const uint32_t* S, const uint32_t* E, const gpsb_epl_req* rq, uint32_t nw, int16_t iq[6])
{
    const uint32_t off[3] = {rq->off_e, rq->off_p, rq->off_l};
    uint32_t RX[EC_RX_WORDS];
    uint32_t total[3] = {0, 0, 0};
    build_rx(E, rq->off_bits, RX);
    for (int w0 = 1; w0 < EC_WORDS - 1; w0 += (int)nw) {
        ec_partial part;
        uint32_t acc[3] = {0, 0, 0};
        const int n = w0 + (int)nw <= EC_WORDS - 1 ? (int)nw : EC_WORDS - 1 - w0;
        ec_epl_phase1(S, RX, off, w0, n, &part);                            /* after the code thread's offsets */
        ec_epl_phase2(rq->acc0, rq->step32, w0, n, &part, acc);             /* after the carrier thread's NCO words */
        for (int a = 0; a < 3; a++) total[a] += acc[a];
    }
    for (int e = 0; e < EC_EDGE_LANES; e++) {
        int w, negative;
        const uint32_t c = ec_epl_edge_phase1(S, RX, off, e, &w, &negative);
        total[e % 3] += ec_epl_edge_phase2(rq->acc0, rq->step32, w, negative, c);
    }
    ec_unpack_sums(total, iq);
}

/* One E/P/L cell with the kernel's work split: 512 / nw worker threads own nw consecutive words each. */
void emu_epl_cell(const uint8_t* chips, const uint8_t* frame2046, uint32_t acc0, uint32_t step32, uint32_t off_e,
                  uint32_t off_p, uint32_t off_l, uint32_t bits, uint32_t nw, int16_t iq[6])
{
    uint32_t E[EC_WORDS], S[EC_WORDS];
    expand_code(chips, E);
    for (int i = 0; i < EC_WORDS; i++) S[i] = 0xDEADBEEFu;      /* the pad bytes of a ring frame must not matter */
    memcpy(S, frame2046, 2046);
    gpsb_epl_req rq;
    memset(&rq, 0, sizeof rq);
    rq.acc0 = acc0;
    rq.step32 = step32;
    rq.off_e = (uint16_t)off_e;
    rq.off_p = (uint16_t)off_p;
    rq.off_l = (uint16_t)off_l;
    rq.off_bits = (uint16_t)bits;
    epl_two_phase(S, E, &rq, nw, iq);
}

/* The kernel's loop for one channel, role by role in the order the barriers impose.  Returns the stop reason;
 * *done_ms = milliseconds completed. */
int emu_track_run_phase(gps_ch_t* ch, gpsb_aux* aux, const uint8_t* signal, uint32_t ms0, uint32_t n_ms, uint32_t nw,
                        uint32_t slot_phase, int16_t* iq_log, int8_t* nav_log, uint32_t* done_ms, int16_t stop_iq[6]);

int emu_track_run(gps_ch_t* ch, gpsb_aux* aux, const uint8_t* signal, uint32_t ms0, uint32_t n_ms, uint32_t nw,
                  int16_t* iq_log, int8_t* nav_log, uint32_t* done_ms, int16_t stop_iq[6])
{
    return emu_track_run_phase(ch, aux, signal, ms0, n_ms, nw, 0, iq_log, nav_log, done_ms, stop_iq);
}

/* slot_phase shifts where the 4-ms slots of the bit synchroniser start (index = (ms + slot_phase) % 4).  0 is what the
 * kernel does today; other values are the per-channel slot phase DESIGN.md section 10 proposes for the next round,
 * pinned here against the reference driven with the same index sequence. */
int emu_track_run_phase(gps_ch_t* ch, gpsb_aux* aux, const uint8_t* signal, uint32_t ms0, uint32_t n_ms, uint32_t nw,
                        uint32_t slot_phase, int16_t* iq_log, int8_t* nav_log, uint32_t* done_ms, int16_t stop_iq[6])
{
    uint32_t E[EC_WORDS], S[EC_WORDS];
    expand_code(ch->prn_code, E);
    gpsb_epl_req rq;
    lc_angle_cache cache;
    memset(&cache, 0, sizeof cache);
    int stop = LC_STOP_NONE;
    uint32_t m = 0;
    if (ch->tracking_data.state == GPS_PRE_TRACK_DONE) ch->tracking_data.state = GPS_TRACKING_RUN;
    if (ch->tracking_data.state != GPS_TRACKING_RUN) stop = LC_STOP_STATE;
    else if (n_ms) {
        if (slot_phase) aux->slot_phase = (uint8_t)slot_phase;
        memset(&rq, 0, sizeof rq);
        if (lc_walk_idle(aux->skip_ms, aux->skip_len, ms0)) lc_plan_code(&ch->tracking_data, &rq);   /* kernel prologue */
        else lc_trk_plan_run(ch, ms0, ms0, &rq);
    }
    /* the control threads' private copies of the walk state (kernel: w_phase, w_skip_ms, w_skip_len) */
    uint32_t w_phase = aux->slot_phase, w_skip_ms = aux->skip_ms, w_skip_len = aux->skip_len;
    for (; m < n_ms && stop == LC_STOP_NONE; m++) {
        const uint32_t ms = ms0 + m;
        const int idle = lc_walk_idle(w_skip_ms, w_skip_len, ms);
        const int idle_next = lc_walk_idle(w_skip_ms, w_skip_len, ms + 1u);
        const uint8_t index = idle ? (uint8_t)LC_IDLE_INDEX : (uint8_t)((ms + w_phase) & (LC_SLOT_LEN - 1u));
        memset(S, 0xA5, sizeof S);
        memcpy(S, signal + (size_t)m * 2046, 2046);
        int16_t iq[6];
        epl_two_phase(S, E, &rq, nw, iq);            /* in an idle gap the workers correlate all the same; nobody looks */
        if (idle) memset(iq, 0, sizeof iq);
        if (iq_log) memcpy(iq_log + 6 * (size_t)m, iq, 12);
        if (nav_log) nav_log[m] = -1;
        if (idle) {
            if (m + 1 < n_ms && !idle_next) lc_plan_carrier(&ch->tracking_data, ch->prn, ms + 1, ms + 1, &rq);
            if (!idle_next) {
                aux->slot_phase = lc_walk_phase_after(ms);
                aux->phase_since_ms = ms + 1u;
                w_phase = aux->slot_phase;
            }
            continue;
        }
        if (lc_dll_is_degenerate(iq)) {
            stop = LC_STOP_DLL_NAN;
            if (stop_iq) memcpy(stop_iq, iq, 12);
            break;
        }
        /* The two control threads run side by side between the barriers; they touch disjoint fields except that
         * the carrier thread READS nav_data.period_sync_ok_flag, which the code thread's tail may rewrite in the
         * same millisecond - so the carrier thread samples the flag right after barrier A (kernel: `sync_ok`).
         * Emulated here by running the carrier thread first. */
        /* carrier thread */
        lc_pll_update(&ch->tracking_data, ch->nav_data.period_sync_ok_flag, index, iq[2], iq[3]);
        lc_fll_update(&ch->tracking_data, aux, ch->acq_data.found_freq_offset_hz, index, iq[2], iq[3], &cache);
        if (m + 1 < n_ms && !idle_next) lc_plan_carrier(&ch->tracking_data, ch->prn, ms + 1, ms + 1, &rq);
        /* nav thread, first part (does not need the DLL) */
        const int refine = lc_nav_new_code(ch, aux, index, iq[2], ms);
        if (nav_log) nav_log[m] = aux->last_nav_bit;
        /* code thread: DLL, next offsets */
        lc_dll_update(&ch->tracking_data, iq[0], iq[1], iq[4], iq[5]);
        if (m + 1 < n_ms) lc_plan_code(&ch->tracking_data, &rq);
        /* nav thread, after the DLL's mbarrier */
        if (refine) lc_refine_edge(ch, aux);
        lc_snr_update(ch, aux, iq[2], iq[3]);
        if (index == LC_SLOT_LEN - 1) lc_walk_policy(ch, aux, ms);
        if (index == LC_SLOT_LEN - 2) {               /* the other control threads re-read the decision here */
            w_skip_ms = aux->skip_ms;
            w_skip_len = aux->skip_len;
        }
    }
    *done_ms = m;
    return stop;
}

void emu_resolve_snr(gps_ch_t* ch, gpsb_aux* aux) { lc_resolve_snr(ch, aux); }

/* --------------------------------------------------------------- device math against the host libm */
static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* Compare the float-path functions the device uses with the host libm over ip in [ip_lo, ip_hi), every
 * qp in [-8184, 8184].  kind 0: atan2f branch of the Costas discriminator (ip > 0); kind 1: FLL angle.
 * Returns the number of mismatching bit patterns; first_bad[2] receives the first offending (ip, qp). */
uint64_t emu_compare_float_math(int kind, int ip_lo, int ip_hi, int32_t first_bad[2])
{
    uint64_t bad = 0;
    for (int ip = ip_lo; ip < ip_hi; ip++) {
        for (int qp = -8184; qp <= 8184; qp++) {
            float dev, host;
            if (kind == 0) {
                if (ip <= 0) continue;
                dev = (float)(lc_atan2f((float)qp, (float)ip) / LC_PI);
                host = (float)(atan2f((float)qp, (float)ip) / LC_PI);
            } else {
                if (ip == 0) continue;
                dev = lc_atanf((float)qp / (float)ip);
                host = atanf((float)qp / (float)ip);
            }
            if (fbits(dev) != fbits(host)) {
                if (!bad && first_bad) { first_bad[0] = ip; first_bad[1] = qp; }
                bad++;
            }
        }
    }
    return bad;
}

/* Host libm values of both discriminators for a block of the domain: what the GPU self-test is compared
 * with.  out[(ip - ip_lo) * 16369 + (qp + 8184)] */
void emu_host_costas(int ip_lo, int ip_hi, float* out)
{
    for (int ip = ip_lo; ip < ip_hi; ip++)
        for (int qp = -8184; qp <= 8184; qp++) {
            float v;
            if (ip > 0) v = (float)(atan2f((float)qp, (float)ip) / LC_PI);
            else v = (float)(atan2((float)-qp, (float)-ip) / LC_PI);
            out[(size_t)(ip - ip_lo) * 16369 + (size_t)(qp + 8184)] = v;
        }
}
void emu_host_fll_angle(int ip_lo, int ip_hi, float* out)
{
    for (int ip = ip_lo; ip < ip_hi; ip++)
        for (int qp = -8184; qp <= 8184; qp++)
            out[(size_t)(ip - ip_lo) * 16369 + (size_t)(qp + 8184)] =
                (ip == 0) ? (float)(LC_PI / 2) : atanf((float)qp / (float)ip);
}

/* the private generator against libc's: n draws after the default seed */
uint32_t emu_rand31_mismatches(uint32_t n)
{
    gpsb_rand31 g;
    memset(&g, 0, sizeof g);
    struct random_data rd;
    char state[128];
    memset(&rd, 0, sizeof rd);
    initstate_r(1u, state, sizeof state, &rd);
    uint32_t bad = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t want = 0;
        random_r(&rd, &want);
        if (lc_rand31_next(&g) != want) bad++;
    }
    return bad;
}

/* lc_fold_half_pi against the reference's literal form (double comparisons) for every float in [lo_bits, hi_bits] */
uint64_t emu_fold_mismatches(uint32_t lo_bits, uint32_t hi_bits)
{
    uint64_t bad = 0;
    for (uint64_t u = lo_bits; u <= hi_bits; u++) {
        for (int neg = 0; neg < 2; neg++) {
            float x;
            uint32_t b = (uint32_t)u | (neg ? 0x80000000u : 0u);
            memcpy(&x, &b, 4);
            float want = x;
            if (want > LC_PI / 2) want = (float)(LC_PI - want);
            if (want < -LC_PI / 2) want = (float)(-LC_PI - want);
            float got = lc_fold_half_pi(x);
            if (fbits(got) != fbits(want)) bad++;
        }
    }
    return bad;
}
