/*
 * rtcm.c - row N4 of SURVEY.md section 8(f), second half: the receiver's observations and ephemerides as RTCM 3
 * frames - message 1019 (GPS ephemeris) and 1075 (GPS MSM5: pseudorange, phase range, rate, C/N0 per satellite and
 * signal) - for an RTK engine downstream.
 *
 * Behaviour follows Firmware/project_main/GPS/obs_publish.c (framing, CRC-24Q, sendrtcmobs / sendrtcmnav) and
 * GPS/RTK/rtcm3e.c (the two message bodies), cited per function; the frames equal the compiled reference's byte for
 * byte (tests/test_rtcm.py).  The reference ships with this output compiled out (config.h:30, ENABLE_RTCM_SEND 0);
 * here it is a run-time switch, off by default (gpsb_host_enable_rtcm).
 *
 * Reproduced on purpose:
 *   - the MSM5 "extended satellite info" nibbles are skipped, not written (rtcm3e.c:625-626), so they carry whatever
 *     the previous frame left at those bit positions of the one static frame buffer; an ephemeris frame starts from a
 *     cleared buffer (obs_publish.c:74), an observation frame does not (obs_publish.c:59);
 *   - every observation frame starts from a re-initialised message state (obs_publish.c:59 -> init_rtcm :130-163), so
 *     the session time is always zero - and so are the lock times of PRN 1..4, because that initialisation clears the
 *     first MAXSAT = 4 of the 32 lock-time stamps; PRN 5..32 keep theirs from frame to frame and report real lock
 *     times, until the next ephemeris frame clears the whole state (obs_publish.c:73).
 * Not reproduced: the reference sizes its per-satellite carrier-phase offsets for MAXSAT = 4 but indexes them by
 * satellite number (rtcm3e.c:351-355), reading past the array for PRN > 4.  This receiver produces no carrier phase
 * (L = 0), the offsets are never written, and the in-bounds reading is "no phase range" - which is what is encoded
 * here for every PRN.
 */
#include <math.h>

#include "host_internal.h"

#define RT_FRAME_MAX   300                         /* rtk_common.h:92 */
#define RT_CLIGHT      299792458.0
#define RT_FREQ1       1.57542E9
#define RT_RANGE_MS    (RT_CLIGHT * 0.001)         /* rtcm3e.c:38 */
#define RT_SC2RAD      3.1415926535898
#define RT_ROUND(x)    ((int)floor((x) + 0.5))
#define RT_ROUND_U(x)  ((unsigned int)floor((x) + 0.5))

/* scale factors spelled as in rtk_common.h:9-32 / rtcm3e.c:39 (three of them are one ulp off the power of two) */
#define RT_2P_5   0.03125
#define RT_2P_10  0.0009765625
#define RT_2P_19  1.907348632812500E-06
#define RT_2P_24  5.960464477539063E-08
#define RT_2P_29  1.862645149230957E-09
#define RT_2P_31  4.656612873077393E-10
#define RT_2P_33  1.164153218269348E-10
#define RT_2P_43  1.136868377216160E-13
#define RT_2P_55  2.775557561562891E-17

static uint8_t g_frame[RT_FRAME_MAX];              /* the one frame buffer (rtcm_data.buff, obs_publish.c:20) */
static int g_enabled = 0;
static void (*g_sink)(const uint8_t*, uint32_t) = NULL;
static int (*g_sink_busy)(void) = NULL;
static uint32_t g_last_obs_ms = 0;
static gtime_t g_locked_since[32];                 /* rtcm_data.lltime, see the file header */

/* ---------------------------------------------------------------------------------------------- bit fields */

/* rtcm3e.c:73-81: len bits of value, MSB first, at bit position pos */
void setbitu(unsigned char* buff, int pos, int len, unsigned int data)
{
    if (len <= 0 || 32 < len) return;
    for (int k = 0; k < len; k++) {
        const int at = pos + k;
        const unsigned char bit = (unsigned char)(0x80u >> (at & 7));
        if ((data >> (len - 1 - k)) & 1u) buff[at >> 3] |= bit; else buff[at >> 3] &= (unsigned char)~bit;
    }
}

/* rtcm3e.c:83-87: sign-magnitude of the top bit only - the low bits stay two's complement */
static void put_signed(unsigned char* buff, int pos, int len, int value)
{
    unsigned int u = (unsigned int)value;
    const unsigned int top = 1u << (len - 1);
    if (value < 0) u |= top; else u &= ~top;
    setbitu(buff, pos, len, u);
}

/* obs_publish.c:82-90, CRC-24Q (generator 0x1864CFB) without the table */
static unsigned int crc24q(const unsigned char* p, int n)
{
    unsigned int crc = 0;
    for (int i = 0; i < n; i++) {
        crc ^= (unsigned int)p[i] << 16;
        for (int b = 0; b < 8; b++) {
            crc <<= 1;
            if (crc & 0x1000000u) crc ^= 0x1864CFBu;
        }
    }
    return crc & 0xFFFFFFu;
}

/* ---------------------------------------------------------------------------------------------- 1019 */

/* rtcm3e.c:156-221: returns the bit position after the body, 0 when the satellite is not a GPS PRN */
static int body_1019(unsigned char* buff, const eph_t* eph, int sat)
{
    int i = 24;
    if (sat <= 0 || sat > 32) return 0;                       /* satsys, rtcm3e.c:90-104 */
    const int week = eph->week % 1024;
    const int toe = RT_ROUND(eph->toes / 16.0);
    const int toc = RT_ROUND(time2gpst(eph->toc, NULL) / 16.0);
    const unsigned int sqrtA = RT_ROUND_U(sqrt(eph->A) / RT_2P_19);
    const unsigned int e = RT_ROUND_U(eph->e / RT_2P_33);
    const int i0 = RT_ROUND(eph->i0 / RT_2P_31 / RT_SC2RAD);
    const int OMG0 = RT_ROUND(eph->OMG0 / RT_2P_31 / RT_SC2RAD);
    const int omg = RT_ROUND(eph->omg / RT_2P_31 / RT_SC2RAD);
    const int M0 = RT_ROUND(eph->M0 / RT_2P_31 / RT_SC2RAD);
    const int deln = RT_ROUND(eph->deln / RT_2P_43 / RT_SC2RAD);
    const int idot = RT_ROUND(eph->idot / RT_2P_43 / RT_SC2RAD);
    const int OMGd = RT_ROUND(eph->OMGd / RT_2P_43 / RT_SC2RAD);
    const int crs = RT_ROUND(eph->crs / RT_2P_5);
    const int crc = RT_ROUND(eph->crc / RT_2P_5);
    const int cus = RT_ROUND(eph->cus / RT_2P_29);
    const int cuc = RT_ROUND(eph->cuc / RT_2P_29);
    const int cis = RT_ROUND(eph->cis / RT_2P_29);
    const int cic = RT_ROUND(eph->cic / RT_2P_29);
    const int af0 = RT_ROUND(eph->f0 / RT_2P_31);
    const int af1 = RT_ROUND(eph->f1 / RT_2P_43);
    const int af2 = RT_ROUND(eph->f2 / RT_2P_55);
    const int tgd = RT_ROUND(eph->tgd[0] / RT_2P_31);

    /* DF002 DF009 DF076 DF077 DF078 DF079 DF071 DF081 DF082 DF083 DF084 DF085 DF086 DF087 DF088 DF089 DF090 DF091
     * DF092 DF093 DF094 DF095 DF096 DF097 DF098 DF099 DF100 DF101 DF102 DF103 DF137 */
    const struct { int bits; int is_signed; long long v; } field[] = {
        { 12, 0, 1019 }, { 6, 0, sat }, { 10, 0, week }, { 4, 0, eph->sva }, { 2, 0, eph->code }, { 14, 1, idot },
        { 8, 0, eph->iode }, { 16, 0, toc }, { 8, 1, af2 }, { 16, 1, af1 }, { 22, 1, af0 }, { 10, 0, eph->iodc },
        { 16, 1, crs }, { 16, 1, deln }, { 32, 1, M0 }, { 16, 1, cuc }, { 32, 0, e }, { 16, 1, cus }, { 32, 0, sqrtA },
        { 16, 0, toe }, { 16, 1, cic }, { 32, 1, OMG0 }, { 16, 1, cis }, { 32, 1, i0 }, { 16, 1, crc }, { 32, 1, omg },
        { 24, 1, OMGd }, { 8, 1, tgd }, { 6, 0, eph->svh }, { 1, 0, eph->flag }, { 1, 0, eph->fit > 0.0 ? 0 : 1 },
    };
    for (unsigned k = 0; k < sizeof field / sizeof field[0]; k++) {
        if (field[k].is_signed) put_signed(buff, i, field[k].bits, (int)field[k].v);
        else setbitu(buff, i, field[k].bits, (unsigned int)field[k].v);
        i += field[k].bits;
    }
    return i;
}

/* ---------------------------------------------------------------------------------------------- 1075 (MSM5) */

/* rtklib_common.c:7-25 with rtcm3e.c:63-68,231-243: observation code -> GPS MSM signal number (0 = none) */
static int msm_signal_gps(unsigned char code)
{
    static const unsigned char id[27] = { 0, 2, 3, 4, 5, 6, 0, 30, 31, 0, 0, 0, 32, 0, 8, 0, 15, 16, 17, 9, 10, 11, 12, 0,
                                          22, 23, 24 };
    return code < 27 ? id[code] : 0;
}

static int msm_satellite_gps(int sat) { return sat >= 1 && sat <= 32 ? sat : 0; }       /* rtcm3e.c:224-229 */

/* rtcm3e.c:121-131 */
static int session_indicator(int lock)
{
    if (lock < 0) return 0;
    if (lock < 24) return lock;
    if (lock < 72) return (lock + 24) / 2;
    if (lock < 168) return (lock + 120) / 4;
    if (lock < 360) return (lock + 408) / 8;
    if (lock < 744) return (lock + 1176) / 16;
    if (lock < 937) return (lock + 3096) / 32;
    return 127;
}

/* rtcm3e.c:133-151: 0 below 32 ms, then one step per octave up to 15 */
static int msm_lock_indicator(int lock)
{
    int ind = 0;
    for (int edge = 32; ind < 15 && lock >= edge; edge <<= 1) ind++;
    return ind;
}

/* rtcm3e.c:611-640 with the header :373-427 and the index / field generators :246-369.  Returns the bit position
 * after the body. */
static int body_1075(unsigned char* buff, const obsd_t* obs, int n, gtime_t frame_time)
{
    unsigned char sat_slot[64] = { 0 }, sig_slot[32] = { 0 };
    static unsigned char cell_slot[32 * 64];
    double rough_range[64], rough_rate[64], fine_range[64], fine_phase[64], fine_rate[64];
    float cnr[64];
    unsigned char half[64];
    int lock[64];
    int nsat = 0, nsig = 0, ncell = 0, i = 24;
    const double lambda = RT_CLIGHT / RT_FREQ1;
    memset(cell_slot, 0, sizeof cell_slot);
    memset(g_locked_since, 0, 4 * sizeof g_locked_since[0]);

    /* which satellites, signals and (satellite, signal) cells are present */
    for (int k = 0; k < n; k++) {
        const int sat = msm_satellite_gps(obs[k].sat), sig = msm_signal_gps(obs[k].code[0]);
        if (sat && sig) sat_slot[sat - 1] = sig_slot[sig - 1] = 1;
    }
    for (int k = 0; k < 64; k++) if (sat_slot[k]) sat_slot[k] = (unsigned char)++nsat;
    for (int k = 0; k < 32; k++) if (sig_slot[k]) sig_slot[k] = (unsigned char)++nsig;
    for (int k = 0; k < n; k++) {
        const int sat = msm_satellite_gps(obs[k].sat), sig = msm_signal_gps(obs[k].code[0]);
        if (sat && sig) cell_slot[sig_slot[sig - 1] - 1 + (sat_slot[sat - 1] - 1) * nsig] = 1;
    }
    for (int k = 0; k < nsat * nsig; k++)
        if (cell_slot[k] && ncell < 64) cell_slot[k] = (unsigned char)++ncell;

    /* header (RTCM 10403.2 table 3.5-78) */
    const unsigned int epoch = RT_ROUND_U(time2gpst(frame_time, NULL) * 1E3);
    const int session_s = 0;                                  /* re-initialised per frame, see the file header */
    setbitu(buff, i, 12, 1075); i += 12;
    setbitu(buff, i, 12, 0); i += 12;                         /* station id */
    setbitu(buff, i, 30, epoch); i += 30;
    setbitu(buff, i, 1, 0); i += 1;                           /* more frames for this epoch: no */
    setbitu(buff, i, 3, 0); i += 3;                           /* issue of data station */
    setbitu(buff, i, 7, (unsigned int)session_indicator(session_s)); i += 7;
    setbitu(buff, i, 2, 0); i += 2;                           /* clock steering */
    setbitu(buff, i, 2, 0); i += 2;                           /* external clock */
    setbitu(buff, i, 1, 0); i += 1;                           /* smoothing */
    setbitu(buff, i, 3, 0); i += 3;
    for (int k = 0; k < 64; k++) { setbitu(buff, i, 1, sat_slot[k] ? 1 : 0); i += 1; }
    for (int k = 0; k < 32; k++) { setbitu(buff, i, 1, sig_slot[k] ? 1 : 0); i += 1; }
    for (int k = 0; k < nsat * nsig && k < 64; k++) { setbitu(buff, i, 1, cell_slot[k] ? 1 : 0); i += 1; }

    /* per satellite: range rounded to 2^-10 ms, range rate rounded to 1 m/s - the first non-zero observation wins */
    for (int k = 0; k < 64; k++) rough_range[k] = rough_rate[k] = 0.0;
    for (int k = 0; k < n; k++) {
        const int sat = msm_satellite_gps(obs[k].sat);
        if (!sat || !msm_signal_gps(obs[k].code[0])) continue;
        const int s = sat_slot[sat - 1] - 1;
        const double range = RT_ROUND(obs[k].P[0] / RT_RANGE_MS / RT_2P_10) * RT_RANGE_MS * RT_2P_10;
        const double rate = RT_ROUND(-obs[k].D[0] * lambda) * 1.0;
        if (rough_range[s] == 0.0 && obs[k].P[0] != 0.0) rough_range[s] = range;
        if (rough_rate[s] == 0.0 && obs[k].D[0] != 0.0) rough_rate[s] = rate;
    }
    /* per cell: what is left after the rough parts */
    for (int k = 0; k < ncell; k++) fine_range[k] = fine_phase[k] = fine_rate[k] = 0.0;
    for (int k = 0; k < n; k++) {
        const int sat = msm_satellite_gps(obs[k].sat), sig = msm_signal_gps(obs[k].code[0]);
        if (!sat || !sig) continue;
        const int s = sat_slot[sat - 1] - 1;
        const int cell = cell_slot[sig_slot[sig - 1] - 1 + s * nsig];
        if (cell >= 64) continue;
        const double dr = obs[k].P[0] == 0.0 ? 0.0 : obs[k].P[0] - rough_range[s];
        const double dp = obs[k].L[0] == 0.0 || lambda <= 0.0 ? 0.0 : obs[k].L[0] * lambda - rough_range[s];
        const double dv = obs[k].D[0] == 0.0 || lambda <= 0.0 ? 0.0 : -obs[k].D[0] * lambda - rough_rate[s];
        /* phase / pseudorange whole-cycle offset: a fresh message state every frame makes it round(dp / lambda) * lambda
         * whenever |dp| > 1171 m or the loss-of-lock bit is set, else 0 (rtcm3e.c:351-356) */
        int lli = obs[k].LLI[0];
        double offset = 0.0;
        if ((lli & 1) || fabs(dp - offset) > 1171.0) { offset = RT_ROUND(dp / lambda) * lambda; lli |= 1; }
        const double phase = dp - offset;
        gtime_t* since = &g_locked_since[obs[k].sat - 1];                   /* rtcm3e.c:112-118 */
        if (!since->time || (lli & 1)) *since = obs[k].time;
        const int locked_s = (int)timediff(obs[k].time, *since);

        if (dr != 0.0) fine_range[cell - 1] = dr;
        if (phase != 0.0) fine_phase[cell - 1] = phase;
        if (dv != 0.0) fine_rate[cell - 1] = dv;
        lock[cell - 1] = locked_s;
        half[cell - 1] = (obs[k].LLI[0] & 2) ? 1 : 0;
        cnr[cell - 1] = (float)(obs[k].SNR[0] * 0.25);
    }

    /* satellite data: whole ms, [extended info: skipped], 2^-10 ms, rate */
    for (int k = 0; k < nsat; k++) {
        unsigned int whole = 255;
        if (rough_range[k] != 0.0 && !(rough_range[k] < 0.0 || rough_range[k] > RT_RANGE_MS * 255.0))
            whole = RT_ROUND_U(rough_range[k] / RT_RANGE_MS / RT_2P_10) >> 10;
        setbitu(buff, i, 8, whole); i += 8;
    }
    i += nsat * 4;
    for (int k = 0; k < nsat; k++) {
        unsigned int part = 0;
        if (!(rough_range[k] <= 0.0 || rough_range[k] > RT_RANGE_MS * 255.0))
            part = RT_ROUND_U(rough_range[k] / RT_RANGE_MS / RT_2P_10) & 0x3FFu;
        setbitu(buff, i, 10, part); i += 10;
    }
    for (int k = 0; k < nsat; k++) {
        const int v = fabs(rough_rate[k]) > 8191.0 ? -8192 : RT_ROUND(rough_rate[k] / 1.0);
        put_signed(buff, i, 14, v); i += 14;
    }
    /* signal data */
    for (int k = 0; k < ncell; k++) {
        const int v = fine_range[k] == 0.0 || fabs(fine_range[k]) > 292.7 ? -16384
                                                                           : RT_ROUND(fine_range[k] / RT_RANGE_MS / RT_2P_24);
        put_signed(buff, i, 15, v); i += 15;
    }
    for (int k = 0; k < ncell; k++) {
        const int v = fine_phase[k] == 0.0 || fabs(fine_phase[k]) > 1171.0 ? -2097152
                                                                            : RT_ROUND(fine_phase[k] / RT_RANGE_MS / RT_2P_29);
        put_signed(buff, i, 22, v); i += 22;
    }
    for (int k = 0; k < ncell; k++) { setbitu(buff, i, 4, (unsigned int)msm_lock_indicator(lock[k])); i += 4; }
    for (int k = 0; k < ncell; k++) { setbitu(buff, i, 1, half[k]); i += 1; }
    for (int k = 0; k < ncell; k++) { setbitu(buff, i, 6, (unsigned int)RT_ROUND(cnr[k] / 1.0)); i += 6; }
    for (int k = 0; k < ncell; k++) {
        const int v = fine_rate[k] == 0.0 || fabs(fine_rate[k]) > 1.6384 ? -16384 : RT_ROUND(fine_rate[k] / 0.0001);
        setbitu(buff, i, 15, (unsigned int)v); i += 15;        /* unsigned writer on purpose, rtcm3e.c:577 */
    }
    return i;
}

/* ---------------------------------------------------------------------------------------------- framing */

/* obs_publish.c:92-128: preamble, 6 reserved bits, 10-bit length, body, zero padding to a byte, CRC-24Q.
 * Returns the frame length in bytes, 0 when there is nothing to send. */
static int finish_frame(unsigned char* buff, int body_end_bit)
{
    if (!body_end_bit) return 0;
    int i = body_end_bit;
    for (; i % 8; i++) setbitu(buff, i, 1, 0);
    const int len = i / 8;
    if (len >= 3 + 1024) return 0;
    setbitu(buff, 14, 10, (unsigned int)(len - 3));
    setbitu(buff, i, 24, crc24q(buff, len));
    return len + 3;
}

static void start_frame(unsigned char* buff)
{
    setbitu(buff, 0, 8, 0xD3);
    setbitu(buff, 8, 6, 0);
    setbitu(buff, 14, 10, 0);
}

/* Ephemeris frame (1019) of one satellite into out[cap]; returns its length, 0 = none / does not fit. */
int gpsb_rtcm_encode_eph(const eph_t* eph, int sat, uint8_t* out, uint32_t cap)
{
    if (!eph) return 0;
    memset(g_frame, 0, sizeof g_frame);                       /* obs_publish.c:73: the whole message state */
    memset(g_locked_since, 0, sizeof g_locked_since);
    start_frame(g_frame);
    const int n = finish_frame(g_frame, body_1019(g_frame, eph, sat));
    if (n <= 0 || (out && (uint32_t)n > cap)) return 0;
    if (out) memcpy(out, g_frame, (size_t)n);
    return n;
}

/* Observation frame (1075) of n observation records; returns its length, 0 = none / does not fit. */
int gpsb_rtcm_encode_obs(const obsd_t* obs, int n_obs, uint8_t* out, uint32_t cap)
{
    if (!obs || n_obs <= 0 || n_obs > GPSB_FIX_MAX_SATS) return 0;
    start_frame(g_frame);
    const int n = finish_frame(g_frame, body_1075(g_frame, obs, n_obs, obs[0].time));
    if (n <= 0 || (out && (uint32_t)n > cap)) return 0;
    if (out) memcpy(out, g_frame, (size_t)n);
    return n;
}

/* ---------------------------------------------------------------------------------------------- reference names */

void gpsb_host_set_rtcm_sink(void (*send)(const uint8_t* frame, uint32_t bytes), int (*busy)(void))
{
    g_sink = send;
    g_sink_busy = busy;
}
void gpsb_host_enable_rtcm(int on) { g_enabled = on != 0; g_last_obs_ms = 0; }
int gpsb_host_rtcm_enabled(void) { return g_enabled; }

void sendrtcmobs(obsd_t* obsd, int nsat)                      /* obs_publish.c:57-69 */
{
    const int n = gpsb_rtcm_encode_obs(obsd, nsat, NULL, 0);
    if (g_sink) g_sink(g_frame, (uint32_t)n);                 /* the reference hands over even an empty frame */
}

void sendrtcmnav(gps_ch_t* channel)                           /* obs_publish.c:71-80 */
{
    if (!channel) return;
    const int n = gpsb_rtcm_encode_eph(&channel->eph_data.eph, channel->prn, NULL, 0);
    if (g_sink) g_sink(g_frame, (uint32_t)n);
}

/* gps_master.c:431-455: one frame per call at most - an ephemeris whenever a channel has a new one, else the
 * observations five times a second.  `received_mask & 0x7 == 0x7` in the reference parses as `received_mask & 1`:
 * the ephemeris goes out as soon as subframe 1 is in, and takes the flags of 2 and 3 with it - reproduced. */
void gps_master_transmit_obs(gps_ch_t* channels)
{
    obsd_t* const obsd = hx_obsd;                  /* the solver's array (gps_master.c:41), see host/fix.c */
    if (!channels || (g_sink_busy && g_sink_busy())) return;
    uint32_t n = gpsb_host_sat_cnt();
    if (n > GPSB_FIX_MAX_SATS) n = GPSB_FIX_MAX_SATS;
    sdrobs2obsd(channels, (int)n, obsd);
    for (uint32_t i = 0; i < n; i++) {
        if (channels[i].eph_data.received_mask & 1) {
            channels[i].eph_data.received_mask &= (uint8_t)~0x7;
            sendrtcmnav(&channels[i]);
            return;
        }
    }
    const uint32_t now = signal_capture_get_packet_cnt();
    if ((now - g_last_obs_ms) > 200) {
        g_last_obs_ms = now;
        sendrtcmobs(obsd, (int)n);
    }
}
