/*
 * fix.c - row N4 of SURVEY.md section 8(f): receiver position from the channels' observations (time of week,
 * pseudorange - master.c) and their decoded broadcast ephemerides (core/gpsb_loop_core.h, lc_decode_subframe).
 *
 * Behaviour follows Firmware/project_main/GPS/RTK/solving.c and rtklib_common.c (cited per function): satellite
 * clock and orbit at the transmission time, pseudorange residuals with broadcast-model ionosphere and standard-
 * atmosphere troposphere corrections, a weighted seven-parameter normal-equation solve (x y z, receiver clock, three
 * system offsets pinned by constraint rows) iterated to convergence, ECEF -> geodetic.  Every floating-point
 * expression keeps the reference's operand order: the solution equals the compiled reference's bit for bit
 * (tests/test_fix.py), which is the only reason some expressions below look the way they do.
 *
 * No GPU work: one fix is a few thousand double operations twice a second.  What changes against the reference is the
 * bookkeeping around the arithmetic:
 *   - the reference cuts a solve into sub-millisecond slices because its caller has 1 ms between DMA interrupts, and
 *     keeps the cut points in function statics; here the slices are steps of one explicit state record (fx_state), so
 *     the sliced driver (gps_pos_solve) and the one-shot driver (gpsb_host_fix_once) run the same arithmetic;
 *   - work arrays are members of that record (the reference mallocs them per solve), sized for 32 satellites instead
 *     of GPS_SAT_CNT = 4; with four channels the results are the reference's.
 *
 * Provenance: two routines below are RTKLIB-derived in the reference itself and follow it operation for operation,
 * because bit-exact doubles leave no freedom in the order of operations: fx_orbit (broadcast-ephemeris Kepler
 * propagation, the reference's eph2pos, RTK/solving.c:1165-1215, RTKLIB ephemeris.c) and fx_lu / fx_lu_solve (Crout LU
 * with partial pivoting, the reference's ludcmp / lubksb, RTK/solving.c:1348-1405, RTKLIB rtkcmn.c).  RTKLIB is
 * BSD-2-Clause (T. Takasu); the rest of this file - the state record, the slicing, the drivers - is this project's.
 */
#include <math.h>
#include <time.h>

#include "../../include/gpsb_flat_state.h"
#include "host_internal.h"

#define FX_MAX_SATS   GPSB_FIX_MAX_SATS
#define FX_NX         7                        /* solving.c:36 */
#define FX_ROWS       (FX_MAX_SATS + 4)
#define FX_MAX_PASSES 10                       /* solving.c:35 */

#define FX_PI         3.1415926535897932      /* solving.h:14 */
#define FX_R2D        (180.0 / FX_PI)
#define FX_CLIGHT     299792458.0             /* rtk_common.h:43 */
#define FX_MU         3.9860050E14            /* solving.c:26 */
#define FX_OMGE       7.2921151467E-5         /* solving.c:27 */
#define FX_RE         6378137.0               /* solving.c:38 */
#define FX_FE         (1.0 / 298.257223563)   /* solving.c:39 */
#define FX_SQ(x)      ((x) * (x))

enum { FX_IDLE = 0, FX_SATELLITES, FX_ESTIMATE };

/* broadcast data: the channels' own ephemeris records (gps_pos_solve_init) and the ionosphere coefficients */
typedef struct {
    const eph_t* eph[FX_MAX_SATS];
    int n_eph;
    double ion[8];
} fx_nav;

typedef struct {
    /* satellite states at the transmission times */
    double pv[6 * FX_MAX_SATS], clk[2 * FX_MAX_SATS], pv_var[FX_MAX_SATS], resid[FX_MAX_SATS];
    int health[FX_MAX_SATS], used[FX_MAX_SATS];
    /* estimator */
    double x[FX_NX], dx[FX_NX], Q[FX_NX * FX_NX];
    double v[FX_ROWS], H[FX_NX * FX_ROWS], w[FX_ROWS];
    double site[3], geo[3];
    int rows;
    /* where the sliced driver stands */
    int stage, sat_slice, sat_ok, pass, op;
    int accepted;                                /* see fx_measure */
    uint8_t solving, converting;
    uint32_t last_request_ms;
} fx_state;

static fx_nav g_nav;
static fx_state g_fx;                            /* the sliced driver's work area (gps_pos_solve) */
static fx_state g_once;                          /* the one-shot driver's: a fix may be taken while a sliced one is in flight,
                                                    as with the reference's pntpos / pntpos_iterative */

/* solving.c:48-52: results, under the names the reference's display code reads them by */
sol_t gps_sol;
double final_pos[3];
double azel[2 * FX_MAX_SATS];

/* ------------------------------------------------------------------------------------------ time (rtklib_common.c) */

double timediff(gtime_t a, gtime_t b) { return difftime(a.time, b.time) + a.sec - b.sec; }      /* :27-30 */

gtime_t gpst2time(int week, double sec) { return lc_gpst2time(week, sec); }                     /* :33-43 */

/* :45-52.  The whole seconds are added to the FRACTION and taken off it again: .time never moves and .sec may leave
 * [0,1).  Every consumer only ever forms differences (timediff), so the receiver works; parity needs it as is. */
gtime_t timeadd(gtime_t t, double sec)
{
    t.sec += sec;
    const double whole = floor(t.sec);
    t.sec += whole;
    t.sec -= whole;
    return t;
}

double time2gpst(gtime_t t, int* week)                                                          /* :62-73 */
{
    const time_t since = t.time - (time_t)315964800;
    const int w = (int)(since / (86400 * 7));
    if (week) *week = w;
    return (double)(since - w * 86400 * 7) + t.sec;
}

/* :75-91: the channels' observations in the solver's record */
void sdrobs2obsd(gps_ch_t* ch, int ns, obsd_t* out)
{
    for (int i = 0; i < ns; i++) {
        out[i].time = lc_gpst2time(ch[i].eph_data.week_gpst, ch[i].obs_data.tow_s);
        out[i].rcv = 1;
        out[i].sat = ch[i].prn;
        out[i].P[0] = ch[i].obs_data.pseudorange_m;
        out[i].L[0] = 0;
        out[i].D[0] = (float)ch[i].tracking_data.if_freq_offset_hz;
        out[i].SNR[0] = (unsigned char)(ch[i].tracking_data.snr_value + 20.0f) * 4;
        out[i].LLI[0] = 0;
        out[i].code[0] = 1;                                   /* CODE_L1C */
    }
}

/* ------------------------------------------------------------------------------------------ small linear algebra */

/* solving.c:332-337: the sum runs from the LAST element down */
static double fx_dot(const double* a, const double* b, int n)
{
    double acc = 0.0;
    while (--n >= 0) acc += a[n] * b[n];
    return acc;
}
static double fx_norm(const double* a, int n) { return sqrt(fx_dot(a, a, n)); }

/* solving.c:1348-1390: LU factors of the n x n column-major A in place, implicit row scaling, partial pivoting */
static int fx_lu(double* A, int n, int* piv)
{
    double scale[FX_NX], big, s, t;
    int top = 0;
    for (int i = 0; i < n; i++) {
        big = 0.0;
        for (int j = 0; j < n; j++)
            if ((t = fabs(A[i + j * n])) > big) big = t;
        if (big > 0.0) scale[i] = 1.0 / big; else return 1;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            s = A[i + j * n];
            for (int k = 0; k < i; k++) s -= A[i + k * n] * A[k + j * n];
            A[i + j * n] = s;
        }
        big = 0.0;
        for (int i = j; i < n; i++) {
            s = A[i + j * n];
            for (int k = 0; k < j; k++) s -= A[i + k * n] * A[k + j * n];
            A[i + j * n] = s;
            if ((t = scale[i] * fabs(s)) >= big) { big = t; top = i; }
        }
        if (j != top) {
            for (int k = 0; k < n; k++) { t = A[top + k * n]; A[top + k * n] = A[j + k * n]; A[j + k * n] = t; }
            scale[top] = scale[j];
        }
        piv[j] = top;
        if (A[j + j * n] == 0.0) return 1;
        if (j != n - 1) {
            t = 1.0 / A[j + j * n];
            for (int i = j + 1; i < n; i++) A[i + j * n] *= t;
        }
    }
    return 0;
}

/* solving.c:1392-1405: forward and back substitution of one right-hand side */
static void fx_lu_solve(const double* A, int n, const int* piv, double* b)
{
    double s;
    int first = -1;
    for (int i = 0; i < n; i++) {
        const int p = piv[i];
        s = b[p]; b[p] = b[i];
        if (first >= 0) { for (int j = first; j < i; j++) s -= A[i + j * n] * b[j]; }
        else if (s) first = i;
        b[i] = s;
    }
    for (int i = n - 1; i >= 0; i--) {
        s = b[i];
        for (int j = i + 1; j < n; j++) s -= A[i + j * n] * b[j];
        b[i] = s / A[i + i * n];
    }
}

/* solving.c:1452-1469 (lsq) with matmul :1308-1331 and matinv :1413-1436 for NX parameters and m rows:
 * x = (A A')^-1 A y, Q = (A A')^-1; A is NX x m column-major (one column per measurement). */
static int fx_normal_solve(const double* A, const double* y, int m, double* x, double* Q)
{
    const int n = FX_NX;
    double Ay[FX_NX], lu[FX_NX * FX_NX], acc;
    int piv[FX_NX];
    if (m < n) return 2;
    for (int i = 0; i < n; i++) {
        acc = 0.0;
        for (int r = 0; r < m; r++) acc += A[i + r * n] * y[r];
        Ay[i] = 1.0 * acc;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            acc = 0.0;
            for (int r = 0; r < m; r++) acc += A[i + r * n] * A[j + r * n];
            Q[i + j * n] = 1.0 * acc;
        }
    memcpy(lu, Q, sizeof lu);
    if (fx_lu(lu, n, piv)) return 3;
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) Q[i + j * n] = 0.0;
        Q[j + j * n] = 1.0;
        fx_lu_solve(lu, n, piv, Q + j * n);
    }
    for (int i = 0; i < n; i++) {
        acc = 0.0;
        for (int k = 0; k < n; k++) acc += Q[i + k * n] * Ay[k];
        x[i] = 1.0 * acc;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ geometry */

/* solving.c:1225-1239: ECEF -> {latitude, longitude (rad), ellipsoidal height (m)}, WGS84 */
void ecef2pos(const double* r, double* pos)
{
    const double e2 = FX_FE * (2.0 - FX_FE), r2 = fx_dot(r, r, 2);
    double z, zk, v = FX_RE, sinp;
    for (z = r[2], zk = 0.0; fabs(z - zk) >= 1E-4;) {
        zk = z;
        sinp = z / sqrt(r2 + z * z);
        v = FX_RE / sqrt(1.0 - e2 * sinp * sinp);
        z = r[2] + v * e2 * sinp;
    }
    pos[0] = r2 > 1E-12 ? atan(z / sqrt(r2)) : (r[2] > 0.0 ? FX_PI / 2.0 : -FX_PI / 2.0);
    pos[1] = r2 > 1E-12 ? atan2(r[1], r[0]) : 0.0;
    pos[2] = sqrt(r2 + z * z) - v;
}

/* solving.c:1248-1258: range with the Sagnac term, unit line of sight; negative when the satellite has no state */
static double fx_range(const double* sat, const double* site, double* los)
{
    if (fx_norm(sat, 3) < FX_RE) return -1.0;
    for (int k = 0; k < 3; k++) los[k] = sat[k] - site[k];
    const double r = fx_norm(los, 3);
    for (int k = 0; k < 3; k++) los[k] /= r;
    return r + FX_OMGE * (sat[0] * site[1] - sat[1] * site[0]) / FX_CLIGHT;
}

/* solving.c:1268-1280 with xyz2enu :1289-1296 and the 3x3 product of ecef2enu :1339-1345: azimuth and elevation of a
 * line of sight seen from geodetic position geo */
static double fx_look_angles(const double* geo, const double* los, double* az_el)
{
    double az = 0.0, el = FX_PI / 2.0;
    if (geo[2] > -FX_RE) {
        const double sinp = sin(geo[0]), cosp = cos(geo[0]), sinl = sin(geo[1]), cosl = cos(geo[1]);
        double E[9], enu[3];
        E[0] = -sinl;        E[3] = cosl;         E[6] = 0.0;
        E[1] = -sinp * cosl; E[4] = -sinp * sinl; E[7] = cosp;
        E[2] = cosp * cosl;  E[5] = cosp * sinl;  E[8] = sinp;
        for (int i = 0; i < 3; i++) {
            double acc = 0.0;
            for (int k = 0; k < 3; k++) acc += E[i + k * 3] * los[k];
            enu[i] = 1.0 * acc;
        }
        az = fx_dot(enu, enu, 2) < 1E-12 ? 0.0 : atan2(enu[0], enu[1]);
        if (az < 0.0) az += 2 * FX_PI;
        el = asin(enu[2]);
    }
    az_el[0] = az;
    az_el[1] = el;
    return el;
}

/* ------------------------------------------------------------------------------------------ propagation delays */

/* solving.c:620-660: broadcast (Klobuchar) ionosphere delay on L1, metres.  The reference never decodes the eight
 * coefficients (subframe 4, page 18), so they are zero and the 2004 defaults below apply; gpsb_host_fix_set_iono
 * lets a caller who has them do better. */
static double fx_iono_delay(gtime_t t, const double* ion, const double* geo, const double* az_el)
{
    static const double ion_2004[8] = {
        0.1118E-07, -0.7451E-08, -0.5961E-07, 0.1192E-06,
        0.1167E+06, -0.2294E+06, -0.1311E+06, 0.1049E+07
    };
    double tt, f, psi, phi, lam, amp, per, x;
    int week;
    if (geo[2] < -1E3 || az_el[1] <= 0) return 0.0;
    if (fx_norm(ion, 8) <= 0.0) ion = ion_2004;

    psi = 0.0137 / (az_el[1] / FX_PI + 0.11) - 0.022;                 /* earth-centred angle, semicircles */
    phi = geo[0] / FX_PI + psi * cos(az_el[0]);                       /* sub-ionospheric point */
    if (phi > 0.416) phi = 0.416;
    else if (phi < -0.416) phi = -0.416;
    lam = geo[1] / FX_PI + psi * sin(az_el[0]) / cos(phi * FX_PI);
    phi += 0.064 * cos((lam - 1.617) * FX_PI);                        /* geomagnetic latitude */
    tt = 43200.0 * lam + time2gpst(t, &week);                         /* local time */
    tt -= floor(tt / 86400.0) * 86400.0;
    f = 1.0 + 16.0 * pow(0.53 - az_el[1] / FX_PI, 3.0);               /* slant factor */
    amp = ion[0] + phi * (ion[1] + phi * (ion[2] + phi * ion[3]));
    per = ion[4] + phi * (ion[5] + phi * (ion[6] + phi * ion[7]));
    amp = amp < 0.0 ? 0.0 : amp;
    per = per < 72000.0 ? 72000.0 : per;
    x = 2.0 * FX_PI * (tt - 50400.0) / per;
    return FX_CLIGHT * f * (fabs(x) < 1.57 ? 5E-9 + amp * (1.0 + x * x * (-0.5 + x * x / 24.0)) : 5E-9);
}

/* solving.c:679-700: Saastamoinen delay in a standard atmosphere, metres */
static double fx_tropo_delay(const double* geo, const double* az_el, double humi)
{
    const double temp0 = 15.0;
    double hgt, pres, temp, e, z, dry, wet;
    if (geo[2] < -100.0 || 1E4 < geo[2] || az_el[1] <= 0) return 0.0;
    hgt = geo[2] < 0.0 ? 0.0 : geo[2];
    pres = 1013.25 * pow(1.0 - 2.2557E-5 * hgt, 5.2568);
    temp = temp0 - 6.5E-3 * hgt + 273.16;
    e = 6.108 * humi * exp((17.15 * temp - 4684.0) / (temp - 38.45));
    z = FX_PI / 2.0 - az_el[1];
    dry = 0.0022768 * pres / (1.0 - 0.00266 * cos(2.0 * geo[0]) - 0.00028 * hgt / 1E3) / cos(z);
    wet = 0.002277 * (1255.0 / temp + 0.05) * e / cos(z);
    return dry + wet;
}

/* ------------------------------------------------------------------------------------------ broadcast orbits */

/* solving.c:1057-1079 with iode < 0: the record of this satellite whose toe is closest to `when`, within 2 h */
static const eph_t* fx_pick_eph(gtime_t when, int sat)
{
    const double limit = 7200.0 + 1.0;
    double best = limit + 1.0, age;
    int pick = -1;
    for (int i = 0; i < g_nav.n_eph; i++) {
        if (g_nav.eph[i]->sat != sat) continue;
        if ((age = fabs(timediff(g_nav.eph[i]->toe, when))) > limit) continue;
        if (age <= best) { pick = i; best = age; }
    }
    return pick < 0 ? NULL : g_nav.eph[pick];
}

/* solving.c:1044-1054: clock polynomial, the argument corrected by its own value twice */
static double fx_clock_bias(gtime_t t, const eph_t* e)
{
    double dt = timediff(t, e->toc);
    for (int k = 0; k < 2; k++) dt -= e->f0 + e->f1 * dt + e->f2 * dt * dt;
    return e->f0 + e->f1 * dt + e->f2 * dt * dt;
}

/* solving.c:1143-1150 */
static double fx_ura_variance(int ura)
{
    static const double metres[] = { 2.4, 3.4, 4.85, 6.85, 9.65, 13.65, 24.0, 48.0, 96.0, 192.0, 384.0, 768.0, 1536.0,
                                     3072.0, 6144.0 };
    return ura < 0 || 15 < ura ? FX_SQ(6144.0) : FX_SQ(metres[ura]);
}

/* solving.c:1165-1215: ECEF position and clock bias (with the relativistic term) of a satellite at time t.  When the
 * Kepler iteration does not settle the outputs are left as they are, like the reference. */
static void fx_orbit(gtime_t t, const eph_t* eph, double* pos, double* bias, double* var)
{
    double tk, M, E, Ek, sinE, cosE, u, r, i, O, sin2u, cos2u, x, y, sinO, cosO, cosi;
    int n;
    if (eph->A <= 0.0) { pos[0] = pos[1] = pos[2] = *bias = *var = 0.0; return; }
    tk = timediff(t, eph->toe);
    M = eph->M0 + (sqrt(FX_MU / (eph->A * eph->A * eph->A)) + eph->deln) * tk;
    for (n = 0, E = M, Ek = 0.0; fabs(E - Ek) > 1E-14 && n < 30; n++) {
        Ek = E;
        E -= (E - eph->e * sin(E) - M) / (1.0 - eph->e * cos(E));
    }
    if (n >= 30) return;
    sinE = sin(E); cosE = cos(E);

    u = atan2(sqrt(1.0 - eph->e * eph->e) * sinE, cosE - eph->e) + eph->omg;
    r = eph->A * (1.0 - eph->e * cosE);
    i = eph->i0 + eph->idot * tk;
    sin2u = sin(2.0 * u); cos2u = cos(2.0 * u);
    u += eph->cus * sin2u + eph->cuc * cos2u;
    r += eph->crs * sin2u + eph->crc * cos2u;
    i += eph->cis * sin2u + eph->cic * cos2u;
    x = r * cos(u); y = r * sin(u); cosi = cos(i);

    O = eph->OMG0 + (eph->OMGd - FX_OMGE) * tk - FX_OMGE * eph->toes;
    sinO = sin(O); cosO = cos(O);
    pos[0] = x * cosO - y * cosi * sinO;
    pos[1] = x * sinO + y * cosi * cosO;
    pos[2] = y * sin(i);

    tk = timediff(t, eph->toc);
    *bias = eph->f0 + eph->f1 * tk + eph->f2 * tk * tk;
    *bias -= 2.0 * sqrt(FX_MU * eph->A) * eph->e * sinE / FX_SQ(FX_CLIGHT);
    *var = fx_ura_variance(eph->sva);
}

/* One satellite of solving.c:910-958 / :966-1040 (with satpos :1109 and ephpos :1118-1140): state at the moment the
 * received signal left it.  Returns 1 when the satellite has a state. */
static int fx_satellite(fx_state* s, gtime_t teph, const obsd_t* obs, int i)
{
    double* pv = s->pv + 6 * i;
    double* clk = s->clk + 2 * i;
    double later_pos[3] = { 0.0, 0.0, 0.0 }, later_bias = 0.0;
    const double tt = 1E-3;
    for (int k = 0; k < 6; k++) pv[k] = 0.0;
    clk[0] = clk[1] = 0.0;
    s->pv_var[i] = 0.0;
    s->health[i] = 0;

    gtime_t t = timeadd(obs[i].time, -obs[i].P[0] / FX_CLIGHT);          /* transmission time by the satellite clock */
    const eph_t* eph = fx_pick_eph(teph, obs[i].sat);
    if (!eph) return 0;
    t = timeadd(t, -fx_clock_bias(t, eph));

    fx_orbit(t, eph, pv, clk, s->pv_var + i);
    fx_orbit(timeadd(t, tt), eph, later_pos, &later_bias, s->pv_var + i);
    s->health[i] = eph->svh;
    for (int k = 0; k < 3; k++) pv[k + 3] = (later_pos[k] - pv[k]) / tt;  /* velocity and drift by difference */
    clk[1] = (later_bias - clk[0]) / tt;

    if (clk[0] == 0.0) {                                                 /* solving.c:948-954; variance slot 0, as there */
        clk[0] = fx_clock_bias(t, eph);
        clk[1] = 0.0;
        s->pv_var[0] = FX_SQ(30.0);
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ estimator */

static void fx_clear_slot(fx_state* s, int i)
{
    s->used[i] = 0;
    azel[i * 2] = azel[1 + i * 2] = s->resid[i] = 0.0;
}

/* The measurement row of satellite i (solving.c:742-781 / :841-873): residual, design row, variance.
 * *accepted is the reference's count of valid satellites.  Its sliced driver never clears it (a function static,
 * solving.c:459 with :866), so the `ns` it reports is a running total modulo 256 - reproduced in fx_state.accepted,
 * because parity is on every field; the one-shot driver counts per pass in a local, like the reference's pntpos. */
static void fx_measure(fx_state* s, const obsd_t* obs, int i, int* accepted)
{
    double los[3], r, dion, vion, dtrp, vtrp;
    double* az_el = azel + i * 2;
    if ((r = fx_range(s->pv + i * 6, s->site, los)) <= 0.0 || fx_look_angles(s->geo, los, az_el) < 0.0) return;

    double P = obs[i].P[0];
    double tgd = 0.0;                                                    /* group delay, solving.c:600-610 */
    for (int k = 0; k < g_nav.n_eph; k++)
        if (g_nav.eph[k]->sat == obs[i].sat) { tgd = FX_CLIGHT * g_nav.eph[k]->tgd[0]; break; }
    P -= tgd;
    if (s->health[i]) return;

    dion = fx_iono_delay(obs[i].time, g_nav.ion, s->geo, az_el);
    vion = FX_SQ(dion * 0.5);
    dtrp = fx_tropo_delay(s->geo, az_el, 0.7);
    vtrp = FX_SQ(0.3 / (sin(az_el[1]) + 0.1));

    const int row = s->rows;
    const double modelled = (r + dion + dtrp + s->x[3] - FX_CLIGHT * s->clk[i * 2]);
    s->v[row] = P - modelled;
    for (int j = 0; j < FX_NX; j++) s->H[j + row * FX_NX] = j < 3 ? -los[j] : (j == 3 ? 1.0 : 0.0);
    s->used[i] = 1;
    s->resid[i] = s->v[row];
    (*accepted)++;

    const double vmeas = 0.0;
    const double elev_var = FX_SQ(1.0) * (FX_SQ(0.003) * (FX_SQ(0.003) + FX_SQ(0.003) / sin(az_el[1])));   /* :591-597 */
    s->w[row] = elev_var + s->pv_var[i] + vmeas + vion + vtrp;
    s->rows = row + 1;
}

/* solving.c:783-791 / :880-886: unit rows that pin the three system offsets this receiver has no data for */
static void fx_pin_offsets(fx_state* s)
{
    for (int i = 1; i < 4; i++) {
        s->v[s->rows] = 0.0;
        for (int j = 0; j < FX_NX; j++) s->H[j + s->rows * FX_NX] = j == i + 3 ? 1.0 : 0.0;
        s->w[s->rows++] = 0.01;
    }
}

static void fx_begin_pass(fx_state* s)
{
    for (int k = 0; k < 3; k++) s->site[k] = s->x[k];
    ecef2pos(s->site, s->geo);
    s->rows = 0;
}

/* solving.c:393-399 / :516-523 */
static void fx_whiten(fx_state* s)
{
    for (int j = 0; j < s->rows; j++) {
        const double sig = sqrt(s->w[j]);
        s->v[j] /= sig;
        for (int k = 0; k < FX_NX; k++) s->H[k + j * FX_NX] /= sig;
    }
}

static int fx_converged(const fx_state* s)
{
    const double* d = s->dx;
    return (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]) < 1E-8;
}

/* solving.c:416-435 / :572-586 */
static void fx_commit(const fx_state* s, const obsd_t* obs, int accepted, sol_t* sol)
{
    sol->type = 0;
    sol->time = timeadd(obs[0].time, -s->x[3] / FX_CLIGHT);
    sol->dtr[0] = s->x[3] / FX_CLIGHT;
    for (int j = 0; j < 6; j++) sol->rr[j] = j < 3 ? s->x[j] : 0.0;
    for (int j = 0; j < 3; j++) sol->qr[j] = (float)s->Q[j + j * FX_NX];
    sol->qr[3] = (float)s->Q[1];
    sol->qr[4] = (float)s->Q[2 + FX_NX];
    sol->qr[5] = (float)s->Q[2];
    sol->ns = (unsigned char)accepted;
    sol->age = sol->ratio = 0.0;
    sol->stat = SOLQ_SINGLE;
}

/* ------------------------------------------------------------------------------------------ sliced driver */

/* solving.c:966-1040: one satellite per call; 0 = more to do, 1 = all have a state, -1 = some have none */
static int fx_satellites_step(fx_state* s, gtime_t teph, const obsd_t* obs, int n)
{
    if (s->sat_slice == 0) s->sat_ok = 0;
    s->sat_ok += fx_satellite(s, teph, obs, s->sat_slice);
    if (++s->sat_slice == n) {
        s->sat_slice = 0;
        return s->sat_ok != n ? -1 : 1;
    }
    return 0;
}

/* solving.c:799-897: the residual of ONE satellite per call; the row count once the last one is in, else -1.  A
 * satellite that repeats its successor's number is dropped (the one-shot driver drops both, like the reference). */
static int fx_residual_step(fx_state* s, const obsd_t* obs, int n, int i)
{
    if (i == 0) fx_begin_pass(s);
    fx_clear_slot(s, i);
    if (!(i < n - 1 && obs[i].sat == obs[i + 1].sat)) fx_measure(s, obs, i, &s->accepted);
    if (i + 1 != n) return -1;
    fx_pin_offsets(s);
    return s->rows;
}

/* solving.c:453-588: n residual slices, then the solve, per pass; 0 = call again, 1 = fix, < 0 = no fix */
static int fx_estimate_step(fx_state* s, const obsd_t* obs, int n, sol_t* sol)
{
    int res = 0;
    if (s->pass == 0 && s->op == 0) {
        memset(s->x, 0, sizeof s->x);
        for (int k = 0; k < 3; k++) s->x[k] = sol->rr[k];              /* start from the previous fix */
    }
    if (s->op < n - 1) {
        if (fx_residual_step(s, obs, n, s->op) < 0) { s->op++; return 0; }
        res = -1;
    } else if (s->op == n - 1) {
        if (fx_residual_step(s, obs, n, s->op) >= FX_NX) { fx_whiten(s); s->op++; return 0; }
        res = -1;                                                        /* fewer than four usable satellites */
    } else {
        if (fx_normal_solve(s->H, s->v, s->rows, s->dx, s->Q) > 0) res = -2;
        else {
            for (int j = 0; j < FX_NX; j++) s->x[j] += s->dx[j];
            if (fx_converged(s)) res = 1;
            else { s->pass++; s->op = 0; }
        }
    }
    if (s->pass > FX_MAX_PASSES) res = -1;
    if (res > 0) fx_commit(s, obs, s->accepted, sol);
    if (res != 0) s->pass = s->op = 0;
    return res;
}

/* solving.c:185-266 (pntpos_iterative) on the registered ephemerides: 0 = call again, 1 = fix in *sol, < 0 = none */
static int fx_step(fx_state* s, const obsd_t* obs, int n, sol_t* sol)
{
    if (!s->solving) {
        memset(s->used, 0, sizeof s->used);
        memset(s->health, 0, sizeof s->health);
        sol->stat = SOLQ_NONE;
        if (n <= 0 || n > FX_MAX_SATS) return -2;
        sol->time = obs[0].time;
        s->solving = 1;
        s->stage = FX_SATELLITES;
    }
    int res = -1;
    if (s->stage == FX_SATELLITES) {
        res = fx_satellites_step(s, sol->time, obs, n);
        if (res == 0) return 0;
        s->stage = res > 0 ? FX_ESTIMATE : FX_IDLE;
        if (res > 0) return 0;
    }
    if (s->stage == FX_ESTIMATE) {
        res = fx_estimate_step(s, obs, n, sol);
        if (res == 0) return 0;
        s->stage = FX_IDLE;
        if (res > 0)
            for (int i = 0; i < 2 * n; i++) azel[i] = azel[i] * FX_R2D;  /* look angles in degrees for the display */
    }
    s->solving = 0;
    return res < 0 ? -1 : 1;
}

/* ------------------------------------------------------------------------------------------ public: reference names */

/* solving.c:104-112 */
void gps_pos_solve_init(gps_ch_t* channels)
{
    uint32_t n = gpsb_host_sat_cnt();
    if (n > FX_MAX_SATS) n = FX_MAX_SATS;
    for (uint32_t i = 0; channels && i < n; i++) g_nav.eph[i] = &channels[i].eph_data.eph;
    g_nav.n_eph = channels ? (int)n : 0;
}

/* solving.c:117-140: one slice per call until solving_is_busy() drops; the call after the fix converts it to
 * latitude / longitude in degrees (final_pos). */
void gps_pos_solve(obsd_t* obs_p)
{
    fx_state* s = &g_fx;
    if (s->converting) {
        ecef2pos(gps_sol.rr, final_pos);
        final_pos[0] = final_pos[0] * FX_R2D;
        final_pos[1] = final_pos[1] * FX_R2D;
        s->converting = 0;
    } else if (obs_p && fx_step(s, obs_p, g_nav.n_eph, &gps_sol) > 0) {
        s->converting = 1;
    }
}

uint8_t solving_is_busy(void) { return (uint8_t)(g_fx.solving | g_fx.converting); }            /* solving.c:268-271 */

/* gps_master.c:394-427: twice a second, once every channel holds subframes 1..3, start a fix from the current
 * observations and keep stepping it on the following calls. */
/* ONE observation array for the solver and the RTCM publisher, as in the reference (obsd, gps_master.c:41): there
 * gps_master_transmit_obs refreshes it on every idle slot, so a sliced solve in flight sees updated observations
 * between its slices when RTCM output is on.  Shared here for the same behaviour (host/rtcm.c uses it too). */
obsd_t hx_obsd[GPSB_FIX_MAX_SATS];

void gps_master_calculate_pos(gps_ch_t* channels)
{
    obsd_t* const obsd = hx_obsd;
    fx_state* s = &g_fx;
    if (solving_is_busy()) { gps_pos_solve(obsd); return; }
    const uint32_t now = signal_capture_get_packet_cnt();
    if ((now - s->last_request_ms) > 500) {
        s->last_request_ms = now;
        uint32_t n = gpsb_host_sat_cnt(), complete = 0;
        if (n > FX_MAX_SATS) n = FX_MAX_SATS;
        for (uint32_t i = 0; i < n; i++) complete += (channels[i].eph_data.received_mask_proc & 0x7) == 0x7;
        sdrobs2obsd(channels, (int)n, obsd);
        if (complete == n) gps_pos_solve(obsd);
    }
}

/* ------------------------------------------------------------------------------------------ public: this library */

/* The whole fix in one call (solving.c:153-181, pntpos, with estpos :376-448 and rescode :711-793), then the geodetic
 * conversion.  Differences from the sliced driver that the reference has too: a repeated satellite number drops both
 * entries, `ns` counts this pass only, and at most 10 passes.  One deliberate difference: the reference ignores a
 * singular normal matrix and goes on with stale increments; this returns "no fix".  Has its own work area: it may be
 * called while a sliced solve is in flight (both write gps_sol / azel when handed them, like the reference's two
 * drivers).  Returns 1 = fix, 0 = none. */
int gpsb_host_fix_once(const obsd_t* obs, int n, sol_t* sol, double pos_deg[3])
{
    fx_state* s = &g_once;
    if (!obs || !sol) return 0;
    sol->stat = SOLQ_NONE;
    if (n <= 0 || n > FX_MAX_SATS) return 0;
    sol->time = obs[0].time;
    for (int i = 0; i < n; i++) fx_satellite(s, sol->time, obs, i);
    memset(s->x, 0, sizeof s->x);
    for (int k = 0; k < 3; k++) s->x[k] = sol->rr[k];
    int fixed = 0;
    for (int pass = 0; pass < FX_MAX_PASSES && !fixed; pass++) {
        fx_begin_pass(s);
        int accepted = 0;
        for (int i = 0; i < n; i++) {
            fx_clear_slot(s, i);
            if (i < n - 1 && obs[i].sat == obs[i + 1].sat) { i++; continue; }
            fx_measure(s, obs, i, &accepted);
        }
        fx_pin_offsets(s);
        if (s->rows < FX_NX) break;
        fx_whiten(s);
        if (fx_normal_solve(s->H, s->v, s->rows, s->dx, s->Q) > 0) break;
        for (int j = 0; j < FX_NX; j++) s->x[j] += s->dx[j];
        if (fx_converged(s)) { fx_commit(s, obs, accepted, sol); fixed = 1; }
    }
    for (int i = 0; i < 2 * n; i++) azel[i] = azel[i] * FX_R2D;
    if (fixed && pos_deg) {
        ecef2pos(sol->rr, pos_deg);
        pos_deg[0] = pos_deg[0] * FX_R2D;
        pos_deg[1] = pos_deg[1] * FX_R2D;
    }
    return fixed;
}

/* Observations of the first n channels -> one fix -> gps_sol / final_pos, in one call: what a host with no 1-ms
 * deadline does instead of stepping gps_pos_solve.  Returns 1 = fix. */
int gpsb_host_fix_channels(gps_ch_t* channels, uint32_t n)
{
    static obsd_t obsd[FX_MAX_SATS];
    if (!channels || n == 0 || n > FX_MAX_SATS) return 0;
    sdrobs2obsd(channels, (int)n, obsd);
    return gpsb_host_fix_once(obsd, (int)n, &gps_sol, final_pos);
}

/* Where the next fix starts iterating from (the reference starts from the previous fix, the first one from the
 * centre of the Earth); an assisted start puts a rough position here. */
void gpsb_host_fix_set_start(const double ecef_m[3])
{
    for (int k = 0; k < 3; k++) gps_sol.rr[k] = ecef_m ? ecef_m[k] : 0.0;
}

void gpsb_host_fix_set_iono(const double coeff[8])
{
    for (int k = 0; k < 8; k++) g_nav.ion[k] = coeff ? coeff[k] : 0.0;
}

/* Forget a solve in flight, the request timer and the last fix (the reference has no such entry: its statics live
 * forever).  The running satellite count of the sliced driver is kept. */
void gpsb_host_fix_reset(void)
{
    fx_state* s = &g_fx;
    s->stage = FX_IDLE;
    s->sat_slice = s->sat_ok = s->pass = s->op = 0;
    s->solving = s->converting = 0;
    s->last_request_ms = 0;
    memset(&gps_sol, 0, sizeof gps_sol);
    memset(final_pos, 0, sizeof final_pos);
    memset(azel, 0, sizeof azel);
}

static uint64_t fx_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static uint32_t fx_bits32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* the solver's visible state as bit patterns (include/gpsb_flat_state.h) */
void gpsb_host_fix_state(gpsb_flat_fix* o)
{
    memset(o, 0, sizeof *o);
    o->stat = gps_sol.stat; o->ns = gps_sol.ns; o->type = gps_sol.type; o->busy = solving_is_busy();
    o->time_time = (int64_t)gps_sol.time.time;
    o->time_sec_bits = fx_bits(gps_sol.time.sec);
    for (int k = 0; k < 6; k++) { o->rr[k] = fx_bits(gps_sol.rr[k]); o->qr[k] = fx_bits32(gps_sol.qr[k]); }
    o->dtr0 = fx_bits(gps_sol.dtr[0]);
    for (int k = 0; k < 3; k++) o->final_pos[k] = fx_bits(final_pos[k]);
    for (int k = 0; k < 2 * GPSB_FLAT_FIX_SATS; k++) o->azel[k] = fx_bits(azel[k]);
}

/* assisted start / tests: write an ephemeris record into a channel from its flat form */
static double fx_dbl(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
void gpsb_host_channel_set_eph(gps_ch_t* ch, const gpsb_flat_eph* in)
{
    sdreph_t* d = &ch->eph_data;
    eph_t* e = &d->eph;
    e->sat = in->sat; e->iode = in->iode; e->iodc = in->iodc; e->sva = in->sva; e->svh = in->svh; e->week = in->week;
    e->code = in->code; e->flag = in->flag;
    e->toe.time = (time_t)in->toe_time; e->toc.time = (time_t)in->toc_time; e->ttr.time = (time_t)in->ttr_time;
    e->toe.sec = fx_dbl(in->toe_sec_bits); e->toc.sec = fx_dbl(in->toc_sec_bits); e->ttr.sec = fx_dbl(in->ttr_sec_bits);
    e->A = fx_dbl(in->A); e->e = fx_dbl(in->e); e->i0 = fx_dbl(in->i0); e->OMG0 = fx_dbl(in->OMG0);
    e->omg = fx_dbl(in->omg); e->M0 = fx_dbl(in->M0); e->deln = fx_dbl(in->deln); e->OMGd = fx_dbl(in->OMGd);
    e->idot = fx_dbl(in->idot); e->crc = fx_dbl(in->crc); e->crs = fx_dbl(in->crs); e->cuc = fx_dbl(in->cuc);
    e->cus = fx_dbl(in->cus); e->cic = fx_dbl(in->cic); e->cis = fx_dbl(in->cis); e->toes = fx_dbl(in->toes);
    e->fit = fx_dbl(in->fit); e->f0 = fx_dbl(in->f0); e->f1 = fx_dbl(in->f1); e->f2 = fx_dbl(in->f2);
    for (int k = 0; k < 4; k++) e->tgd[k] = fx_dbl(in->tgd[k]);
    d->ctype = in->ctype; d->week_gpst = in->week_gpst; d->cnt = in->cnt; d->cntth = in->cntth; d->update = in->update;
    d->prn = in->prn; d->week_gst = in->week_gst; d->sub_cnt = (uint16_t)in->sub_cnt;
    d->received_mask = (uint8_t)in->received_mask; d->received_mask_proc = (uint8_t)in->received_mask_proc;
    d->tow_gpst = fx_dbl(in->tow_gpst);
}
void gpsb_host_channel_set_obs(gps_ch_t* ch, double pseudorange_m, double tow_s)
{
    ch->obs_data.pseudorange_m = pseudorange_m;
    ch->obs_data.tow_s = tow_s;
}
uint32_t gpsb_host_sizeof_obsd(void) { return (uint32_t)sizeof(obsd_t); }
uint32_t gpsb_host_sizeof_sol(void) { return (uint32_t)sizeof(sol_t); }
