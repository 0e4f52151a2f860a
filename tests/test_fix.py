"""Row N4 of SURVEY.md section 8(f): position fix from pseudoranges and broadcast ephemerides (RTK/solving.c,
rtklib_common.c) - libgpsb_host.so against the UNMODIFIED reference compiled into oracle/_ref, every output double by
bit pattern.  The scenarios are synthetic: satellites placed on plausible GPS orbits above a chosen receiver site,
pseudoranges computed from an independent numpy propagation of the same elements."""
import ctypes as C
import struct

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, load_host_library
from test_nav_decode import FlatEph

MU, OMGE, CLIGHT = 3.9860050E14, 7.2921151467E-5, 299792458.0
RE, FE = 6378137.0, 1.0 / 298.257223563
WEEK = 2290


class FlatFix(C.Structure):
    """include/gpsb_flat_state.h, gpsb_flat_fix"""
    _fields_ = [("stat", C.c_int32), ("ns", C.c_int32), ("type", C.c_int32), ("busy", C.c_int32),
                ("time_time", C.c_int64), ("time_sec_bits", C.c_uint64), ("rr", C.c_uint64 * 6), ("qr", C.c_uint32 * 6),
                ("dtr0", C.c_uint64), ("final_pos", C.c_uint64 * 3), ("azel", C.c_uint64 * 64)]


def bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def dbl(u: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(u)))[0]


def geodetic_to_ecef(lat_deg, lon_deg, h):
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    e2 = FE * (2 - FE)
    v = RE / np.sqrt(1 - e2 * np.sin(lat) ** 2)
    return np.array([(v + h) * np.cos(lat) * np.cos(lon), (v + h) * np.cos(lat) * np.sin(lon),
                     (v * (1 - e2) + h) * np.sin(lat)])


def propagate(el, tow):
    """Satellite ECEF position and clock bias at GPS time of week `tow` (IS-GPS-200 table 20-IV), numpy."""
    tk = tow - el["toes"]
    n = np.sqrt(MU / el["A"] ** 3) + el["deln"]
    M = el["M0"] + n * tk
    E = M
    for _ in range(30):
        E = E - (E - el["e"] * np.sin(E) - M) / (1 - el["e"] * np.cos(E))
    u = np.arctan2(np.sqrt(1 - el["e"] ** 2) * np.sin(E), np.cos(E) - el["e"]) + el["omg"]
    r = el["A"] * (1 - el["e"] * np.cos(E))
    i = el["i0"] + el["idot"] * tk
    s2, c2 = np.sin(2 * u), np.cos(2 * u)
    u += el["cus"] * s2 + el["cuc"] * c2
    r += el["crs"] * s2 + el["crc"] * c2
    i += el["cis"] * s2 + el["cic"] * c2
    x, y = r * np.cos(u), r * np.sin(u)
    O = el["OMG0"] + (el["OMGd"] - OMGE) * tk - OMGE * el["toes"]
    pos = np.array([x * np.cos(O) - y * np.cos(i) * np.sin(O), x * np.sin(O) + y * np.cos(i) * np.cos(O), y * np.sin(i)])
    dt = tow - el["toc"]
    clk = el["f0"] + el["f1"] * dt + el["f2"] * dt * dt - 2 * np.sqrt(MU * el["A"]) * el["e"] * np.sin(E) / CLIGHT ** 2
    return pos, clk


def elements_above(rng, site, tow, toes, az_deg, el_deg):
    """Orbital elements of a satellite that stands near azimuth / elevation (deg) over `site` at time `tow`."""
    A = 26559800.0 + rng.uniform(-3e4, 3e4)
    lat = np.arctan2(site[2], np.hypot(site[0], site[1]))
    lon = np.arctan2(site[1], site[0])
    east = np.array([-np.sin(lon), np.cos(lon), 0.0])
    north = np.array([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)])
    up = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
    az, elv = np.radians(az_deg), np.radians(el_deg)
    d = np.cos(elv) * (np.sin(az) * east + np.cos(az) * north) + np.sin(elv) * up
    b = site @ d
    rho = -b + np.sqrt(b * b - (site @ site - A * A))
    S = site + rho * d
    inc = np.radians(55.0 + rng.uniform(-1.5, 1.5))
    sz = S[2] / (A * np.sin(inc))
    if abs(sz) > 0.98:
        return None
    u = np.arcsin(sz)
    if rng.integers(0, 2):
        u = np.pi - u
    x, y = A * np.cos(u), A * np.sin(u)
    O = np.arctan2(S[1], S[0]) - np.arctan2(y * np.cos(inc), x)
    tk = tow - toes
    ecc = rng.uniform(0.001, 0.015)
    omg = rng.uniform(-np.pi, np.pi)
    nu = u - omg
    Ea = 2 * np.arctan2(np.sqrt(1 - ecc) * np.sin(nu / 2), np.sqrt(1 + ecc) * np.cos(nu / 2))
    deln = rng.uniform(3e-9, 6e-9)
    M0 = (Ea - ecc * np.sin(Ea)) - (np.sqrt(MU / A ** 3) + deln) * tk
    OMGd = rng.uniform(-8.5e-9, -7.5e-9)
    return dict(A=A, e=ecc, i0=inc, OMG0=O - (OMGd - OMGE) * tk + OMGE * toes, omg=omg, M0=M0, deln=deln, OMGd=OMGd,
                idot=rng.uniform(-5e-10, 5e-10), crc=rng.uniform(-300, 300), crs=rng.uniform(-100, 100),
                cuc=rng.uniform(-5e-6, 5e-6), cus=rng.uniform(-5e-6, 1e-5), cic=rng.uniform(-2e-7, 2e-7),
                cis=rng.uniform(-2e-7, 2e-7), toes=float(toes), toc=float(toes), f0=rng.uniform(-5e-4, 5e-4),
                f1=rng.uniform(-1e-11, 1e-11), f2=0.0, tgd=rng.uniform(-1.2e-8, 1.2e-8))


def flat_eph(prn, el, sva=0, svh=0) -> FlatEph:
    f = FlatEph()
    f.sat = f.prn = prn
    f.iode, f.iodc, f.sva, f.svh, f.week, f.week_gpst = 17, 17, sva, svh, WEEK, WEEK
    for stamp, sec in (("toe", el["toes"]), ("toc", el["toc"]), ("ttr", el["toes"])):
        setattr(f, stamp + "_time", 315964800 + 604800 * WEEK + int(sec))
        setattr(f, stamp + "_sec_bits", bits(sec - int(sec)))
    for name in ("A", "e", "i0", "OMG0", "omg", "M0", "deln", "OMGd", "idot", "crc", "crs", "cuc", "cus", "cic", "cis",
                 "toes", "f0", "f1", "f2"):
        setattr(f, name, bits(el[name]))
    f.fit = bits(0.0)
    f.tgd[0] = bits(el["tgd"])
    f.received_mask = f.received_mask_proc = 7
    return f


def pseudorange(el, site, t_rx, rx_clock_s):
    """What a receiver at `site` whose clock reads t_rx (true time t_rx - rx_clock_s) measures."""
    t_true = t_rx - rx_clock_s
    tau = 0.075
    for _ in range(4):
        pos, clk = propagate(el, t_true - tau)
        th = OMGE * tau
        rot = np.array([pos[0] * np.cos(th) + pos[1] * np.sin(th), -pos[0] * np.sin(th) + pos[1] * np.cos(th), pos[2]])
        tau = np.linalg.norm(rot - site) / CLIGHT
    return CLIGHT * (tau + rx_clock_s - clk + el["tgd"])


def make_sky(rng, site, tow, n):
    toes = float(int(tow) // 7200 * 7200)
    out = []
    while len(out) < n:
        el = elements_above(rng, site, tow, toes, rng.uniform(0, 360), rng.uniform(15, 85))
        if el is not None:
            out.append(el)
    return out


class Pair:
    """The same channels on both sides."""

    def __init__(self, reference, prns):
        self.lib = lib = load_host_library()
        self.rl = rl = reference.lib
        self.reference = reference
        lib.gpsb_host_set_sat_cnt(len(prns))
        lib.gpsb_host_channel_set_eph.argtypes = [C.c_void_p, C.c_void_p]
        lib.gpsb_host_channel_set_obs.argtypes = [C.c_void_p, C.c_double, C.c_double]
        lib.gps_pos_solve_init.argtypes = [C.c_void_p]
        lib.gps_pos_solve.argtypes = [C.c_void_p]
        lib.solving_is_busy.restype = C.c_uint8
        lib.sdrobs2obsd.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.gpsb_host_fix_state.argtypes = [C.c_void_p]
        lib.gpsb_host_fix_channels.argtypes = [C.c_void_p, C.c_uint32]
        lib.gpsb_host_fix_set_start.argtypes = [C.c_void_p]
        rl.ref_channel_set_eph.argtypes = [C.c_void_p, C.c_void_p]
        rl.ref_channel_set_obs.argtypes = [C.c_void_p, C.c_double, C.c_double]
        rl.ref_fix_init.argtypes = [C.c_void_p]
        rl.ref_fix_run.argtypes = [C.c_void_p, C.c_uint32]
        rl.ref_fix_run.restype = C.c_uint32
        rl.ref_fix_once.argtypes = [C.c_void_p]
        rl.ref_fix_state.argtypes = [C.c_void_p]
        rl.ref_fix_set_start.argtypes = [C.c_void_p]
        rl.ref_obsd.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        self.n = len(prns)
        self.ch = Channels(prns)
        self.rchans = None
        if self.n == 4:                                        # the reference is compiled for GPS_SAT_CNT = 4
            self.rchans = reference.channels(4)
            for i, p in enumerate(prns):
                reference.channel_init(reference.channel_at(self.rchans, i), p, 0)
            rl.ref_fix_init(self.rchans)
            rl.ref_fix_clear()
        lib.gpsb_host_fix_reset()
        lib.gps_pos_solve_init(self.ch.base)
        self.start((0.0, 0.0, 0.0))

    def start(self, ecef):
        v = (C.c_double * 3)(*ecef)
        NS_OFFSET[0] = None
        self.lib.gpsb_host_fix_set_start(v)
        if self.rchans:
            self.rl.ref_fix_set_start(v)

    def set_eph(self, i, f: FlatEph):
        self.lib.gpsb_host_channel_set_eph(self.ch.at(i), C.byref(f))
        if self.rchans:
            self.rl.ref_channel_set_eph(self.reference.channel_at(self.rchans, i), C.byref(f))

    def set_obs(self, i, pr, tow):
        self.lib.gpsb_host_channel_set_obs(self.ch.at(i), pr, tow)
        if self.rchans:
            self.rl.ref_channel_set_obs(self.reference.channel_at(self.rchans, i), pr, tow)

    def run_sliced(self, max_calls=400):
        """sdrobs2obsd + gps_pos_solve until not busy; returns (calls, state) of this library."""
        obsd = (C.c_uint8 * (48 * self.n))()
        self.lib.sdrobs2obsd(self.ch.base, self.n, obsd)
        calls = 0
        while True:
            self.lib.gps_pos_solve(obsd)
            calls += 1
            if not self.lib.solving_is_busy() or calls >= max_calls:
                break
        return calls, self.state()

    def state(self) -> FlatFix:
        f = FlatFix()
        self.lib.gpsb_host_fix_state(C.byref(f))
        return f

    def ref_state(self) -> FlatFix:
        f = FlatFix()
        self.rl.ref_fix_state(C.byref(f))
        return f

    def free(self):
        self.lib.gpsb_host_fix_reset()
        self.lib.gps_pos_solve_init(None)                      # the solver holds pointers into the channel records
        self.ch.free()


NS_OFFSET = [None]


def fix_diff(a: FlatFix, b: FlatFix, n_azel=8):
    """Field-by-field differences.  `ns` is the sliced driver's never-cleared running count of accepted satellites
    (a function static in the reference that nothing can reset): the two sides' totals may start apart - tests that
    exercise only this library move only its counter - so they are compared modulo 256 up to an offset that must not
    change while both sides run the same solves."""
    out = []
    for name, _ in a._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        if name == "ns":
            if a.stat == 5 and b.stat == 5:
                off = (va - vb) % 256
                if NS_OFFSET[0] is None:
                    NS_OFFSET[0] = off
                if off != NS_OFFSET[0]:
                    out.append((name, va, vb, NS_OFFSET[0]))
            continue
        if hasattr(va, "__len__"):
            va, vb = list(va), list(vb)
            if name == "azel":
                va, vb = va[:n_azel], vb[:n_azel]
        if va != vb:
            out.append((name, va, vb))
    return out


def load_scene(pair, sky, site, t_rx, rx_clock, prns, sva=None, svh=None):
    for i, el in enumerate(sky):
        pair.set_eph(i, flat_eph(prns[i], el, sva[i] if sva else 0, svh[i] if svh else 0))
        pair.set_obs(i, pseudorange(el, site, t_rx, rx_clock), t_rx)


def test_record_layouts(reference):
    lib = load_host_library()
    assert reference.lib.ref_sizeof_obsd() == lib.gpsb_host_sizeof_obsd() == 48
    assert reference.lib.ref_sizeof_sol() == lib.gpsb_host_sizeof_sol() == 152
    assert C.sizeof(FlatFix) == 16 + 16 + 48 + 24 + 8 + 24 + 512


@pytest.mark.parametrize("seed", range(6))
def test_sliced_fix_equals_reference(reference, seed):
    """Cold start from the centre of the Earth, then four more fixes half a second apart, each starting from the
    previous one, stepped slice by slice like the reference's idle loop: number of calls until the solver reports idle,
    gps_sol (time, ECEF position, covariance, clock bias, status, the never-cleared satellite count), final_pos and the
    look angles equal the reference's bit for bit - and the position is where the receiver was put."""
    rng = np.random.default_rng(9000 + seed)
    prns = [int(p) for p in rng.choice(np.arange(1, 33), 4, replace=False)]
    pair = Pair(reference, prns)
    lat, lon, h = rng.uniform(-70, 70), rng.uniform(-180, 180), rng.uniform(0, 2500)
    site = geodetic_to_ecef(lat, lon, h)
    tow0 = float(rng.integers(20000, 580000)) + 0.37
    sky = make_sky(rng, site, tow0, 4)
    sva = [int(x) for x in rng.integers(0, 8, 4)]
    rx_clock = rng.uniform(-2e-3, 2e-3)
    for k in range(5):
        t_rx = tow0 + 0.5 * k
        load_scene(pair, sky, site, t_rx, rx_clock, prns, sva)
        want_calls = pair.rl.ref_fix_run(pair.rchans, 400)
        calls, got = pair.run_sliced()
        want = pair.ref_state()
        assert calls == want_calls, (k, calls, want_calls)
        assert not fix_diff(got, want), (k, fix_diff(got, want))
        assert got.stat == 5 and got.busy == 0
        fixed = np.array([dbl(u) for u in got.rr[:3]])
        assert np.linalg.norm(fixed - site) < 50.0             # ~7 m: the atmosphere models the synthetic truth does not have
        assert abs(dbl(got.final_pos[0]) - lat) < 0.01 and abs(dbl(got.dtr0) - rx_clock) < 2e-6
    # the observation records themselves
    mine = (C.c_uint8 * 192)()
    theirs = (C.c_uint8 * 192)()
    pair.lib.sdrobs2obsd(pair.ch.base, 4, mine)
    pair.rl.ref_obsd(pair.rchans, theirs, 192)
    assert bytes(mine) == bytes(theirs)
    pair.free()


def test_no_fix_cases_equal_reference(reference):
    """The ways a solve ends without a fix, each followed by a good one (the statics must be back in step):
    an unhealthy satellite, an ephemeris older than two hours, a satellite number missing from the ephemerides,
    a satellite below the horizon, the same satellite in two adjacent channels."""
    rng = np.random.default_rng(4242)
    prns = [3, 11, 19, 27]
    pair = Pair(reference, prns)
    site = geodetic_to_ecef(48.1, 11.6, 520.0)
    tow = 302400.25
    sky = make_sky(rng, site, tow, 4)
    low = None
    while low is None:
        low = elements_above(rng, site, tow, sky[0]["toes"], rng.uniform(0, 360), -25.0)

    def both(label, want_stat):
        want_calls = pair.rl.ref_fix_run(pair.rchans, 400)
        calls, got = pair.run_sliced()
        want = pair.ref_state()
        assert calls == want_calls, (label, calls, want_calls)
        assert not fix_diff(got, want), (label, fix_diff(got, want))
        assert got.stat == want_stat, label
        return got

    load_scene(pair, sky, site, tow, 1e-4, prns)
    both("good", 5)
    load_scene(pair, sky, site, tow + 1, 1e-4, prns, svh=[0, 0, 1, 0])
    both("unhealthy", 0)
    load_scene(pair, sky, site, tow + 2, 1e-4, prns)
    both("good again", 5)
    stale = dict(sky[1]); stale["toes"] = stale["toc"] = sky[1]["toes"] - 3 * 7200.0
    load_scene(pair, [sky[0], stale, sky[2], sky[3]], site, tow + 3, 1e-4, prns)
    both("stale ephemeris", 0)
    load_scene(pair, sky, site, tow + 4, 1e-4, prns)
    f = flat_eph(9, sky[2])                                    # channel 2 tracks PRN 19 but its record says 9
    pair.set_eph(2, f)
    both("no ephemeris for the satellite", 0)
    load_scene(pair, [sky[0], sky[1], sky[2], low], site, tow + 5, 1e-4, prns)
    both("below the horizon", 0)
    load_scene(pair, sky, site, tow + 6, 1e-4, prns)
    both("good", 5)
    pair.free()

    pair = Pair(reference, [8, 8, 21, 30])                     # the same satellite twice, adjacent
    load_scene(pair, [sky[0], sky[0], sky[2], sky[3]], site, tow + 7, 1e-4, [8, 8, 21, 30])
    both("duplicate", 0)
    pair.free()


@pytest.mark.parametrize("seed", range(3))
def test_one_shot_fix_equals_reference(reference, seed):
    """gpsb_host_fix_channels (the fix without the slicing) against the reference's pntpos."""
    rng = np.random.default_rng(7100 + seed)
    prns = [2, 13, 24, 31]
    pair = Pair(reference, prns)
    site = geodetic_to_ecef(rng.uniform(-60, 60), rng.uniform(-180, 180), rng.uniform(0, 1000))
    tow = float(rng.integers(20000, 580000))
    sky = make_sky(rng, site, tow, 4)
    for k in range(3):
        load_scene(pair, sky, site, tow + k, -3e-4, prns, svh=[0, 0, 0, 1] if k == 1 else None)
        want_ok = pair.rl.ref_fix_once(pair.rchans)
        got_ok = pair.lib.gpsb_host_fix_channels(pair.ch.base, 4)
        got, want = pair.state(), pair.ref_state()
        assert got_ok == want_ok == (0 if k == 1 else 1)
        assert not fix_diff(got, want), (k, fix_diff(got, want))
    pair.free()


def test_one_shot_fix_while_a_sliced_one_is_in_flight(reference):
    """The reference's pntpos and pntpos_iterative do not share work arrays, only the outputs: a one-shot fix taken
    in the middle of a sliced solve (here after 3, 6 and 11 of its slices) succeeds, and the sliced solve then runs
    to its end exactly as the reference's does - same number of calls, same outputs."""
    rng = np.random.default_rng(5150)
    prns = [1, 10, 20, 30]
    pair = Pair(reference, prns)
    pair.rl.ref_fix_steps.argtypes = [C.c_void_p, C.c_uint32]
    site = geodetic_to_ecef(-12.0, 130.8, 30.0)
    sky = make_sky(rng, site, 250000.0, 4)
    for k, after in enumerate((3, 6, 11)):
        load_scene(pair, sky, site, 250000.0 + k, 2e-4, prns)
        obsd = (C.c_uint8 * 192)()
        pair.lib.sdrobs2obsd(pair.ch.base, 4, obsd)
        for _ in range(after):
            pair.lib.gps_pos_solve(obsd)
        pair.rl.ref_fix_steps(pair.rchans, after)
        assert pair.lib.solving_is_busy() and pair.ref_state().busy
        assert pair.lib.gpsb_host_fix_channels(pair.ch.base, 4) == pair.rl.ref_fix_once(pair.rchans) == 1
        assert not fix_diff(pair.state(), pair.ref_state()), (after, fix_diff(pair.state(), pair.ref_state()))
        want_calls = pair.rl.ref_fix_run(pair.rchans, 400)
        calls, got = pair.run_sliced()
        assert calls == want_calls and not fix_diff(got, pair.ref_state()), (after, calls, want_calls)
        assert got.stat == 5 and not got.busy
    pair.free()


@pytest.mark.parametrize("n", [5, 8, 12])
def test_more_than_four_satellites(reference, n):
    """Beyond the reference's GPS_SAT_CNT: sliced and one-shot drivers agree with each other, and the fix tightens."""
    rng = np.random.default_rng(60 + n)
    prns = [int(p) for p in rng.choice(np.arange(1, 33), n, replace=False)]
    pair = Pair(reference, prns)
    site = geodetic_to_ecef(35.7, 139.7, 40.0)
    tow = 123456.5
    sky = make_sky(rng, site, tow, n)
    load_scene(pair, sky, site, tow, 5e-4, prns)
    calls, sliced = pair.run_sliced()
    assert sliced.stat == 5 and calls > n
    pair.start((0.0, 0.0, 0.0))
    assert pair.lib.gpsb_host_fix_channels(pair.ch.base, n) == 1
    once = pair.state()
    assert [x for x in fix_diff(sliced, once, 2 * n) if x[0] != "ns"] == []
    fixed = np.array([dbl(u) for u in once.rr[:3]])
    assert np.linalg.norm(fixed - site) < 100.0
    # the Python-side wrappers
    pair.start((0.0, 0.0, 0.0))
    fix = pair.ch.position_fix()
    assert fix is not None and np.array_equal(fix["ecef_m"], fixed) and abs(fix["lat_deg"] - 35.7) < 1e-3
    assert fix["azel_deg"].shape == (n, 2) and np.all(fix["azel_deg"][:, 1] > 5.0)
    frame = pair.ch.rtcm_observations()
    assert frame[0] == 0xD3 and len(frame) == 6 + ((frame[1] & 3) << 8 | frame[2])
    pair.lib.gpsb_host_set_sat_cnt(4)
    pair.free()


def test_idle_loop_observations_to_fix(reference):
    _idle_loop_observations_to_fix(reference, rtcm=False)


def test_idle_loop_observations_to_fix_with_rtcm_output_on():
    """The same timeline with the RTCM output switched ON on both sides (the reference ships with it compiled out,
    config.h:30; oracle/ref_master_rtcm_unit.c compiles its gps_master.c with the switch on into _ref/libgpsref_rtcm.so).
    gps_master_transmit_obs then runs ahead of gps_master_calculate_pos on every idle slot and refreshes the ONE observation
    array the sliced solver reads between its slices (gps_master.c:41, :279-285, :439): channel records, observations and the
    solver's state still equal the reference's after every call."""
    from oracle_lib import REF_SO, Reference
    so = REF_SO.parent / "libgpsref_rtcm.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libgpsref_rtcm.so not built (make -C oracle ref)")
    _idle_loop_observations_to_fix(Reference(so), rtcm=True)


def _idle_loop_observations_to_fix(reference, rtcm):
    """Rows N3 + N4 chained the way the reference's main loop runs them: gps_master_nav_handling every 17 ms on four
    channels whose subframe stamps and code phases follow a physically consistent scene (satellites on orbits, a
    receiver on the ground), ephemerides in place.  The code-phase filter closes windows, pseudoranges and times of week
    are assembled, gps_master_calculate_pos requests a fix twice a second and steps it one slice per call: the channel
    records, the observations and the solver's state equal the reference's after every call, and the fixes land on the
    receiver."""
    lib = load_host_library()
    rl = reference.lib
    lib.gps_master_nav_handling.argtypes = [C.c_void_p]
    lib.gpsb_host_enable_rtcm.argtypes = [C.c_int]
    lib.gpsb_host_enable_rtcm(1 if rtcm else 0)
    lib.gpsb_host_channel_set_tow.argtypes = [C.c_void_p, C.c_double]
    rl.ref_nav_handling.argtypes = [C.c_void_p, C.c_uint32]
    rl.ref_channel_set_tow.argtypes = [C.c_void_p, C.c_double]
    rng = np.random.default_rng(31337)
    prns = [6, 12, 22, 29]
    pair = Pair(reference, prns)
    ch, rchans = pair.ch, pair.rchans
    lat, lon, h = 59.93, 30.31, 15.0
    site = geodetic_to_ecef(lat, lon, h)
    tow0 = 388806.0                                             # the subframe edge every satellite sends at this time
    c_ms = CLIGHT / 1000.0
    while True:
        sky = make_sky(rng, site, tow0, 4)
        flight0 = np.array([pseudorange(el, site, tow0 + 0.075, 0.0) for el in sky]) / c_ms     # ms, incl. satellite clocks
        a0 = 9000.3 - flight0.min()                             # receiver ms counter at which a zero-delay signal would arrive
        arrival = a0 + flight0
        stamp = np.floor(arrival).astype(int)
        if np.all((arrival - stamp > 0.1) & (arrival - stamp < 0.9)):       # no code-epoch wrap during the run
            break
    ref_i = int(np.argmin(arrival))
    for i, el in enumerate(sky):
        pair.set_eph(i, flat_eph(prns[i], el))
        lib.gpsb_host_channel_set_tow(ch.at(i), tow0)
        rl.ref_channel_set_tow(reference.channel_at(rchans, i), tow0)

    def true_time(ms):                                          # GPS time at receiver millisecond ms
        return tow0 + flight0[ref_i] / 1000.0 + (ms - arrival[ref_i]) / 1000.0

    fixes = []
    busy_calls = 0
    for now in range(10000, 17000, 17):
        t = true_time(now)
        for i, el in enumerate(sky):
            st = ch.snapshot(i)
            flight = pseudorange(el, site, t, 0.0) / c_ms
            fine = np.float32((a0 + flight - stamp[i]) * 16368.0)
            assert 0 < fine < 16368
            st.code_phase_fine_bits = int(fine.view(np.uint32))
            if now == 10000:
                st.old_code_phase_fine_bits = st.code_phase_fine_bits      # tracking has been running: no false wrap
            filt = np.uint32(st.code_phase_fine_filt_bits).view(np.float32)
            for _ in range(16):
                filt = np.float32(filt + fine)
            st.code_phase_fine_filt_bits = int(np.float32(filt).view(np.uint32))
            st.code_filt_cnt += 16
            st.last_subframe_time = int(stamp[i])
            ch.restore(i, st)
            rch = reference.channel_at(rchans, i)
            reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
        lib.gpsb_host_set_packet_cnt(now)
        lib.gps_master_nav_handling(ch.at(0))
        rl.ref_nav_handling(rchans, now)
        for i in range(4):
            assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(reference.channel_at(rchans, i))), (now, i)
        got, want = pair.state(), pair.ref_state()
        assert not fix_diff(got, want), (now, fix_diff(got, want))
        busy_calls += 1 if got.busy else 0
        if got.stat == 5 and not got.busy:
            fixes.append(np.array([dbl(u) for u in got.rr[:3]]))
    if rtcm:
        # The point of this variant is the equality after every call, above.  (With its RTCM output on, the reference
        # refreshes the observations under a solve in flight on every idle slot; on this timeline its sliced solver then
        # keeps starting over - here exactly as there.)
        assert busy_calls > 100
        lib.gpsb_host_enable_rtcm(0)
        pair.free()
        return
    assert len(fixes) > 100                                     # a fix stood for most of the run
    distinct = {tuple(f) for f in fixes}
    assert len(distinct) >= 5                                   # and was renewed twice a second
    worst = max(np.linalg.norm(f - site) for f in fixes)
    assert worst < 2000.0, worst        # exact observations give ~7 m; the assembly's time tags run a flight time late
    lib.gpsb_host_enable_rtcm(0)
    pair.free()
