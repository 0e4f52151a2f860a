"""Rows N2 -> N3 -> N4 of SURVEY.md section 8(f) chained on the CPU: navigation BITS in, latitude / longitude out.

Four satellites on GPS-like orbits; their ephemerides are quantised and packed into subframes 1-3 by an encoder written
here from IS-GPS-200 (figure 20-1, table 20-III) - independently of the decoder under test - with parity per table
20-XIV; the bit streams go through the word assembler and the subframe decoder (nav.c / lc_decode_subframe), the
subframe stamps and code phases of a physically consistent scene through gps_master_nav_handling (master.c), and the
resulting pseudoranges through the position solver (fix.c).  Checked: every channel record, observation and solver
output equals the UNMODIFIED reference's along the way, the decoded ephemerides reproduce the orbits, and the fix lands
on the receiver."""
import ctypes as C

import numpy as np

from stm32f4_sdr_gps_b200 import load_host_library
from test_fix import (CLIGHT, WEEK, Pair, dbl, fix_diff, geodetic_to_ecef, make_sky, pseudorange)
from test_nav_decode import encode_word, host_eph

SC = 3.1415926535898                                            # semicircle -> radian, IS-GPS-200


def put(bits, first, n, value):
    """n bits of value, MSB first, at 1-based subframe bit `first` (two's complement for negative values)."""
    value &= (1 << n) - 1
    for k in range(n):
        bits[first - 1 + k] = (value >> (n - 1 - k)) & 1


def quantise(el, sva):
    """Orbit elements -> the integers the navigation message carries, and the elements those integers stand for."""
    def semi(x):
        return ((x / SC + 1.0) % 2.0) - 1.0
    raw = dict(
        crs=round(el["crs"] * 2 ** 5), deln=round(el["deln"] / SC * 2 ** 43), M0=round(semi(el["M0"]) * 2 ** 31),
        cuc=round(el["cuc"] * 2 ** 29), e=round(el["e"] * 2 ** 33), cus=round(el["cus"] * 2 ** 29),
        sqrtA=round(np.sqrt(el["A"]) * 2 ** 19), toe=int(el["toes"]) // 16, cic=round(el["cic"] * 2 ** 29),
        OMG0=round(semi(el["OMG0"]) * 2 ** 31), cis=round(el["cis"] * 2 ** 29), i0=round(el["i0"] / SC * 2 ** 31),
        crc=round(el["crc"] * 2 ** 5), omg=round(semi(el["omg"]) * 2 ** 31), OMGd=round(el["OMGd"] / SC * 2 ** 43),
        idot=round(el["idot"] / SC * 2 ** 43), af0=round(el["f0"] * 2 ** 31), af1=round(el["f1"] * 2 ** 43), af2=0,
        tgd=round(el["tgd"] * 2 ** 31), toc=int(el["toc"]) // 16, sva=sva, iode=77, iodc=0x100 | 77)
    back = dict(
        crs=raw["crs"] / 2 ** 5, deln=raw["deln"] / 2 ** 43 * SC, M0=raw["M0"] / 2 ** 31 * SC, cuc=raw["cuc"] / 2 ** 29,
        e=raw["e"] / 2 ** 33, cus=raw["cus"] / 2 ** 29, A=(raw["sqrtA"] / 2 ** 19) ** 2, toes=raw["toe"] * 16.0,
        cic=raw["cic"] / 2 ** 29, OMG0=raw["OMG0"] / 2 ** 31 * SC, cis=raw["cis"] / 2 ** 29, i0=raw["i0"] / 2 ** 31 * SC,
        crc=raw["crc"] / 2 ** 5, omg=raw["omg"] / 2 ** 31 * SC, OMGd=raw["OMGd"] / 2 ** 43 * SC,
        idot=raw["idot"] / 2 ** 43 * SC, f0=raw["af0"] / 2 ** 31, f1=raw["af1"] / 2 ** 43, f2=0.0, tgd=raw["tgd"] / 2 ** 31,
        toc=raw["toc"] * 16.0)
    return raw, back


def subframe_bits(rng, sf_id, tow_count, raw):
    """300 transmitted bits of subframe sf_id (IS-GPS-200 figure 20-1); spare / reserved bits random."""
    b = rng.integers(0, 2, 300, dtype=np.uint8)
    put(b, 1, 8, 0x8B)                                           # TLM preamble
    put(b, 31, 17, tow_count)                                    # HOW: time-of-week count of the NEXT subframe
    put(b, 50, 3, sf_id)
    if sf_id == 1:
        put(b, 61, 10, WEEK % 1024); put(b, 71, 2, 1); put(b, 73, 4, raw["sva"]); put(b, 77, 6, 0)
        put(b, 83, 2, raw["iodc"] >> 8); put(b, 197, 8, raw["tgd"]); put(b, 211, 8, raw["iodc"] & 255)
        put(b, 219, 16, raw["toc"]); put(b, 241, 8, raw["af2"]); put(b, 249, 16, raw["af1"]); put(b, 271, 22, raw["af0"])
    elif sf_id == 2:
        put(b, 61, 8, raw["iode"]); put(b, 69, 16, raw["crs"]); put(b, 91, 16, raw["deln"])
        put(b, 107, 8, (raw["M0"] >> 24) & 255); put(b, 121, 24, raw["M0"] & 0xFFFFFF)
        put(b, 151, 16, raw["cuc"]); put(b, 167, 8, raw["e"] >> 24); put(b, 181, 24, raw["e"] & 0xFFFFFF)
        put(b, 211, 16, raw["cus"]); put(b, 227, 8, raw["sqrtA"] >> 24); put(b, 241, 24, raw["sqrtA"] & 0xFFFFFF)
        put(b, 271, 16, raw["toe"]); put(b, 287, 1, 0)
    elif sf_id == 3:
        put(b, 61, 16, raw["cic"]); put(b, 77, 8, (raw["OMG0"] >> 24) & 255); put(b, 91, 24, raw["OMG0"] & 0xFFFFFF)
        put(b, 121, 16, raw["cis"]); put(b, 137, 8, (raw["i0"] >> 24) & 255); put(b, 151, 24, raw["i0"] & 0xFFFFFF)
        put(b, 181, 16, raw["crc"]); put(b, 197, 8, (raw["omg"] >> 24) & 255); put(b, 211, 24, raw["omg"] & 0xFFFFFF)
        put(b, 241, 24, raw["OMGd"]); put(b, 271, 8, raw["iode"]); put(b, 279, 14, raw["idot"])
    out = []
    prev29 = prev30 = 0
    for w in range(10):
        d = [int(x) for x in b[30 * w:30 * w + 24]]
        if w in (1, 9):                                          # the two non-information bits make D29 = D30 = 0
            for t in range(4):
                d[22], d[23] = t >> 1, t & 1
                tx = encode_word(d, prev29, prev30)
                if tx[28] == 0 and tx[29] == 0:
                    break
        tx = encode_word(d, prev29, prev30)
        prev29, prev30 = tx[28], tx[29]
        out += tx
    return np.array(out, np.uint8)


def test_navigation_bits_to_position(reference):
    lib = load_host_library()
    rl = reference.lib
    lib.gpsb_host_feed_nav_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    rl.ref_feed_nav_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.gps_master_nav_handling.argtypes = [C.c_void_p]
    rl.ref_nav_handling.argtypes = [C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(20200)
    prns = [7, 16, 21, 30]
    pair = Pair(reference, prns)
    ch, rchans = pair.ch, pair.rchans
    lat, lon, h = 37.42, -122.08, 12.0
    site = geodetic_to_ecef(lat, lon, h)
    first_tow_count = 40000                                      # HOW of the first subframe sent
    ids = [1, 2, 3, 4, 5, 1, 2, 3]
    t_end = (first_tow_count + len(ids) - 1) * 6.0              # GPS time at which the last of these subframes ends
    c_ms = CLIGHT / 1000.0
    while True:
        sky, raws = [], []
        for el in make_sky(rng, site, t_end, 4):
            el["toes"] = el["toc"] = float(int(t_end) // 7200 * 7200)
            raw, back = quantise(el, int(rng.integers(0, 4)))
            sky.append(back); raws.append(raw)
        flight0 = np.array([pseudorange(el, site, t_end + 0.075, 0.0) for el in sky]) / c_ms
        a0 = 50000.3 - flight0.min()
        arrival = a0 + flight0
        stamp = np.floor(arrival).astype(int)
        if np.all((arrival - stamp > 0.1) & (arrival - stamp < 0.9)) and np.all(flight0 > 60) and np.all(flight0 < 95):
            break
    ref_i = int(np.argmin(arrival))

    # ---- N2: the bit streams through the word assembler and the decoder, the bit edges known from tracking
    for i in range(4):
        stream = np.concatenate([rng.integers(0, 2, 40, dtype=np.uint8)] +
                                [subframe_bits(rng, sf, first_tow_count + k, raws[i]) for k, sf in enumerate(ids)])
        st = ch.snapshot(i)
        st.accurate_swap_ok, st.accurate_swap_time = 1, int(stamp[i]) % 20
        ch.restore(i, st)
        rch = reference.channel_at(rchans, i)
        reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
        ms0 = int(stamp[i]) + 3 - 20 * (stream.size - 1)         # the last bit closes 3 ms after its edge
        lib.gpsb_host_feed_nav_bits(ch.at(i), stream.ctypes.data, stream.size, ms0)
        rl.ref_feed_nav_bits(rch, stream.ctypes.data, stream.size, ms0)
        assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(rch))
        st = ch.snapshot(i)
        assert st.last_subframe_time == stamp[i] and st.subframe_cnt >= len(ids) - 1   # random lead-in bits may cost the first one
        e = host_eph(lib, ch.at(i))
        assert e.received_mask_proc & 7 == 7 and e.week == WEEK and dbl(e.tow_gpst) == t_end and e.sva == raws[i]["sva"]
        # the decoded record IS the orbit that was encoded (scale factors, bit positions, sign extension)
        for name in ("crs", "deln", "M0", "cuc", "e", "cus", "A", "toes", "cic", "OMG0", "cis", "i0", "crc", "omg", "OMGd",
                     "idot", "f0", "f1", "f2"):
            got, want = dbl(getattr(e, name)), sky[i][name]
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (i, name, got, want)
        assert abs(dbl(e.tgd[0]) - sky[i]["tgd"]) < 1e-18 and e.iode == 77 and e.iodc == 0x100 | 77

    # ---- N3 + N4: the idle loop
    def true_time(ms):
        return t_end + flight0[ref_i] / 1000.0 + (ms - arrival[ref_i]) / 1000.0

    fixes = []
    for now in range(50100, 53500, 17):
        t = true_time(now)
        for i, el in enumerate(sky):
            st = ch.snapshot(i)
            fine = np.float32((a0 + pseudorange(el, site, t, 0.0) / c_ms - stamp[i]) * 16368.0)
            assert 0 < fine < 16368
            st.code_phase_fine_bits = int(fine.view(np.uint32))
            if now == 50100:
                st.old_code_phase_fine_bits = st.code_phase_fine_bits
            filt = np.uint32(st.code_phase_fine_filt_bits).view(np.float32)
            for _ in range(16):
                filt = np.float32(filt + fine)
            st.code_phase_fine_filt_bits = int(np.float32(filt).view(np.uint32))
            st.code_filt_cnt += 16
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
        if got.stat == 5 and not got.busy:
            fixes.append((np.array([dbl(u) for u in got.rr[:3]]), dbl(got.final_pos[0]), dbl(got.final_pos[1])))
    assert len(fixes) > 60
    worst = max(np.linalg.norm(f[0] - site) for f in fixes)
    print('worst fix error, m:', worst)
    assert worst < 2000.0, worst
    assert abs(fixes[-1][1] - lat) < 0.02 and abs(fixes[-1][2] - lon) < 0.02
    pair.free()
