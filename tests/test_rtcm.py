"""Row N4 of SURVEY.md section 8(f), second half: RTCM 3 frames 1019 (ephemeris) and 1075 (MSM5 observations) -
libgpsb_host.so against the UNMODIFIED reference (GPS/obs_publish.c with its compile-time switch turned on by
oracle/ref_rtcm_unit.c, GPS/RTK/rtcm3e.c), byte for byte, plus an independent parse of the frames."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, load_host_library
from test_fix import FlatEph, Pair, bits, flat_eph, WEEK
from test_nav_decode import host_eph


class Obsd(C.Structure):
    """obsd_t, rtk_common.h:50-59"""
    _fields_ = [("time", C.c_int64), ("sec", C.c_double), ("sat", C.c_uint8), ("rcv", C.c_uint8), ("SNR", C.c_uint8),
                ("LLI", C.c_uint8), ("code", C.c_uint8), ("L", C.c_double), ("P", C.c_double), ("D", C.c_float)]


SINK = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_uint32)


def crc24q(data: bytes) -> int:
    crc = 0
    for b in data:
        crc ^= b << 16
        for _ in range(8):
            crc <<= 1
            if crc & 0x1000000:
                crc ^= 0x1864CFB
    return crc & 0xFFFFFF


def field(frame: bytes, pos: int, n: int, signed=False) -> int:
    v = 0
    for k in range(pos, pos + n):
        v = (v << 1) | ((frame[k >> 3] >> (7 - (k & 7))) & 1)
    if signed and v >> (n - 1):
        v -= 1 << n
    return v


def check_frame(frame: bytes, msg: int):
    assert frame[0] == 0xD3 and field(frame, 8, 6) == 0
    assert field(frame, 14, 10) == len(frame) - 6
    assert crc24q(frame[:-3]) == int.from_bytes(frame[-3:], "big")
    assert field(frame, 24, 12) == msg


class Frames:
    """Frames of this library (through the registered sink) and of the reference (through its UART stub)."""

    def __init__(self, reference):
        self.lib = lib = load_host_library()
        self.rl = rl = reference.lib
        self.got = []
        self._cb = SINK(lambda p, n: self.got.append(bytes(p[:n])))
        lib.gpsb_host_set_rtcm_sink.argtypes = [SINK, C.c_void_p]
        lib.gpsb_host_set_rtcm_sink(self._cb, None)
        lib.sendrtcmnav.argtypes = [C.c_void_p]
        lib.sendrtcmobs.argtypes = [C.c_void_p, C.c_int]
        lib.sdrobs2obsd.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for fn in (rl.ref_rtcm_obs, rl.ref_rtcm_nav):
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
            fn.restype = C.c_uint32
        rl.ref_rtcm_obs_records.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        rl.ref_rtcm_obs_records.restype = C.c_uint32

    def mine_nav(self, ch_ptr) -> bytes:
        self.got.clear()
        self.lib.sendrtcmnav(ch_ptr)
        return self.got[-1]

    def mine_obs(self, records, n) -> bytes:
        self.got.clear()
        self.lib.sendrtcmobs(records, n)
        return self.got[-1]

    def ref(self, fn, *args) -> bytes:
        out = (C.c_uint8 * 1200)()
        n = fn(*args, out, 1200)
        return bytes(out[:n])

    def close(self):
        self.lib.gpsb_host_set_rtcm_sink(SINK(0), None)


def test_ephemeris_frames_equal_reference(reference):
    """Message 1019 from ephemerides decoded out of random subframe images (every field at arbitrary values, extremes
    included) and from the orbit-like records of the fix tests: equal to the reference's frame, well formed, and the
    fields that the frame carries unscaled come back out of it."""
    fr = Frames(reference)
    lib, rl = fr.lib, reference.lib
    lib.gpsb_host_channel_set_eph.argtypes = [C.c_void_p, C.c_void_p]
    rl.ref_channel_set_eph.argtypes = [C.c_void_p, C.c_void_p]
    lib.gps_nav_data_decode_subframe.argtypes = [C.c_void_p]
    rng = np.random.default_rng(1019)
    prns = [1, 9, 23, 32]
    ch = Channels(prns)
    rchans = reference.channels(4)
    for i, p in enumerate(prns):
        reference.channel_init(reference.channel_at(rchans, i), p, 0)
    for trial in range(60):
        i = trial % 4
        if trial % 3 == 2:
            from test_fix import elements_above, geodetic_to_ecef
            el = None
            while el is None:
                el = elements_above(rng, geodetic_to_ecef(10.0, 20.0, 0.0), 200000.0, 194400.0, rng.uniform(0, 360), 40.0)
            f = flat_eph(prns[i], el, sva=int(rng.integers(0, 16)), svh=int(rng.integers(0, 64)))
        else:                                                    # decode three random subframe images into the record
            for sf_id in (1, 2, 3):
                img = rng.integers(0, 256, 38, dtype=np.uint8)
                if trial % 5 == 4:
                    img[:] = 0xFF if sf_id != 2 else 0x00
                for k in range(3):
                    bit = (sf_id >> (2 - k)) & 1
                    img[(49 + k) >> 3] = (int(img[(49 + k) >> 3]) & ~(1 << ((49 + k) & 7))) | (bit << ((49 + k) & 7))
                st = ch.snapshot(i)
                st.subframe_data[:] = list(img)
                ch.restore(i, st)
                lib.gps_nav_data_decode_subframe(ch.at(i))
            f = host_eph(lib, ch.at(i))
        lib.gpsb_host_channel_set_eph(ch.at(i), C.byref(f))
        rl.ref_channel_set_eph(reference.channel_at(rchans, i), C.byref(f))
        got = fr.mine_nav(ch.at(i))
        want = fr.ref(rl.ref_rtcm_nav, reference.channel_at(rchans, i))
        assert got == want, (trial, got.hex(), want.hex())
        check_frame(got, 1019)
        assert len(got) == 67                                    # 488 body bits
        assert field(got, 36, 6) == prns[i] and field(got, 42, 10) == f.week % 1024
        assert field(got, 52, 4) == f.sva & 15 and field(got, 72, 8) == f.iode & 255
    ch.free()
    fr.close()


def make_records(rows):
    rec = (Obsd * len(rows))()
    for r, row in zip(rec, rows):
        r.time = 315964800 + 604800 * WEEK + int(row["tow"])
        r.sec = row["tow"] - int(row["tow"])
        r.sat, r.rcv, r.SNR, r.LLI, r.code = row["sat"], 1, row.get("snr", 160), row.get("lli", 0), row.get("code", 1)
        r.L, r.P, r.D = row.get("L", 0.0), row["P"], row.get("D", 0.0)
    return rec


def aliasing_free(sats, stamped=()):
    """The reference indexes a 4-entry table by satellite number (rtcm3e.c:351); for PRN p > 5 the slot it reads
    overlays the lock-time stamp of satellite (p - 6) // 2 + 1 - set when that satellite came earlier in the same frame
    or, for satellites above 4 (whose stamps survive from frame to frame), in any frame since the last ephemeris frame.
    This library reads the in-bounds value instead (rtcm.c header); the parity cases stay clear of the overlay."""
    for k, p in enumerate(sats):
        if p > 5:
            other = (p - 6) // 2 + 1
            if other in sats[:k] or other in stamped:
                return False
    return True


def test_observation_frames_equal_reference(reference):
    """Message 1075 from observation records: random pseudoranges / Doppler / C/N0 on random satellite sets, then the
    edge cases - zero pseudorange, zero Doppler, values beyond the field ranges, carrier phase with and without the
    loss-of-lock and half-cycle bits, a repeated satellite, satellite numbers outside GPS, signals without an MSM
    number.  The frame buffer is shared between messages (stale bits show through the skipped info nibbles), so both
    sides go through the same sequence."""
    fr = Frames(reference)
    lib, rl = fr.lib, reference.lib
    rng = np.random.default_rng(1075)
    pair = Pair(reference, [3, 4, 17, 28])
    # both buffers start from the same contents: an ephemeris frame clears them
    pair.set_eph(0, flat_eph(3, dict(A=2.65e7 ** 1, e=0.01, i0=0.96, OMG0=1.0, omg=0.5, M0=0.2, deln=4e-9, OMGd=-8e-9,
                                     idot=1e-10, crc=200.0, crs=-50.0, cuc=1e-6, cus=2e-6, cic=1e-7, cis=-1e-7,
                                     toes=194400.0, toc=194400.0, f0=1e-4, f1=1e-12, f2=0.0, tgd=5e-9)))
    assert fr.mine_nav(pair.ch.at(0)) == fr.ref(rl.ref_rtcm_nav, reference.channel_at(pair.rchans, 0))
    nav_pair, nav_ch = pair, pair.ch

    def both(rows, label):
        rec = make_records(rows)
        got = fr.mine_obs(rec, len(rows))
        want = fr.ref(rl.ref_rtcm_obs_records, rec, len(rows))
        assert got == want, (label, got.hex(), want.hex())
        check_frame(got, 1075)
        return got

    stamped = set()                                              # satellites above 4 stamped since the last 1019
    for trial in range(200):
        if trial % 8 == 7:                                       # an ephemeris frame clears the message state
            assert fr.mine_nav(nav_ch.at(0)) == fr.ref(rl.ref_rtcm_nav, reference.channel_at(nav_pair.rchans, 0))
            stamped.clear()
        while True:
            sats = [int(s) for s in rng.choice(np.arange(1, 33), 4, replace=False)]
            if aliasing_free(sats, stamped):
                break
        stamped |= {s for s in sats if s > 4}
        tow = float(rng.integers(0, 604800)) + float(rng.integers(0, 1000)) / 1000.0
        rows = [dict(sat=s, tow=tow + 0.004 * k, P=float(rng.uniform(1.9e7, 2.6e7)), D=float(rng.uniform(-5000, 5000)),
                     snr=int(rng.integers(0, 256))) for k, s in enumerate(sats)]
        frame = both(rows, ("random", trial))
        if trial == 0:
            assert field(frame, 48, 30) == round(tow * 1000)
            mask = field(frame, 73 + 24, 32) << 32 | field(frame, 73 + 24 + 32, 32)
            assert {64 - b for b in range(64) if mask >> b & 1} == set(sats)
            assert field(frame, 73 + 24 + 64, 32) == 1 << 30       # signal 2 = L1 C/A
    assert fr.mine_nav(nav_ch.at(0)) == fr.ref(rl.ref_rtcm_nav, reference.channel_at(nav_pair.rchans, 0))
    t = 345600.5
    both([dict(sat=2, tow=t, P=0.0, D=0.0), dict(sat=3, tow=t, P=2.2e7, D=0.0), dict(sat=4, tow=t, P=2.3e7, D=1200.0),
          dict(sat=5, tow=t, P=2.4e7, D=-1200.0)], "zero pseudorange / Doppler")
    both([dict(sat=1, tow=t, P=8.0e7, D=60000.0), dict(sat=2, tow=t, P=-5.0, D=-60000.0), dict(sat=3, tow=t, P=7.64e7, D=0.5),
          dict(sat=4, tow=t, P=1.0, D=-0.5)], "beyond the field ranges")
    lam = 299792458.0 / 1.57542e9
    both([dict(sat=1, tow=t, P=2.1e7, L=2.1e7 / lam + 50.3, D=100.0), dict(sat=2, tow=t, P=2.2e7, L=2.2e7 / lam - 4000.2, lli=1),
          dict(sat=3, tow=t, P=2.3e7, L=2.3e7 / lam + 1.0e4, lli=2), dict(sat=4, tow=t, P=2.4e7, L=-3.0, lli=3)], "carrier phase")
    both([dict(sat=4, tow=t, P=2.1e7, D=10.0), dict(sat=4, tow=t + 2.0, P=2.2e7, D=20.0), dict(sat=1, tow=t, P=2.3e7),
          dict(sat=3, tow=t, P=2.4e7)], "repeated satellite")
    both([dict(sat=0, tow=t, P=2.1e7), dict(sat=33, tow=t, P=2.2e7), dict(sat=200, tow=t, P=2.3e7), dict(sat=5, tow=t, P=2.4e7)],
         "outside GPS")
    both([dict(sat=1, tow=t, P=2.1e7, code=0), dict(sat=2, tow=t, P=2.2e7, code=6), dict(sat=3, tow=t, P=2.3e7, code=14),
          dict(sat=4, tow=t, P=2.4e7, code=49)], "signals")
    both([dict(sat=1, tow=t, P=2.1e7, code=0), dict(sat=2, tow=t, P=2.2e7, code=9), dict(sat=3, tow=t, P=2.3e7, code=60),
          dict(sat=4, tow=t, P=2.4e7, code=23)], "nothing to send")
    fr.close()


def test_channels_to_frames_through_the_master(reference):
    """gps_master_transmit_obs as the idle slot runs it when RTCM is enabled: a channel whose subframe 1 just arrived
    sends its ephemeris (and loses the flags of subframes 1-3, the reference's precedence slip), otherwise the
    observations go out every 200 ms; the frames equal sendrtcmnav / sendrtcmobs of the reference on the same channels."""
    fr = Frames(reference)
    lib, rl = fr.lib, reference.lib
    lib.gps_master_transmit_obs.argtypes = [C.c_void_p]
    lib.gpsb_host_enable_rtcm.argtypes = [C.c_int]
    prns = [2, 3, 4, 5]
    pair = Pair(reference, prns)
    rng = np.random.default_rng(5)
    from test_fix import make_sky, geodetic_to_ecef, load_scene
    site = geodetic_to_ecef(-33.9, 151.2, 30.0)
    sky = make_sky(rng, site, 100000.0, 4)
    load_scene(pair, sky, site, 100000.25, 1e-4, prns)
    for i in range(4):
        st = pair.ch.snapshot(i)
        st.if_freq_offset_hz_bits = int(np.float32(rng.uniform(-4000, 4000)).view(np.uint32))
        st.snr_value_bits = int(np.float32(rng.uniform(5, 25)).view(np.uint32))
        pair.ch.restore(i, st)
        rch = reference.channel_at(pair.rchans, i)
        reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
    lib.gpsb_host_enable_rtcm(1)
    sent = []
    for now in range(1000, 3000, 17):
        lib.gpsb_host_set_packet_cnt(now)
        fr.got.clear()
        lib.gps_master_transmit_obs(pair.ch.base)
        if fr.got:
            sent.append((now, fr.got[-1]))
    lib.gpsb_host_enable_rtcm(0)
    kinds = [field(f, 24, 12) for _, f in sent]
    assert kinds[:4] == [1019] * 4 and set(kinds[4:]) == {1075}           # four ephemerides first, one per call
    gaps = np.diff([t for t, f in sent[4:]])
    assert gaps.min() > 200 and gaps.max() <= 217
    for i in range(4):
        assert host_eph(lib, pair.ch.at(i)).received_mask == 0
        assert sent[i][1] == fr.ref(rl.ref_rtcm_nav, reference.channel_at(pair.rchans, i))
    assert sent[4][1] == fr.ref(rl.ref_rtcm_obs, pair.rchans)
    pair.free()
    fr.close()


def test_ephemeris_frame_round_trip_without_the_reference():
    """Navigation-message integers -> subframes 1-3 (encoder written from IS-GPS-200) -> word assembler + decoder ->
    eph_t -> message 1019 -> a field parser written from RTCM 10403 table 3.5-21: the integers come back.  The
    navigation message and message 1019 use the same scale factors, so this checks the decoder and the encoder against
    the standards, not against the reference."""
    from test_bits_to_position import quantise, subframe_bits
    from test_fix import geodetic_to_ecef, make_sky
    lib = load_host_library()
    lib.gpsb_host_feed_nav_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.sendrtcmnav.argtypes = [C.c_void_p]
    lib.gpsb_host_set_rtcm_sink.argtypes = [SINK, C.c_void_p]
    frames = []
    cb = SINK(lambda p, n: frames.append(bytes(p[:n])))
    lib.gpsb_host_set_rtcm_sink(cb, None)
    rng = np.random.default_rng(3)
    site = geodetic_to_ecef(0.0, 0.0, 0.0)
    layout = [("msg", 12, 0), ("sat", 6, 0), ("week", 10, 0), ("sva", 4, 0), ("code", 2, 0), ("idot", 14, 1), ("iode", 8, 0),
              ("toc", 16, 0), ("af2", 8, 1), ("af1", 16, 1), ("af0", 22, 1), ("iodc", 10, 0), ("crs", 16, 1), ("deln", 16, 1),
              ("M0", 32, 1), ("cuc", 16, 1), ("e", 32, 0), ("cus", 16, 1), ("sqrtA", 32, 0), ("toe", 16, 0), ("cic", 16, 1),
              ("OMG0", 32, 1), ("cis", 16, 1), ("i0", 32, 1), ("crc", 16, 1), ("omg", 32, 1), ("OMGd", 24, 1), ("tgd", 8, 1),
              ("svh", 6, 0), ("flag", 1, 0), ("fit", 1, 0)]
    for trial, prn in enumerate((2, 17, 31)):
        el = make_sky(rng, site, 90000.0, 1)[0]
        el["toes"] = el["toc"] = 86400.0 + 7200.0 * trial
        raw, _ = quantise(el, sva=trial + 1)
        stream = np.concatenate([rng.integers(0, 2, 30, dtype=np.uint8)] +
                                [subframe_bits(rng, sf, 15000 + k, raw) for k, sf in enumerate((1, 2, 3))])
        ch = Channels([prn])
        lib.gpsb_host_feed_nav_bits(ch.at(0), stream.ctypes.data, stream.size, 1000)
        assert host_eph(lib, ch.at(0)).received_mask & 7 == 7
        frames.clear()
        lib.sendrtcmnav(ch.at(0))
        frame = frames[-1]
        check_frame(frame, 1019)
        got, at = {}, 24
        for name, bits_, signed in layout:
            got[name] = field(frame, at, bits_, bool(signed))
            at += bits_
        assert at == 24 + 488
        assert got["sat"] == prn and got["week"] == WEEK % 1024 and got["sva"] == trial + 1 and got["code"] == 1
        for name in ("idot", "iode", "toc", "af2", "af1", "af0", "iodc", "crs", "deln", "M0", "cuc", "e", "cus", "sqrtA",
                     "toe", "cic", "OMG0", "cis", "i0", "crc", "omg", "OMGd", "tgd"):
            assert got[name] == raw[name], (prn, name, got[name], raw[name])
        # the decoder keeps the raw fit-interval FLAG in eph->fit (nav_data_decode.c) where the encoder expects hours
        # (rtcm3e.c:218, fit > 0 ? 0 : 1): a flag of 0 goes out as DF137 = 1 - the reference's behaviour, kept
        assert got["svh"] == 0 and got["fit"] == 1
        ch.free()
    lib.gpsb_host_set_rtcm_sink(SINK(0), None)
