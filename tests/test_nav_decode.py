"""Row N2 of SURVEY.md section 8(f): subframe -> ephemeris / clock fields (nav_data_decode.c) and the word assembler
that triggers it (nav_data.c:257-352), libgpsb_host.so against the UNMODIFIED reference on the CPU.

The decode is one source (core/gpsb_loop_core.h, lc_decode_subframe) compiled into the host library and into the
device-resident loop; the device side is checked in tests/test_gpu_loop.py on a recording that carries real
subframes."""
import ctypes as C

import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import Channels, load_host_library


class FlatEph(C.Structure):
    """include/gpsb_flat_state.h, gpsb_flat_eph"""
    _fields_ = ([(n, C.c_int32) for n in ("sat", "iode", "iodc", "sva", "svh", "week", "code", "flag")] +
                [(n, C.c_int64) for n in ("toe_time", "toc_time", "ttr_time")] +
                [(n, C.c_uint64) for n in ("toe_sec_bits", "toc_sec_bits", "ttr_sec_bits", "A", "e", "i0", "OMG0", "omg",
                                           "M0", "deln", "OMGd", "idot", "crc", "crs", "cuc", "cus", "cic", "cis", "toes",
                                           "fit", "f0", "f1", "f2")] +
                [("tgd", C.c_uint64 * 4), ("ctype", C.c_int32)] +
                [(n, C.c_int32) for n in ("week_gpst", "cnt", "cntth", "update", "prn", "week_gst")] +
                [(n, C.c_uint32) for n in ("sub_cnt", "received_mask", "received_mask_proc")] +
                [("tow_gpst", C.c_uint64)])


def host_eph(lib, ch_ptr) -> FlatEph:
    e = FlatEph()
    lib.gpsb_host_channel_eph(C.c_void_p(ch_ptr), C.byref(e))
    return e


def ref_eph(reference, rch) -> FlatEph:
    e = FlatEph()
    reference.lib.ref_channel_eph(C.c_void_p(rch), C.byref(e))
    return e


def eph_diff(a: FlatEph, b: FlatEph):
    out = []
    for name, _ in a._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        va, vb = (list(va), list(vb)) if hasattr(va, "__len__") else (va, vb)
        if va != vb:
            out.append((name, va, vb))
    return out


# ---- IS-GPS-200 navigation words, written from the ICD (table 20-XIV) independently of the code under test
PARITY_TAPS = {
    25: [1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23],
    26: [2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24],
    27: [1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22],
    28: [2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23],
    29: [1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24],
    30: [3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24],
}


def encode_word(d24, prev29, prev30):
    """24 source bits -> 30 transmitted bits (data complemented by the previous D30, six parity bits)."""
    d = [0] + list(d24)                                        # 1-based
    par = []
    for k in range(25, 31):
        seed = prev29 if k in (25, 27, 30) else prev30
        v = seed
        for t in PARITY_TAPS[k]:
            v ^= d[t]
        par.append(v)
    return [b ^ prev30 for b in d24] + par


def make_subframe(rng, sf_id, tow):
    """300 transmitted bits of one subframe with random payload: TLM preamble, HOW with TOW count and subframe id,
    the two non-information bits of the HOW (and of word 10) solved so that D29 = D30 = 0 there, as the ICD requires -
    which is what lets every subframe start with an uncomplemented preamble."""
    words = []
    prev29 = prev30 = 0
    for w in range(10):
        d = list(rng.integers(0, 2, 24))
        if w == 0:
            d[:8] = [1, 0, 0, 0, 1, 0, 1, 1]
        if w == 1:
            d[:17] = [(tow >> (16 - i)) & 1 for i in range(17)]
            d[19:22] = [(sf_id >> 2) & 1, (sf_id >> 1) & 1, sf_id & 1]
        if w in (1, 9):                                        # choose d23, d24 so that the last two parity bits are 0
            for t in range(4):
                d[22], d[23] = t >> 1, t & 1
                tx = encode_word(d, prev29, prev30)
                if tx[28] == 0 and tx[29] == 0:
                    break
            else:
                raise AssertionError("no solution for the non-information bits")
        tx = encode_word(d, prev29, prev30)
        prev29, prev30 = tx[28], tx[29]
        words += tx
    return np.array(words, np.uint8)


def test_decode_random_subframe_images(reference):
    """gps_nav_data_decode_subframe on random 300-bit images, every subframe id incl. the invalid ones, several
    decodes accumulating in the same record (the toe of subframe 2 uses the week of an earlier subframe 1): every field
    of eph_t / sdreph_t, doubles by bit pattern, equals the reference's."""
    lib = load_host_library()
    rng = np.random.default_rng(2290)
    ch = Channels([7])
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, 7, 0)
    lib.gps_nav_data_decode_subframe.restype = C.c_uint8
    lib.gps_nav_data_decode_subframe.argtypes = [C.c_void_p]
    reference.lib.ref_decode_subframe.restype = C.c_uint32
    reference.lib.ref_decode_subframe.argtypes = [C.c_void_p, C.c_void_p]
    sub = (C.c_uint8 * 38)
    ids = [2, 3, 1, 2, 3, 4, 5, 0, 6, 7] + [int(x) for x in rng.integers(0, 8, 300)]
    for n, sf_id in enumerate(ids):
        img = rng.integers(0, 256, 38, dtype=np.uint8)
        if n % 7 == 3:
            img[:] = 0xFF                                      # all fields at their extremes, every sign bit set
        for k in range(3):                                     # subframe id: bits 49..51, MSB first
            bit = (sf_id >> (2 - k)) & 1
            img[(49 + k) >> 3] = (int(img[(49 + k) >> 3]) & ~(1 << ((49 + k) & 7))) | (bit << ((49 + k) & 7))
        st = ch.snapshot(0)
        st.subframe_data[:] = list(img)
        ch.restore(0, st)
        got_id = lib.gps_nav_data_decode_subframe(ch.at(0))
        want_id = reference.lib.ref_decode_subframe(rch, sub(*img))
        assert got_id == want_id == sf_id
        d = eph_diff(host_eph(lib, ch.at(0)), ref_eph(reference, rch))
        assert not d, (n, sf_id, d)
    ch.free()


@pytest.mark.parametrize("inverted", [False, True])
def test_word_assembler_and_decode_on_a_valid_bit_stream(reference, inverted):
    """A bit stream of correctly encoded subframes 1..5 (two frames' worth, random payload, noise bits in front, one
    word corrupted in the middle) handed bit by bit to gps_nav_data_words_detection: nav_data (word / subframe
    bookkeeping, time stamps, subframe image) and eph_data equal the reference's after every subframe; with the
    stream inverted neither side finds a word (the polarity is resolved upstream, nav_data.c:63)."""
    lib = load_host_library()
    lib.gpsb_host_feed_nav_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    reference.lib.ref_feed_nav_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(77 + inverted)
    chunks = [rng.integers(0, 2, 83, dtype=np.uint8)]
    tow = 0x1A2B
    for k in range(10):
        chunks.append(make_subframe(rng, k % 5 + 1, tow + k))
    stream = np.concatenate(chunks)
    stream[83 + 3 * 300 + 95] ^= 1                             # one bad bit: that word fails parity, the hunt restarts
    if inverted:
        stream ^= 1
    ch = Channels([11])
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, 11, 0)
    ms0 = 12345
    step = 150
    for at in range(0, stream.size, step):
        part = np.ascontiguousarray(stream[at:at + step])
        lib.gpsb_host_feed_nav_bits(ch.at(0), part.ctypes.data, part.size, ms0 + 20 * at)
        reference.lib.ref_feed_nav_bits(rch, part.ctypes.data, part.size, ms0 + 20 * at)
        assert bytes(ch.snapshot(0)) == bytes(reference.snapshot(rch)), at
        d = eph_diff(host_eph(lib, ch.at(0)), ref_eph(reference, rch))
        assert not d, (at, d)
    e = host_eph(lib, ch.at(0))
    if inverted:
        assert e.sub_cnt == 0
    else:
        assert e.sub_cnt >= 8 and e.received_mask == 0x1F and ch.snapshot(0).word_cnt_test >= 70
    ch.free()
