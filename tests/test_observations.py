"""Row N3 of SURVEY.md section 8(f): gps_master_nav_handling - subframe-time bookkeeping, code-phase filter,
pseudorange and time-of-week assembly (gps_master.c:159-388) - libgpsb_host.so against the UNMODIFIED reference."""
import ctypes as C

import numpy as np

from stm32f4_sdr_gps_b200 import Channels, load_host_library


def _obs(fn, ch_ptr):
    out = (C.c_uint64 * 2)()
    fn(C.c_void_p(ch_ptr), out)
    return int(out[0]), int(out[1])


def test_nav_handling_timeline_equals_reference(reference):
    """Four channels over a simulated 40-s timeline of idle-slot calls every 17 ms: subframe stamps arriving within
    and outside the 100-ms epoch window, the one-time zero moment, code-phase samples accumulating in the filter,
    filter windows that are too long, code-epoch wraps (negative filter sum, jump of more than half the range, the swap
    flag cleared by the next subframe), negative Doppler, a reference time difference that goes negative.  After every
    call the channel records and the observation pair (doubles by bit pattern) equal the reference's."""
    lib = load_host_library()
    lib.gpsb_host_set_sat_cnt(4)
    lib.gps_master_nav_handling.argtypes = [C.c_void_p]
    lib.gpsb_host_channel_obs.argtypes = [C.c_void_p, C.c_void_p]
    lib.gpsb_host_channel_set_tow.argtypes = [C.c_void_p, C.c_double]
    rl = reference.lib
    rl.ref_nav_handling.argtypes = [C.c_void_p, C.c_uint32]
    rl.ref_channel_obs.argtypes = [C.c_void_p, C.c_void_p]
    rl.ref_channel_set_tow.argtypes = [C.c_void_p, C.c_double]
    rl.ref_channel_at.restype = C.c_void_p
    for seed in range(6):
        rng = np.random.default_rng(500 + seed)
        prns = [5, 14, 20, 30]
        ch = Channels(prns)
        rchans = reference.channels(4)
        for i, p in enumerate(prns):
            reference.channel_init(reference.channel_at(rchans, i), p, 0)
        arrival = rng.integers(0, 60, 4)                       # ms offsets of the subframe ends inside an epoch
        fine = rng.uniform(0, 16368, 4).astype(np.float32)
        drift = rng.uniform(-60, 60, 4).astype(np.float32)     # samples per second: forces wraps at both ends
        doppler = rng.uniform(-4000, 4000, 4).astype(np.float32)
        first_epoch = int(rng.integers(2000, 9000))
        n_obs = 0
        for now in range(1000, 41000, 17):
            # what tracking / the word assembler would have done to the records since the previous idle slot
            for i in range(4):
                st = ch.snapshot(i)
                fine[i] = np.float32(fine[i] + drift[i] * np.float32(0.017))
                wrapped = False
                if fine[i] < 0:
                    fine[i] = np.float32(16368.0) + fine[i]
                    wrapped = True
                elif fine[i] > 16368:
                    fine[i] = fine[i] - np.float32(16368.0)
                    wrapped = True
                st.code_phase_fine_bits = int(np.float32(fine[i]).view(np.uint32))
                st.if_freq_offset_hz_bits = int(np.float32(doppler[i]).view(np.uint32))
                filt = np.uint32(st.code_phase_fine_filt_bits).view(np.float32)
                if wrapped:
                    filt = np.float32(-1.0)
                elif filt >= 0:
                    for _ in range(int(rng.integers(12, 17))):            # the DLL's per-ms accumulation, tracking.c:372-384
                        filt = np.float32(filt + fine[i])
                        st.code_filt_cnt += 1
                st.code_phase_fine_filt_bits = int(np.float32(filt).view(np.uint32))
                k = (now - first_epoch - int(arrival[i])) // 6000
                if now >= first_epoch + arrival[i] and k >= 0:
                    stamp = first_epoch + int(arrival[i]) + 6000 * k
                    if seed == 3 and i == 2:
                        stamp += 250 * (k % 2)                            # one channel falls outside the 100-ms window now and then
                    if st.last_subframe_time != stamp:
                        st.last_subframe_time = stamp
                        st.subframe_cnt += 1
                        st.new_subframe_flag = 1
                ch.restore(i, st)
                rch = reference.channel_at(rchans, i)
                reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
                tow = 6.0 * (1000 + (now - first_epoch) // 6000)
                lib.gpsb_host_channel_set_tow(ch.at(i), tow)
                rl.ref_channel_set_tow(rch, tow)
            lib.gpsb_host_set_packet_cnt(now)
            lib.gps_master_nav_handling(ch.at(0))
            rl.ref_nav_handling(rchans, now)
            for i in range(4):
                rch = reference.channel_at(rchans, i)
                assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(rch)), (seed, now, i)
                got, want = _obs(lib.gpsb_host_channel_obs, ch.at(i)), _obs(rl.ref_channel_obs, rch)
                assert got == want, (seed, now, i, got, want)
                n_obs += got[0] != 0
        assert n_obs > 1000, n_obs                              # observations were really produced along the way
        ch.free()
