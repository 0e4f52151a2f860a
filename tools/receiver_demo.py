#!/usr/bin/env python
"""A whole receiver pass on one B200, the way a user of the library runs it (needs a GPU):

    cold start over a list of PRNs (one launch)  ->  the reference's code-search rounds and pre-track for the
    satellites found  ->  closed-loop tracking with the loop filters on the device, the recording streamed from
    host memory  ->  data bits, subframes, ephemeris fields, pseudoranges (gps_master_nav_handling)

Input: a raw recording in the reference's capture format (1 bit per sample, sign of I, 16.368 Msps, IF 4.092 MHz,
LSB first, 2046 bytes per millisecond; Firmware/project_main/signal_capture.c) given with --file, or - by default - a
synthetic scene with three satellites.

    python tools/receiver_demo.py [--file capture.bin] [--seconds 3] [--prns 1-32]
"""
import argparse
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver, load_host_library  # noqa: E402
from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize  # noqa: E402


def parse_prns(text):
    out = []
    for part in text.split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--file")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--prns", default="1-16")
    args = ap.parse_args()
    n_ms = int(args.seconds * 1000)
    if args.file:
        raw = np.fromfile(args.file, dtype=np.uint8)
        n_ms = min(n_ms, raw.size // 2046)
        sig = np.ascontiguousarray(raw[:n_ms * 2046].reshape(n_ms, 2046))
    else:
        rng = np.random.default_rng(2024)
        sats = [Satellite(prn=p, doppler_hz=d, code_phase_samples=c, cn0_dbhz=48.0, carrier_phase_rad=float(rng.uniform(0, 6.28)),
                          nav_bit_offset_ms=int(rng.integers(0, 20))) for p, d, c in ((3, -3210.0, 4000.5), (7, 1475.0, 12001.2),
                                                                                      (11, 4620.0, 777.7))]
        print("synthesising %d ms with PRNs 3, 7, 11 ..." % n_ms)
        sig = synthesize(Scene(sats=sats, n_ms=n_ms, seed=4040))
    searched = parse_prns(args.prns)
    lib = load_host_library()
    with Engine(device=0, max_sv=211, ring_ms=1024) as eng:
        assert lib.gpsb_host_attach(eng.handle) == 0
        eng.upload_signal(0, sig[:min(n_ms, 1024)])

        t0 = time.perf_counter()
        ch_all = Channels(searched)
        rx_all = Receiver(eng, ch_all)
        votes, _ = rx_all.cold_sweep(-5000, 500, 21, 0, 10)
        found = {}
        for i, p in enumerate(searched):
            v = votes[i].astype(int)
            pair = v[:-1] + v[1:]
            b = int(np.argmax(pair))
            if ch_all.snapshot(i).acq_state == 2 and pair[b] >= 6:
                found[p] = (-5000 + 500 * b) + 500.0 * v[b + 1] / pair[b]
        rx_all.close()
        ch_all.free()
        print("cold start over %d PRNs x 21 Doppler bins x 10 ms x 2046 phases: %.2f ms -> found %s"
              % (len(searched), (time.perf_counter() - t0) * 1e3, {p: round(f) for p, f in found.items()}))
        if not found:
            return

        prns = sorted(found)
        ch = Channels(prns, [int(found[p]) for p in prns])
        rx = Receiver(eng, ch)
        lib.gpsb_host_set_sat_cnt(len(prns))
        lib.gpsb_host_master_reset()
        lib.gps_master_handling.argtypes = [C.c_void_p, C.c_uint8]
        ms = 10
        t0 = time.perf_counter()
        while ms < min(n_ms, 1000):
            lib.gpsb_host_set_packet_cnt(ms)
            lib.gps_master_handling(ch.at(0), ms % 4)
            if not lib.gps_master_need_acq():
                break
            rx.acquire_ms(ms)
            ms += 1
        print("code-phase search rounds 1..3: %d snapshots, %.1f ms of wall time; code phases (half chips): %s"
              % (ms - 10, (time.perf_counter() - t0) * 1e3, [ch.snapshot(i).found_code_phase for i in range(len(prns))]))
        if lib.gps_master_need_acq():
            print("acquisition did not complete")
            return
        lib.gps_master_handling(ch.at(0), ms % 4)
        iq_a, _ = rx.track_run(ms, 200)                       # pre-track, then the device-resident loop
        t1 = ms + 200
        t0 = time.perf_counter()
        iq, nav = rx.track_stream(t1, np.ascontiguousarray(sig[t1:n_ms]), chunk_ms=64)
        dt = time.perf_counter() - t0
        dev_ms, host_ms = rx.loop_stats()
        print("tracking %d ms x %d satellites, streamed from host memory: %.2f ms of wall time (%.0f x real time), "
              "%d channel-ms on the device, %d on the host path" % (n_ms - t1, len(prns), dt * 1e3, (n_ms - t1) * 1e-3 / dt, dev_ms, host_ms))
        lib.gpsb_host_set_packet_cnt(n_ms - 1)
        lib.gps_master_handling(ch.at(0), 0xFF)               # idle slot: observations, if subframes have been seen
        obs = (C.c_uint64 * 2)()
        lib.gpsb_host_channel_obs.argtypes = [C.c_void_p, C.c_void_p]
        print(" PRN  Doppler Hz  code phase   |IP| mean   SNR dB  bit sync  data bits  words ok  subframes  pseudorange m")
        for i, p in enumerate(prns):
            st = ch.snapshot(i)
            lib.gpsb_host_channel_obs(ch.at(i), obs)
            pr = float(np.uint64(obs[0]).view(np.float64))
            print(" %3d  %10.1f  %10.1f  %9.0f  %7.1f  %8d  %9d  %8d  %9d  %13.1f" % (
                p, float(np.uint32(st.if_freq_offset_hz_bits).view(np.float32)),
                float(np.uint32(st.code_phase_fine_bits).view(np.float32)), np.abs(iq[-500:, i, 2]).mean(),
                float(np.uint32(st.snr_value_bits).view(np.float32)), st.period_sync_ok_flag, int((nav[:, i] >= 0).sum()),
                st.word_cnt_test, st.subframe_cnt, pr))
        rx.close()
        ch.free()
        lib.gpsb_host_attach(None)


if __name__ == "__main__":
    main()
