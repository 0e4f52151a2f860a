#!/usr/bin/env python
"""Raw IF samples -> latitude / longitude on the GPU, with timings (the scene of tests/test_if_to_position.py).

    python tools/if_to_position_demo.py          # needs a B200; oracle/_ref is used only to pick the carrier phases

Prints one JSON line: where the fix landed, how far from the true site, and how long the device-resident loop took for
the two tracking legs (four channels, 20.2 s + 0.3 s of signal, streamed from host memory; slot-phase walk on)."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]


def main():
    from oracle_lib import Reference
    from stm32f4_sdr_gps_b200 import Channels, Engine, Receiver, load_host_library
    import test_if_to_position as T
    reference = Reference()
    lib = load_host_library()
    T.protos(lib, reference.lib)
    t0 = time.time()
    sc, sig = T.scene_and_signal(reference, 83)          # bit edges at all four slot alignments
    t_synth = time.time() - t0
    with Engine(device=0, max_sv=211, ring_ms=1024) as eng:
        assert lib.gpsb_host_attach(eng.handle) == 0
        lib.gpsb_host_set_sat_cnt(4)
        lib.gpsb_host_fix_reset()
        def tracking_channels():                                  # as acquisition and pre-track would leave them
            ch = Channels(sc.prns)
            for i in range(4):
                st = ch.snapshot(i)
                st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, int(sc.doppler[i])
                st.if_freq_offset_hz_bits = int(np.float32(sc.doppler[i]).view(np.uint32))
                st.code_phase_fine_bits = int(np.float32(sc.code_phase[i]).view(np.uint32))
                ch.restore(i, st)
            return ch

        ch = tracking_channels()                                  # warm-up launch on channels that are thrown away
        rx = Receiver(eng, ch)
        rx.track_stream(0, np.ascontiguousarray(sig[:64]), log=False)
        rx.close()
        ch.free()
        ch = tracking_channels()
        rx = Receiver(eng, ch)
        rx.set_slot_walk(True)                                    # every satellite gets its bit edges refined
        legs = []
        for ms0, n in ((0, sc.n_first), (sc.n_first, sc.n_second)):
            part = np.ascontiguousarray(sig[ms0:ms0 + n])
            t0 = time.perf_counter()
            rx.track_stream(ms0, part, log=False)
            legs.append(time.perf_counter() - t0)
            lib.gpsb_host_set_packet_cnt(ms0 + n)
            lib.gps_master_nav_handling(ch.base)
        t0 = time.perf_counter()
        fix = ch.position_fix()
        t_fix = time.perf_counter() - t0
        rx_walks = [rx.sync_status(i).walks for i in range(4)]
        rx.close()
        lib.gpsb_host_attach(None)
    out = {"truth": {"lat_deg": sc.lat, "lon_deg": sc.lon, "height_m": sc.h},
           "fix": None if fix is None else {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in fix.items()},
           "error_m": None if fix is None else float(np.linalg.norm(fix["ecef_m"] - sc.site)),
           "signal_s": (sc.n_first + sc.n_second) / 1000.0, "channels": 4,
           "tracking_leg_ms": [round(1e3 * x, 2) for x in legs], "fix_us": round(1e6 * t_fix, 1),
           "times_real_time": round((sc.n_first + sc.n_second) / 1000.0 / sum(legs), 1),
           "slot_walks": [int(rx_walks[i]) for i in range(4)], "bit_edge_alignments": [int(x) for x in sc.flip_ms % 4],
           "synthesis_s": round(t_synth, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
