"""A receiver on the ground under four GPS satellites, as a raw IF recording: orbits -> quantised ephemerides -> subframes
1-3 as navigation bits -> Doppler, code phase and bit timing of each satellite consistent with the geometry ->
stm32f4_sdr_gps_b200.signal_synth.  Shared by the IF-samples-to-position tests."""
import numpy as np

from stm32f4_sdr_gps_b200.signal_synth import (CHIP_RATE_HZ, FS_HZ, IF_HZ, MS_SAMPLES, Satellite, ca_code)
from test_bits_to_position import quantise, subframe_bits
from test_fix import CLIGHT, geodetic_to_ecef, make_sky, pseudorange

L1_HZ = 1575.42e6
LEAD_BITS, TAIL_BITS = 100, 45      # 2 s ahead of subframe 1: time for the slot-phase walk and the bit synchroniser
FIRST_TOW_COUNT = 52000


class PositionScene:
    def __init__(self, seed=77, lat=52.52, lon=13.40, h=40.0, prns=(4, 11, 19, 26), cn0=50.0):
        rng = np.random.default_rng(seed)
        self.prns, self.lat, self.lon, self.h = list(prns), lat, lon, h
        self.site = geodetic_to_ecef(lat, lon, h)
        ids = [1, 2, 3]
        t_sf = (FIRST_TOW_COUNT - 1) * 6.0                      # GPS time at which subframe 1 starts
        self.t0 = t_sf - 0.02 * LEAD_BITS                       # ... and the first bit of the stream
        self.t_end = (FIRST_TOW_COUNT + len(ids) - 1) * 6.0     # end of subframe 3 = the time of week its HOW carries
        t_meas = self.t_end + 0.35
        c_ms = CLIGHT / 1000.0
        n_stream = LEAD_BITS + 300 * len(ids) + TAIL_BITS
        # The geometry is taken as drawn.  Where each satellite's data-bit edges fall relative to the 4-ms channel slots
        # (flip_ms % 4 below) is whatever the flight times give: the batched paths bring them to slot position 2 - the
        # only one the reference's bit synchroniser refines, nav_data.c:87-138 - with the slot-phase walk
        # (gpsb_rx_set_slot_walk), as the MCU's 17-ms channel schedule does on the hardware.
        self.sky, self.raws = [], []
        for el in make_sky(rng, self.site, self.t_end, 4):
            el["toes"] = el["toc"] = float(int(self.t_end) // 7200 * 7200)
            raw, back = quantise(el, 1)
            self.sky.append(back); self.raws.append(raw)
        flight_m = np.array([pseudorange(el, self.site, t_meas, 0.0) for el in self.sky]) / c_ms       # ms at t_meas
        rate = np.array([pseudorange(el, self.site, t_meas + 0.5, 0.0) - pseudorange(el, self.site, t_meas - 0.5, 0.0)
                         for el in self.sky])               # m/s
        # linear model anchored at the measurement: a signal sent at GPS time tau arrives flight(tau) later
        flight_at = lambda tau: flight_m + rate / CLIGHT * (tau - t_meas) * 1000.0                    # ms
        a0 = 102.3 - flight_at(self.t0).min()               # receiver ms of a zero-delay arrival of the stream start
        first = a0 + flight_at(self.t0)                     # receiver ms at which stream bit 1 starts arriving
        whole, part = np.floor(first).astype(int), first % 1.0
        self.flip_ms = np.where(part < 0.5, whole, whole + 1)   # the millisecond that shows the sign flip of an edge
        self.a0, self.rate, self.flight_m, self.t_meas = a0, rate, flight_m, t_meas
        self.doppler = -rate / (CLIGHT / L1_HZ)
        self.offset_ms = np.floor(first).astype(int)
        self.code_phase = (first - self.offset_ms) * 16368.0
        sats = []
        for i, prn in enumerate(self.prns):
            stream = np.concatenate([rng.integers(0, 2, 1 + LEAD_BITS, dtype=np.uint8)] +
                                    [subframe_bits(rng, sf, FIRST_TOW_COUNT + k, self.raws[i]) for k, sf in enumerate(ids)] +
                                    [rng.integers(0, 2, TAIL_BITS + 40, dtype=np.uint8)])
            sats.append(Satellite(prn=prn, doppler_hz=float(self.doppler[i]), code_phase_samples=float(self.code_phase[i]),
                                  cn0_dbhz=cn0, nav_bits=stream, nav_bit_offset_ms=int(self.offset_ms[i])))
        self.sats = sats
        # receiver ms at which the end of subframe 3 of the latest satellite has arrived, plus the first filter window
        last_edge = (a0 + (self.t_end - self.t0) * 1000.0 + flight_m).max()
        # (the reference keeps the samples of the slot in progress in function statics shared by all channels,
        # nav_data.c:48-51; the checker's driver ref_track_run_walk puts a channel's own samples back when a leg resumes
        # inside a slot, so legs may end anywhere)
        self.n_first = int(last_edge) + 121
        self.n_second = 300
        self.n_ms = self.n_first + self.n_second + 8
        self.seed = seed

    def with_carrier_phases(self, phases):
        for s, p in zip(self.sats, phases):
            s.carrier_phase_rad = float(p)

    def synthesize(self, n_ms=None, only=None):
        sats = self.sats if only is None else [self.sats[only]]
        return synthesize_long(sats, n_ms or self.n_ms, self.seed)


def synthesize_long(sats, n_ms, seed, block_ms=50, workers=None):
    """Recordings of tens of seconds: stm32f4_sdr_gps_b200.signal_synth.synthesize_blocks (independent 50-ms blocks, a
    thread per core)."""
    from stm32f4_sdr_gps_b200.signal_synth import synthesize_blocks
    return synthesize_blocks(sats, n_ms, seed, block_ms, workers)
