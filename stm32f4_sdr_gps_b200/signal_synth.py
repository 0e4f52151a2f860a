"""Synthetic MAX2769-style front-end recordings (fixture generator for tests and bench.py).

The reference ships no recording; its only generator is the single-satellite,
single-millisecond simulator of Firmware/project_single_sat/GPS/simulator.c:88-146.  This module
produces what the hot path actually consumes - a stream of 1-bit samples (sign of I), 16.368 Msps,
IF 4.092 MHz, packed LSB-first, 2046 bytes per millisecond (Firmware/project_main/config.h:23-28,
signal_capture.c:9,169) - for several satellites with carrier *and* code Doppler, 50 bps data bits
and white Gaussian noise, from a fixed seed.

It is input plumbing only: nothing here takes part in the correlator arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FS_HZ = 16_368_000
IF_HZ = 4_092_000
CHIP_RATE_HZ = 1_023_000
L1_HZ = 1_575_420_000
MS_SAMPLES = 16368
MS_BYTES = 2046

# G2 output delays in chips for PRN 1..37 (IS-GPS-200 table 3-Ia expressed as delays).
_G2_DELAY = (5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471, 472,
             473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862, 863, 950, 947, 948, 950)


def ca_code(prn: int) -> np.ndarray:
    """C/A Gold code of PRN 1..37 as 0/1 bytes (IS-GPS-200: G1 = 1+x^3+x^10, G2 = 1+x^2+x^3+x^6+x^8+x^9+x^10)."""
    if not 1 <= prn <= len(_G2_DELAY):
        raise ValueError("prn %d outside 1..%d" % (prn, len(_G2_DELAY)))
    g1 = np.zeros(1023, np.uint8)
    g2 = np.zeros(1023, np.uint8)
    r1 = [1] * 10
    r2 = [1] * 10
    for i in range(1023):
        g1[i] = r1[9]
        g2[i] = r2[9]
        f1 = r1[2] ^ r1[9]
        f2 = r2[1] ^ r2[2] ^ r2[5] ^ r2[7] ^ r2[8] ^ r2[9]
        r1 = [f1] + r1[:9]
        r2 = [f2] + r2[:9]
    return g1 ^ np.roll(g2, _G2_DELAY[prn - 1])


@dataclass
class Satellite:
    prn: int
    doppler_hz: float
    code_phase_samples: float          # position of the code epoch inside ms 0, in samples (0..16368)
    cn0_dbhz: float = 45.0
    carrier_phase_rad: float = 0.0
    nav_bits: np.ndarray | None = None  # 0/1 per 20 ms; random from the scene seed when None
    nav_bit_offset_ms: int = 0          # ms index (mod 20) of a data-bit edge


@dataclass
class Scene:
    sats: list[Satellite]
    n_ms: int
    seed: int = 0x5D120001
    noise: bool = True
    truth: dict = field(default_factory=dict)


def synthesize(scene: Scene, chunk_ms: int = 50) -> np.ndarray:
    """Return the packed recording, shape (n_ms, 2046) uint8.  Sample n of a ms is bit n%8 of byte n//8."""
    rng = np.random.default_rng(scene.seed)
    n_ms = scene.n_ms
    n_bits = n_ms // 20 + 2
    sat_state = []
    for s in scene.sats:
        bits = s.nav_bits if s.nav_bits is not None else rng.integers(0, 2, n_bits, dtype=np.uint8)
        chips = ca_code(s.prn).astype(np.int8) * 2 - 1          # 0/1 -> -1/+1
        # C/N0 for a real signal in noise of unit variance over fs/2: A = sqrt(4*CN0/fs)
        amp = np.sqrt(4.0 * 10.0 ** (s.cn0_dbhz / 10.0) / FS_HZ) if scene.noise else 1.0
        sat_state.append((s, np.asarray(bits, np.uint8), chips, amp))
        scene.truth[s.prn] = {"doppler_hz": s.doppler_hz, "code_phase_samples": s.code_phase_samples,
                              "nav_bits": np.asarray(bits, np.uint8).copy(), "nav_bit_offset_ms": s.nav_bit_offset_ms}
    out = np.empty((n_ms, MS_BYTES), np.uint8)
    for m0 in range(0, n_ms, chunk_ms):
        m1 = min(n_ms, m0 + chunk_ms)
        n = np.arange(m0 * MS_SAMPLES, m1 * MS_SAMPLES, dtype=np.float64)
        t = n / FS_HZ
        acc = rng.standard_normal(n.size) if scene.noise else np.zeros(n.size)
        for s, bits, chips, amp in sat_state:
            code_rate = CHIP_RATE_HZ * (1.0 + s.doppler_hz / L1_HZ)
            # chip index: the code epoch sits code_phase_samples into ms 0
            chip_pos = (t - s.code_phase_samples / FS_HZ) * code_rate
            chip_idx = np.floor(chip_pos).astype(np.int64)
            epoch = np.floor_divide(chip_idx, 1023)               # code periods since the epoch
            bit_idx = np.floor_divide(epoch - s.nav_bit_offset_ms, 20) + 1
            d = bits[np.clip(bit_idx, 0, bits.size - 1)].astype(np.int8) * 2 - 1
            c = chips[np.mod(chip_idx, 1023)]
            phase = 2.0 * np.pi * (IF_HZ + s.doppler_hz) * t + s.carrier_phase_rad
            acc += amp * (d * c) * np.cos(phase)
        sign = (acc < 0.0).astype(np.uint8)
        out[m0:m1] = np.packbits(sign.reshape(m1 - m0, MS_SAMPLES), axis=1, bitorder="little")
    return out


def _block(args):
    sats, m0, m1, seed = args
    rng = np.random.default_rng([seed, m0])
    t = np.arange(m0 * MS_SAMPLES, m1 * MS_SAMPLES, dtype=np.float64) / FS_HZ
    acc = rng.standard_normal(t.size, dtype=np.float32)
    for s in sats:
        chips = np.concatenate([ca_code(s.prn), ca_code(s.prn)[:1]]).astype(np.float32) * 2 - 1
        data = np.asarray(s.nav_bits, np.float32) * 2 - 1
        amp = np.float32(np.sqrt(4.0 * 10.0 ** (s.cn0_dbhz / 10.0) / FS_HZ))
        code_rate = CHIP_RATE_HZ * (1.0 + s.doppler_hz / L1_HZ)
        chip_pos = (t - s.code_phase_samples / FS_HZ) * code_rate
        epoch = np.floor(chip_pos * (1.0 / 1023.0))
        chip = (chip_pos - epoch * 1023.0).astype(np.int32)                 # 0..1023 (1023 only by rounding: chip 0 again)
        bit = np.floor((epoch - s.nav_bit_offset_ms) * 0.05) + 1.0
        np.clip(bit, 0, data.size - 1, out=bit)
        carrier = np.cos((2.0 * np.pi * (IF_HZ + s.doppler_hz)) * t + s.carrier_phase_rad).astype(np.float32)
        carrier *= chips[chip]
        carrier *= data[bit.astype(np.int32)]
        carrier *= amp
        acc += carrier
    return m0, np.packbits((acc < 0).reshape(m1 - m0, MS_SAMPLES), axis=1, bitorder="little")


def synthesize_blocks(sats, n_ms, seed, block_ms=50, workers=None):
    """The signal model of synthesize() (1-bit samples of code x data x carrier in white
    noise, packed LSB first) for recordings of tens of seconds: independent 50-ms blocks (noise seeded per block) on a
    thread per core.  Returns (n_ms, 2046) uint8; the result does not depend on the number of workers."""
    import os
    from multiprocessing.pool import ThreadPool
    for s in sats:                                              # data bits of a satellite that brings none: seeded per PRN
        if s.nav_bits is None:
            s.nav_bits = np.random.default_rng([seed, 1000 + s.prn]).integers(0, 2, n_ms // 20 + 3, dtype=np.uint8)
    jobs = [(sats, m0, min(n_ms, m0 + block_ms), seed) for m0 in range(0, n_ms, block_ms)]
    out = np.empty((n_ms, MS_SAMPLES // 8), np.uint8)
    workers = workers or max(1, min(32, (os.cpu_count() or 2) - 1, len(jobs)))
    if workers == 1:
        results = map(_block, jobs)
    else:
        pool = ThreadPool(workers)                              # numpy releases the GIL inside the array operations
        results = pool.imap_unordered(_block, jobs)
    for m0, block in results:
        out[m0:m0 + block.shape[0]] = block
    if workers > 1:
        pool.close()
        pool.join()
    return out


def iq2_from_packed(packed: np.ndarray, seed: int = 7) -> np.ndarray:
    """Expand a packed recording to the MAX2769-native 2-bit I / 2-bit Q container, one byte per sample
    (bit0 I sign, bit1 I magnitude, bit2 Q sign, bit3 Q magnitude).  Only the I sign carries the signal
    the reference front end wires up (config.h:16); the other three bits are filled with noise."""
    rng = np.random.default_rng(seed)
    bits = np.unpackbits(np.ascontiguousarray(packed).reshape(-1, MS_BYTES), axis=1, bitorder="little")
    junk = rng.integers(0, 8, bits.shape, dtype=np.uint8) << 1
    return (bits | junk).astype(np.uint8)


def config2_scene(n_ms: int = 1000, prns=(5, 14, 20, 30), seed: int = 0x5D120001) -> Scene:
    """four satellites (PRNs of Firmware/project_main/main.c:59-71)."""
    rng = np.random.default_rng(seed ^ 0xA5A5)
    sats = [Satellite(prn=p, doppler_hz=float(rng.uniform(-5000, 5000)),
                      code_phase_samples=float(rng.uniform(0, MS_SAMPLES)),
                      carrier_phase_rad=float(rng.uniform(0, 2 * np.pi)),
                      nav_bit_offset_ms=int(rng.integers(0, 20))) for p in prns]
    return Scene(sats=sats, n_ms=n_ms, seed=seed)


def config3_scene(n_ms: int = 10, n_present: int = 10, seed: int = 0x5D120003) -> Scene:
    """32 PRNs searched, n_present of them in the sky."""
    rng = np.random.default_rng(seed ^ 0x3C3C)
    prns = sorted(rng.choice(np.arange(1, 33), size=n_present, replace=False).tolist())
    sats = [Satellite(prn=int(p), doppler_hz=float(rng.integers(-9, 10) * 500 + rng.uniform(-100, 100)),
                      code_phase_samples=float(rng.uniform(0, MS_SAMPLES)),
                      carrier_phase_rad=float(rng.uniform(0, 2 * np.pi)), cn0_dbhz=47.0,
                      nav_bit_offset_ms=int(rng.integers(0, 20))) for p in prns]
    return Scene(sats=sats, n_ms=n_ms, seed=seed)
