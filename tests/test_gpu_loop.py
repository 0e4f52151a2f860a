"""The device-resident tracking loop (k_track_run) on a real B200, through the C ABI, against the unmodified
reference closed loop and against the host-resident loop of the same library - bit-exact."""
import ctypes as C

import numpy as np
import pytest

from emu_lib import load_emulator
from stm32f4_sdr_gps_b200 import Channels, FlatState, Receiver, load_host_library

pytestmark = pytest.mark.gpu


def test_loop_math_certificate_whole_domain(host_engine):
    """Every (IP, QP) in [-8184, 8184]^2: the device's Costas error (fdlibm atan2f on one branch, CUDA double
    atan2 on the other, then / pi in double, rounded to float) and FLL angle (fdlibm atanf) carry the host
    libm's bit pattern.  2 x 268 M values, compared inside libgpsb_host.so on all host cores."""
    lib = load_host_library()
    assert lib.gpsb_host_certify_loop_math(host_engine.handle, 0) == 0


def test_loop_math_rows_against_emulator_tables(host_engine):
    """Independent path to the same statement for a few rows: device values vs libm values computed by the
    test-side emulator library (tests/emu/loop_emu.c)."""
    emu = load_emulator()
    for ip_lo in (-8184, -4097, -3, 0, 1, 2, 777, 8180):
        n = min(5, 8185 - ip_lo)
        want = np.empty((n, 16369), np.float32)
        for kind, fn in ((0, emu.emu_host_costas), (1, emu.emu_host_fll_angle)):
            fn(ip_lo, ip_lo + n, want.ctypes.data)
            got = host_engine.l0_loop_math(kind, ip_lo, n)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (kind, ip_lo)


def _locked(ch, i, found_freq, freq_hz, fine, master=0, bad=0):
    st = ch.snapshot(i)
    st.acq_state, st.trk_state, st.found_freq_offset_hz = 9, 4, found_freq
    st.if_freq_offset_hz_bits = int(np.float32(freq_hz).view(np.uint32))
    st.code_phase_fine_bits = int(np.float32(fine).view(np.uint32))
    st.pll_bad_state_cnt, st.pll_bad_state_master_cnt = bad, master
    ch.restore(i, st)
    return st


def test_device_loop_many_channels_ring_wrap_equals_reference(host_engine, golden, reference):
    """Six channels (two satellites present, each tracked by three channels from different starting points) in
    ONE k_track_run launch, over a span of the signal ring that wraps around its end: per-millisecond sums, nav
    bits and the final records equal the reference run of each channel."""
    sig = golden["scene_signal"]
    n_ms = 560
    ms0 = host_engine.ring_ms * 5 - 200                 # frames ring-200 .. ring-1, 0 .. 359
    host_engine.upload_signal(ms0, sig[:n_ms])
    prns = [5, 14, 5, 14, 5, 14]
    ch = Channels(prns)
    rchans = reference.channels(len(prns))
    for i, prn in enumerate(prns):
        s = i % 2
        fine = float(golden["track_found_phase"][s]) * 8.0 + (i // 2) * 3.0
        freq = float(golden["track_found_freq"][s]) + (i // 2) * 40.0
        st = _locked(ch, i, int(golden["track_found_freq"][s]), freq, fine)
        rch = reference.channel_at(rchans, i)
        reference.channel_init(rch, prn, 0)
        reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
    rx = Receiver(host_engine, ch)
    launches0 = host_engine.launch_count
    iq, nav = rx.track_run(ms0, n_ms)
    assert host_engine.launch_count - launches0 == 1                  # the whole run is one launch
    assert rx.loop_stats() == (len(prns) * n_ms, 0)
    for i in range(len(prns)):
        rch = reference.channel_at(rchans, i)
        want_iq, want_nav, _ = reference.track_run(rch, sig[:n_ms], ms0, n_ms)
        assert np.array_equal(iq[:, i, :], want_iq), i
        assert np.array_equal(nav[:, i], want_nav), i
        assert bytes(ch.snapshot(i)) == bytes(reference.snapshot(rch)), i
    rx.close()
    ch.free()


def test_device_loop_false_lock_reseed_equals_reference(host_engine, reference):
    """Noise-only input: the false-lock kicker fires inside the kernel and draws from the channel's private
    generator exactly what the reference draws from a freshly seeded libc rand() (tracking.c:300-326)."""
    rng = np.random.default_rng(77)
    n_ms = 1000
    sig = rng.integers(0, 256, (n_ms, 2046), dtype=np.uint8)
    host_engine.upload_signal(0, sig)
    ch = Channels([9])
    st = _locked(ch, 0, 1500, 1500.0, 4000.0, master=80, bad=10)
    C.CDLL(None).srand(1)
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, 9, 0)
    reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
    want_iq, want_nav, _ = reference.track_run(rch, sig, 0, n_ms)
    rx = Receiver(host_engine, ch)
    iq, nav = rx.track_run(0, n_ms)
    assert rx.loop_stats()[0] == n_ms
    assert np.array_equal(iq[:, 0, :], want_iq)
    assert np.array_equal(nav[:, 0], want_nav)
    got, want = ch.snapshot(0), reference.snapshot(rch)
    assert bytes(got) == bytes(want)
    assert got.if_freq_offset_hz_bits != st.if_freq_offset_hz_bits
    rx.close()
    ch.free()


def test_device_loop_split_runs_equal_one_run(host_engine, golden):
    """State (NCO phase, filters, nav-bit buffers, rand stream, deferred SNR) survives the trip through host
    memory: 600 ms as 1 + 3 + 96 + 500 ms in four calls equals one call."""
    sig = golden["scene_signal"]
    host_engine.upload_signal(0, sig)
    runs = []
    for parts in ((600,), (1, 3, 96, 500)):
        ch = Channels([5, 14])
        for i in range(2):
            _locked(ch, i, int(golden["track_found_freq"][i]), float(golden["track_found_freq"][i]),
                    float(golden["track_found_phase"][i]) * 8.0)
        rx = Receiver(host_engine, ch)
        rx.set_loop_site(2)
        ms, iqs = 0, []
        for n in parts:
            iq, _ = rx.track_run(ms, n)
            iqs.append(iq)
            ms += n
        runs.append((np.concatenate(iqs), [bytes(ch.snapshot(i)) for i in range(2)]))
        rx.close()
        ch.free()
    assert np.array_equal(runs[0][0], runs[1][0])
    assert runs[0][1] == runs[1][1]


def test_device_loop_refuses_what_it_cannot_do(host_engine, golden):
    """Through the raw C ABI: wrong record sizes and a satellite slot without a code are argument / state
    errors; a channel that is not tracking comes back untouched with stop == 1."""
    lib = host_engine.lib
    host_engine.set_code_prn(5, 5)
    ch = Channels([5])
    ch_b, aux_b = host_engine.record_bytes()
    aux = C.create_string_buffer(aux_b)
    res = (C.c_uint8 * 24)()
    assert lib.gpsb_track_loop(host_engine.handle, 1, ch.at(0), ch_b - 8, aux, aux_b, 0, 4, None, None, res) == -1
    before = bytes(ch.snapshot(0))
    assert lib.gpsb_track_loop(host_engine.handle, 1, ch.at(0), ch_b, aux, aux_b, 0, 4, None, None, res) == 0
    done, stop = np.frombuffer(bytes(res), np.uint32, 2)
    assert (int(done), int(stop)) == (0, 1) and bytes(ch.snapshot(0)) == before
    ch.free()


def _two_locked_channels(golden):
    ch = Channels([5, 14])
    for i in range(2):
        _locked(ch, i, int(golden["track_found_freq"][i]), float(golden["track_found_freq"][i]),
                float(golden["track_found_phase"][i]) * 8.0)
    return ch


def test_streaming_run_equals_resident_run(host_engine, golden):
    """gpsb_rx_track_stream: the samples start in host memory and are DMA-ed into the ring in chunks WHILE the loop
    launch is already tracking (watermark in device memory).  Sums, nav bits and final records equal the run on a
    ring uploaded beforehand - for chunk sizes from one millisecond to the whole recording, and across a ring wrap."""
    sig = np.ascontiguousarray(golden["scene_signal"][:600])
    ms0 = host_engine.ring_ms * 3 - 77
    host_engine.upload_signal(ms0, sig)
    ch = _two_locked_channels(golden)
    rx = Receiver(host_engine, ch)
    want_iq, want_nav = rx.track_run(ms0, 600)
    want_rec = [bytes(ch.snapshot(i)) for i in range(2)]
    rx.close()
    ch.free()
    for chunk in (1, 7, 128, 0, 600, 5000):
        host_engine.upload_signal(ms0, np.zeros_like(sig))          # nothing of the recording is left in the ring
        ch = _two_locked_channels(golden)
        rx = Receiver(host_engine, ch)
        launches0 = host_engine.launch_count
        iq, nav = rx.track_stream(ms0, sig, chunk_ms=chunk)
        assert host_engine.launch_count - launches0 == 1, chunk      # still one launch for the whole run
        assert rx.loop_stats() == (2 * 600, 0), chunk
        assert np.array_equal(iq, want_iq) and np.array_equal(nav, want_nav), chunk
        assert [bytes(ch.snapshot(i)) for i in range(2)] == want_rec, chunk
        assert host_engine.stream_progress(2) >= ms0 + 512
        rx.close()
        ch.free()


def test_streaming_call_where_copies_cannot_overlap_the_kernel():
    """gpsb_rx_track_stream in a process whose launches are synchronous (CUDA_LAUNCH_BLOCKING=1; a profiler's kernel
    replay does the same): the frames pushed behind the launch can never arrive while the loop kernel runs.  The call
    notices (the loop ends starved after ~0.1 s instead of the 2-s producer time-out), finishes upload-then-run, and
    the sums equal those of the overlapped run (tools/stream_blocked_probe.py compares them itself)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", PROBE_MS="600")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "stream_blocked_probe.py")], env=env, cwd=root,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sums identical: True" in out.stdout
    calls = [float(l.split(":")[1].split("ms")[0]) for l in out.stdout.splitlines() if l.startswith("ring")]
    assert len(calls) == 2 and max(calls) < 600.0, out.stdout        # nowhere near the 2-s time-out


def test_streaming_run_starved_producer_ends_cleanly(host_engine, golden):
    """Raw C ABI: a loop started in streaming mode whose producer stops after 100 ms neither hangs nor touches
    frames that never arrived: it ends by its time-out with stop == 3 and the milliseconds it reports as done are
    bit-identical to the same milliseconds of a resident run; the rest can be run afterwards."""
    lib = host_engine.lib
    sig = np.ascontiguousarray(golden["scene_signal"][:300])
    host_engine.upload_signal(0, sig)
    ch = _two_locked_channels(golden)
    rx = Receiver(host_engine, ch)
    want_iq, _ = rx.track_run(0, 300)
    want_rec = [bytes(ch.snapshot(i)) for i in range(2)]
    rx.close()
    ch.free()

    host_engine.upload_signal(0, np.zeros_like(sig))
    ch = _two_locked_channels(golden)
    ch_b, aux_b = host_engine.record_bytes()
    aux = np.zeros(2 * aux_b, np.uint8)
    rx = Receiver(host_engine, ch)                       # loads the codes
    res = np.zeros((2, 6), np.uint32)
    iq = np.zeros((300, 2, 6), np.int16)
    host_engine.stream_set_timeout_ms(50)
    host_engine.stream_reset(0)
    host_engine.stream_push(0, sig[:100])
    assert lib.gpsb_track_loop_begin(host_engine.handle, 2, ch.at(0), ch_b, aux.ctypes.data, aux_b, 0, 300,
                                     iq.ctypes.data, None, res.ctypes.data, 1) == 0
    assert lib.gpsb_track_loop_end(host_engine.handle) == 0
    host_engine.stream_wait()
    host_engine.stream_set_timeout_ms(2000)
    for i in range(2):
        done, stop = int(res[i, 0]), int(res[i, 1])
        assert stop == 3 and 90 <= done <= 100, (done, stop)
        assert np.array_equal(iq[:done, i, :], want_iq[:done, i, :])
    # the rest of the recording, resident: the records end where the uninterrupted run ended
    host_engine.upload_signal(0, sig)
    res2 = np.zeros((1, 6), np.uint32)
    for i in range(2):
        done = int(res[i, 0])
        a = aux.reshape(2, aux_b)[i]
        assert lib.gpsb_track_loop(host_engine.handle, 1, ch.at(i), ch_b, a.ctypes.data, aux_b, done, 300 - done, None,
                                   None, res2.ctypes.data) == 0
        assert (int(res2[0, 0]), int(res2[0, 1])) == (300 - done, 0)
    got = [bytes(ch.snapshot(i)) for i in range(2)]
    # snr_value is finished on the host by the receiver layer (log10f); compare everything else via the flat state
    for i in range(2):
        a, b = FlatState.from_buffer_copy(got[i]), FlatState.from_buffer_copy(want_rec[i])
        a.snr_value_bits = b.snr_value_bits = 0
        assert bytes(a) == bytes(b), i
    rx.close()
    ch.free()


def test_stream_abort_ends_a_waiting_loop_at_once(host_engine, golden):
    """gpsb_stream_abort (raw C ABI): a streaming loop waiting for frames that will never come ends when the producer says
    so - not after the stream time-out (5 s here) - with stop == 3 and its milliseconds complete and exact."""
    import time
    lib = host_engine.lib
    sig = np.ascontiguousarray(golden["scene_signal"][:300])
    host_engine.upload_signal(0, sig)
    ch = _two_locked_channels(golden)
    rx = Receiver(host_engine, ch)
    want_iq, _ = rx.track_run(0, 300)
    rx.close()
    ch.free()

    host_engine.upload_signal(0, np.zeros_like(sig))
    ch = _two_locked_channels(golden)
    ch_b, aux_b = host_engine.record_bytes()
    aux = np.zeros(2 * aux_b, np.uint8)
    rx = Receiver(host_engine, ch)                       # loads the codes
    res = np.zeros((2, 6), np.uint32)
    iq = np.zeros((300, 2, 6), np.int16)
    host_engine.stream_set_timeout_ms(5000)
    host_engine.stream_reset(0)
    host_engine.stream_push(0, sig[:100])
    assert lib.gpsb_track_loop_begin(host_engine.handle, 2, ch.at(0), ch_b, aux.ctypes.data, aux_b, 0, 300,
                                     iq.ctypes.data, None, res.ctypes.data, 1) == 0
    time.sleep(0.05)                                     # the loop has long run out of frames and is waiting
    assert lib.gpsb_stream_loop_running(host_engine.handle) == 1
    t0 = time.perf_counter()
    host_engine.stream_abort()
    assert lib.gpsb_track_loop_end(host_engine.handle) == 0
    waited = time.perf_counter() - t0
    host_engine.stream_wait()
    host_engine.stream_set_timeout_ms(2000)
    assert waited < 1.0, waited
    for i in range(2):
        done, stop = int(res[i, 0]), int(res[i, 1])
        assert stop == 3 and 90 <= done <= 100, (done, stop)
        assert np.array_equal(iq[:done, i, :], want_iq[:done, i, :])
    host_engine.stream_reset(0)                          # clears the abort: the next streamed run is unaffected
    host_engine.stream_push(0, sig)
    ch2 = _two_locked_channels(golden)
    aux2 = np.zeros(2 * aux_b, np.uint8)
    assert lib.gpsb_track_loop_begin(host_engine.handle, 2, ch2.at(0), ch_b, aux2.ctypes.data, aux_b, 0, 300,
                                     iq.ctypes.data, None, res.ctypes.data, 1) == 0
    assert lib.gpsb_track_loop_end(host_engine.handle) == 0
    host_engine.stream_wait()
    assert [int(res[i, 0]) for i in range(2)] == [300, 300] and [int(res[i, 1]) for i in range(2)] == [0, 0]
    assert np.array_equal(iq, want_iq)
    rx.close()
    ch.free()
    ch2.free()


@pytest.mark.parametrize("ring_ms,chunk", [(256, 32), (192, 0), (64, 16)])
def test_streaming_run_longer_than_the_ring(host_engine, golden, ring_ms, chunk):
    """A 600-ms run through a ring of 256 / 192 ms: the producer refills the ring behind the loop (flow control on
    the progress words), still one launch; a 64-ms ring cannot be refilled behind a loop that reports progress every
    64 ms and is served ring-full by ring-full.  Same sums, nav bits and records as the resident run."""
    from stm32f4_sdr_gps_b200 import Engine
    sig = np.ascontiguousarray(golden["scene_signal"][:600])
    host_engine.upload_signal(0, sig)
    ch = _two_locked_channels(golden)
    rx = Receiver(host_engine, ch)
    want_iq, want_nav = rx.track_run(0, 600)
    want_rec = [bytes(ch.snapshot(i)) for i in range(2)]
    rx.close()
    ch.free()
    with Engine(device=0, max_sv=211, ring_ms=ring_ms) as eng:
        ch = _two_locked_channels(golden)
        rx = Receiver(eng, ch)
        launches0 = eng.launch_count
        iq, nav = rx.track_stream(0, sig, chunk_ms=chunk)
        if ring_ms >= 192:
            assert eng.launch_count - launches0 == 1
        assert np.array_equal(iq, want_iq) and np.array_equal(nav, want_nav)
        assert [bytes(ch.snapshot(i)) for i in range(2)] == want_rec
        assert rx.loop_stats() == (2 * 600, 0)
        rx.close()
        ch.free()


def test_device_loop_decodes_subframes_like_the_reference(host_engine, reference):
    """Row N2 on the device: a 13-s recording of one satellite whose data bits are correctly encoded subframes 1 and 2
    (tests/test_nav_decode.py builds them from the ICD), streamed through the 1024-ms ring into ONE k_track_run
    launch.  Bit synchronisation, preamble hunt, parity, subframe assembly AND the ephemeris / clock field decode
    (nav_data_decode.c, doubles) all happen in the kernel's nav thread: sums, nav bits, the whole channel record and
    every field of eph_t equal the unmodified reference's run."""
    from test_nav_decode import eph_diff, host_eph, make_subframe, ref_eph
    from stm32f4_sdr_gps_b200.signal_synth import Satellite, Scene, synthesize
    rng = np.random.default_rng(31)
    bits = np.concatenate([rng.integers(0, 2, 50, dtype=np.uint8), make_subframe(rng, 1, 0x0F00F), make_subframe(rng, 2, 0x0F010),
                           rng.integers(0, 2, 20, dtype=np.uint8)])
    n_ms = 13200
    sat = Satellite(prn=5, doppler_hz=1234.5, code_phase_samples=7001.3, cn0_dbhz=50.0, nav_bits=bits, nav_bit_offset_ms=7,
                    carrier_phase_rad=3.14159)          # the Costas loop locks upright: no wait for two inverted preambles
    sig = synthesize(Scene(sats=[sat], n_ms=n_ms, seed=99))
    ch = Channels([5])
    st = _locked(ch, 0, 1000, 1234.5, 7001.3)
    rchans = reference.channels(1)
    rch = reference.channel_at(rchans, 0)
    reference.channel_init(rch, 5, 0)
    reference.restore(rch, type(reference.snapshot(rch)).from_buffer_copy(bytes(st)))
    want_iq, want_nav, _ = reference.track_run(rch, sig, 0, n_ms)
    rx = Receiver(host_engine, ch)
    launches0 = host_engine.launch_count
    iq, nav = rx.track_stream(0, sig, chunk_ms=100)
    assert host_engine.launch_count - launches0 == 1
    assert rx.loop_stats() == (n_ms, 0)
    assert np.array_equal(iq[:, 0, :], want_iq)
    assert np.array_equal(nav[:, 0], want_nav)
    assert bytes(ch.snapshot(0)) == bytes(reference.snapshot(rch))
    lib = load_host_library()
    got, want = host_eph(lib, ch.at(0)), ref_eph(reference, rch)
    assert not eph_diff(got, want), eph_diff(got, want)
    assert got.sub_cnt == 2 and got.received_mask == 3 and got.sat == 5      # both subframes went through the decode
    rx.close()
    ch.free()


@pytest.mark.parametrize("ring_ms,chunk", [(1024, 0), (256, 48)])
def test_streaming_run_from_the_2bit_iq_container(golden, ring_ms, chunk):
    """gpsb_rx_track_stream_iq2: the recording arrives as MAX2769-native 2-bit I / 2-bit Q samples (one byte each, only
    the I sign carries signal); every chunk is copied and packed to the ring format on the copy stream behind the
    running loop.  Same sums, nav bits and records as the packed recording - also through a ring shorter than the run."""
    from stm32f4_sdr_gps_b200 import Engine
    from stm32f4_sdr_gps_b200.signal_synth import iq2_from_packed
    sig = np.ascontiguousarray(golden["scene_signal"][:600])
    samples = iq2_from_packed(sig)
    with Engine(device=0, max_sv=211, ring_ms=1024) as eng:
        eng.upload_signal(0, sig)
        ch = _two_locked_channels(golden)
        rx = Receiver(eng, ch)
        want_iq, want_nav = rx.track_run(0, 600)
        want_rec = [bytes(ch.snapshot(i)) for i in range(2)]
        rx.close()
        ch.free()
    with Engine(device=0, max_sv=211, ring_ms=ring_ms) as eng:
        ch = _two_locked_channels(golden)
        rx = Receiver(eng, ch)
        iq, nav = rx.track_stream_iq2(0, samples, chunk_ms=chunk)
        assert np.array_equal(iq, want_iq) and np.array_equal(nav, want_nav)
        assert [bytes(ch.snapshot(i)) for i in range(2)] == want_rec
        assert rx.loop_stats() == (2 * 600, 0)
        rx.close()
        ch.free()


def test_sweep_scratch_survives_an_iq2_stream(golden):
    """Round-1 advisor finding: gpsb_stream_push_iq2 used to free the acquisition scratch when it grew its own staging
    buffer, so a sweep AFTER an iq2-streamed run worked on freed memory.  Cold sweep -> iq2-streamed tracking -> the same
    sweep again: identical cells (this sequence is also what tools/profile_r2.sh puts under compute-sanitizer)."""
    from stm32f4_sdr_gps_b200 import Engine, nco_step32
    from stm32f4_sdr_gps_b200.signal_synth import iq2_from_packed
    sig = np.ascontiguousarray(golden["scene_signal"][:300])
    samples = iq2_from_packed(sig)
    step = np.array([nco_step32(np.float32(4092000 - 5000 + 500 * b)) for b in range(21)], np.uint32)
    with Engine(device=0, max_sv=211, ring_ms=256) as eng:
        eng.upload_signal(0, sig[:16])
        for prn in (5, 14, 20):
            eng.set_code_prn(prn, prn)
        first = eng.sweep([5, 14, 20], step, 0, 10)
        ch = _two_locked_channels(golden)
        rx = Receiver(eng, ch)
        rx.track_stream_iq2(0, samples, chunk_ms=32)
        rx.close()
        ch.free()
        eng.upload_signal(0, sig[:16])
        again = eng.sweep([5, 14, 20], step, 0, 10)
        assert np.array_equal(first, again)


def test_loop_begin_end_call_sequence_errors(host_engine, golden):
    """gpsb_track_loop_begin / _end: a second begin before the end, and an end without a begin, are state errors
    (negative status, message available), never a dead lock; the open loop still ends normally afterwards."""
    lib = host_engine.lib
    sig = np.ascontiguousarray(golden["scene_signal"][:64])
    host_engine.upload_signal(0, sig)
    ch = _two_locked_channels(golden)
    rx = Receiver(host_engine, ch)                       # loads the codes
    ch_b, aux_b = host_engine.record_bytes()
    aux = np.zeros(2 * aux_b, np.uint8)
    res = np.zeros((2, 6), np.uint32)
    assert lib.gpsb_track_loop_end(host_engine.handle) == -4                      # GPSB_ERR_STATE
    args = (host_engine.handle, 2, ch.at(0), ch_b, aux.ctypes.data, aux_b, 0, 64, None, None, res.ctypes.data, 0)
    assert lib.gpsb_track_loop_begin(*args) == 0
    assert lib.gpsb_track_loop_begin(*args) == -4
    assert b"has not been ended" in lib.gpsb_last_error()
    assert lib.gpsb_track_loop_end(host_engine.handle) == 0
    assert [int(res[i, 0]) for i in range(2)] == [64, 64] and [int(res[i, 1]) for i in range(2)] == [0, 0]
    rx.close()
    ch.free()


def test_code_rounds_arguments_and_idle_channels(host_engine, golden):
    """gpsb_code_rounds (raw C ABI): bad arguments are refused with a message; mode 0 leaves a channel's records
    untouched and reports 0 snapshots; mode 1 on a channel whose state is not in busy_mask consumes nothing."""
    lib = host_engine.lib
    sig = np.ascontiguousarray(golden["scene_signal"][:64])
    host_engine.upload_signal(0, sig)
    ch = _two_locked_channels(golden)                    # GPS_ACQ_DONE: not in any code round
    rx = Receiver(host_engine, ch)                       # loads the codes
    ch_b, aux_b = host_engine.record_bytes()
    aux = np.zeros(2 * aux_b, np.uint8)
    before = [bytes(ch.snapshot(i)) for i in range(2)]
    busy12 = (1 << 3) | (1 << 4) | (1 << 5)              # CODE_PHASE_SEARCH1, _1_DONE, _2 (gps_acq_state_t)
    used = host_engine.code_rounds(2, ch.at(0), aux, 0, 16, busy12, [0, 1])
    assert used.tolist() == [0, 0]
    assert [bytes(ch.snapshot(i)) for i in range(2)] == before and not aux.any()
    mode = np.array([3, 0], np.uint8)
    used = np.zeros(2, np.uint32)
    assert lib.gpsb_code_rounds(host_engine.handle, 2, ch.at(0), ch_b, aux.ctypes.data, aux_b, 0, 16, busy12,
                                mode.ctypes.data, used.ctypes.data) == -1          # GPSB_ERR_ARG
    assert b"mode 3" in lib.gpsb_last_error()
    mode[0] = 1
    assert lib.gpsb_code_rounds(host_engine.handle, 2, ch.at(0), ch_b, aux.ctypes.data, aux_b, 0, 1 << 20, busy12,
                                mode.ctypes.data, used.ctypes.data) == -1
    assert b"exceed the ring" in lib.gpsb_last_error()
    assert lib.gpsb_code_rounds(host_engine.handle, 2, ch.at(0), ch_b + 8, aux.ctypes.data, aux_b, 0, 16, busy12,
                                mode.ctypes.data, used.ctypes.data) == -1
    rx.close()
    ch.free()


def test_slot_walk_all_alignments_on_the_device(reference):
    """Four satellites whose data-bit edges sit at all four slot alignments, tracked by ONE k_track_run launch with the
    slot-phase walk enabled (gpsb_rx_set_slot_walk): every channel ends with a refined bit edge, and sums, nav bits,
    the idle schedule and the channel records equal the UNMODIFIED reference driven on the walked (millisecond, slot
    index) schedule by the checker's own restatement of the policy (oracle/ref_shim.c, ref_track_run_walk).  The same
    recording streamed through a ring shorter than the run, split into several launches, and with the loop filters on
    the host gives the same bytes."""
    from stm32f4_sdr_gps_b200 import Engine
    from test_slot_walk import N_MS, locked, reference_walk, walk_scene
    scene, sig = walk_scene()
    refs = [reference_walk(reference, sat, sig, N_MS) for sat in scene.sats]
    eng = Engine(device=0, max_sv=40, ring_ms=4096)
    eng.upload_signal(0, sig)

    def fresh():
        ch = Channels([s.prn for s in scene.sats])
        for i, sat in enumerate(scene.sats):
            ch.restore(i, locked(ch.snapshot(i), sat))
        rx = Receiver(eng, ch)
        rx.set_slot_walk(True)
        return ch, rx

    def check(ch, rx, iq, nav, what):
        for i, (want, iq_ref, nav_ref, idx_ref, walk) in enumerate(refs):
            assert np.array_equal(iq[:, i, :], iq_ref), (what, i)
            assert np.array_equal(nav[:, i], nav_ref), (what, i)
            assert np.array_equal(~iq[:, i, :].any(axis=1), idx_ref == 0xFF), (what, i)
            assert bytes(ch.snapshot(i)) == bytes(want), (what, i)
            s = rx.sync_status(i)
            assert s.bit_edge_refined == 1 and s.walks == walk.gaps_taken and s.slot_phase == walk.slot_phase, (what, i)
        assert sorted(int((~iq[:, i, :].any(axis=1)).sum()) for i in range(4)) == [0, 1, 2, 3], what

    ch, rx = fresh()
    launches0 = eng.launch_count
    iq, nav = rx.track_run(0, N_MS)
    assert eng.launch_count - launches0 == 1 and rx.loop_stats() == (4 * N_MS, 0)
    check(ch, rx, iq, nav, "one launch")
    rx.close(); ch.free()

    ch, rx = fresh()                                   # cut inside slots and around the idle gaps
    parts, at = [], 0
    for end in (1203, 1207, 1810, 1811, 2404, N_MS):
        parts.append(rx.track_run(at, end - at))
        at = end
    check(ch, rx, np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), "split launches")
    rx.close(); ch.free()

    ch, rx = fresh()                                   # loop filters on the host, one round trip per millisecond
    rx.set_loop_site(1)
    n_host = 2000
    iq_h, nav_h = rx.track_run(0, n_host)
    assert np.array_equal(iq_h, iq[:n_host]) and np.array_equal(nav_h, nav[:n_host])
    rx.close(); ch.free()
    eng.close()

    small = Engine(device=0, max_sv=40, ring_ms=256)   # streamed through a ring shorter than the run
    ch = Channels([s.prn for s in scene.sats])
    for i, sat in enumerate(scene.sats):
        ch.restore(i, locked(ch.snapshot(i), sat))
    rx = Receiver(small, ch)
    rx.set_slot_walk(True)
    iq_s, nav_s = rx.track_stream(0, sig, chunk_ms=64)
    check(ch, rx, iq_s, nav_s, "streamed")
    rx.close(); ch.free()
    small.close()
