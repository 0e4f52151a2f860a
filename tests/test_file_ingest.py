"""Row N1 of SURVEY.md section 8(f), the file side (host/ingest.c): a recording on disk into a tracking run."""
import numpy as np
import pytest

from stm32f4_sdr_gps_b200 import load_host_library


def test_file_length_and_argument_errors(tmp_path):
    """No GPU needed: what the file holds, and the refusals that come before any device work."""
    lib = load_host_library()
    rec = tmp_path / "rec_file.bin"
    rec.write_bytes(bytes(2046 * 7 + 100))
    assert lib.gpsb_file_ms(str(rec).encode(), 0, 0) == 7
    assert lib.gpsb_file_ms(str(rec).encode(), 101, 0) == 6             # an offset that leaves 6 whole milliseconds
    assert lib.gpsb_file_ms(str(rec).encode(), 10 ** 9, 0) == 0
    assert lib.gpsb_file_ms(str(rec).encode(), 0, 2) == 0               # as a 2-bit I/Q container: 16368 bytes per ms
    assert lib.gpsb_file_ms(str(tmp_path / "missing.bin").encode(), 0, 0) == -1
    assert lib.gpsb_file_ms(str(tmp_path).encode(), 0, 0) == -1         # a directory
    assert lib.gpsb_file_ms(None, 0, 0) == -1
    assert lib.gpsb_rx_track_file(None, str(rec).encode(), 0, 0, 4, 0, None, None) == -1
    assert lib.gpsb_host_last_status() == -1


@pytest.mark.gpu
def test_tracking_from_a_file_equals_tracking_from_memory(golden, tmp_path):
    """The three containers - the MCU's memory image at a byte offset that is not page aligned, the same stream MSB
    first, the 2-bit I/Q bytes - give the sums, nav bits and channel records of the run fed from memory; a file that
    is shorter than the run, or missing, is refused before anything is launched."""
    from stm32f4_sdr_gps_b200 import Engine, GpsbError, Receiver
    from stm32f4_sdr_gps_b200.signal_synth import iq2_from_packed
    from test_gpu_loop import _two_locked_channels
    sig = np.ascontiguousarray(golden["scene_signal"][:600])
    lead = 4099
    (tmp_path / "lsb.bin").write_bytes(bytes(lead) + sig.tobytes())
    rev = np.array([int(f"{b:08b}"[::-1], 2) for b in range(256)], np.uint8)
    (tmp_path / "msb.bin").write_bytes(rev[sig].tobytes())
    (tmp_path / "iq2.bin").write_bytes(iq2_from_packed(sig).tobytes())

    def run(how):
        with Engine(device=0, max_sv=211, ring_ms=256) as eng:
            ch = _two_locked_channels(golden)
            rx = Receiver(eng, ch)
            try:
                out = how(rx, eng)
                return out[0], out[1], [bytes(ch.snapshot(i)) for i in range(2)], eng.launch_count
            finally:
                rx.close()
                ch.free()

    want = run(lambda rx, eng: rx.track_stream(0, sig))
    for name, kw in (("lsb.bin", dict(first_byte=lead)), ("msb.bin", dict(msb_first=True)), ("iq2.bin", dict(iq2=True))):
        got = run(lambda rx, eng: rx.track_file(tmp_path / name, 0, 600, **kw))
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2], name

    def refused(rx, eng):
        before = eng.launch_count
        for path, n_ms in ((tmp_path / "lsb.bin", 603), (tmp_path / "nothing.bin", 10)):
            with pytest.raises(GpsbError):
                rx.track_file(path, 0, n_ms)
        assert eng.launch_count == before
        return None, None
    run(refused)
