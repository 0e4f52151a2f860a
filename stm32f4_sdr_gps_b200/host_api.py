"""ctypes binding of the host-side C mirror (``include/gpsb_host.h``, ``libgpsb_host.so``).

Plumbing only.  The library holds the reference-named acquisition / tracking / gps_master entry
points, the batched receiver (``gpsb_rx_*``) and the split-phase plan / finish API; all correlations
behind them run on the GPU through ``libgpsb_cuda.so``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build
from .engine import Engine, GpsbError, load_library

WANT_NOTHING, WANT_SEARCH, WANT_EPL = 0, 1, 2
ACQ_NEED_FREQ_SEARCH, ACQ_FREQ_SEARCH_RUN, ACQ_FREQ_SEARCH_DONE = 0, 1, 2
ACQ_DONE = 9
TRK_IDLE, TRK_NEED_PRE_TRACK, TRK_PRE_TRACK_RUN, TRK_PRE_TRACK_DONE, TRK_RUN = range(5)


class SearchReq(C.Structure):
    _fields_ = [("sv_slot", C.c_uint32), ("ms_index", C.c_uint32), ("acc0", C.c_uint32), ("step32", C.c_uint32),
                ("off_bits", C.c_uint16), ("start", C.c_uint16), ("stop", C.c_uint16), ("flags", C.c_uint16)]


class EplReq(C.Structure):
    _fields_ = [("sv_slot", C.c_uint32), ("ms_index", C.c_uint32), ("acc0", C.c_uint32), ("step32", C.c_uint32),
                ("off_e", C.c_uint16), ("off_p", C.c_uint16), ("off_l", C.c_uint16), ("off_bits", C.c_uint16)]


class SearchRes(C.Structure):
    _fields_ = [("max", C.c_uint16), ("phase", C.c_uint16), ("avg", C.c_uint16), ("reserved", C.c_uint16)]


class Plan(C.Structure):
    _fields_ = [("want", C.c_int), ("search", SearchReq), ("epl", EplReq), ("stage", C.c_int)]


class FlatState(C.Structure):
    """include/gpsb_flat_state.h"""
    _fields_ = [
        ("prn", C.c_uint32),
        ("acq_state", C.c_uint32), ("freq_index", C.c_uint32),
        ("found_freq_offset_hz", C.c_int32), ("given_freq_offset_hz", C.c_int32),
        ("found_code_phase", C.c_uint32), ("acq_code_search_start", C.c_uint32),
        ("acq_code_search_stop", C.c_uint32), ("code_hist_step", C.c_uint32),
        ("acq_start_timestamp", C.c_uint32), ("hist_ratio_bits", C.c_uint32),
        ("code_phase_histogram", C.c_uint8 * 32),
        ("trk_state", C.c_uint32), ("trk_code_search_start", C.c_uint32),
        ("trk_code_search_stop", C.c_uint32), ("if_freq_offset_hz_bits", C.c_uint32),
        ("if_freq_accum", C.c_uint32), ("pre_track_count", C.c_uint32),
        ("prev_track_timestamp", C.c_uint32), ("code_phase_fine_bits", C.c_uint32),
        ("old_code_phase_fine_bits", C.c_uint32), ("code_phase_swap_flag", C.c_uint32),
        ("dll_code_err_bits", C.c_uint32), ("pll_code_err_bits", C.c_uint32),
        ("fll_old_i", C.c_int32), ("fll_old_q", C.c_int32), ("fll_err_bits", C.c_uint32),
        ("pll_bad_state_cnt", C.c_uint32), ("pll_bad_state_master_cnt", C.c_uint32),
        ("i_part_summ", C.c_uint32), ("q_part_summ", C.c_uint32), ("snr_summ_cnt", C.c_uint32),
        ("snr_value_bits", C.c_uint32), ("filt_start_time_ms", C.c_uint32),
        ("code_filt_cnt", C.c_uint32), ("code_phase_fine_filt_bits", C.c_uint32),
        ("pre_track_phases", C.c_uint16 * 30), ("pll_check_buf", C.c_int16 * 4),
        ("period_sync_ok_flag", C.c_uint32), ("right_period_cnt", C.c_uint32),
        ("old_swap_time", C.c_uint32), ("old_reminder", C.c_uint32),
        ("accurate_swap_time", C.c_uint32), ("accurate_swap_ok", C.c_uint32),
        ("last_bit_pos_cnt", C.c_uint32), ("last_bit_neg_cnt", C.c_uint32),
        ("inv_polarity_flag", C.c_uint32), ("polarity_found", C.c_uint32),
        ("inv_preabmle_cnt", C.c_uint32), ("word_cnt", C.c_uint32), ("word_bit_cnt", C.c_uint32),
        ("old_D29", C.c_uint32), ("old_D30", C.c_uint32),
        ("word_detection_timestamp", C.c_uint32), ("word_cnt_test", C.c_uint32),
        ("last_subframe_time", C.c_uint32), ("first_subframe_time", C.c_uint32),
        ("subframe_cnt", C.c_uint32), ("new_subframe_flag", C.c_uint32),
        ("word_buf", C.c_uint8 * 30), ("subframe_data", C.c_uint8 * 38),
    ]


class SyncStatus(C.Structure):
    """include/gpsb_host.h, gpsb_sync_status"""
    _fields_ = [("tracking", C.c_uint8), ("bit_period_found", C.c_uint8), ("bit_edge_refined", C.c_uint8),
                ("polarity_found", C.c_uint8), ("slot_phase", C.c_uint8), ("walk_enabled", C.c_uint8),
                ("walk_pending", C.c_uint8), ("reserved", C.c_uint8), ("walks", C.c_uint16), ("subframes", C.c_uint16),
                ("words_ok", C.c_uint32)]


class ColdStartOpts(C.Structure):
    """include/gpsb_host.h, gpsb_cold_start_opts (zero = the reference's defaults)"""
    _fields_ = [("first_bin_hz", C.c_int32), ("bin_step_hz", C.c_int32), ("n_bins", C.c_uint32), ("sweep_ms", C.c_uint32),
                ("round_timeout_ms", C.c_uint32), ("window_ms", C.c_uint32), ("sweeps", C.c_uint32), ("serve_rank", C.c_uint32),
                ("serve_world", C.c_uint32), ("window_max_ms", C.c_uint32), ("code_rounds", C.c_uint32)]


class ColdStartReport(C.Structure):
    """include/gpsb_host.h, gpsb_cold_start_report"""
    _fields_ = [(n, C.c_uint32) for n in ("ms_sweep0", "ms_code0", "ms_code12_last", "ms_code3_first", "ms_last", "ms_next",
                                          "n_sweeps", "n_searched", "n_doppler_found", "n_served", "n_acquired", "launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_hostlib = None


class FlatFix(C.Structure):
    """include/gpsb_flat_state.h, gpsb_flat_fix: the position solver's outputs, doubles / floats as bit patterns"""
    _fields_ = [("stat", C.c_int32), ("ns", C.c_int32), ("type", C.c_int32), ("busy", C.c_int32),
                ("time_time", C.c_int64), ("time_sec_bits", C.c_uint64), ("rr", C.c_uint64 * 6), ("qr", C.c_uint32 * 6),
                ("dtr0", C.c_uint64), ("final_pos", C.c_uint64 * 3), ("azel", C.c_uint64 * 64)]


def load_host_library(path: Path | None = None) -> C.CDLL:
    """Load libgpsb_host.so (in-tree).  Raises if it has not been built."""
    global _hostlib
    if _hostlib is not None and path is None:
        return _hostlib
    load_library()                      # libgpsb_cuda.so first (also resolved through the rpath)
    p = Path(path) if path else _build.HOST_LIB
    if not p.exists():
        raise ImportError("%s is missing: build it with `python -m stm32f4_sdr_gps_b200.build`" % p)
    lib = C.CDLL(str(p))
    vp, u32, i32, u8, u16 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint8, C.c_uint16
    protos = {
        "gpsb_host_attach": (i32, [vp]), "gpsb_host_last_status": (i32, []),
        "gpsb_host_set_sat_cnt": (None, [u32]), "gpsb_host_sat_cnt": (u32, []),
        "gpsb_host_set_packet_cnt": (None, [u32]), "signal_capture_get_packet_cnt": (u32, []),
        "gpsb_host_master_reset": (None, []), "gpsb_host_last_nav_bit": (i32, []),
        "gps_fill_summ_table": (None, []), "gps_channell_prepare": (None, [vp]),
        "acquisition_process": (None, [vp, vp]), "acquisition_process_channel": (None, [vp, vp]),
        "acquisition_get_hist": (C.POINTER(u32), []),
        "acquisition_start_channel": (None, [vp]), "acquisition_start_code_search_channel": (None, [vp]),
        "acquisition_start_code_search3_channel": (None, [vp]),
        "gps_tracking_process": (None, [vp, vp, u8]),
        "gps_nav_data_analyse_new_code": (None, [vp, u8, C.c_int16]),
        "gps_nav_data_words_detection": (None, [vp, u8]),
        "gps_master_handling": (None, [vp, u8]), "gps_master_need_acq": (u8, []),
        "gps_master_need_freq_search": (u8, [vp]), "gps_master_is_code_search3": (u8, [vp]),
        "gps_master_reset_to_aqc_start": (None, [vp]),
        "gps_correlation8": (C.c_int16, [vp, vp, vp, u16]),
        "gps_correlation_iq": (None, [vp, vp, vp, u16, C.POINTER(C.c_int16), C.POINTER(C.c_int16)]),
        "correlation_search": (u16, [vp, vp, vp, u16, u16, C.POINTER(u16), C.POINTER(u16)]),
        "gps_shift_to_zero_freq": (None, [vp, vp, vp, C.c_float]),
        "gps_shift_to_zero_freq_track": (None, [vp, vp, vp, vp]),
        "gps_generate_prn_data2": (None, [vp, vp, u16]), "gps_rewind_if_phase": (None, [vp, u8]),
        "gps_generate_prn": (None, [vp, i32]),
        "gpsb_rx_create": (i32, [C.POINTER(vp), vp, vp, u32]), "gpsb_rx_destroy": (None, [vp]),
        "gpsb_rx_track_ms": (i32, [vp, u32]), "gpsb_rx_track_run": (i32, [vp, u32, u32, vp, vp]),
        "gpsb_rx_track_stream": (i32, [vp, u32, u32, vp, u32, vp, vp]),
        "gpsb_rx_track_stream_iq2": (i32, [vp, u32, u32, vp, u32, vp, vp]),
        "gpsb_rx_track_file": (i32, [vp, C.c_char_p, C.c_uint64, u32, u32, u32, vp, vp]),
        "gpsb_file_ms": (C.c_int64, [C.c_char_p, C.c_uint64, u32]),
        "gpsb_rx_acquire_ms": (i32, [vp, u32]),
        "gpsb_rx_set_threads": (None, [vp, u32]),
        "gpsb_rx_set_loop_site": (None, [vp, i32]),
        "gpsb_rx_loop_stats": (None, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "gpsb_rx_cold_start": (i32, [vp, u32, C.POINTER(ColdStartOpts), C.POINTER(ColdStartReport)]),
        "gpsb_rx_set_slot_walk": (i32, [vp, i32, u32]),
        "gpsb_rx_channel_sync": (i32, [vp, u32, C.POINTER(SyncStatus)]),
        "gpsb_host_aux_walk": (None, [vp, u32, u32]),
        "gpsb_host_aux_walk_state": (None, [vp, C.POINTER(u32 * 6)]),
        "gpsb_host_certify_loop_math": (C.c_int64, [vp, u32]),
        "gpsb_rx_cold_sweep": (i32, [vp, C.c_int32, C.c_int32, u32, u32, u32, vp, vp]),
        "gpsb_host_plan_acq": (i32, [vp, u32, C.POINTER(Plan)]),
        "gpsb_host_finish_acq": (i32, [vp, C.POINTER(Plan), C.POINTER(SearchRes)]),
        "gpsb_host_plan_track": (i32, [vp, u32, u8, C.POINTER(Plan)]),
        "gpsb_host_finish_track": (i32, [vp, u8, C.POINTER(Plan), C.POINTER(SearchRes), vp]),
        "gpsb_host_snapshot": (None, [vp, C.POINTER(FlatState)]),
        "gpsb_host_restore": (None, [vp, C.POINTER(FlatState)]),
        "gpsb_host_sizeof_channel": (u32, []),
        "gpsb_host_channels_alloc": (vp, [u32]), "gpsb_host_channels_free": (None, [vp]),
        "gpsb_host_channel_at": (vp, [vp, u32]), "gpsb_host_channel_init": (None, [vp, u32, C.c_int32]),
        "gpsb_host_channel_code": (C.POINTER(u8), [vp]),
        # position fix and RTCM frames (rows N4; host/fix.c, host/rtcm.c)
        "gps_pos_solve_init": (None, [vp]), "gps_pos_solve": (None, [vp]), "solving_is_busy": (u8, []),
        "gps_master_calculate_pos": (None, [vp]), "sdrobs2obsd": (None, [vp, i32, vp]),
        "gpsb_host_fix_channels": (i32, [vp, u32]), "gpsb_host_fix_state": (None, [C.POINTER(FlatFix)]),
        "gpsb_host_fix_reset": (None, []), "gpsb_host_fix_set_start": (None, [vp]),
        "gpsb_host_fix_set_iono": (None, [vp]),
        "gpsb_host_enable_rtcm": (None, [i32]), "gpsb_host_rtcm_enabled": (i32, []),
        "gps_master_transmit_obs": (None, [vp]),
        "gpsb_rtcm_encode_obs": (i32, [vp, i32, vp, u32]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._gpsb_protos = tuple(protos)
    if path is None:
        _hostlib = lib
    return lib


class Channels:
    """A caller-owned array of gps_ch_t records (allocated by the library so Python needs no layout)."""

    def __init__(self, prns, given_freq_hz=None):
        self.lib = load_host_library()
        self.n = len(prns)
        self.base = self.lib.gpsb_host_channels_alloc(self.n)
        for i, prn in enumerate(prns):
            self.lib.gpsb_host_channel_init(self.at(i), int(prn), int(given_freq_hz[i]) if given_freq_hz else 0)

    def at(self, i: int):
        return self.lib.gpsb_host_channel_at(self.base, i)

    def snapshot(self, i: int) -> FlatState:
        s = FlatState()
        self.lib.gpsb_host_snapshot(self.at(i), C.byref(s))
        return s

    def restore(self, i: int, s: FlatState) -> None:
        self.lib.gpsb_host_restore(self.at(i), C.byref(s))

    def code(self, i: int) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.gpsb_host_channel_code(self.at(i)), (1023,)).copy()

    def position_fix(self):
        """One position fix from the channels' current observations and ephemerides (gpsb_host_fix_channels, the
        reference's pntpos without the slicing).  Returns None without a fix, else a dict: latitude / longitude (deg),
        height (m), ECEF position (m), receiver clock bias (s), azimuth / elevation per channel (deg)."""
        self.lib.gps_pos_solve_init(self.base)
        self._fix_registered = True
        if self.lib.gpsb_host_fix_channels(self.base, self.n) != 1:
            return None
        f = FlatFix()
        self.lib.gpsb_host_fix_state(C.byref(f))
        dbl = lambda arr: np.array(list(arr), np.uint64).view(np.float64)
        pos, ecef, azel = dbl(f.final_pos), dbl(f.rr)[:3], dbl(f.azel)[:2 * self.n].reshape(self.n, 2)
        return dict(lat_deg=float(pos[0]), lon_deg=float(pos[1]), height_m=float(pos[2]), ecef_m=ecef,
                    clock_bias_s=float(dbl([f.dtr0])[0]), azel_deg=azel)

    def rtcm_observations(self) -> bytes:
        """The channels' current observations as one RTCM 3 frame, message 1075 (empty when there is nothing to send)."""
        obsd = (C.c_uint8 * (48 * self.n))()
        self.lib.sdrobs2obsd(self.base, self.n, obsd)
        out = (C.c_uint8 * 1200)()
        n = self.lib.gpsb_rtcm_encode_obs(obsd, self.n, out, 1200)
        return bytes(out[:n])

    def free(self) -> None:
        if self.base:
            if getattr(self, "_fix_registered", False):          # the solver holds pointers into these records
                self.lib.gps_pos_solve_init(None)
            self.lib.gpsb_host_channels_free(self.base)
            self.base = None


class Receiver:
    """Batched receiver (gpsb_rx_*): all channels of a millisecond in one launch."""

    def __init__(self, engine: Engine, channels: Channels):
        self.lib = load_host_library()
        self.engine, self.channels = engine, channels
        self._rx = C.c_void_p()
        rc = self.lib.gpsb_rx_create(C.byref(self._rx), engine.handle, channels.base, channels.n)
        if rc != 0:
            raise GpsbError(rc, engine.lib.gpsb_last_error().decode())

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise GpsbError(rc, self.engine.lib.gpsb_last_error().decode())

    def track_ms(self, ms: int) -> None:
        self._check(self.lib.gpsb_rx_track_ms(self._rx, ms))

    def track_run(self, ms0: int, n_ms: int, log: bool = True):
        n = self.channels.n
        iq = np.zeros((n_ms, n, 6), np.int16) if log else None
        nav = np.zeros((n_ms, n), np.int8) if log else None
        self._check(self.lib.gpsb_rx_track_run(self._rx, ms0, n_ms, iq.ctypes.data if log else None,
                                               nav.ctypes.data if log else None))
        return iq, nav

    def track_stream_iq2(self, ms0: int, samples: np.ndarray, chunk_ms: int = 0, log: bool = True):
        """gpsb_rx_track_stream_iq2: the 2-bit I/Q container (one byte per sample) streamed and packed behind the loop."""
        samples = np.ascontiguousarray(samples, dtype=np.uint8)
        n_ms = samples.size // 16368
        n = self.channels.n
        iq = np.zeros((n_ms, n, 6), np.int16) if log else None
        nav = np.zeros((n_ms, n), np.int8) if log else None
        self._check(self.lib.gpsb_rx_track_stream_iq2(self._rx, ms0, n_ms, samples.ctypes.data, chunk_ms,
                                                      iq.ctypes.data if log else None, nav.ctypes.data if log else None))
        return iq, nav

    def track_stream(self, ms0: int, packed: np.ndarray, chunk_ms: int = 0, log: bool = True):
        """gpsb_rx_track_stream: the samples (n_ms x 2046 bytes, host memory) are streamed into the HBM ring while
        the device-resident loop is already tracking."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        n_ms = packed.size // 2046
        n = self.channels.n
        iq = np.zeros((n_ms, n, 6), np.int16) if log else None
        nav = np.zeros((n_ms, n), np.int8) if log else None
        self._check(self.lib.gpsb_rx_track_stream(self._rx, ms0, n_ms, packed.ctypes.data, chunk_ms,
                                                  iq.ctypes.data if log else None, nav.ctypes.data if log else None))
        return iq, nav

    def track_file(self, path, ms0: int, n_ms: int, first_byte: int = 0, msb_first: bool = False, iq2: bool = False,
                   log: bool = True):
        """gpsb_rx_track_file: n_ms milliseconds of a recording on disk, streamed into the ring behind the running loop.
        msb_first: first sample of each byte in bit 7; iq2: the 2-bit I / 2-bit Q container (one byte per sample)."""
        n = self.channels.n
        iq = np.zeros((n_ms, n, 6), np.int16) if log else None
        nav = np.zeros((n_ms, n), np.int8) if log else None
        flags = (1 if msb_first else 0) | (2 if iq2 else 0)
        self._check(self.lib.gpsb_rx_track_file(self._rx, str(path).encode(), first_byte, ms0, n_ms, flags,
                                                iq.ctypes.data if log else None, nav.ctypes.data if log else None))
        return iq, nav

    def set_threads(self, n: int) -> None:
        self.lib.gpsb_rx_set_threads(self._rx, n)

    def set_loop_site(self, site: int) -> None:
        """0 automatic (device-resident loop for runs), 1 host loop filters, 2 device."""
        self.lib.gpsb_rx_set_loop_site(self._rx, site)

    def set_slot_walk(self, enable: bool = True, period_ms: int = 0) -> None:
        """gpsb_rx_set_slot_walk: let channels without a refined bit edge move their 4-ms slots (idle gaps between two
        slots, like the MCU's 17-ms schedule) until every satellite delivers subframe time stamps."""
        self._check(self.lib.gpsb_rx_set_slot_walk(self._rx, 1 if enable else 0, period_ms))

    def sync_status(self, i: int) -> SyncStatus:
        s = SyncStatus()
        self._check(self.lib.gpsb_rx_channel_sync(self._rx, i, C.byref(s)))
        return s

    def loop_stats(self):
        """(channel-milliseconds run by k_track_run, channel-milliseconds run on the per-ms host path)"""
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.gpsb_rx_loop_stats(self._rx, C.byref(a), C.byref(b))
        return a.value, b.value

    def acquire_ms(self, ms: int) -> None:
        self._check(self.lib.gpsb_rx_acquire_ms(self._rx, ms))

    def cold_sweep(self, first_bin_hz: int, bin_step_hz: int, n_bins: int, ms0: int, n_ms: int):
        n = self.channels.n
        votes = np.zeros((n, n_bins), np.uint8)
        phases = np.zeros((n, n_bins), np.uint16)
        self._check(self.lib.gpsb_rx_cold_sweep(self._rx, first_bin_hz, bin_step_hz, n_bins, ms0, n_ms,
                                                votes.ctypes.data, phases.ctypes.data))
        return votes, phases

    def cold_start(self, ms0: int, **opts) -> dict:
        """gpsb_rx_cold_start: Doppler sweep, code-phase rounds 1..3 with the cells of coming snapshots computed ahead,
        acquired channels handed to pre-track.  opts: fields of gpsb_cold_start_opts.  Returns the report as a dict."""
        o, r = ColdStartOpts(), ColdStartReport()
        for k, v in opts.items():
            setattr(o, k, v)
        self._check(self.lib.gpsb_rx_cold_start(self._rx, ms0, C.byref(o), C.byref(r)))
        return r.as_dict()

    def close(self) -> None:
        if self._rx:
            self.lib.gpsb_rx_destroy(self._rx)
            self._rx = C.c_void_p()
