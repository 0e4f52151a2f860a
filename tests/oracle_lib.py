"""ctypes access to the checkers under oracle/ (TEST INFRASTRUCTURE - never imported by the product).

* ``Oracle``    - oracle/liboracle.so, the C restatement (always available once built)
* ``Reference`` - oracle/_ref/libgpsref.so, the unmodified reference C compiled from /root/reference
                  (built in the authoring container, travels to the GPU box as a prebuilt .so)
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ORACLE_DIR = REPO / "oracle"
ORACLE_SO = ORACLE_DIR / "liboracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libgpsref.so"
REF_O3 = {"x86-64-v4": ORACLE_DIR / "_ref" / "libgpsref_o3v4.so", "x86-64-v3": ORACLE_DIR / "_ref" / "libgpsref_o3v3.so"}
# CPU flags (as /proc/cpuinfo spells them) each x86-64 micro-architecture level adds over the one below
_LEVEL_FLAGS = {"x86-64-v3": ("avx", "avx2", "bmi1", "bmi2", "f16c", "fma", "abm", "movbe", "xsave"),
                "x86-64-v4": ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")}


def best_o3_variant():
    """(march level, path) of the -O3 build of the reference this machine's CPU can run, or (None, None)."""
    try:
        flags = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
    except OSError:
        return None, None
    v3 = all(f in flags for f in _LEVEL_FLAGS["x86-64-v3"])
    v4 = v3 and all(f in flags for f in _LEVEL_FLAGS["x86-64-v4"])
    for level, ok in (("x86-64-v4", v4), ("x86-64-v3", v3)):
        if ok and REF_O3[level].exists():
            return level, REF_O3[level]
    return None, None

sys.path.insert(0, str(REPO))

MS_BYTES = 2046
CHIPS = 1023


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def ensure_built() -> None:
    if not ORACLE_SO.exists() or (Path("/root/reference").exists() and
                                  not all(p.exists() for p in (REF_SO, *REF_O3.values()))):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "all"], check=True, capture_output=True)


class FlatState(C.Structure):
    """Mirror of include/gpsb_flat_state.h."""
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

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else int(v)
        return out


def f32_bits(x: float) -> int:
    return int(np.float32(x).view(np.uint32))


def bits_f32(u: int) -> float:
    return float(np.uint32(u).view(np.float32))


class Oracle:
    def __init__(self):
        ensure_built()
        self.lib = lib = C.CDLL(str(ORACLE_SO))
        vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int, C.c_float
        lib.orc_ca_code.argtypes = [i32, vp]
        lib.orc_replica.argtypes = [vp, u32, vp]
        lib.orc_nco_step.argtypes = [f32]; lib.orc_nco_step.restype = u32
        lib.orc_nco_step32.argtypes = [u32]; lib.orc_nco_step32.restype = u32
        lib.orc_mix.argtypes = [vp, u32, u32, vp, vp]; lib.orc_mix.restype = u32
        lib.orc_corr_sums.argtypes = [vp, vp, vp, u32, C.POINTER(i32), C.POINTER(i32)]
        lib.orc_correlation_iq.argtypes = [vp, vp, vp, u32, C.POINTER(C.c_int16), C.POINTER(C.c_int16)]
        lib.orc_correlation8.argtypes = [vp, vp, vp, u32]; lib.orc_correlation8.restype = C.c_int16
        lib.orc_correlation_search.argtypes = [vp, vp, vp, u32, u32, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
        lib.orc_correlation_search.restype = C.c_uint16
        lib.orc_rewind_if_phase.argtypes = [u32, f32, u32]; lib.orc_rewind_if_phase.restype = u32
        lib.orc_search_cell.argtypes = [vp, vp, f32, u32, u32, u32, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
        lib.orc_search_cell.restype = C.c_uint16
        lib.orc_track_epl.argtypes = [vp, vp, f32, u32, f32, vp]; lib.orc_track_epl.restype = u32
        lib.orc_epl_explicit.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, vp]
        lib.orc_epl_offsets.argtypes = [f32] + [C.POINTER(u32)] * 4
        lib.orc_time_epl.argtypes = [vp, u32, vp, u32, vp, vp, vp, vp, vp]; lib.orc_time_epl.restype = C.c_double
        lib.orc_time_sweep.argtypes = [vp, u32, vp, u32, i32, i32, u32, u32, vp]
        lib.orc_time_sweep.restype = C.c_double

    def ca_code(self, prn: int) -> np.ndarray:
        chips = np.zeros(CHIPS, np.uint8)
        if self.lib.orc_ca_code(prn, _p(chips)) != 0:
            raise ValueError("bad prn %d" % prn)
        return chips

    def replica(self, chips: np.ndarray, bits: int) -> np.ndarray:
        rep = np.zeros(2048, np.uint8)
        self.lib.orc_replica(_p(chips), bits, _p(rep))
        return rep

    def nco_step(self, freq_hz: float) -> int:
        return int(self.lib.orc_nco_step(np.float32(freq_hz)))

    def nco_step32(self, freq_hz: float) -> int:
        return int(self.lib.orc_nco_step32(self.nco_step(freq_hz)))

    def mix(self, sig: np.ndarray, acc0: int, step32: int):
        di = np.zeros(2048, np.uint8)
        dq = np.zeros(2048, np.uint8)
        acc = self.lib.orc_mix(_p(sig), acc0, step32, _p(di), _p(dq))
        return di, dq, int(acc)

    def correlation_iq(self, rep, di, dq, off: int):
        a, b = C.c_int16(), C.c_int16()
        self.lib.orc_correlation_iq(_p(rep), _p(di), _p(dq), off, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def correlation8(self, rep, di, dq, off: int) -> int:
        return int(self.lib.orc_correlation8(_p(rep), _p(di), _p(dq), off))

    def correlation_search(self, rep, di, dq, start: int, stop: int):
        avg, ph = C.c_uint16(), C.c_uint16()
        mx = self.lib.orc_correlation_search(_p(rep), _p(di), _p(dq), start, stop, C.byref(avg), C.byref(ph))
        return int(mx), int(ph.value), int(avg.value)

    def search_cell(self, chips, sig, freq_hz: float, bits: int, start: int, stop: int):
        avg, ph = C.c_uint16(), C.c_uint16()
        mx = self.lib.orc_search_cell(_p(chips), _p(sig), np.float32(freq_hz), bits, start, stop,
                                      C.byref(avg), C.byref(ph))
        return int(mx), int(ph.value), int(avg.value)

    def epl_explicit(self, chips, sig, acc0, step32, off_e, off_p, off_l, bits) -> np.ndarray:
        out = np.zeros(6, np.int16)
        self.lib.orc_epl_explicit(_p(chips), _p(sig), acc0, step32, off_e, off_p, off_l, bits, _p(out))
        return out

    def track_epl(self, chips, sig, if_freq_offset_hz: float, accum_in: int, code_phase_fine: float):
        out = np.zeros(6, np.int16)
        acc = self.lib.orc_track_epl(_p(chips), _p(sig), np.float32(if_freq_offset_hz), accum_in,
                                     np.float32(code_phase_fine), _p(out))
        return out, int(acc)

    def epl_offsets(self, code_phase_fine: float):
        v = [C.c_uint32() for _ in range(4)]
        self.lib.orc_epl_offsets(np.float32(code_phase_fine), *[C.byref(x) for x in v])
        return tuple(int(x.value) for x in v)  # e, p, l, bits


class Reference:
    """The compiled, unmodified reference (oracle/_ref/libgpsref.so)."""

    def __init__(self, so_path=None):
        ensure_built()
        so_path = Path(so_path) if so_path else REF_SO
        if not so_path.exists():
            raise FileNotFoundError(str(so_path))
        self.lib = lib = C.CDLL(str(so_path))
        vp, u32, i32, f32, u16 = C.c_void_p, C.c_uint32, C.c_int32, C.c_float, C.c_uint16
        lib.gps_fill_summ_table()
        lib.ref_set_packet_cnt.argtypes = [u32]
        lib.ref_sizeof_channel.restype = u32
        lib.ref_channels_alloc.argtypes = [u32]; lib.ref_channels_alloc.restype = vp
        lib.ref_channels_free.argtypes = [vp]
        lib.ref_channel_at.argtypes = [vp, u32]; lib.ref_channel_at.restype = vp
        lib.ref_channel_init.argtypes = [vp, u32, i32]
        lib.ref_channel_prn_code.argtypes = [vp]; lib.ref_channel_prn_code.restype = C.POINTER(C.c_uint8)
        lib.ref_channel_snapshot.argtypes = [vp, C.POINTER(FlatState)]
        lib.ref_channel_restore.argtypes = [vp, C.POINTER(FlatState)]
        for n in ("ref_tmp_prn_data", "ref_tmp_data_i", "ref_tmp_data_q"):
            getattr(lib, n).restype = C.POINTER(C.c_uint16)
        lib.ref_sim_buffer.argtypes = [vp, u32, u32]
        lib.ref_search_cell.argtypes = [vp, vp, i32, u32, u32, u32, C.POINTER(u16), C.POINTER(u16)]
        lib.ref_search_cell.restype = u16
        lib.ref_search_cell_f.argtypes = [vp, vp, f32, u32, u32, u32, C.POINTER(u16), C.POINTER(u16)]
        lib.ref_search_cell_f.restype = u16
        lib.ref_iq_cell.argtypes = [vp, vp, f32, u32, u32, u32, vp]
        lib.ref_sweep_cells.argtypes = [vp, u32, vp, u32, i32, i32, u32, u32, vp]
        lib.ref_track_run.argtypes = [vp, vp, u32, u32, vp, vp, vp]
        lib.ref_track_run_walk.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp]
        lib.ref_sizeof_walk.restype = u32
        lib.ref_epl_cell.argtypes = [vp, vp, f32, u32, f32, vp, C.POINTER(u32)]
        lib.ref_now_s.restype = C.c_double
        # plain reference primitives (gps_misc.h:198-216)
        lib.gps_correlation8.argtypes = [vp, vp, vp, u16]; lib.gps_correlation8.restype = C.c_int16
        lib.gps_correlation_iq.argtypes = [vp, vp, vp, u16, C.POINTER(C.c_int16), C.POINTER(C.c_int16)]
        lib.correlation_search.argtypes = [vp, vp, vp, u16, u16, C.POINTER(u16), C.POINTER(u16)]
        lib.correlation_search.restype = u16
        lib.gps_shift_to_zero_freq.argtypes = [vp, vp, vp, f32]
        lib.gps_generate_prn_data2.argtypes = [vp, vp, u16]
        lib.gps_tracking_process.argtypes = [vp, vp, C.c_uint8]
        lib.acquisition_process_channel.argtypes = [vp, vp]
        lib.acquisition_start_channel.argtypes = [vp]
        lib.acquisition_start_code_search_channel.argtypes = [vp]
        lib.acquisition_start_code_search3_channel.argtypes = [vp]

    # channels ----------------------------------------------------------------
    def channels(self, n: int):
        return self.lib.ref_channels_alloc(n)

    def channel_at(self, base, i: int):
        return self.lib.ref_channel_at(base, i)

    def channel_init(self, ch, prn: int, given_freq: int = 0) -> None:
        self.lib.ref_channel_init(ch, prn, given_freq)

    def prn_code(self, ch) -> np.ndarray:
        return np.ctypeslib.as_array(self.lib.ref_channel_prn_code(ch), (CHIPS,)).copy()

    def snapshot(self, ch) -> FlatState:
        s = FlatState()
        self.lib.ref_channel_snapshot(ch, C.byref(s))
        return s

    def restore(self, ch, s: FlatState) -> None:
        self.lib.ref_channel_restore(ch, C.byref(s))

    def set_ms(self, ms: int) -> None:
        self.lib.ref_set_packet_cnt(ms)

    # fixtures ----------------------------------------------------------------
    def sim_buffer(self, noise: int = 0, seed: int = 1) -> np.ndarray:
        out = np.zeros(MS_BYTES, np.uint8)
        self.lib.ref_sim_buffer(_p(out), noise, seed)
        return out

    # cells -------------------------------------------------------------------
    def search_cell(self, ch, sig, freq_offset_hz: int, bits: int, start: int, stop: int):
        avg, ph = C.c_uint16(), C.c_uint16()
        mx = self.lib.ref_search_cell(ch, _p(sig), freq_offset_hz, bits, start, stop, C.byref(avg), C.byref(ph))
        return int(mx), int(ph.value), int(avg.value)

    def search_cell_f(self, ch, sig, freq_hz: float, bits: int, start: int, stop: int):
        avg, ph = C.c_uint16(), C.c_uint16()
        mx = self.lib.ref_search_cell_f(ch, _p(sig), np.float32(freq_hz), bits, start, stop,
                                        C.byref(avg), C.byref(ph))
        return int(mx), int(ph.value), int(avg.value)

    def iq_cell(self, ch, sig, freq_hz: float, bits: int, start: int, stop: int) -> np.ndarray:
        out = np.zeros((stop - start, 2), np.int16)
        self.lib.ref_iq_cell(ch, _p(sig), np.float32(freq_hz), bits, start, stop, _p(out))
        return out

    def epl_cell(self, ch, sig, if_freq_offset_hz: float, accum_in: int, code_phase_fine: float):
        out = np.zeros(6, np.int16)
        acc = C.c_uint32()
        self.lib.ref_epl_cell(ch, _p(sig), np.float32(if_freq_offset_hz), accum_in,
                              np.float32(code_phase_fine), _p(out), C.byref(acc))
        return out, int(acc.value)

    def sweep_cells(self, chans, n_sv: int, signal: np.ndarray, n_ms: int, first_bin_hz: int,
                    bin_step_hz: int, n_bins: int, bits: int = 0) -> np.ndarray:
        out = np.zeros((n_sv, n_bins, n_ms, 3), np.uint16)
        self.lib.ref_sweep_cells(chans, n_sv, _p(signal), n_ms, first_bin_hz, bin_step_hz, n_bins, bits, _p(out))
        return out

    def track_run(self, ch, signal: np.ndarray, ms_first: int, n_ms: int):
        iq = np.zeros((n_ms, 6), np.int16)
        nav = np.zeros(n_ms, np.int8)
        st = np.zeros((n_ms, 2), np.float32)
        self.lib.ref_track_run(ch, _p(signal), ms_first, n_ms, _p(iq), _p(nav), _p(st))
        return iq, nav, st


class RefWalk(C.Structure):
    """oracle/ref_shim.c, ref_walk: the checker's restatement of the slot-phase walk"""
    _fields_ = [("enable", C.c_uint32), ("period_ms", C.c_uint32), ("slot_phase", C.c_uint32), ("gap_first", C.c_uint32),
                ("gap_len", C.c_uint32), ("phase_since", C.c_uint32), ("gaps_taken", C.c_uint32), ("edge_pos", C.c_uint32),
                ("edges_at", C.c_uint32 * 4), ("armed", C.c_uint32),
                ("slot_first_ms", C.c_uint32), ("sign", C.c_uint8 * 4), ("ip", C.c_int16 * 4), ("slot_fill", C.c_uint32)]


def _track_run_walk(self, ch, signal: np.ndarray, ms_first: int, n_ms: int, walk: RefWalk):
    """The unmodified reference on the walked (millisecond, slot index) schedule; returns iq, nav, index per ms."""
    assert self.lib.ref_sizeof_walk() == C.sizeof(RefWalk)
    iq = np.zeros((n_ms, 6), np.int16)
    nav = np.zeros(n_ms, np.int8)
    idx = np.zeros(n_ms, np.uint8)
    self.lib.ref_track_run_walk(ch, _p(signal), ms_first, n_ms, C.byref(walk), _p(iq), _p(nav), _p(idx))
    return iq, nav, idx


Reference.track_run_walk = _track_run_walk


def have_reference() -> bool:
    try:
        ensure_built()
    except Exception:
        pass
    return REF_SO.exists()
