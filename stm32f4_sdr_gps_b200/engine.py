"""ctypes binding of the C ABI in ``include/gpsb.h`` (``libgpsb_cuda.so``).

This module is plumbing only: it loads the in-tree shared library, declares the prototypes and moves
numpy buffers across the boundary.  All arithmetic happens in the CUDA kernels; there is no Python or
CPU implementation of any of it here, and importing / constructing fails loudly when the native
library or a usable GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

CHIPS = 1023
MS_BYTES = 2046
MS_SAMPLES = 16368
FRAME_BYTES = 2048
OFFSETS = 2046
IF_FREQ_HZ = 4092000            # PM/config.h:23
SAMPLE_RATE_HZ = 16368000       # PM/config.h:24
IF_NCO_STEP_HZ = np.float32(0.003810972)   # PM/config.h:53

EPL_REQ = np.dtype([("sv_slot", "<u4"), ("ms_index", "<u4"), ("acc0", "<u4"), ("step32", "<u4"),
                    ("off_e", "<u2"), ("off_p", "<u2"), ("off_l", "<u2"), ("off_bits", "<u2")])
SEARCH_REQ = np.dtype([("sv_slot", "<u4"), ("ms_index", "<u4"), ("acc0", "<u4"), ("step32", "<u4"),
                       ("off_bits", "<u2"), ("start", "<u2"), ("stop", "<u2"), ("flags", "<u2")])
SEARCH_RES = np.dtype([("max", "<u2"), ("phase", "<u2"), ("avg", "<u2"), ("reserved", "<u2")])
assert EPL_REQ.itemsize == 24 and SEARCH_REQ.itemsize == 24 and SEARCH_RES.itemsize == 8


class GpsbError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__("gpsb error %d: %s" % (code, message))
        self.code = code


_lib = None


def _p(arr: np.ndarray):
    return arr.ctypes.data_as(C.c_void_p)


def load_library(path: Path | None = None) -> C.CDLL:
    """Load libgpsb_cuda.so (in-tree).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else _build.CUDA_LIB
    if not p.exists():
        raise ImportError(
            "%s is missing: build it with `python -m stm32f4_sdr_gps_b200.build` "
            "(or __graft_entry__.build()); the engine has no non-CUDA implementation" % p)
    lib = C.CDLL(str(p))
    vp, u32, i32, u16, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint16, C.c_uint64
    protos = {
        "gpsb_create": (i32, [C.POINTER(vp), i32, u32, u32]),
        "gpsb_destroy": (None, [vp]),
        "gpsb_last_error": (C.c_char_p, []),
        "gpsb_abi_version": (u32, []),
        "gpsb_launch_count": (u64, [vp]),
        "gpsb_ring_ms": (u32, [vp]),
        "gpsb_set_stream": (i32, [vp, vp]),
        "gpsb_synchronize": (i32, [vp]),
        "gpsb_timer_start": (i32, [vp, u32]),
        "gpsb_timer_stop": (i32, [vp, u32]),
        "gpsb_timer_elapsed_ms": (i32, [vp, u32, C.POINTER(C.c_float)]),
        "gpsb_set_code": (i32, [vp, u32, vp]),
        "gpsb_set_code_prn": (i32, [vp, u32, u32]),
        "gpsb_get_code": (i32, [vp, u32, vp]),
        "gpsb_upload_signal": (i32, [vp, u32, u32, vp]),
        "gpsb_upload_signal_async": (i32, [vp, u32, u32, vp]),
        "gpsb_upload_signal_iq2": (i32, [vp, u32, u32, vp]),
        "gpsb_download_signal": (i32, [vp, u32, u32, vp]),
        "gpsb_track_epl": (i32, [vp, u32, vp, vp]),
        "gpsb_search": (i32, [vp, u32, vp, vp]),
        "gpsb_search_iq": (i32, [vp, vp, vp]),
        "gpsb_sweep": (i32, [vp, vp, u32, vp, u32, u32, u32, u32, vp]),
        "gpsb_set_sweep_method": (i32, [vp, i32]),
        "gpsb_set_realtime": (i32, [vp, i32]),
        "gpsb_session_begin": (i32, [vp, u32]),
        "gpsb_session_end": (i32, [vp]),
        "gpsb_session_slots": (u32, [vp]),
        "gpsb_session_post": (i32, [vp, u32, vp, C.POINTER(u32)]),
        "gpsb_session_wait": (i32, [vp, u32, u32, vp]),
        "gpsb_track_epl_dev": (i32, [vp, u32, vp, vp]),
        "gpsb_prompt_iq_dev": (i32, [vp, u32, vp, vp]),
        "gpsb_prompt_iq": (i32, [vp, u32, vp, vp]),
        "gpsb_set_epl_batch_min": (i32, [vp, u32]),
        "gpsb_search_dev": (i32, [vp, u32, vp, vp]),
        "gpsb_sweep_dev": (i32, [vp, vp, u32, vp, u32, u32, u32, u32, vp]),
        "gpsb_set_epl_batch_kernel": (i32, [vp, i32]),
        "gpsb_comm_unique_id": (i32, [vp]),
        "gpsb_comm_init": (i32, [vp, i32, i32, vp]),
        "gpsb_comm_destroy": (i32, [vp]),
        "gpsb_comm_rank": (i32, [vp]),
        "gpsb_comm_size": (i32, [vp]),
        "gpsb_sweep_gather": (i32, [vp, vp, u32, vp, u32, u32, u32, u32, vp]),
        "gpsb_sweep_gather_dev": (i32, [vp, vp, u32, vp, u32, u32, u32, u32, C.POINTER(vp)]),
        "gpsb_track_loop": (i32, [vp, u32, vp, u32, vp, u32, u32, u32, vp, vp, vp]),
        "gpsb_track_loop_dev": (i32, [vp, u32, vp, vp, u32, u32, vp, vp, vp]),
        "gpsb_track_loop_dev_ex": (i32, [vp, u32, vp, vp, u32, u32, vp, vp, vp, u32]),
        "gpsb_track_loop_begin": (i32, [vp, u32, vp, u32, vp, u32, u32, u32, vp, vp, vp, u32]),
        "gpsb_track_loop_end": (i32, [vp]),
        "gpsb_stream_reset": (i32, [vp, u32]),
        "gpsb_stream_push": (i32, [vp, u32, u32, vp]),
        "gpsb_stream_push_iq2": (i32, [vp, u32, u32, vp]),
        "gpsb_stream_wait": (i32, [vp]),
        "gpsb_stream_progress": (u32, [vp, u32]),
        "gpsb_stream_set_timeout_ms": (i32, [vp, u32]),
        "gpsb_stream_loop_running": (i32, [vp]),
        "gpsb_stream_timeout_ms": (u32, [vp]),
        "gpsb_stream_abort": (i32, [vp]),
        "gpsb_stream_copies_pending": (i32, [vp]),
        "gpsb_code_rounds": (i32, [vp, u32, vp, u32, vp, u32, u32, u32, u32, vp, vp]),
        "gpsb_track_loop_record_bytes": (None, [C.POINTER(u32), C.POINTER(u32)]),
        "gpsb_l0_loop_math": (i32, [vp, i32, C.c_int32, u32, vp]),
        "gpsb_l0_generate_prn_data2": (i32, [vp, vp, vp, u16]),
        "gpsb_l0_shift_to_zero_freq": (i32, [vp, vp, vp, vp, u32, u32, C.POINTER(u32)]),
        "gpsb_l0_correlation_iq": (i32, [vp, vp, vp, vp, u16, C.POINTER(C.c_int16), C.POINTER(C.c_int16)]),
        "gpsb_l0_correlation8": (i32, [vp, vp, vp, vp, u16, C.POINTER(C.c_int16)]),
        "gpsb_l0_correlation_search": (i32, [vp, vp, vp, vp, u16, u16, C.POINTER(u16), C.POINTER(u16),
                                             C.POINTER(u16)]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)          # AttributeError here == ABI/header mismatch
        fn.restype = res
        fn.argtypes = args
    lib._gpsb_protos = tuple(protos)
    if path is None:
        _lib = lib
    return lib


def nco_step(freq_hz) -> int:
    """The reference NCO word, PM/GPS/gps_misc.c:219: (uint32_t)(freq_hz / IF_NCO_STEP_HZ) in fp32."""
    q = np.float32(freq_hz) / IF_NCO_STEP_HZ
    return int(np.uint32(q))


def nco_step32(freq_hz) -> int:
    """Per-32-sample phase advance, PM/GPS/gps_misc.c:220-221."""
    return (nco_step(freq_hz) * 32) & 0xFFFFFFFF


class Engine:
    """One context per GPU (``gpsb_ctx``)."""

    def __init__(self, device: int = 0, max_sv: int = 32, ring_ms: int = 1024):
        self.lib = load_library()
        self._ctx = C.c_void_p()
        self.device, self.max_sv, self.ring_ms = device, max_sv, ring_ms
        self._check(self.lib.gpsb_create(C.byref(self._ctx), device, max_sv, ring_ms))

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise GpsbError(rc, self.lib.gpsb_last_error().decode("utf-8", "replace"))

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.gpsb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._ctx

    @property
    def launch_count(self) -> int:
        return int(self.lib.gpsb_launch_count(self._ctx))

    def set_stream(self, cuda_stream: int | None) -> None:
        self._check(self.lib.gpsb_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    def synchronize(self) -> None:
        self._check(self.lib.gpsb_synchronize(self._ctx))

    def timer_start(self, slot: int = 0) -> None:
        self._check(self.lib.gpsb_timer_start(self._ctx, slot))

    def timer_stop(self, slot: int = 0) -> None:
        self._check(self.lib.gpsb_timer_stop(self._ctx, slot))

    def timer_elapsed_ms(self, slot: int = 0) -> float:
        ms = C.c_float()
        self._check(self.lib.gpsb_timer_elapsed_ms(self._ctx, slot, C.byref(ms)))
        return float(ms.value)

    # ------------------------------------------------------------------ resident data
    def set_code(self, slot: int, chips: np.ndarray) -> None:
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        if chips.shape != (CHIPS,):
            raise ValueError("chips must have shape (1023,)")
        self._check(self.lib.gpsb_set_code(self._ctx, slot, _p(chips)))

    def set_code_prn(self, slot: int, prn: int) -> None:
        self._check(self.lib.gpsb_set_code_prn(self._ctx, slot, prn))

    def get_code(self, slot: int) -> np.ndarray:
        out = np.zeros(CHIPS, np.uint8)
        self._check(self.lib.gpsb_get_code(self._ctx, slot, _p(out)))
        return out

    def upload_signal(self, ms0: int, packed: np.ndarray) -> None:
        packed = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1)
        if packed.size % MS_BYTES:
            raise ValueError("signal length must be a multiple of 2046 bytes")
        self._check(self.lib.gpsb_upload_signal(self._ctx, ms0, packed.size // MS_BYTES, _p(packed)))

    def upload_signal_iq2(self, ms0: int, samples: np.ndarray) -> None:
        samples = np.ascontiguousarray(samples, dtype=np.uint8).reshape(-1)
        if samples.size % MS_SAMPLES:
            raise ValueError("sample count must be a multiple of 16368")
        self._check(self.lib.gpsb_upload_signal_iq2(self._ctx, ms0, samples.size // MS_SAMPLES, _p(samples)))

    def download_signal(self, ms0: int, n_ms: int) -> np.ndarray:
        out = np.zeros(n_ms * MS_BYTES, np.uint8)
        self._check(self.lib.gpsb_download_signal(self._ctx, ms0, n_ms, _p(out)))
        return out.reshape(n_ms, MS_BYTES)

    # ------------------------------------------------------------------ level 1
    def track_epl(self, reqs: np.ndarray) -> np.ndarray:
        reqs = np.ascontiguousarray(reqs, dtype=EPL_REQ).reshape(-1)
        out = np.zeros((reqs.size, 6), np.int16)
        self._check(self.lib.gpsb_track_epl(self._ctx, reqs.size, _p(reqs), _p(out)))
        return out

    def prompt_iq(self, reqs: np.ndarray) -> np.ndarray:
        """The prompt arm alone (I, Q) of every request: gpsb_prompt_iq."""
        reqs = np.ascontiguousarray(reqs, dtype=EPL_REQ).reshape(-1)
        out = np.zeros((reqs.size, 2), np.int16)
        self._check(self.lib.gpsb_prompt_iq(self._ctx, reqs.size, _p(reqs), _p(out)))
        return out

    def set_epl_batch_min(self, n_cells: int) -> None:
        self._check(self.lib.gpsb_set_epl_batch_min(self._ctx, n_cells))

    def search(self, reqs: np.ndarray) -> np.ndarray:
        reqs = np.ascontiguousarray(reqs, dtype=SEARCH_REQ).reshape(-1)
        res = np.zeros(reqs.size, SEARCH_RES)
        self._check(self.lib.gpsb_search(self._ctx, reqs.size, _p(reqs), _p(res)))
        return res

    def search_iq(self, req: np.ndarray) -> np.ndarray:
        req = np.ascontiguousarray(req, dtype=SEARCH_REQ).reshape(-1)[:1]
        n = max(0, int(req["stop"][0]) - int(req["start"][0]))
        iq = np.zeros((n, 2), np.int16)
        self._check(self.lib.gpsb_search_iq(self._ctx, _p(req), _p(iq)))
        return iq

    def sweep(self, sv_slots, step32, ms0: int, n_ms: int, off_bits: int = 0) -> np.ndarray:
        sv = np.ascontiguousarray(sv_slots, dtype=np.uint32)
        st = np.ascontiguousarray(step32, dtype=np.uint32)
        res = np.zeros((sv.size, st.size, n_ms), SEARCH_RES)
        self._check(self.lib.gpsb_sweep(self._ctx, _p(sv), sv.size, _p(st), st.size, ms0, n_ms, off_bits, _p(res)))
        return res

    def set_epl_batch_kernel(self, kernel: int) -> None:
        """0 = TMA ring in shared memory (default), 1 = frames staged in registers (kept for comparison)."""
        self._check(self.lib.gpsb_set_epl_batch_kernel(self._ctx, kernel))

    def set_sweep_method(self, method: int) -> None:
        """0 = direct XOR/popcount, 1 = byte-popcount dp4a correlation (default)."""
        self._check(self.lib.gpsb_set_sweep_method(self._ctx, method))

    def set_realtime(self, enabled: bool) -> None:
        self._check(self.lib.gpsb_set_realtime(self._ctx, int(bool(enabled))))

    def session_begin(self, n_slots: int) -> None:
        self._check(self.lib.gpsb_session_begin(self._ctx, n_slots))

    def session_end(self) -> None:
        self._check(self.lib.gpsb_session_end(self._ctx))

    # device-resident variants: raw device pointers (ints), asynchronous on the context stream
    def prompt_iq_dev(self, n: int, d_req: int, d_out: int) -> None:
        self._check(self.lib.gpsb_prompt_iq_dev(self._ctx, n, C.c_void_p(d_req), C.c_void_p(d_out)))

    def track_epl_dev(self, n: int, d_req: int, d_out: int) -> None:
        self._check(self.lib.gpsb_track_epl_dev(self._ctx, n, C.c_void_p(d_req), C.c_void_p(d_out)))

    def search_dev(self, n: int, d_req: int, d_res: int) -> None:
        self._check(self.lib.gpsb_search_dev(self._ctx, n, C.c_void_p(d_req), C.c_void_p(d_res)))

    def sweep_dev(self, d_sv: int, n_sv: int, d_step32: int, n_bins: int, ms0: int, n_ms: int,
                  off_bits: int, d_res: int) -> None:
        self._check(self.lib.gpsb_sweep_dev(self._ctx, C.c_void_p(d_sv), n_sv, C.c_void_p(d_step32), n_bins,
                                            ms0, n_ms, off_bits, C.c_void_p(d_res)))

    # ------------------------------------------------------------------ multi-GPU
    def comm_unique_id(self) -> bytes:
        """gpsb_comm_unique_id: the 128-byte NCCL rendezvous id (made on rank 0, handed to every rank by the caller)."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.gpsb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes) -> None:
        """gpsb_comm_init: NCCL communicator over all ranks' contexts (collective)."""
        assert len(unique_id) == 128
        self._check(self.lib.gpsb_comm_init(self._ctx, rank, n_ranks, C.create_string_buffer(unique_id, 128)))

    def comm_init_torch(self) -> None:
        """The same with the id handed round by torch.distributed (must be initialised; any backend)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(rank, world, box[0])

    def comm_destroy(self) -> None:
        self._check(self.lib.gpsb_comm_destroy(self._ctx))

    @property
    def comm_size(self) -> int:
        return int(self.lib.gpsb_comm_size(self._ctx))

    def sweep_gather(self, sv_slots, step32, ms0: int, n_ms: int, off_bits: int = 0) -> np.ndarray:
        """gpsb_sweep_gather: the sweep sharded over the communicator's ranks and all-gathered; every rank gets the whole
        grid, shape (n_sv, n_bins, n_ms) of SEARCH_RES."""
        sv = np.ascontiguousarray(sv_slots, dtype=np.uint32)
        st = np.ascontiguousarray(step32, dtype=np.uint32)
        res = np.zeros((sv.size, st.size, n_ms), SEARCH_RES)
        self._check(self.lib.gpsb_sweep_gather(self._ctx, sv.ctypes.data, sv.size, st.ctypes.data, st.size, ms0, n_ms,
                                               off_bits, res.ctypes.data))
        return res

    def sweep_gather_dev(self, d_sv: int, n_sv: int, d_step32: int, n_bins: int, ms0: int, n_ms: int, off_bits: int = 0) -> int:
        """gpsb_sweep_gather_dev: only enqueued; returns the device address of the gathered (sv, bin, ms) grid."""
        out = C.c_void_p()
        self._check(self.lib.gpsb_sweep_gather_dev(self._ctx, C.c_void_p(d_sv), n_sv, C.c_void_p(d_step32), n_bins, ms0,
                                                   n_ms, off_bits, C.byref(out)))
        return int(out.value or 0)

    # ------------------------------------------------------------------ level 0
    def track_loop_dev(self, n_ch: int, d_channels: int, d_aux: int, ms0: int, n_ms: int, d_iq_log: int,
                       d_nav_log: int, d_results: int) -> None:
        """k_track_run on device-resident channel records (raw device pointers; only enqueued)."""
        self._check(self.lib.gpsb_track_loop_dev(self._ctx, n_ch, d_channels, d_aux, ms0, n_ms, d_iq_log or None,
                                                 d_nav_log or None, d_results))

    def track_loop_dev_ex(self, n_ch: int, d_channels: int, d_aux: int, ms0: int, n_ms: int, d_iq_log: int,
                          d_nav_log: int, d_results: int, flags: int) -> None:
        self._check(self.lib.gpsb_track_loop_dev_ex(self._ctx, n_ch, d_channels, d_aux, ms0, n_ms, d_iq_log or None,
                                                    d_nav_log or None, d_results, flags))

    # streaming ingest (include/gpsb.h): frames DMA-ed into the ring while a loop launch consumes them
    def stream_reset(self, ms_valid_upto: int) -> None:
        self._check(self.lib.gpsb_stream_reset(self._ctx, ms_valid_upto))

    def stream_push(self, ms0: int, packed: np.ndarray) -> None:
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self._check(self.lib.gpsb_stream_push(self._ctx, ms0, packed.size // MS_BYTES, _p(packed)))

    def stream_wait(self) -> None:
        self._check(self.lib.gpsb_stream_wait(self._ctx))

    def stream_progress(self, n_ch: int) -> int:
        return int(self.lib.gpsb_stream_progress(self._ctx, n_ch))

    def stream_set_timeout_ms(self, ms: int) -> None:
        self._check(self.lib.gpsb_stream_set_timeout_ms(self._ctx, ms))

    def stream_timeout_ms(self) -> int:
        return int(self.lib.gpsb_stream_timeout_ms(self._ctx))

    def stream_abort(self) -> None:
        """gpsb_stream_abort: the producer gives up - a streaming loop waiting for a frame ends at once (stop == 3)."""
        self._check(self.lib.gpsb_stream_abort(self._ctx))

    def code_rounds(self, n_ch: int, channels: int, aux: np.ndarray, ms0: int, n_ms: int, busy_mask: int, mode) -> np.ndarray:
        """gpsb_code_rounds: the code-phase rounds of the acquisition device-resident (k_code_rounds_run), one CTA per
        channel.  channels: address of n_ch host channel records; aux: their aux records (uint8, in place);
        mode[i] 0 skip / 1 until the state leaves busy_mask / 2 exactly n_ms snapshots.  Returns snapshots used per channel."""
        ch_b, aux_b = self.record_bytes()
        mode = np.ascontiguousarray(mode, dtype=np.uint8)
        used = np.zeros(n_ch, np.uint32)
        assert mode.size == n_ch and aux.dtype == np.uint8 and aux.size == n_ch * aux_b
        self._check(self.lib.gpsb_code_rounds(self._ctx, n_ch, C.c_void_p(channels), ch_b, aux.ctypes.data, aux_b, ms0, n_ms,
                                              busy_mask, mode.ctypes.data, used.ctypes.data))
        return used

    def record_bytes(self):
        a, b = C.c_uint32(), C.c_uint32()
        self.lib.gpsb_track_loop_record_bytes(C.byref(a), C.byref(b))
        return a.value, b.value

    def l0_loop_math(self, kind: int, ip_lo: int, n_ip: int) -> np.ndarray:
        """Device values of the Costas error (kind 0) / FLL angle (kind 1) for ip_lo.. x every qp."""
        out = np.empty((n_ip, 16369), np.float32)
        self._check(self.lib.gpsb_l0_loop_math(self._ctx, kind, ip_lo, n_ip, out.ctypes.data))
        return out

    def l0_generate_prn_data2(self, chips: np.ndarray, offset_bits: int) -> np.ndarray:
        chips = np.ascontiguousarray(chips, dtype=np.uint8)
        data = np.zeros(1023, np.uint16)
        self._check(self.lib.gpsb_l0_generate_prn_data2(self._ctx, _p(chips), _p(data), offset_bits))
        return data

    def l0_shift_to_zero_freq(self, signal: np.ndarray, acc0: int, step32: int, data_i: np.ndarray,
                              data_q: np.ndarray) -> int:
        signal = np.ascontiguousarray(signal, dtype=np.uint8)
        acc = C.c_uint32()
        self._check(self.lib.gpsb_l0_shift_to_zero_freq(self._ctx, _p(signal), _p(data_i), _p(data_q), acc0,
                                                        step32, C.byref(acc)))
        return int(acc.value)

    def l0_correlation_iq(self, prn, data_i, data_q, offset: int):
        ri, rq = C.c_int16(), C.c_int16()
        self._check(self.lib.gpsb_l0_correlation_iq(self._ctx, _p(prn), _p(data_i), _p(data_q), offset,
                                                    C.byref(ri), C.byref(rq)))
        return int(ri.value), int(rq.value)

    def l0_correlation8(self, prn, data_i, data_q, offset: int) -> int:
        r = C.c_int16()
        self._check(self.lib.gpsb_l0_correlation8(self._ctx, _p(prn), _p(data_i), _p(data_q), offset, C.byref(r)))
        return int(r.value)

    def l0_correlation_search(self, prn, data_i, data_q, start: int, stop: int):
        avg, ph, mx = C.c_uint16(), C.c_uint16(), C.c_uint16()
        self._check(self.lib.gpsb_l0_correlation_search(self._ctx, _p(prn), _p(data_i), _p(data_q), start, stop,
                                                        C.byref(avg), C.byref(ph), C.byref(mx)))
        return int(mx.value), int(ph.value), int(avg.value)
