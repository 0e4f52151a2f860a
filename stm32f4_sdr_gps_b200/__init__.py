"""gpsb200 - B200-native GPS L1 C/A acquisition search + E/P/L correlator tracking engine.

Scope: exactly the hot path of iliasam/STM32F4_SDR_GPS named in BASELINE.json:
hand-written sm_100a CUDA kernels behind a C ABI (``include/gpsb.h``), a host-side C mirror of the
reference's acquisition / tracking API (``include/gpsb_host.h``) and this thin Python binding.
"""
from .engine import (CHIPS, EPL_REQ, FRAME_BYTES, IF_FREQ_HZ, MS_BYTES, MS_SAMPLES, OFFSETS, SEARCH_REQ,
                     SEARCH_RES, Engine, GpsbError, load_library, nco_step, nco_step32)

__all__ = ["CHIPS", "EPL_REQ", "FRAME_BYTES", "IF_FREQ_HZ", "MS_BYTES", "MS_SAMPLES", "OFFSETS",
           "SEARCH_REQ", "SEARCH_RES", "Engine", "GpsbError", "load_library", "nco_step", "nco_step32"]
from .host_api import Channels, FlatFix, FlatState, Plan, Receiver, SearchRes, load_host_library  # noqa: E402

__all__ += ["Channels", "FlatFix", "FlatState", "Plan", "Receiver", "SearchRes", "load_host_library"]
