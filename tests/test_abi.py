"""The C-ABI library loads and exports every symbol include/gpsb.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback).  CPU only - no compute calls."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _declared(header: str):
    text = (REPO / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpsb_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    from stm32f4_sdr_gps_b200 import build, load_library
    build.build_cuda()
    lib = load_library()
    names = _declared("gpsb.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libgpsb_cuda.so does not export %s" % n
    assert lib.gpsb_abi_version() >= 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stm32f4_sdr_gps_b200 import Engine, GpsbError
    with pytest.raises(GpsbError) as e:
        Engine(device=0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_bad_arguments_do_not_touch_the_gpu():
    from stm32f4_sdr_gps_b200 import load_library
    lib = load_library()
    assert lib.gpsb_create(None, 0, 1, 1) == -1
    ctx = C.c_void_p()
    assert lib.gpsb_create(C.byref(ctx), 0, 0, 16) == -1
    assert lib.gpsb_track_epl(None, 0, None, None) == -1
    assert b"null" in lib.gpsb_last_error() or b"bad" in lib.gpsb_last_error()


def test_record_layouts_match_header():
    from stm32f4_sdr_gps_b200 import EPL_REQ, SEARCH_REQ, SEARCH_RES
    assert EPL_REQ.itemsize == 24 and EPL_REQ.fields["off_e"][1] == 16
    assert SEARCH_REQ.itemsize == 24 and SEARCH_REQ.fields["off_bits"][1] == 16
    assert SEARCH_RES.itemsize == 8


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    for p in (REPO / "stm32f4_sdr_gps_b200").rglob("*"):
        if p.suffix in (".py", ".c", ".h", ".cu", ".cuh"):
            txt = p.read_text()
            if p.name == "build.py":
                continue  # build_oracle() compiles the checker; it does not load it
            assert "liboracle" not in txt and "libgpsref" not in txt and "gps_oracle" not in txt, p
