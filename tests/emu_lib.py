"""TEST INFRASTRUCTURE: builds and binds tests/emu/loop_emu.c, the CPU emulation of the device-resident
tracking loop (same core sources as the kernel, compiled by gcc with the device math selected)."""
import ctypes as C
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
SRC = HERE / "emu" / "loop_emu.c"
OUT = HERE / "emu" / "_build" / "libloop_emu.so"
DEPS = [SRC, *sorted((REPO / "stm32f4_sdr_gps_b200" / "core").glob("*.h")), *sorted((REPO / "include").glob("*.h"))]

_lib = None


def load_emulator() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not OUT.exists() or any(p.stat().st_mtime > OUT.stat().st_mtime for p in DEPS):
        OUT.parent.mkdir(exist_ok=True)
        cmd = ["gcc", "-std=gnu11", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-ffp-contract=off",
               "-fno-fast-math", "-D_GNU_SOURCE", "-Wall", "-Wextra", "-I", str(REPO / "include"), "-o", str(OUT),
               str(SRC), "-lm"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("emulator build failed:\n" + proc.stderr)
    lib = C.CDLL(str(OUT))
    vp, u32, i32 = C.c_void_p, C.c_uint32, C.c_int32
    lib.emu_sizeof_aux.restype = u32
    lib.emu_sizeof_channel.restype = u32
    lib.emu_epl_cell.restype = None
    lib.emu_epl_cell.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, u32, vp]
    lib.emu_track_run.restype = i32
    lib.emu_track_run.argtypes = [vp, vp, vp, u32, u32, u32, vp, vp, C.POINTER(u32), vp]
    lib.emu_track_run_phase.restype = i32
    lib.emu_track_run_phase.argtypes = [vp, vp, vp, u32, u32, u32, u32, vp, vp, C.POINTER(u32), vp]
    lib.emu_resolve_snr.restype = None
    lib.emu_resolve_snr.argtypes = [vp, vp]
    lib.emu_compare_float_math.restype = C.c_uint64
    lib.emu_compare_float_math.argtypes = [i32, i32, i32, C.POINTER(i32)]
    lib.emu_host_costas.restype = None
    lib.emu_host_costas.argtypes = [i32, i32, vp]
    lib.emu_host_fll_angle.restype = None
    lib.emu_host_fll_angle.argtypes = [i32, i32, vp]
    lib.emu_rand31_mismatches.restype = u32
    lib.emu_rand31_mismatches.argtypes = [u32]
    lib.emu_fold_mismatches.restype = C.c_uint64
    lib.emu_fold_mismatches.argtypes = [u32, u32]
    _lib = lib
    return lib
