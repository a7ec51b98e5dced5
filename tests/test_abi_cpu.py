"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol include/mppi_b200.h
declares, refuses to create an engine without a CUDA device (there is no CPU fallback), and the
C++ facade compiles against the header."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from assistedmanipulation_b200 import abi

HEADER = os.path.join(ol.ROOT, "include", "mppi_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mppi_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(abi.library_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(abi.EXPORTS) == names  # the ctypes mirror covers the whole header


def test_struct_mirrors_match_the_header_sizes():
    # sizes implied by the header's field lists (all members are 4 or 8 bytes, no surprises in padding)
    assert C.sizeof(abi.Barrier) == 24 and C.sizeof(abi.Quadratic) == 24 and C.sizeof(abi.ToyObjective) == 40
    assert C.sizeof(abi.TrackPoint) == 24 + 16 + 24 * 24 + 24 + 64 + 24 + 8
    assert C.sizeof(abi.AssistedManipulation) == 32 + 24 * 24 + 24 + 64 + 3 * 24 + 24 + 2 * 24 + 12 * 24 + 16 + 24 + 8 + 24 + 24 + 24
    # defaults provided by the library equal the Python mirror (reference track_point.hpp:77-114, assisted_manipulation.hpp:133-206)
    lib = abi.load_library()
    tp, am, toy = abi.TrackPoint(), abi.AssistedManipulation(), abi.ToyObjective()
    lib.mppi_b200_default_track_point(C.byref(tp))
    lib.mppi_b200_default_assisted_manipulation(C.byref(am))
    lib.mppi_b200_default_toy_objective(C.byref(toy))
    for a, b in ((tp, abi.default_track_point()), (am, abi.default_assisted_manipulation()), (toy, abi.default_toy_objective())):
        assert bytes(a) == bytes(b)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = abi.load_library()
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 8, 0.1)
    p = abi.default_toy_objective()
    out = C.c_void_p()
    rc = lib.mppi_b200_create(C.byref(h.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p), C.byref(out))
    assert rc == abi.ERR_CUDA and not out.value
    assert "no CPU fallback" in lib.mppi_b200_last_error(None).decode()
    # configuration errors are reported before any device is touched (mppi.cpp:18-69)
    h = abi.make_config(abi.SYSTEM_TOY, abi.OBJECTIVE_TOY, 0, 0.1)
    assert lib.mppi_b200_create(C.byref(h.cfg), C.cast(C.byref(p), C.c_void_p), C.sizeof(p), C.byref(out)) == abi.ERR_INVALID
    assert lib.mppi_b200_last_error(None).decode() == "trajectory rollouts must be greater than zero"


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load_library(str(tmp_path / "libmppi_b200.so"))


def build_facade_demo(name="facade_demo"):
    exe = os.path.join(ol.ROOT, "tests", "cpp", name)
    src = exe + ".cpp"
    hdrs = [os.path.join(ol.ROOT, "assistedmanipulation_b200", "cpp", "mppi_b200", f) for f in ("trajectory.hpp", "systems.hpp", "linalg.hpp", "forecast.hpp", "logging.hpp")] + [HEADER]
    if not os.path.exists(exe) or any(os.path.getmtime(f) > os.path.getmtime(exe) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-Wall", "-Werror", "-DMPPI_B200_NO_EIGEN", "-I" + os.path.join(ol.ROOT, "include"),
                               "-I" + os.path.join(ol.ROOT, "assistedmanipulation_b200", "cpp"), src, "-L" + os.path.dirname(abi.library_path()), "-lmppi_b200",
                               "-Wl,-rpath," + os.path.dirname(abi.library_path()), "-o", exe])
    return exe


def test_cpp_facade_compiles_and_refuses_without_gpu(tmp_path):
    import torch
    exe = build_facade_demo()
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "toy", "16", "0.2", "1", "-", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr  # Trajectory::create returned nullptr with a reason


def test_python_host_side_keeps_pointers_of_reused_buffers_only():
    """engine.py _host_pointer: the ctypes pointer of a state / wrench array is built once per array OBJECT (the caller
    refills it in place); arrays that need a conversion are converted — and their copy kept alive — on every call; a
    new object in the same argument slot replaces the cached one. No device needed: the method touches no library."""
    import numpy as np
    from assistedmanipulation_b200 import engine
    e = engine.Engine.__new__(engine.Engine)
    e._pointers, e._converted = {}, {}
    a = np.arange(4.0)
    p = e._host_pointer("state", a)
    assert C.addressof(p.contents) == a.ctypes.data and e._host_pointer("state", a) is p      # cached by identity
    a[1] = 7.0
    assert e._host_pointer("state", a)[1] == 7.0                                               # the same memory, refilled in place
    b = np.arange(4.0) + 10
    q = e._host_pointer("state", b)
    assert C.addressof(q.contents) == b.ctypes.data and e._pointers["state"][0] is b          # a new object replaces the entry
    strided = np.arange(8.0)[::2]
    r = e._host_pointer("state", strided)
    assert [r[i] for i in range(4)] == [0.0, 2.0, 4.0, 6.0] and "state" not in e._pointers     # converted, not cached ...
    strided[1] = -1.0
    assert e._host_pointer("state", strided)[1] == -1.0                                        # ... so a change is seen by the next call
    assert e._host_pointer("wrench", None) is None
    w32 = np.ones((3, 6), dtype=np.float32)
    assert e._host_pointer("wrench", w32)[17] == 1.0 and "wrench" not in e._pointers and e._converted["wrench"].dtype == np.float64
