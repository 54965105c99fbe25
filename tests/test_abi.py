"""The C-ABI boundary: libdtmpc_b200.so loads, exports every symbol include/dtmpc.h declares, the ctypes mirrors of
the POD structs have the C layout, and argument validation fails cleanly before any CUDA work (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dtmpc.h")


@pytest.fixture(scope="module")
def lib():
    from differentiable_tube_mpc_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def _declared_functions():
    src = open(HEADER, encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dtmpc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    from differentiable_tube_mpc_b200 import _lib

    assert _declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    from differentiable_tube_mpc_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (dtmpc_[a-z0-9_]+)", out))
    missing = [f for f in _declared_functions() if f not in exported]
    assert not missing, missing
    assert lib.dtmpc_abi_version() == _lib.ABI_VERSION == 3


def test_library_is_sm100a_native():
    from differentiable_tube_mpc_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_struct_layouts_match_the_header(tmp_path):
    from differentiable_tube_mpc_b200 import _lib

    structs = {"dtmpc_env": _lib.Env, "dtmpc_ocp": _lib.Ocp, "dtmpc_solve_io": _lib.SolveIO, "dtmpc_sens_io": _lib.SensIO,
               "dtmpc_disturbance": _lib.Disturbance}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        cname, field, val = line.split()
        cls = structs[cname]
        if field == "size":
            assert C.sizeof(cls) == int(val), (cname, C.sizeof(cls), val)
        else:
            assert getattr(cls, field).offset == int(val), (cname, field)


def test_bad_arguments_are_rejected_before_any_launch(lib):
    from differentiable_tube_mpc_b200 import _lib

    env, ocp, io = _lib.Env(), _lib.Ocp(), _lib.SolveIO()
    # null description
    assert lib.dtmpc_ilqr_solve_f64(None, None, 4, None, None, 0, None) == _lib.E_BADARG
    env.dt = 0.01
    env.n_obs = 99  # more obstacles than DTMPC_MAX_OBS
    ocp.horizon, ocp.max_iter, ocp.n_alpha = 10, 5, 1
    assert lib.dtmpc_ilqr_solve_f64(C.byref(env), C.byref(ocp), 4, C.byref(io), None, 0, None) == _lib.E_BADARG
    env.n_obs = 0
    ocp.n_alpha = 9  # more step sizes than DTMPC_MAX_ALPHAS
    assert lib.dtmpc_ilqr_solve_f64(C.byref(env), C.byref(ocp), 4, C.byref(io), None, 0, None) == _lib.E_BADARG
    ocp.n_alpha = 1
    # valid description, missing buffers
    assert lib.dtmpc_ilqr_solve_f64(C.byref(env), C.byref(ocp), 4, C.byref(io), None, 0, None) == _lib.E_BADARG
    assert b"buffer" in lib.dtmpc_last_error()
    # empty batch is a no-op
    assert lib.dtmpc_ilqr_solve_f64(C.byref(env), C.byref(ocp), 0, C.byref(io), None, 0, None) == _lib.OK
    assert lib.dtmpc_sensitivity_f32(C.byref(env), C.byref(ocp), 3, None, None, 0, None) == _lib.E_BADARG
    assert lib.dtmpc_reduce_columns_f64(None, 8, 3, None, None, 0, None) == _lib.E_BADARG
    with pytest.raises(ValueError):
        _lib.check(_lib.E_BADARG, "x")
    with pytest.raises(MemoryError):
        _lib.check(_lib.E_WORKSPACE, "x")


def test_missing_library_fails_loudly(monkeypatch):
    from differentiable_tube_mpc_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdtmpc_b200.so")
    with pytest.raises(_lib.LibraryMissing):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or mention it."""
    pkg = os.path.join(ROOT, "differentiable_tube_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "dtmpc_oracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
