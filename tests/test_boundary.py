"""CPU: the C-ABI library loads, exports every symbol include/vpf.h declares, and the
product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vipformer_b200 import build

    so = build.build()
    lib = ctypes.CDLL(so)
    syms = _declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in vpf.h but not exported: {missing}"
    lib.vpf_abi_version.restype = ctypes.c_int
    assert lib.vpf_abi_version() == 1


def test_argument_errors_do_not_need_a_gpu():
    from vipformer_b200 import _lib

    lib = _lib.lib()
    rc = lib.vpf_fps(None, 1, 10, 3, 4, None, None, None)
    assert rc == -1
    assert b"null" in lib.vpf_last_error_string()
    with pytest.raises(_lib.VpfError):
        _lib.call("vpf_knn_point", ctypes.c_int(64), None, 1, 128, 3, None, 4, 3, None, None)


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "vipformer_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libvpf_oracle|include\s*[<\"].*oracle", txt, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from vipformer_b200 import _lib
    from vipformer_b200.preproc import divide_patches

    with pytest.raises(_lib.VpfError):
        divide_patches(torch.randn(1, 64, 3), 8, 4)


def test_every_exported_symbol_is_declared_in_the_header():
    """The reverse of the export check: nothing reachable in the .so is missing from include/vpf.h."""
    import re
    import subprocess

    from vipformer_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True, check=True).stdout
    syms = sorted({l.split()[-1] for l in out.splitlines() if re.search(r"\bT vpf_", l)})
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "vpf.h")).read()
    assert len(syms) > 50
    missing = [s for s in syms if not re.search(r"\b" + s + r"\s*\(", hdr)]
    assert not missing, missing
