"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/svs_b200.h declares; without a GPU the entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "svs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svs_\w+)\s*\(", src)))


def test_header_symbols_are_exported(svs):
    L = svs.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and the python binding table covers the header
    assert set(names) <= set(svs.EXPORTS) | {"svs_last_error"}, set(names) - set(svs.EXPORTS)


def test_no_cpu_fallback(svs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(svs.SvsError):
        svs.BundleAdjuster()
    with pytest.raises(svs.SvsError):
        svs.FastGrid(640, 480, 222, 74, 25, 3, 3)
    buf = ctypes.create_string_buffer(64)
    assert svs.lib().svs_device_info(buf, 64) == -5   # SVS_ERR_NOGPU


def test_product_does_not_import_the_oracle():
    """scavislam_b200/ must never import, include or link anything under oracle/ (the checker is
    not the product); comments may of course mention it."""
    pkg = os.path.join(ROOT, "scavislam_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|#\s*include\s*[\"<][^\">]*oracle|liboracle|pyoracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), os.path.join(dirpath, f)
