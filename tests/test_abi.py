"""CPU: the C-ABI library loads and exports every symbol include/slam2d_b200.h declares (no compute calls)."""
import os
import re

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "slam2d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(slam_[a-z_0-9]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    import slam_2d_lidar_scan_b200 as S
    nat = S._native
    decl = declared_symbols()
    assert decl, "no declarations parsed"
    assert decl == set(nat.SYMBOLS), (decl ^ set(nat.SYMBOLS))
    for name in decl:
        assert hasattr(nat.lib, name), name
    assert nat.lib.slam_version() == 100


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "slam-2d-lidar-scan_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle's", "") or fn == "README.md", (dirpath, fn)


def test_no_cuda_means_loud_failure():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import slam_2d_lidar_scan_b200 as S
    with pytest.raises(RuntimeError):
        S.OccupancyGrid(10, 10, {"x": 0.0, "y": 0.0}, 0.1, 3.14159, 180, 10, 0.5)
