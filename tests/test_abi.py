"""The C-ABI library loads and exports every symbol include/rs_knn.h declares (no compute)."""
import ctypes as C
import re
from pathlib import Path

import pytest

import recommend_sys_b200 as rs

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "rs_knn.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = rs.core.knn_lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(rs.core.ABI_SYMBOLS) == names


def test_params_default_matches_reference_defaults():
    p = rs.core.RsKnnParams()
    assert rs.core.knn_lib().rs_knn_params_default(C.byref(p)) == 0
    assert p.sim == rs.core.RS_SIM["msd"]       # core/knn.go:145
    assert p.k == 40 and p.min_k == 1           # core/knn.go:80-81
    assert p.knn_type == 0 and p.device == -1 and p.row_begin == 0 and p.row_end == 0


def test_struct_sizes_match_header():
    # 10 int32 + 2 int64 + 1 double + 2 int32
    assert C.sizeof(rs.core.RsKnnParams) == 10 * 4 + 2 * 8 + 8 + 2 * 4
    assert C.sizeof(rs.core.RsKnnProfile) == 3 * 8 + 3 * 8 + 2 * 4 + 8


def test_no_cpu_fallback_without_device():
    L = rs.core.knn_lib()
    if L.rs_knn_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rs.core.RsError) as e:
        rs.core._Handle()
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(rs.core.RsError):
        rs.Cosine(rs.NewSortedIdRatings([(1, 4)]), rs.NewSortedIdRatings([(1, 5)]))


def test_product_never_imports_oracle():
    pkg = ROOT / "recommend-sys_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cc")) + list(pkg.rglob("*.cuh")) \
            + list(pkg.rglob("*.hpp")) + list(pkg.rglob("*.h")):
        text = f.read_text()
        assert "knn_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
