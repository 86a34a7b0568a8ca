"""The C-ABI library loads on a CPU-only box and exports every symbol include/lakeside_b200.h declares; compute entry
points fail loudly (LK_ERR_CUDA) instead of falling back when no device is visible."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lakeside_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lakeside_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.lk_version().decode().startswith("lakeside_b200")


def test_no_cpu_fallback_without_device():
    import torch

    from lakeside_b200 import _lib, api

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert api.device_count() == 0
    with pytest.raises(api.LakesideError) as e:
        api.init()
    assert e.value.code == _lib.LK_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(api.LakesideError) as e:
        api.merge_streams_index([__import__("numpy").arange(4)])
    assert e.value.code == _lib.LK_ERR_CUDA


def test_segment_cache_bookkeeping_needs_no_device():
    # capacity / statistics of the HBM-resident segment cache are host-side bookkeeping: configurable before any GPU is touched
    from lakeside_b200 import api

    st = api.cache_stats()
    try:
        api.cache_configure(123456)
        assert api.cache_stats()["capacity_bytes"] == 123456
        api.cache_configure(0)
        api.cache_clear()
        s0 = api.cache_stats()
        assert s0["capacity_bytes"] == 0 and s0["resident_bytes"] == 0 and s0["segments"] == 0
        with pytest.raises(api.LakesideError):
            api.cache_configure(-2)
    finally:
        api.cache_configure(st["capacity_bytes"] if st["capacity_bytes"] > 0 else -1)  # (-1: the default, resolved by lk_init)


def test_file_segments_without_a_device_fail_loudly(tmp_path):
    import torch

    from lakeside_b200 import _lib, api

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    f = tmp_path / "x.parquet"
    f.write_bytes(b"PAR1" + b"\0" * 32 + b"PAR1")
    with pytest.raises(api.LakesideError) as e:
        api.eval_glob('{"baseExpr": {"dataset": "logs", "filter": {"k": "a", "v": ["b"], "op": "eq"}, "chart": {"aggregation": "sum"}}, '
                      '"segmentRequests": [{"stepInMillis": 1, "startTs": 0, "endTs": 1}]}', [str(f)])
    assert e.value.code == _lib.LK_ERR_CUDA


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lakeside_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "lakeside_oracle" not in txt and "oracle/" not in txt.replace("lakeside_oracle", ""), f
                assert "lk_emul" not in txt, f
