"""CPU: the C-ABI library loads and exports every symbol include/cagroup3d_b200.h declares; host-only helpers
answer without a GPU.  (No compute entry point is called here.)"""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(lib):
    from cagroup3d_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (cg3d_\w+)", out))
    assert set(protos) <= exported, sorted(set(protos) - exported)
    # nothing undeclared leaks out of the library either
    assert exported <= set(protos), sorted(exported - set(protos))
    for name, types in protos.items():
        fn = getattr(lib, name)
        assert fn.restype is ctypes.c_int and list(fn.argtypes) == types


def test_every_stream_entry_point_ends_with_stream():
    from cagroup3d_b200 import _lib
    _lib.parse_header()
    assert _lib._host_only == {"cg3d_hash_capacity", "cg3d_scan_workspace_ints", "cg3d_sort_workspace_ints",
                               "cg3d_spconv_tc_ntile",
                               "cg3d_segment_mean_workspace", "cg3d_spconv_tc_splitk", "cg3d_spconv_wgrad_slabs",
                               "cg3d_bn_train_workspace", "cg3d_focal_loss_workspace", "cg3d_loss_workspace", "cg3d_knn_grid_workspace"}


def test_host_only_helpers(lib):
    from cagroup3d_b200 import _lib
    assert _lib.hash_capacity(0) == 1024 and _lib.hash_capacity(513) == 2048 and _lib.hash_capacity(100000) == 262144
    for n in (0, 1, 1000, 10 ** 6):
        assert _lib.scan_workspace_ints(n) >= 1
        assert _lib.sort_workspace_ints(n) >= 1
    assert [_lib.host("cg3d_spconv_tc_ntile", c) for c in (64, 128, 192, 256, 512, 18)] == [64, 128, 64, 256, 256, 0]
    assert 2 * _lib.host("cg3d_segment_mean_workspace", 1000, 1) >= 1000 + 6          # n points in a single segment


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "cagroup3d_b200.h")).read()
    for needle in ("iou3d_nms.cpp", "cagroup_head.py", "cagroup_roi_head.py", "biresnet.py", "cagroup3d.py"):
        assert needle in src


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cagroup3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    """the oracle and the C-ABI emulator (tests/cabi_emulator.py) are test infrastructure: nothing under cagroup3d_b200/,
    pcdet/ or tools/ may import them."""
    bad = []
    for top in ("cagroup3d_b200", "pcdet", "tools"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    s = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+(oracle|tests)\b", s, flags=re.M) or "tests.golden" in s or "cabi_emulator" in s:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_splitk_factor_selection(lib):
    """cg3d_spconv_tc_splitk: only launches with fewer CTAs than the 296 slots and a long K loop are split; grouped
    launches, 1x1 convs and well-filled launches never are (host-only helper, no GPU needed)."""
    from cagroup3d_b200 import _lib
    sk = lambda *a: _lib.host("cg3d_spconv_tc_splitk", *a)
    assert sk(3192, 128, 128, 343, 0, 0) > 1                   # 7^3 RoI pooling contraction: 50 CTAs x 1372 stages
    assert sk(2719, 512, 512, 27, 0, 0) > 1                    # stride-32 bottleneck conv
    assert sk(153503, 128, 128, 27, 0, 0) == 1                 # 1200 tiles: well filled
    assert sk(42500, 256, 256, 27, 0, 0) == 1                  # 1.1 waves: not split (see spconv_tc.cu)
    assert sk(2719, 1024, 128, 1, 0, 0) == 1                   # K = 1: nothing to split
    assert sk(3192, 128, 128, 343, 1, 25) == 1                 # grouped launches are never split
    for n in (1, 100, 5000):
        k = sk(n, 64, 128, 125, 0, 0)
        assert 1 <= k <= 8 and (125 * 2) // k >= 24            # >= 24 stages left per CTA
