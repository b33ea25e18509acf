"""The C-ABI library loads, exports every symbol include/plume_b200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "plume_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plume_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import plume_b200
    if not os.path.exists(plume_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(plume_b200.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libplume_b200.so does not export " + n
    assert set(names) == set(plume_b200.SYMBOLS), "python binding and header disagree"
    assert plume_b200.load().plume_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import plume_b200
    with pytest.raises(plume_b200.PlumeError):
        plume_b200.PlumeContext(0)
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "zk-nullifier-sig_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "c_oracle" not in txt and "plume_ref" not in txt and "plume_oracle" not in txt, f


def test_sm100a_sass_present():
    """The shipped library carries sm_100a SASS whose multiplier is IMAD.WIDE.U32 carry chains."""
    import shutil
    import subprocess
    import plume_b200
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", plume_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_self_test_vectors_are_the_golden_ones():
    """The constants plume_self_test compares against (csrc/selftest.cu) are the reference's vectors of
    tests/golden/reference_vectors.json, so a typo there cannot hide behind a GPU-only test."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")))
    src = open(os.path.join(ROOT, "zk-nullifier-sig_b200", "csrc", "selftest.cu")).read()
    body = src[src.index("struct Kat {"):src.index("\n};", src.index("struct Kat {"))]
    fields = {}
    for name, val in re.findall(r"const char\* (\w+)(?:\[2\])? = (.*?);", body, flags=re.S):
        fields[name] = ["".join(re.findall(r'"([^"]*)"', part)) for part in (val.strip("{} \n").split(",") if val.lstrip().startswith("{") else [val])]
    kat, mid = g["sign_kat"], g["intermediates"]
    xy = lambda p: p["x"] + p["y"]
    assert fields["msg"] == [kat["message_ascii"]]
    assert fields["sk"] == [kat["sk"]["hex"]] and fields["r"] == [kat["r"]["hex"]]
    assert fields["c"] == [kat["v1_c"]["hex"], kat["v2_c"]["hex"]]
    assert fields["s"] == [kat["v1_s"]["hex"], kat["v2_s"]["hex"]]
    assert fields["pk"] == [xy(mid["pk"])] and fields["g_r"] == [xy(mid["g_r"])]
    assert fields["h"] == [xy(mid["h"])] and fields["h_r"] == [xy(mid["h_r"])] and fields["h_sk"] == [xy(mid["h_sk"])]
    assert fields["h2c_abc"] == [xy(g["h2c_abc"])]
