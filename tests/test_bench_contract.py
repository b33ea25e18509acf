"""bench.py's JSON-line contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`, the CPU
path alone) prints one line with the keys the driver reads, on our arm's metric / unit / config; input synthesis follows
SURVEY.md 8(d) and is reproducible by index range; the algorithmic work table reproduces the survey's totals."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "64"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "ops/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("PLUME sigs+verifies/sec") and "configs[1]+[2]" in d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_input_synthesis_rule():
    sys.path.insert(0, ROOT)
    import bench
    msgs, sk, r = bench.synth_inputs(2, 5, 4)
    S = (2).to_bytes(8, "big")
    for j in range(4):
        i = (5 + j).to_bytes(8, "big")
        assert bytes(msgs[j]) == hashlib.sha256(b"plume-b200/m" + S + i).digest()
        assert bytes(sk[j]) == hashlib.sha256(b"plume-b200/sk" + S + i + bytes(8)).digest()
        assert bytes(r[j]) == hashlib.sha256(b"plume-b200/r" + S + i + bytes(8)).digest()
    a = bench.synth_inputs(2, 0, 9)
    b = bench.synth_inputs(2, 3, 6)
    for x, y in zip(a, b):
        assert np.array_equal(x[3:], y)                       # a rank's range equals the same range of the global batch
    pre = bench.synth_h2c_inputs(5, 0, 3)
    m, s, _ = bench.synth_inputs(5, 0, 3)
    assert pre.shape == (3, 65) and np.array_equal(pre[:, :32], m) and (pre[:, 32] == 2).all() and np.array_equal(pre[:, 33:], s)


def test_algorithmic_work_table():
    sys.path.insert(0, ROOT)
    import bench
    tot = lambda k: sum(bench.WORK_MS[k])
    assert tot("sign_varbase") == 2838 and tot("verify_mul_a") == 1766 and tot("verify_mul_b") + tot("verify_tab_b") == 2214
    assert tot("sign") == 4478 and tot("verify") == 4916 and tot("h2c") == 906
    assert bench.work_lp("sign_varbase") == 1270 * 72 + 1568 * 44 and bench.work_lp("sign_varbase", True) == 2838 * 72
