"""Multi-GPU development probe (run under `gpurun --gpus N`): time plume_ctx_create_multi with the generator table built on
every device (default), peer-copied from device 0 (PLUME_GTAB_BCAST=p2p) and broadcast with NCCL (=nccl), and check that a
small batch signs identically under each.  Prints one JSON line per mode."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "zk-nullifier-sig_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch   # loads the bundled libnccl.so.2 into the process (the nccl mode finds it through the loader)
import plume_b200
import c_oracle

G = torch.cuda.device_count()
rng = np.random.default_rng(1)
n = 4096
msgs = rng.integers(0, 256, (n, 32), dtype=np.uint8)
sk = rng.integers(0, 256, (n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
r = rng.integers(0, 256, (n, 32), dtype=np.uint8); r[:, 0] &= 0x7F
want = c_oracle.sign_batch(1, msgs, sk, r, threads=os.cpu_count() or 1)
try:
    import nvidia.nccl
    os.environ.setdefault("PLUME_NCCL_LIB", os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2"))
except Exception:
    pass
plume_b200.PlumeContext(0).close()   # CUDA context creation and module load out of the timings
for mode in ("none", "p2p", "nccl", "none"):
    if mode == "none":
        os.environ.pop("PLUME_GTAB_BCAST", None)
    else:
        os.environ["PLUME_GTAB_BCAST"] = mode
    t0 = time.perf_counter()
    try:
        ctx = plume_b200.PlumeContext(list(range(G)))
    except plume_b200.PlumeError as e:
        print(json.dumps({"mode": mode, "gpus": G, "error": str(e)}))
        continue
    dt = time.perf_counter() - t0
    got = ctx.sign_batch(1, msgs, sk, r)
    same = all(np.array_equal(got[k], want[k]) for k in want)
    ctx.close()
    print(json.dumps({"mode": mode, "gpus": G, "ctx_create_s": round(dt, 4), "bit_exact": bool(same)}))
