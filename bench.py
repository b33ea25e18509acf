#!/usr/bin/env python3
"""bench.py -- PLUME sigs+verifies/sec (batch, bit-exact) on N B200s, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sign_verify|sign|verify|h2c|config4]
                    [--log2-batch B] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU; N > 1)

One "step" = one pass of the hot path over one batch of synthetic input per GPU.  Default workload
(BASELINE.json configs[1] + configs[2]): PLUME V1 sign of 2^20 items (32-byte messages, random sk, r)
followed by V1 verify of those 2^20 signatures, i.e. 2^21 operations per step per GPU; the metric is
operations (signatures + verifications) per second, whole job.

Numbers on the JSON line:
  value     device-resident: inputs already in HBM, `_device` C-ABI calls on one stream, CUDA events.
  e2e       same work through the host-pointer C-ABI calls with pinned host buffers: H2D of the inputs
            and D2H of every output inside the timed region (wall clock between synchronisations).
  roofline  dominant kernel (variable-base scalar multiplications) against the integer-ALU peak measured
            on the spot by plume_measure_imad_peak (IMAD.WIDE.U32 issue rate); SURVEY.md 8(d).
  cpu_baseline  the C oracle ("port" of the rust-k256 path; cargo is not available) on the host cores,
            bounded sample.
Inputs follow SURVEY.md 8(d): m_i = SHA256("plume-b200/m" || S || i), sk_i / r_i by SHA-256 counter
with rejection into [1, n-1]; S and i are u64 big-endian.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "zk-nullifier-sig_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
# Algorithmic work per item, SURVEY.md 8(d): the canonical algorithm (GLV + width-5 wNAF Jacobian, a = 0 doubling 2M + 5S,
# mixed addition 8M + 3S, 8-bit fixed window for G, RFC 9380 straight-line SSWU, batched inversion) counted in field
# multiplications M and squarings S.  A multiplication is 64 + 8 = 72 limb products (8 x 8 schoolbook + fold).  SURVEY 8d
# charges a squaring the same 72; a squaring needs only 36 + 8 = 44, and with 72 the squaring-dominated hash_to_curve kernel
# would sit at 1.4 x the measured IMAD.WIDE peak.  `frac` therefore counts 44 per squaring and
# `frac_survey_units` keeps the 72-for-both figure next to it.
LP_PER_M, LP_PER_S = 72, 44
WORK_MS = {                                # (M, S); M + S reproduces the totals of SURVEY 8(d)
    "h2c_map": (98, 538),                  # 2 x SSWU (254 S + 12 M exponentiation, + 20 M + 9 S) + 2 x isogeny + addition  = 636
    "inversion": (15, 255),                # 270
    "fixed_pair": (512, 192),              # g^r, g^sk: 2 x 32 mixed additions                                             = 704
    "sign_varbase": (1270, 1568),          # h^r, h^sk sharing one table: 2 x (128 dbl + 43 add) + table                    = 2838
    "verify_mul_a": (886, 880),            # G*s - pk*c: 128 dbl + 70 add + table                                           = 1766
    "verify_mul_b": (1030, 984),           # h*s - nul*c, the ladder: 128 dbl + 86 add                                      = 2014
    "verify_tab_b": (140, 60),             #              its two window tables                                              = 200
    "affine_out": (300, 0),                # batched conversion of 3-5 output points
}
WORK_MS["verify_muls"] = tuple(a + b + c for a, b, c in zip(WORK_MS["verify_mul_a"], WORK_MS["verify_mul_b"], WORK_MS["verify_tab_b"]))
WORK_MS["sign"] = tuple(sum(WORK_MS[k][i] for k in ("h2c_map", "sign_varbase", "fixed_pair", "affine_out")) for i in (0, 1))   # 4478
WORK_MS["verify"] = tuple(sum(WORK_MS[k][i] for k in ("h2c_map", "verify_muls", "affine_out")) for i in (0, 1))               # 4916
WORK_MS["h2c"] = tuple(sum(WORK_MS[k][i] for k in ("h2c_map", "inversion")) for i in (0, 1))                                   # 906
for _k in ("sign_h2c", "verify_h2c"):
    WORK_MS[_k] = WORK_MS["h2c_map"]
WORK_MS["sign_fixed"] = WORK_MS["fixed_pair"]


def work_lp(kind, survey_units=False):
    m, sq = WORK_MS[kind]
    return (m + sq) * LP_PER_M if survey_units else m * LP_PER_M + sq * LP_PER_S


# algorithmic bytes per item of the dominant kernels (what they must read + write in HBM)
BYTES = {"sign_varbase": 3 * 32 + 64 + 2 * 32 + 6 * 32, "verify_muls": 3 * 32 + 64 * 2 + 64 + 2 * 32 + 6 * 32,
         "verify_mul_a": 64 + 2 * 32 + 3 * 32, "verify_mul_b": 2 * 32 + 32 + 3 * 32, "verify_tab_b": 3 * 32 + 64 + 3 * 32}


_K256 = np.array([
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2], dtype=np.uint32)
_H256 = (0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19)


def _sha256_short(prefix, idx, suffix=b""):
    """SHA-256(prefix || u64_be(i) || suffix) for every i of the u64 array idx (one 64-byte block each), vectorised
    over the batch with numpy; returns u8[len(idx), 32].  Input synthesis only (bit-for-bit hashlib, checked below)."""
    n = idx.shape[0]
    mlen = len(prefix) + 8 + len(suffix)
    assert mlen <= 55
    blk = np.zeros((n, 64), dtype=np.uint8)
    blk[:, :len(prefix)] = np.frombuffer(prefix, dtype=np.uint8)
    blk[:, len(prefix):len(prefix) + 8] = idx.astype(">u8").view(np.uint8).reshape(n, 8)
    if suffix:
        blk[:, len(prefix) + 8:mlen] = np.frombuffer(suffix, dtype=np.uint8)
    blk[:, mlen] = 0x80
    blk[:, 62] = (mlen * 8) >> 8
    blk[:, 63] = (mlen * 8) & 0xFF
    w = [blk.view(">u4")[:, t].astype(np.uint32) for t in range(16)]
    rotr = lambda x, k: (x >> np.uint32(k)) | (x << np.uint32(32 - k))
    for t in range(16, 64):
        s0 = rotr(w[t - 15], 7) ^ rotr(w[t - 15], 18) ^ (w[t - 15] >> np.uint32(3))
        s1 = rotr(w[t - 2], 17) ^ rotr(w[t - 2], 19) ^ (w[t - 2] >> np.uint32(10))
        w.append(w[t - 16] + s0 + w[t - 7] + s1)
    a, b, c, d, e, f, g, h = (np.full(n, v, dtype=np.uint32) for v in _H256)
    for t in range(64):
        t1 = h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + _K256[t] + w[t]
        t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))
        h, g, f, e, d, c, b, a = g, f, e, d + t1, c, b, a, t1 + t2
    out = np.empty((n, 8), dtype=">u4")
    for k, (v, iv) in enumerate(zip((a, b, c, d, e, f, g, h), _H256)):
        out[:, k] = v + np.uint32(iv)
    return out.view(np.uint8).reshape(n, 32)


def synth_inputs(seed, first, count):
    """(msgs u8[count,32], sk u8[count,32], r u8[count,32]) for global item indices first..first+count."""
    S = seed.to_bytes(8, "big")
    pm, ps, pr = b"plume-b200/m" + S, b"plume-b200/sk" + S, b"plume-b200/r" + S
    msgs = np.empty((count, 32), dtype=np.uint8)
    sk = np.empty((count, 32), dtype=np.uint8)
    r = np.empty((count, 32), dtype=np.uint8)
    sha = hashlib.sha256
    order_top = ORDER >> 128
    with np.errstate(over="ignore"):
        for c0 in range(0, count, 1 << 20):
            c1 = min(count, c0 + (1 << 20))
            idx = np.arange(first + c0, first + c1, dtype=np.uint64)
            msgs[c0:c1] = _sha256_short(pm, idx)
            for dst, pre in ((sk, ps), (r, pr)):
                d = _sha256_short(pre, idx, bytes(8))           # ctr = 0
                dst[c0:c1] = d
                # rejection into [1, n-1], mirrors SecretKey::random: redo (scalar code) the ~2^-128 that fall outside
                top = d[:, :16].view(">u8")
                suspect = np.nonzero((top[:, 0] >= np.uint64(order_top >> 64)) | ((top[:, 0] | top[:, 1]) == 0))[0]
                for j in suspect:
                    ib, ctr = int(idx[j]).to_bytes(8, "big"), 0
                    while True:
                        v = sha(pre + ib + ctr.to_bytes(8, "big")).digest()
                        if 1 <= int.from_bytes(v, "big") < ORDER:
                            break
                        ctr += 1
                    dst[c0 + j] = np.frombuffer(v, dtype=np.uint8)
    return msgs, sk, r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_run(workload, version, msgs, sk, r, threads):
    """One pass of the workload through the C oracle; returns (ops, seconds, outputs)."""
    import c_oracle
    n = msgs.shape[0]
    t0 = time.perf_counter()
    ops = 0
    out = None
    if workload == "h2c":
        pre = np.concatenate([msgs, np.full((n, 1), 2, np.uint8), sk], axis=1)   # 65-byte preimages m || 02 || x
        c_oracle.h2c_batch(pre, threads=threads)
        ops = n
    else:
        out = c_oracle.sign_batch(version, msgs, sk, r, threads=threads)
        if workload in ("sign", "sign_verify", "config4"):
            ops += n
        if workload in ("verify", "sign_verify", "config4"):
            t1 = time.perf_counter()
            if workload == "verify":
                t0 = t1    # signing only prepared the verifier's input
            c_oracle.verify_batch(version, msgs, out["pk"], out["nullifier"], out["c"], out["s"], out["r_point"],
                                  out["hashed_to_curve_r"], threads=threads)
            ops += n
    return ops, time.perf_counter() - t0, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sign_verify", choices=["sign_verify", "sign", "verify", "h2c", "config4", "sec1"])
    ap.add_argument("--log2-batch", type=int, default=None, help="items per GPU per step = 2^B")
    ap.add_argument("--cpu-sample", type=int, default=0, help="items of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    version = 2 if args.workload == "config4" else 1
    if args.workload == "sec1" and args.impl == "reference":
        raise SystemExit("--workload sec1 has no reference arm (the CPU path is timed on the 64-byte form)")
    lg = args.log2_batch if args.log2_batch is not None else {"config4": 21, "h2c": 22}.get(args.workload, 20)
    n = 1 << lg
    seed = {"sign": 2, "verify": 3, "config4": 4, "h2c": 5}.get(args.workload, 2)
    ops_per_item = 2 if args.workload in ("sign_verify", "config4", "sec1") else 1
    metric = "PLUME sigs+verifies/sec (batch, bit-exact)"
    cfg = {"workload": {"sign_verify": "BASELINE configs[1]+[2]: batch 2^%d PLUME V1 sign then V1 verify of the same batch" % lg,
                        "sign": "BASELINE configs[1]: batch 2^%d PLUME V1 sign" % lg,
                        "verify": "BASELINE configs[2]: batch 2^%d PLUME V1 verify" % lg,
                        "config4": "BASELINE configs[3]: batch 2^%d per GPU PLUME V2 sign+verify, range-split" % lg,
                        "h2c": "BASELINE configs[4]: hash_to_curve-only, 2^%d 65-byte preimages" % lg,
                        "sec1": "SURVEY 8f-2: batch 2^%d PLUME V1 sign then verify on SEC1-compressed (33-byte) points, host API only" % lg}[args.workload],
           "version": "V%d" % version, "items_per_gpu_per_step": n, "msg_bytes": 32,
           "parallelism": "range-split x%d, no data-path collective" % world,
           "l2": "working set per step (inputs+outputs+workspace > 500 MB) exceeds the 126 MB L2; no explicit flush"}

    # ------------------------------------------------------------------ reference arm: the CPU path
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = host_threads()
        if args.cpu_sample:
            sample = args.cpu_sample
        else:   # calibrate, then size each step for ~4 s so that warmup + steps stay within a couple of minutes
            cal = 16 * threads
            m0, s0, r0 = synth_inputs(seed, 0, cal)
            _, dt, _ = cpu_run(args.workload, version, m0, s0, r0, threads)
            sample = int(max(cal, min(n, 4.0 * cal / max(dt, 1e-3))))
        msgs, sk, r = synth_inputs(seed, 0, sample)
        times, ops = [], 0
        for it in range(args.warmup + args.steps):
            ops, dt, _ = cpu_run(args.workload, version, msgs, sk, r, threads)
            if it >= args.warmup:
                times.append(dt)
        val = ops * len(times) / sum(times)
        line = {"metric": metric, "value": val, "unit": "ops/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 limbs (256-bit modular integer), bit-exact", "data": "synthetic", "config": cfg, "impl": "reference",
                "cpu_baseline": {"value": val, "unit": "ops/s", "cores": threads, "kind": "port",
                                 "sample": "%d items per step (%d ops) of the same synthetic workload; C restatement of the "
                                           "rust-k256 path (cargo/rustc unavailable), %d pthreads" % (sample, ops, threads)},
                "e2e": {"value": val, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import plume_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the PLUME kernels have no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ctx = plume_b200.PlumeContext(local_rank)
    chunk = ctx.chunk_items

    first, last = plume_b200.shard_range(n * world, rank, world)   # this rank's contiguous range of the global batch
    assert last - first == n
    msgs_h, sk_h, r_h = synth_inputs(seed, first, n)

    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t
    # pinned host buffers (e2e arm)
    H, D = {}, {}
    if args.workload != "h2c":   # (the hash_to_curve-only workload has its own two buffers below)
        H = {"msgs": pinned(msgs_h), "sk": pinned(sk_h), "r": pinned(r_h)}
        for k, w in (("pk", 64), ("nullifier", 64), ("c", 32), ("s", 32), ("r_point", 64), ("hashed_to_curve_r", 64)):
            H[k] = torch.empty((n, w), dtype=torch.uint8).pin_memory()
        H["status"] = torch.empty(n, dtype=torch.uint8).pin_memory()
        H["ok"] = torch.empty(n, dtype=torch.uint8).pin_memory()
        # device-resident buffers (value arm)
        D = {k: H[k].to(dev) for k in ("msgs", "sk", "r")}
        for k in ("pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r", "status", "ok"):
            D[k] = torch.empty_like(H[k], device=dev)
    if args.workload == "sec1":   # 33-byte slots next to the 64-byte buffers
        for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r"):
            H[k + "33"] = torch.empty((n, 33), dtype=torch.uint8).pin_memory()
            D[k + "33"] = torch.empty((n, 33), dtype=torch.uint8, device=dev)
            D[k + "ok"] = torch.empty(n, dtype=torch.uint8, device=dev)
    if args.workload == "h2c":
        pre_h = np.ascontiguousarray(np.concatenate([msgs_h, np.full((n, 1), 2, np.uint8), sk_h], axis=1))
        H["pre"] = pinned(pre_h); H["h"] = torch.empty((n, 64), dtype=torch.uint8).pin_memory()
        D["pre"] = H["pre"].to(dev); D["h"] = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    do_sign = args.workload in ("sign", "sign_verify", "config4", "verify", "sec1")
    do_verify = args.workload in ("verify", "sign_verify", "config4", "sec1")
    PTS = ("pk", "nullifier", "r_point", "hashed_to_curve_r")

    def ptr(t, off_items=0, width=None):
        return t.data_ptr() + off_items * (width if width is not None else (t.shape[1] if t.dim() > 1 else 1))

    def step_device(sign=True, verify=True):
        for i0 in range(0, n, chunk):
            cn = min(chunk, n - i0)
            if args.workload == "h2c":
                ctx.hash_to_curve_batch_device(cn, ptr(D["pre"], i0), 0, 65, ptr(D["h"], i0), sp)
                continue
            if sign:
                ctx.sign_batch_device(version, cn, ptr(D["msgs"], i0), 0, 32, ptr(D["sk"], i0), ptr(D["r"], i0), ptr(D["pk"], i0),
                                      ptr(D["nullifier"], i0), ptr(D["c"], i0), ptr(D["s"], i0), ptr(D["r_point"], i0),
                                      ptr(D["hashed_to_curve_r"], i0), ptr(D["status"], i0), sp)
            if args.workload == "sec1":   # sign -> compress the four points -> decompress them -> verify
                for k in PTS:
                    ctx.points_compress_device(cn, ptr(D[k], i0), ptr(D[k + "33"], i0), sp)
                for k in PTS:
                    ctx.points_decompress_device(cn, ptr(D[k + "33"], i0), ptr(D[k], i0), ptr(D[k + "ok"], i0), sp)
            if verify:
                ctx.verify_batch_device(version, cn, ptr(D["msgs"], i0), 0, 32, ptr(D["pk"], i0), ptr(D["nullifier"], i0),
                                        ptr(D["c"], i0), ptr(D["s"], i0), ptr(D["r_point"], i0), ptr(D["hashed_to_curve_r"], i0),
                                        ptr(D["ok"], i0), sp)

    def step_host(sign=True, verify=True):
        if args.workload == "h2c":
            ctx.hash_to_curve_batch_ptr(n, ptr(H["pre"]), 0, 65, ptr(H["h"]))
            return
        if args.workload == "sec1":
            ctx.sign_batch_sec1_ptr(version, n, ptr(H["msgs"]), 0, 32, ptr(H["sk"]), ptr(H["r"]), ptr(H["pk33"]), ptr(H["nullifier33"]),
                                    ptr(H["c"]), ptr(H["s"]), ptr(H["r_point33"]), ptr(H["hashed_to_curve_r33"]), ptr(H["status"]))
            ctx.verify_batch_sec1_ptr(version, n, ptr(H["msgs"]), 0, 32, ptr(H["pk33"]), ptr(H["nullifier33"]), ptr(H["c"]), ptr(H["s"]),
                                      ptr(H["r_point33"]), ptr(H["hashed_to_curve_r33"]), ptr(H["ok"]))
            return
        if sign:
            ctx.sign_batch_ptr(version, n, ptr(H["msgs"]), 0, 32, ptr(H["sk"]), ptr(H["r"]), ptr(H["pk"]), ptr(H["nullifier"]),
                               ptr(H["c"]), ptr(H["s"]), ptr(H["r_point"]), ptr(H["hashed_to_curve_r"]), ptr(H["status"]))
        if verify:
            ctx.verify_batch_ptr(version, n, ptr(H["msgs"]), 0, 32, ptr(H["pk"]), ptr(H["nullifier"]), ptr(H["c"]), ptr(H["s"]),
                                 ptr(H["r_point"]), ptr(H["hashed_to_curve_r"]), ptr(H["ok"]))

    timed_sign = args.workload != "verify"      # "verify" times the verifier only; signing prepares its input
    if args.workload == "verify":
        step_device(sign=True, verify=False); step_host(sign=True, verify=False)
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(args.warmup):
        step_device(sign=timed_sign, verify=do_verify)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.set_profiling(True)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device(sign=timed_sign, verify=do_verify)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    stage = {}
    for st in ("sign_fixed", "sign_h2c", "sign_varbase", "sign_final", "verify_h2c", "verify_muls", "verify_mul_a", "verify_tab_b", "verify_mul_b", "verify_final", "h2c_map",
               "h2c_out", "binv", "sec1_compress", "sec1_decompress"):
        ms, k = ctx.stage_ms(st)
        if k:
            stage[st] = {"ms_total": round(ms, 3), "launches": k}
    ctx.set_profiling(False)

    # ---- e2e: host pointers (pinned), H2D + D2H inside the timed region, wall clock between syncs
    for _ in range(max(1, args.warmup // 2)):
        step_host(sign=timed_sign, verify=do_verify)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host(sign=timed_sign, verify=do_verify)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    dev_ms, e2e_ms = plume_b200.reduce_max([dev_ms, e2e_s * 1e3], dist, dev)   # max over ranks
    ops_step_gpu = n * ops_per_item
    value = world * ops_step_gpu * args.steps / (dev_ms * 1e-3)
    e2e_val = world * ops_step_gpu * args.steps / (e2e_ms * 1e-3)

    # bytes crossing PCIe per step per GPU (counted from the buffers handed to the host API)
    if args.workload == "h2c":
        h2d, d2h = n * 65, n * 64
    else:
        h2d = d2h = 0
        pt = 33 if args.workload == "sec1" else 64
        if timed_sign:
            h2d += n * 96; d2h += n * (pt * 4 + 64 + 1)
        if do_verify:
            h2d += n * (32 + pt * 4 + 64); d2h += n

    # ---- correctness inside the run: every status OK, every signature verifies; strided bit-exact sample vs the oracle
    checks = {}
    if args.workload != "h2c":
        if do_sign:
            checks["sign_status_ok"] = int((D["status"] == 0).sum().item()) == n and int((H["status"] == 0).sum().item()) == n
        if do_verify:
            checks["verify_all_true"] = int(D["ok"].sum().item()) == n and int(H["ok"].sum().item()) == n
        for k in (("c", "s", "pk33", "nullifier33", "r_point33", "hashed_to_curve_r33") if args.workload == "sec1"
                  else ("pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r")):
            if not torch.equal(D[k].cpu(), H[k]):
                checks["device_vs_host_api_" + k] = False
    else:
        checks["h2c_device_vs_host_api"] = bool(torch.equal(D["h"].cpu(), H["h"]))
        k = min(n, 4096)   # and a sample against an independent SHA-256 + big-integer restatement is left to tests/ (oracle)
        checks["h2c_nonzero"] = bool((H["h"][:k] != 0).any(dim=1).all().item())
    if dist is not None:
        checks["all_ranks"] = plume_b200.all_ranks_true(all(checks.values()), dist, dev)
        checks["shards_tile_batch"] = sum(plume_b200.gather_counts(n, dist, dev)) == n * world

    line = {"metric": metric, "value": value, "unit": "ops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (256-bit modular integer), bit-exact", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_val, "unit": "ops/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "checks": checks}

    if rank == 0:
        # roofline of the dominant kernel against the integer-ALU peak measured here
        plain_lp, carry_lp = ctx.measure_imad_rates(4096)
        peak_lp = max(plain_lp, carry_lp)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        dom = max((s for s in stage if s != "binv"), key=lambda s: stage[s]["ms_total"])
        per_launch_items = min(chunk, n)
        avg_ms = stage[dom]["ms_total"] / stage[dom]["launches"]
        kind = dom if dom in WORK_MS else "h2c_map"
        ach = per_launch_items * work_lp(kind) / (avg_ms * 1e-3)
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full summary
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if dom in tj:   # per item, scaled to this launch (the kernel's traffic is proportional to the items)
                traffic = tj[dom]["dram_bytes_per_launch"] * per_launch_items / tj[dom]["items_per_launch"]
        except Exception:
            pass
        line["roofline"] = {"bound": "int-alu", "kernel": dom, "achieved": ach, "peak": peak_lp, "unit": "limb-products/s",
                            "frac": ach / peak_lp,
                            "frac_survey_units": per_launch_items * work_lp(kind, True) / (avg_ms * 1e-3) / peak_lp,
                            "traffic": traffic,
                            "peak_source": "measured in this run: faster of two independent-chain microbenchmarks, plain "
                                           "IMAD.WIDE.U32 columns (%.3e LP/s) and carry-chain IMAD.WIDE.U32.X rows (%.3e LP/s); "
                                           "MEASURED_PEAKS.json carries no INT32 figure" % (plain_lp, carry_lp),
                            "algorithmic_work": "%d multiplications x %d + %d squarings x %d limb products per item "
                                                "(SURVEY.md 8d counts, squarings at 44 instead of 72), %d items per launch"
                                                % (WORK_MS[kind][0], LP_PER_M, WORK_MS[kind][1], LP_PER_S, per_launch_items),
                            "avg_launch_ms": avg_ms}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        gbs = per_launch_items * BYTES.get(dom, 160) / (avg_ms * 1e-3) / 1e9
        line["roofline"]["hbm"] = {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                                   "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
        # whole-step view: all kernels of the step against the same peak
        kinds = {"sign_verify": ("sign", "verify"), "config4": ("sign", "verify"), "sign": ("sign",), "verify": ("verify",),
                 "h2c": ("h2c",), "sec1": ("sign", "verify") + ("inversion",) * 4}[args.workload]   # sec1: + four square roots
        for key, su in (("whole_step_frac", False), ("whole_step_frac_survey_units", True)):
            line["roofline"][key] = (n * sum(work_lp(k, su) for k in kinds) * args.steps / (dev_ms * 1e-3)) / peak_lp
        line["stages"] = stage
        line["clocks"] = clocks
        # cpu baseline, bounded sample, rank 0 only at N = 1
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            if args.cpu_sample:
                sample = args.cpu_sample
            else:   # calibrate on a small slice, then size the sample for ~15 s of wall time on all host threads
                cal = min(n, 16 * threads)
                ops, dt, _ = cpu_run(args.workload, version, msgs_h[:cal], sk_h[:cal], r_h[:cal], threads)
                sample = int(max(cal, min(n, 15.0 * cal / max(dt, 1e-3))))
            ops, dt, out = cpu_run(args.workload, version, msgs_h[:sample], sk_h[:sample], r_h[:sample], threads)
            line["cpu_baseline"] = {"value": ops / dt, "unit": "ops/s", "cores": threads, "kind": "port",
                                    "sample": "first %d items of this workload (%d ops), C restatement of the rust-k256 path "
                                              "(cargo/rustc unavailable), %d pthreads, %.1f s" % (sample, ops, threads, dt)}
            if out is not None and do_sign and args.workload != "sec1":
                same = all(np.array_equal(H[k][:sample].numpy(), out[k]) for k in
                           ("pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r", "status"))
                line["checks"]["bit_exact_vs_oracle_first_%d" % sample] = bool(same)
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
