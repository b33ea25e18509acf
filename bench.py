#!/usr/bin/env python3
"""bench.py -- PLUME sigs+verifies/sec (batch, bit-exact) on N B200s, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sign_verify|sign|verify|h2c|config4|sec1]
                    [--log2-batch B] [--impl ours|reference] [--single-process] [--no-extra-configs]
    torchrun ... bench.py --gpus N ...          (one rank per GPU; N > 1)

One "step" = one pass of the hot path over one batch of synthetic input per GPU.  Default workload
(BASELINE.json configs[1] + configs[2]): PLUME V1 sign of 2^20 items (32-byte messages, random sk, r)
followed by V1 verify of those 2^20 signatures, i.e. 2^21 operations per step per GPU; the metric is
operations (signatures + verifications) per second, whole job.

Numbers on the JSON line:
  value     device-resident: inputs already in HBM, `_device` C-ABI calls on one stream, CUDA events.
  e2e       same work through the host-pointer C-ABI calls with pinned host buffers: H2D of the inputs
            and D2H of every output inside the timed region (wall clock between synchronisations).
  e2e_pageable  the same calls on ordinary (pageable) host memory, staged by the library's copy threads.
  roofline  dominant kernel (variable-base scalar multiplications) against the integer-ALU peak measured
            on the spot by plume_measure_imad_peak (IMAD.WIDE.U32 issue rate); SURVEY.md 8(d).
  cpu_baseline  the C oracle ("port" of the rust-k256 path; cargo is not available) on the host cores,
            bounded sample.
  configs   the other BASELINE.json configs measured in the same run, each with its own value / e2e / roofline /
            checks: "config4" (configs[3]: V2 sign + verify, 2^21 items per GPU, whenever N > 1) and "h2c"
            (configs[4]: hash_to_curve-only on 2^26 65-byte preimages, at N = 1).
  checks    every rank compares a strided sample of ALL output bytes with the C oracle (bit-exact), every status is OK,
            every signature verifies, and a tamper set (one flipped bit in 1/1024 of the items) is
            rejected item for item; the flags are AND-reduced over the ranks.
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


# algorithmic bytes per item of the big kernels: what each must move through HBM given the stage split -- its inputs and outputs
# (32-byte scalars / coordinates, 64-byte points, 32-byte workspace slots) plus, ONCE, the per-item table that travels from a
# table kernel to its ladder kernel through HBM (16 or 2 x 8 entries of 128 bytes = 2 KiB).  Tables a kernel builds and walks
# itself (verify_mul_a's) and re-reads of entries are not algorithmic; the measured traffic (profiles/traffic.json) exceeds
# these figures by 2-7 x because ~110 000 resident threads x 2-4 KiB of table do not fit the 126 MB L2.
BYTES = {"sign_tab": 3 * 32 + 2048 + 2 * 32, "sign_varbase": 2048 + 2 * 32 + 2 * 32 + 6 * 32,
         "sign_tab+sign_varbase": 3 * 32 + 2 * 2048 + 2 * 32 + 6 * 32,
         "verify_mul_a": 64 + 2 * 32 + 12 * 64 + 3 * 32, "verify_mul_b": 2048 + 2 * 32 + 2 * 32 + 3 * 32,
         "verify_tab_b": 3 * 32 + 64 + 2048 + 2 * 32, "h2c_map": 65 + 3 * 32, "sign_h2c": 32 + 6 * 32 + 3 * 32, "verify_h2c": 32 + 64 * 4 + 64 + 3 * 32}


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


FIELDS = ("pk", "nullifier", "c", "s", "r_point", "hashed_to_curve_r")
WIDTH = {"pk": 64, "nullifier": 64, "c": 32, "s": 32, "r_point": 64, "hashed_to_curve_r": 64}
METRIC = "PLUME sigs+verifies/sec (batch, bit-exact)"
DTYPE = "u32 limbs (256-bit modular integer), bit-exact"


def synth_h2c_inputs(seed, first, count):
    """65-byte hash_to_curve preimages m_i || 02 || x_i (BASELINE configs[4]).  Items below 2^20 are SHA-256 derived like
    every other input (m_i and the sk stream standing in for x_i); beyond that the 2^20-item block repeats with the block
    number (i >> 20, u64 big-endian) XORed into the first eight bytes, so that 2^26 distinct preimages cost seconds, not
    minutes, of host time.  hash_to_curve accepts any byte string, so nothing depends on x_i being a real coordinate."""
    out = np.empty((count, 65), dtype=np.uint8)
    blk = 1 << 20
    base_cache = {}
    pos = 0
    while pos < count:
        i = first + pos
        b, o = i >> 20, i & (blk - 1)
        take = min(count - pos, blk - o)
        key = (o, take)
        if key not in base_cache:
            m, sk, _ = synth_inputs(seed, o, take)
            base_cache.clear()
            base_cache[key] = np.concatenate([m, np.full((take, 1), 2, np.uint8), sk], axis=1)
        seg = out[pos:pos + take]
        seg[:] = base_cache[key]
        if b:
            seg[:, :8] ^= np.frombuffer(int(b).to_bytes(8, "big"), dtype=np.uint8)
        pos += take
    return out


def make_cfg(workload, lg, world):
    version = 2 if workload == "config4" else 1
    return {"workload": {"sign_verify": "BASELINE configs[1]+[2]: batch 2^%d PLUME V1 sign then V1 verify of the same batch" % lg,
                         "sign": "BASELINE configs[1]: batch 2^%d PLUME V1 sign" % lg,
                         "verify": "BASELINE configs[2]: batch 2^%d PLUME V1 verify" % lg,
                         "config4": "BASELINE configs[3]: batch 2^%d per GPU PLUME V2 sign+verify, range-split (2^24 in all at 8 GPUs)" % lg,
                         "h2c": "BASELINE configs[4]: hash_to_curve-only, 2^%d 65-byte preimages per GPU" % lg,
                         "sec1": "SURVEY 8f-2: batch 2^%d PLUME V1 sign then verify on SEC1-compressed (33-byte) points, host API only" % lg}[workload],
            "version": "V%d" % version, "items_per_gpu_per_step": 1 << lg, "msg_bytes": 65 if workload == "h2c" else 32,
            "parallelism": "range-split x%d, no data-path collective" % world,
            "l2": "working set per step (inputs+outputs+workspace > 500 MB) exceeds the 126 MB L2; no explicit flush"}


def cpu_threads_for(world):
    return max(1, host_threads() // max(1, world))


class Env:
    """What every leg of the run shares: torch, the device, the context, the process group."""

    def __init__(self, args):
        import torch
        import plume_b200
        self.torch, self.plume = torch, plume_b200
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device visible; the PLUME kernels have no CPU fallback")
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        self.ctx = plume_b200.PlumeContext(self.local_rank)
        self.chunk = self.ctx.chunk_items
        self.stream = torch.cuda.Stream(device=self.dev)
        self.args = args

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def close(self):
        self.ctx.close()
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def ptr(t, off_items=0):
    if isinstance(t, np.ndarray):
        return t.ctypes.data + off_items * (t.shape[1] if t.ndim > 1 else 1)
    return t.data_ptr() + off_items * (t.shape[1] if t.dim() > 1 else 1)


def run_leg(env, workload, lg, steps, warmup, with_cpu_baseline=False, with_pageable=False, clocks=False):
    """One workload on this rank's shard: device-resident timing, end-to-end timing, checks, roofline.  Returns the record
    (rank 0 fills the rank-0-only parts)."""
    torch, ctx, dev, dist, rank, world = env.torch, env.ctx, env.dev, env.dist, env.rank, env.world
    version = 2 if workload == "config4" else 1
    n = 1 << lg
    seed = {"sign": 2, "verify": 3, "config4": 4, "h2c": 5}.get(workload, 2)
    ops_per_item = 2 if workload in ("sign_verify", "config4", "sec1") else 1
    cfg = make_cfg(workload, lg, world)
    chunk = env.chunk
    first, last = env.plume.shard_range(n * world, rank, world)   # this rank's contiguous range of the global batch
    assert last - first == n
    pinned = lambda a: torch.from_numpy(a).pin_memory()
    H, D = {}, {}
    if workload == "h2c":
        pre_h = synth_h2c_inputs(seed, first, n)
        H["pre"] = pinned(pre_h); H["h"] = torch.empty((n, 64), dtype=torch.uint8).pin_memory()
        D["pre"] = H["pre"].to(dev); D["h"] = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    else:
        msgs_h, sk_h, r_h = synth_inputs(seed, first, n)
        H = {"msgs": pinned(msgs_h), "sk": pinned(sk_h), "r": pinned(r_h)}
        for k in FIELDS:
            H[k] = torch.empty((n, WIDTH[k]), dtype=torch.uint8).pin_memory()
        H["status"] = torch.empty(n, dtype=torch.uint8).pin_memory()
        H["ok"] = torch.empty(n, dtype=torch.uint8).pin_memory()
        D = {k: H[k].to(dev) for k in ("msgs", "sk", "r")}
        for k in FIELDS + ("status", "ok"):
            D[k] = torch.empty_like(H[k], device=dev)
    if workload == "sec1":   # 33-byte slots next to the 64-byte buffers
        for k in ("pk", "nullifier", "r_point", "hashed_to_curve_r"):
            H[k + "33"] = torch.empty((n, 33), dtype=torch.uint8).pin_memory()
            D[k + "33"] = torch.empty((n, 33), dtype=torch.uint8, device=dev)
            D[k + "ok"] = torch.empty(n, dtype=torch.uint8, device=dev)
    sp = env.stream.cuda_stream
    do_sign = workload in ("sign", "sign_verify", "config4", "verify", "sec1")
    do_verify = workload in ("verify", "sign_verify", "config4", "sec1")
    PTS = ("pk", "nullifier", "r_point", "hashed_to_curve_r")

    def verify_device(B, ok, i0, cn):
        ctx.verify_batch_device(version, cn, ptr(B["msgs"], i0), 0, 32, ptr(B["pk"], i0), ptr(B["nullifier"], i0), ptr(B["c"], i0),
                                ptr(B["s"], i0), ptr(B["r_point"], i0), ptr(B["hashed_to_curve_r"], i0), ptr(ok, i0), sp)

    def step_device(sign=True, verify=True):
        for i0 in range(0, n, chunk):
            cn = min(chunk, n - i0)
            if workload == "h2c":
                ctx.hash_to_curve_batch_device(cn, ptr(D["pre"], i0), 0, 65, ptr(D["h"], i0), sp)
                continue
            if sign:
                ctx.sign_batch_device(version, cn, ptr(D["msgs"], i0), 0, 32, ptr(D["sk"], i0), ptr(D["r"], i0), ptr(D["pk"], i0),
                                      ptr(D["nullifier"], i0), ptr(D["c"], i0), ptr(D["s"], i0), ptr(D["r_point"], i0),
                                      ptr(D["hashed_to_curve_r"], i0), ptr(D["status"], i0), sp)
            if workload == "sec1":   # sign -> compress the four points -> decompress them -> verify
                for k in PTS:
                    ctx.points_compress_device(cn, ptr(D[k], i0), ptr(D[k + "33"], i0), sp)
                for k in PTS:
                    ctx.points_decompress_device(cn, ptr(D[k + "33"], i0), ptr(D[k], i0), ptr(D[k + "ok"], i0), sp)
            if verify:
                verify_device(D, D["ok"], i0, cn)

    def step_host(B, sign=True, verify=True):
        if workload == "h2c":
            ctx.hash_to_curve_batch_ptr(n, ptr(B["pre"]), 0, 65, ptr(B["h"]))
            return
        if workload == "sec1":
            ctx.sign_batch_sec1_ptr(version, n, ptr(B["msgs"]), 0, 32, ptr(B["sk"]), ptr(B["r"]), ptr(B["pk33"]), ptr(B["nullifier33"]),
                                    ptr(B["c"]), ptr(B["s"]), ptr(B["r_point33"]), ptr(B["hashed_to_curve_r33"]), ptr(B["status"]))
            ctx.verify_batch_sec1_ptr(version, n, ptr(B["msgs"]), 0, 32, ptr(B["pk33"]), ptr(B["nullifier33"]), ptr(B["c"]), ptr(B["s"]),
                                      ptr(B["r_point33"]), ptr(B["hashed_to_curve_r33"]), ptr(B["ok"]))
            return
        if sign:
            ctx.sign_batch_ptr(version, n, ptr(B["msgs"]), 0, 32, ptr(B["sk"]), ptr(B["r"]), ptr(B["pk"]), ptr(B["nullifier"]),
                               ptr(B["c"]), ptr(B["s"]), ptr(B["r_point"]), ptr(B["hashed_to_curve_r"]), ptr(B["status"]))
        if verify:
            ctx.verify_batch_ptr(version, n, ptr(B["msgs"]), 0, 32, ptr(B["pk"]), ptr(B["nullifier"]), ptr(B["c"]), ptr(B["s"]),
                                 ptr(B["r_point"]), ptr(B["hashed_to_curve_r"]), ptr(B["ok"]))

    timed_sign = workload != "verify"      # "verify" times the verifier only; signing prepares its input
    if workload == "verify":
        step_device(sign=True, verify=False)
        torch.cuda.synchronize()
        step_host(H, sign=True, verify=False)
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(warmup):
        step_device(sign=timed_sign, verify=do_verify)
    env.barrier()
    sampler = ClockSampler(env.local_rank) if (clocks and rank == 0) else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(env.stream)
    for _ in range(steps):
        step_device(sign=timed_sign, verify=do_verify)
    e1.record(env.stream)
    env.barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    # per-stage times: the same K steps once more with an event pair around every stage launch (plume_ctx_set_profiling).
    # With the pairs on, the library runs each batch as ONE kernel sequence (no half-batch overlap), so a stage's time is the
    # time of its kernel running alone on the GPU -- what the roofline of that kernel is about.
    ctx.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(env.stream)
    for _ in range(steps):
        step_device(sign=timed_sign, verify=do_verify)
    p1.record(env.stream)
    env.barrier()
    prof_ms = p0.elapsed_time(p1)
    stage = {}
    for st in ("sign_fixed", "sign_h2c", "sign_tab", "sign_varbase", "sign_final", "verify_h2c", "verify_tab_a", "verify_mul_a", "verify_tab_b", "verify_mul_b",
               "verify_final", "h2c_map", "h2c_out", "binv", "sec1_compress", "sec1_decompress"):
        ms, k = ctx.stage_ms(st)
        if k:
            stage[st] = {"ms_total": round(ms, 3), "launches": k}
    ctx.set_profiling(False)

    # ---- e2e: host pointers (pinned), H2D + D2H inside the timed region, wall clock between syncs
    def time_host(B):
        for _ in range(max(1, warmup // 2)):
            step_host(B, sign=timed_sign, verify=do_verify)
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_host(B, sign=timed_sign, verify=do_verify)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        env.barrier()
        return dt

    e2e_s = time_host(H)
    page_s = None
    if with_pageable and workload != "h2c":   # the same buffers as ordinary numpy memory (what a Rust Vec<u8> is)
        P = {k: np.array(v.numpy(), copy=True) for k, v in H.items()}
        page_s = time_host(P)
        for k in FIELDS + ("status", "ok"):
            if k in P and not np.array_equal(P[k], H[k].numpy()):
                page_s = -1.0   # flagged below
    clk = sampler.stop() if sampler else None

    red = [dev_ms, e2e_s * 1e3] + ([page_s * 1e3] if page_s is not None else [])
    red = env.plume.reduce_max(red, dist, dev)   # max over ranks
    dev_ms, e2e_ms = red[0], red[1]
    ops_step_gpu = n * ops_per_item
    value = world * ops_step_gpu * steps / (dev_ms * 1e-3)
    e2e_val = world * ops_step_gpu * steps / (e2e_ms * 1e-3)

    # bytes crossing PCIe per step per GPU (counted from the buffers handed to the host API)
    if workload == "h2c":
        h2d, d2h = n * 65, n * 64
    else:
        h2d = d2h = 0
        pt = 33 if workload == "sec1" else 64
        if timed_sign:
            h2d += n * 96; d2h += n * (pt * 4 + 64 + 1)
        if do_verify:
            h2d += n * (32 + pt * 4 + 64); d2h += n

    # ---- correctness inside the run, on every rank
    import c_oracle
    checks = {}
    thr = cpu_threads_for(world)
    m_sample = min(n, max(4096, (1 << 16) // world))
    idx = np.arange(0, n, max(1, n // m_sample))[:m_sample]
    if workload == "h2c":
        checks["h2c_device_vs_host_api"] = bool(torch.equal(D["h"].cpu(), H["h"]))
        want = c_oracle.h2c_batch(H["pre"].numpy()[idx], threads=thr)
        checks["bit_exact_vs_oracle_strided_%d" % len(idx)] = bool(np.array_equal(H["h"].numpy()[idx], want))
    else:
        if do_sign:
            checks["sign_status_ok"] = int((D["status"] == 0).sum().item()) == n and int((H["status"] == 0).sum().item()) == n
        if do_verify:
            checks["verify_all_true"] = int(D["ok"].sum().item()) == n and int(H["ok"].sum().item()) == n
        for k in (("c", "s", "pk33", "nullifier33", "r_point33", "hashed_to_curve_r33") if workload == "sec1" else FIELDS):
            if not torch.equal(D[k].cpu(), H[k]):
                checks["device_vs_host_api_" + k] = False
        if page_s is not None:
            checks["pageable_equals_pinned"] = page_s > 0
        if workload != "sec1":
            # strided sample of every output byte against the oracle (all ranks, whatever N)
            want = c_oracle.sign_batch(version, H["msgs"].numpy()[idx], H["sk"].numpy()[idx], H["r"].numpy()[idx], threads=thr)
            same = all(np.array_equal(H[k].numpy()[idx], want[k]) for k in FIELDS + ("status",))
            checks["bit_exact_vs_oracle_strided_%d" % len(idx)] = bool(same)
            if do_verify:
                # tamper set: one flipped bit in 1/1024 of the items, every field in turn; exactly those fail
                fields = FIELDS if version == 1 else ("pk", "nullifier", "c", "s")
                T = {k: D[k].clone() for k in FIELDS}
                T["msgs"] = D["msgs"]
                tam = np.arange(7, n, 1024)
                for j, f in enumerate(fields):
                    rows = tam[j::len(fields)]
                    col = (rows * 13 + j) % WIDTH[f]
                    bit = 1 << ((rows // 7) % 8)
                    r_t = torch.from_numpy(rows).to(dev)
                    T[f][r_t, torch.from_numpy(col).to(dev)] ^= torch.from_numpy(bit.astype(np.uint8)).to(dev)
                ok_t = torch.empty(n, dtype=torch.uint8, device=dev)
                for i0 in range(0, n, chunk):
                    verify_device(T, ok_t, i0, min(chunk, n - i0))
                torch.cuda.synchronize()
                expect = np.ones(n, dtype=np.uint8)
                expect[tam] = 0
                checks["tamper_set_%d_rejected_exactly" % len(tam)] = bool(np.array_equal(ok_t.cpu().numpy(), expect))
    if dist is not None:
        checks["all_ranks"] = env.plume.all_ranks_true(all(checks.values()), dist, dev)
        checks["shards_tile_batch"] = sum(env.plume.gather_counts(n, dist, dev)) == n * world

    rec = {"metric": METRIC, "value": value, "unit": "ops/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": DTYPE, "data": "synthetic", "config": cfg,
           "e2e": {"value": e2e_val, "unit": "ops/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                   "ms_per_step": e2e_ms / steps},
           "gpu_launches": launches, "checks": checks}
    if page_s is not None:
        rec["e2e_pageable"] = {"value": world * ops_step_gpu * steps / (red[2] * 1e-3), "unit": "ops/s", "ms_per_step": red[2] / steps,
                               "note": "same calls on pageable numpy memory, staged by the library's copy threads (PLUME_STAGE_THREADS, default 8)"}
    if rank != 0:
        return rec

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
    if dom in ("sign_varbase", "sign_tab") and "sign_tab" in stage and "sign_varbase" in stage:
        # the signer's h^r, h^sk is two kernels (comb table, comb ladders) with ONE algorithmic count: they are rated together
        avg_ms = sum(stage[k]["ms_total"] / stage[k]["launches"] for k in ("sign_tab", "sign_varbase"))
        dom, kind = "sign_tab+sign_varbase", "sign_varbase"
    ach = per_launch_items * work_lp(kind) / (avg_ms * 1e-3)
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full summary
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        parts = ("sign_tab", "sign_varbase") if dom.startswith("sign_tab+") else (dom,)
        if all(k in tj for k in parts):   # per item, scaled to this launch (the kernel's traffic is proportional to the items)
            traffic = sum(tj[k]["dram_bytes_per_launch"] * per_launch_items / tj[k]["items_per_launch"] for k in parts)
    except Exception:
        pass
    rec["roofline"] = {"bound": "int-alu", "kernel": dom, "achieved": ach, "peak": peak_lp, "unit": "limb-products/s",
                       "frac": ach / peak_lp,
                       "frac_survey_units": per_launch_items * work_lp(kind, True) / (avg_ms * 1e-3) / peak_lp,
                       "traffic": traffic,
                       "peak_source": "measured in this run: faster of two independent-chain microbenchmarks, plain "
                                      "IMAD.WIDE.U32 columns (%.3e LP/s) and carry-chain IMAD.WIDE.U32.X rows (%.3e LP/s); "
                                      "MEASURED_PEAKS.json carries no INT32 figure; SASS and ncu evidence of the two kernels: "
                                      "profiles/r02_imad_peak.md" % (plain_lp, carry_lp),
                       "algorithmic_work": "%d multiplications x %d + %d squarings x %d limb products per item "
                                           "(SURVEY.md 8d counts, squarings at 44 instead of 72), %d items per launch"
                                           % (WORK_MS[kind][0], LP_PER_M, WORK_MS[kind][1], LP_PER_S, per_launch_items),
                       "avg_launch_ms": avg_ms}
    # the same fraction for every kernel with an algorithmic count (the two kernels of the signer's h^r, h^sk together)
    per_kernel = {}
    for st_name, st_kind in (("sign_varbase", "sign_varbase"), ("verify_mul_a", "verify_mul_a"), ("verify_mul_b", "verify_mul_b"),
                             ("verify_tab_b", "verify_tab_b"), ("sign_h2c", "h2c_map"), ("verify_h2c", "h2c_map"), ("h2c_map", "h2c_map")):
        if st_name in stage:
            t = stage[st_name]["ms_total"] / stage[st_name]["launches"]
            label = st_name
            if st_name == "sign_varbase" and "sign_tab" in stage:
                t += stage["sign_tab"]["ms_total"] / stage["sign_tab"]["launches"]
                label = "sign_tab+sign_varbase"
            if st_name == "verify_mul_a" and "verify_tab_a" in stage:
                t += stage["verify_tab_a"]["ms_total"] / stage["verify_tab_a"]["launches"]
                label = "verify_tab_a+verify_mul_a"
            per_kernel[label] = {"ms": round(t, 3), "frac": round(per_launch_items * work_lp(st_kind) / (t * 1e-3) / peak_lp, 4)}
    rec["roofline"]["per_kernel"] = per_kernel
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = per_launch_items * BYTES.get(dom, 160)
    gbs = alg_bytes / (avg_ms * 1e-3) / 1e9
    rec["roofline"]["hbm"] = {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                              "algorithmic_bytes": alg_bytes, "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                              "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
    # whole-step view: all kernels of the step against the same peak
    kinds = {"sign_verify": ("sign", "verify"), "config4": ("sign", "verify"), "sign": ("sign",), "verify": ("verify",),
             "h2c": ("h2c",), "sec1": ("sign", "verify") + ("inversion",) * 4}[workload]   # sec1: + four square roots
    for key, su in (("whole_step_frac", False), ("whole_step_frac_survey_units", True)):
        rec["roofline"][key] = (n * sum(work_lp(k, su) for k in kinds) * steps / (dev_ms * 1e-3)) / peak_lp
    rec["stages"] = stage
    rec["stages_note"] = ("per-stage CUDA-event times of a second pass of the same %d steps with an event pair around every stage "
                          "launch; in that pass each batch is one kernel sequence (%.2f ms per step), in the timed region above large "
                          "device-resident batches run as two half-batches on two streams so that one's kernels fill the other's "
                          "partly filled last waves" % (steps, prof_ms / steps))
    if clk is not None:
        rec["clocks"] = clk
    # cpu baseline, bounded sample, rank 0 only at N = 1
    if with_cpu_baseline and world == 1:
        threads = host_threads()
        if workload == "h2c":
            m0 = s0 = r0 = None
            pre = H["pre"].numpy()
            cal = min(n, 64 * threads)
            t0 = time.perf_counter(); c_oracle.h2c_batch(pre[:cal], threads=threads); dt = time.perf_counter() - t0
            sample = int(max(cal, min(n, 10.0 * cal / max(dt, 1e-3))))
            t0 = time.perf_counter(); out = c_oracle.h2c_batch(pre[:sample], threads=threads); dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": sample / dt, "unit": "ops/s", "cores": threads, "kind": "port",
                                   "sample": "first %d preimages of this workload, C restatement of the rust-k256 path "
                                             "(cargo/rustc unavailable), %d pthreads, %.1f s" % (sample, threads, dt)}
            rec["checks"]["bit_exact_vs_oracle_first_%d" % sample] = bool(np.array_equal(H["h"].numpy()[:sample], out))
        else:
            msgs_h, sk_h, r_h = H["msgs"].numpy(), H["sk"].numpy(), H["r"].numpy()
            if env.args.cpu_sample:
                sample = env.args.cpu_sample
            else:   # calibrate on a small slice, then size the sample for ~15 s of wall time on all host threads
                cal = min(n, 16 * threads)
                ops, dt, _ = cpu_run(workload, version, msgs_h[:cal], sk_h[:cal], r_h[:cal], threads)
                sample = int(max(cal, min(n, 15.0 * cal / max(dt, 1e-3))))
            ops, dt, out = cpu_run(workload, version, msgs_h[:sample], sk_h[:sample], r_h[:sample], threads)
            rec["cpu_baseline"] = {"value": ops / dt, "unit": "ops/s", "cores": threads, "kind": "port",
                                   "sample": "first %d items of this workload (%d ops), C restatement of the rust-k256 path "
                                             "(cargo/rustc unavailable), %d pthreads, %.1f s" % (sample, ops, threads, dt)}
            if out is not None and do_sign and workload != "sec1":
                same = all(np.array_equal(H[k][:sample].numpy(), out[k]) for k in FIELDS + ("status",))
                rec["checks"]["bit_exact_vs_oracle_first_%d" % sample] = bool(same)
    return rec


def reference_arm(args, workload, lg):
    """--impl reference: the CPU path alone (the C port of the rust-k256 path: cargo/rustc are not in the image), all host
    threads, each step a bounded sample of the workload of at least 2^16 items."""
    version = 2 if workload == "config4" else 1
    seed = {"sign": 2, "verify": 3, "config4": 4, "h2c": 5}.get(workload, 2)
    n = 1 << lg
    threads = host_threads()
    if args.cpu_sample:
        sample = args.cpu_sample
    else:   # calibrate; ~4 s per step but never fewer than 2^16 items
        cal = 16 * threads
        m0, s0, r0 = synth_inputs(seed, 0, cal)
        _, dt, _ = cpu_run(workload, version, m0, s0, r0, threads)
        sample = int(max(min(n, 1 << 16), min(n, 4.0 * cal / max(dt, 1e-3))))
    msgs, sk, r = synth_inputs(seed, 0, sample)
    times, ops = [], 0
    for it in range(args.warmup + args.steps):
        ops, dt, _ = cpu_run(workload, version, msgs, sk, r, threads)
        if it >= args.warmup:
            times.append(dt)
    val = ops * len(times) / sum(times)
    cfg = make_cfg(workload, lg, int(os.environ.get("WORLD_SIZE", "1")))   # our arm's config; the sample is in cpu_baseline
    return {"metric": METRIC, "value": val, "unit": "ops/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": cfg, "impl": "reference",
            "cpu_baseline": {"value": val, "unit": "ops/s", "cores": threads, "kind": "port",
                             "sample": "%d items per step (%d ops) of the same synthetic workload; C restatement of the "
                                       "rust-k256 path (cargo/rustc unavailable), %d pthreads" % (sample, ops, threads)},
            "e2e": {"value": val, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def single_process(args):
    """--single-process --gpus N: ONE process, ONE pinned host batch, N GPUs behind one multi-device context
    (plume_ctx_create_multi): strong scaling of the host-pointer API, reported next to the same batch on one GPU."""
    import torch
    import plume_b200
    import c_oracle
    lg = args.log2_batch if args.log2_batch is not None else 23
    n = 1 << lg
    G = args.gpus
    # items beyond 2^20 repeat the SHA-derived block with the block number XORed into the message (distinct messages, hence
    # distinct h, nullifiers and signatures, at seconds instead of minutes of single-process input synthesis)
    blk = min(n, 1 << 20)
    m0, s0, r0 = synth_inputs(2, 0, blk)
    msgs_h, sk_h, r_h = np.tile(m0, (n // blk, 1)), np.tile(s0, (n // blk, 1)), np.tile(r0, (n // blk, 1))
    for b in range(1, n // blk):
        msgs_h[b * blk:(b + 1) * blk, :8] ^= np.frombuffer(b.to_bytes(8, "big"), dtype=np.uint8)
    pinned = lambda a: torch.from_numpy(a).pin_memory()
    H = {"msgs": pinned(msgs_h), "sk": pinned(sk_h), "r": pinned(r_h)}
    for k in FIELDS:
        H[k] = torch.empty((n, WIDTH[k]), dtype=torch.uint8).pin_memory()
    H["status"] = torch.empty(n, dtype=torch.uint8).pin_memory()
    H["ok"] = torch.empty(n, dtype=torch.uint8).pin_memory()

    def step(ctx):
        ctx.sign_batch_ptr(1, n, ptr(H["msgs"]), 0, 32, ptr(H["sk"]), ptr(H["r"]), ptr(H["pk"]), ptr(H["nullifier"]), ptr(H["c"]),
                           ptr(H["s"]), ptr(H["r_point"]), ptr(H["hashed_to_curve_r"]), ptr(H["status"]))
        ctx.verify_batch_ptr(1, n, ptr(H["msgs"]), 0, 32, ptr(H["pk"]), ptr(H["nullifier"]), ptr(H["c"]), ptr(H["s"]),
                             ptr(H["r_point"]), ptr(H["hashed_to_curve_r"]), ptr(H["ok"]))

    res = {}
    sampler = ClockSampler(0)
    for label, devs in (("1gpu", [0]), ("%dgpu" % G, list(range(G)))):
        t0 = time.perf_counter()
        ctx = plume_b200.PlumeContext(devs)
        t_create = time.perf_counter() - t0
        for _ in range(max(1, args.warmup // 2)):
            step(ctx)
        if label != "1gpu":
            sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(ctx)
        dt = time.perf_counter() - t0
        res[label] = {"ops_per_s": 2 * n * args.steps / dt, "ms_per_step": 1e3 * dt / args.steps, "ctx_create_s": round(t_create, 3),
                      "gpu_launches": ctx.launch_count}
        ctx.close()
    clk = sampler.stop()
    m = 1 << 16
    idx = np.arange(0, n, n // m)[:m]
    want = c_oracle.sign_batch(1, msgs_h[idx], sk_h[idx], r_h[idx], threads=host_threads())
    checks = {"sign_status_ok": bool((H["status"] == 0).all().item()), "verify_all_true": bool((H["ok"] == 1).all().item()),
              "bit_exact_vs_oracle_strided_%d" % m: all(np.array_equal(H[k].numpy()[idx], want[k]) for k in FIELDS + ("status",))}
    big = res["%dgpu" % G]
    line = {"metric": METRIC, "value": big["ops_per_s"], "unit": "ops/s", "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": big["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic",
            "config": {"workload": "single process, one pinned host batch of 2^%d items, PLUME V1 sign then verify through the "
                                   "host-pointer C ABI of a %d-device context (plume_ctx_create_multi)" % (lg, G),
                       "parallelism": "range-split x%d inside the library, one worker thread per GPU, no data-path collective" % G},
            "e2e": {"value": big["ops_per_s"], "unit": "ops/s", "h2d_bytes_per_step": n * (96 + 32 + 64 * 4 + 64),
                    "d2h_bytes_per_step": n * (64 * 4 + 64 + 1 + 1)},
            "single_process": res, "speedup_vs_1gpu": big["ops_per_s"] / res["1gpu"]["ops_per_s"], "gpu_launches": big["gpu_launches"],
            "checks": checks, "clocks": clk,
            "gtab_broadcast": os.environ.get("PLUME_GTAB_BCAST", "none (every device builds its own table concurrently)")}
    print(json.dumps(line))
    return 0


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
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the configs[3] / configs[4] sub-records")
    ap.add_argument("--h2c-log2", type=int, default=26, help="size of the hash_to_curve-only sub-record at N = 1")
    ap.add_argument("--single-process", action="store_true", help="one process, --gpus N devices behind one multi-device context")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "sec1" and args.impl == "reference":
        raise SystemExit("--workload sec1 has no reference arm (the CPU path is timed on the 64-byte form)")
    lg = args.log2_batch if args.log2_batch is not None else {"config4": 21, "h2c": 22}.get(args.workload, 20)

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps(reference_arm(args, args.workload, lg)))
        return 0
    if args.single_process:
        return single_process(args)

    env = Env(args)
    line = run_leg(env, args.workload, lg, args.steps, args.warmup, with_cpu_baseline=not args.no_cpu_baseline,
                   with_pageable=True, clocks=True)
    # the other BASELINE configs in the same run (default workload only)
    if args.workload == "sign_verify" and not args.no_extra_configs and args.log2_batch is None:
        extra = {}
        sub_steps = max(2, min(args.steps, 3))
        if env.world > 1:     # configs[3]: V2 sign + verify, 2^21 items per GPU (2^24 over 8 GPUs)
            extra["config4"] = run_leg(env, "config4", 21, sub_steps, 3)
        else:                 # configs[4]: hash_to_curve-only, 2^26 preimages on one GPU
            extra["h2c"] = run_leg(env, "h2c", args.h2c_log2, sub_steps, 3, with_cpu_baseline=not args.no_cpu_baseline)
        if env.rank == 0:
            line["configs"] = extra
            for k, r in extra.items():
                line["checks"]["configs_" + k] = all(r["checks"].values())
    if env.rank == 0:
        print(json.dumps(line))
    env.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
