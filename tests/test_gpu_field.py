"""The field layer's carry-chain assembly, in isolation on the GPU (plume_debug_fe_op), against Python
integers: edge representatives (0, p, 2^256-1, values that trigger the rare wrap-around branches) and
random values, for every operation the kernels use."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 2**32 - 977


def _limbs(vals):
    return np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in vals], dtype=np.uint32)


def _ints(arr):
    return [sum(int(row[i]) << (32 * i) for i in range(8)) for row in arr]


def test_field_ops(gpu_ctx):
    rnd = random.Random(12)
    C = 2**32 + 977
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, P - 2, 2**255, C, C - 1, C + 1, 2**256 - C, 2**256 - C - 1, 0xFFFFFFFF,
            2**224 - 1, (2**256 - 1) ^ (2**128), (1 << 256) - (1 << 224), 2**64 - 1, 2**64, (2**256 - 1) ^ 0xFFFFFFFF, 977, 2**33,
            # shifted left by 1..3 these leave limb 1 (and limbs 2..7) all ones, so that folding the bits shifted out ripples
            (2**256 - 1) >> 1, (2**256 - 1) >> 2, (2**256 - 1) >> 3, (7 << 253) | ((2**253 - 1) ^ 0x1FFFFFFF), (3 << 254) | (2**254 - 2**29),
            (1 << 255) | (2**255 - 2**31), (7 << 253) | (2**253 - 2**29), (7 << 253) | (2**61 - 2**29)]
    A, B = [], []
    for x in edge:           # every edge value against every edge value
        for y in edge:
            A.append(x); B.append(y)
    for _ in range(20000):
        A.append(rnd.randrange(2**256)); B.append(rnd.randrange(2**256))
    # values engineered so that sums / differences land next to the wrap-around boundaries
    for _ in range(2000):
        x = rnd.randrange(2**256)
        for t in (2**256 - 1 - x, 2**256 - x, (2**256 - x + rnd.randrange(2 * C)) % 2**256, (x + rnd.randrange(2 * C)) % 2**256):
            A.append(x); B.append(t % 2**256)
    a, b = _limbs(A), _limbs(B)
    mul = _ints(gpu_ctx.debug_fe_op(0, a, b))
    add = _ints(gpu_ctx.debug_fe_op(2, a, b))
    sub = _ints(gpu_ctx.debug_fe_op(3, a, b))
    sqr = _ints(gpu_ctx.debug_fe_op(1, a, b))
    neg = _ints(gpu_ctx.debug_fe_op(8, a, b))
    nrm = _ints(gpu_ctx.debug_fe_op(5, a, b))
    isz = _ints(gpu_ctx.debug_fe_op(10, a, b))
    sh = {k: _ints(gpu_ctx.debug_fe_op(op, a, b)) for k, op in ((1, 12), (2, 13), (3, 11))}
    for i, (x, y) in enumerate(zip(A, B)):
        assert mul[i] % P == x * y % P, ("mul", hex(x), hex(y))
        assert add[i] % P == (x + y) % P, ("add", hex(x), hex(y))
        assert sub[i] % P == (x - y) % P, ("sub", hex(x), hex(y))
        assert sqr[i] % P == x * x % P, ("sqr", hex(x))
        assert neg[i] % P == (-x) % P, ("neg", hex(x))
        assert nrm[i] == x % P, ("norm", hex(x))
        assert isz[i] == (1 if x % P == 0 else 0)
        for k in (1, 2, 3):
            assert sh[k][i] % P == (x << k) % P, ("shl", k, hex(x))
    for k in (0, 1, 2, 3, 8, 11, 1771, 65535, 65536, 2**32 - 1):
        kb = _limbs([k] * len(A))
        ms = _ints(gpu_ctx.debug_fe_op(6, a, kb))
        for i, x in enumerate(A):
            assert ms[i] % P == x * k % P, ("mul_small", hex(x), k)
    m = 600
    inv = _ints(gpu_ctx.debug_fe_op(4, a[:m], b[:m]))
    p34 = _ints(gpu_ctx.debug_fe_op(7, a[:m], b[:m]))
    sq = _ints(gpu_ctx.debug_fe_op(9, a[:m], b[:m]))
    invv = _ints(gpu_ctx.debug_fe_op(14, a, b))     # the division-step inversion of the small-batch path, every value
    for i, x in enumerate(A):
        assert invv[i] % P == (pow(x, -1, P) if x % P else 0), ("inv_var", hex(x))
    for i, x in enumerate(A[:m]):
        if x % P:
            assert inv[i] % P == pow(x, -1, P)
        assert p34[i] % P == pow(x, (P - 3) // 4, P)
        assert sq[i] % P == pow(x, (P + 1) // 4, P)
