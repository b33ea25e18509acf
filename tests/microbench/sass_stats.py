#!/usr/bin/env python3
"""sass_stats.py -- static instruction statistics of one kernel from `cuobjdump -sass` (no GPU needed).

    python tests/microbench/sass_stats.py <object-or-so> <kernel-substring> [--loops] [--range LO HI]

Prints the opcode histogram of the kernel, the out-of-line callees (fe_mul / fe_sqr live at the end of every
kernel's text) with their own histograms, and -- with --loops -- every backward branch (loop) with the histogram
of its body and the number of calls it makes.  An issue-cost estimate uses the per-instruction sub-partition
cycles measured in profiles/r01_imad_rates.md (IMAD.WIDE 4.2, IMAD.HI 4.6, two-input IADD3/VIADD 1.1, everything
else on the integer pipes 2).  Development tool: it is how loop bodies were compared before spending GPU time.
"""
import collections
import re
import subprocess
import sys

COST = {"IMAD.WIDE.U32": 4.2, "IMAD.WIDE.U32.X": 4.2, "IMAD.WIDE": 4.2, "IMAD.HI.U32": 4.6, "VIADD": 1.1}


def cost(op):
    if op.startswith("IMAD.WIDE"):
        return 4.2
    if op.startswith("IMAD.HI"):
        return 4.6
    if op in ("VIADD",):
        return 1.1
    if op.startswith(("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "NOP")):
        return 2.0
    return 2.0


def parse(path, kernel):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and cur is not None:
            addr = int(m.group(1), 16)
            txt = m.group(2).strip()
            pred = None
            pm = re.match(r"(@!?U?P\d+)\s+(.*)", txt)
            if pm:
                pred, txt = pm.group(1), pm.group(2)
            op = txt.split()[0]
            funcs[cur].append((addr, op, txt, pred))
    names = [f for f in funcs if kernel in f]
    if not names:
        raise SystemExit("no kernel matching %r; have: %s" % (kernel, ", ".join(funcs)))
    return names[0], funcs[names[0]]


def hist(instrs):
    h = collections.Counter(op for _, op, _, _ in instrs)
    return h


def show(title, instrs, top=14):
    h = hist(instrs)
    n = sum(h.values())
    cyc = sum(cost(op) * k for op, k in h.items())
    wide = sum(k for op, k in h.items() if op.startswith("IMAD.WIDE"))
    print("%s: %d instructions, %d IMAD.WIDE, est. %.0f issue cycles" % (title, n, wide, cyc))
    print("   " + ", ".join("%s %d" % (op, k) for op, k in h.most_common(top)))
    return n, wide, cyc


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    name, ins = parse(path, kernel)
    print("kernel", name, "text bytes", (ins[-1][0] + 16) if ins else 0)
    # callees: targets of CALL.REL
    targets = sorted({int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) for _, op, t, _ in ins if op.startswith("CALL.REL")})
    # a callee runs from its entry to the RET before the next callee / end
    bounds = targets + [ins[-1][0] + 16]
    callee = {}
    for a, b in zip(bounds, bounds[1:]):
        body = [i for i in ins if a <= i[0] < b]
        callee[a] = body
    main_end = targets[0] if targets else ins[-1][0] + 16
    show("whole kernel (static)", ins)
    for a, body in callee.items():
        show("  callee @0x%x" % a, body)
    ccost = {a: sum(cost(op) for _, op, _, _ in b) for a, b in callee.items()}
    cn = {a: len(b) for a, b in callee.items()}
    cw = {a: sum(1 for _, op, _, _ in b if op.startswith("IMAD.WIDE")) for a, b in callee.items()}
    if "--range" in sys.argv:
        k = sys.argv.index("--range")
        lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        body = [i for i in ins if lo <= i[0] <= hi]
        report(body, "range 0x%x..0x%x" % (lo, hi), ccost, cn, cw)
    if "--loops" in sys.argv:
        loops = []
        for addr, op, txt, pred in ins:
            if op.startswith("BRA") and addr < main_end:
                m = re.search(r"0x([0-9a-f]+)", txt)
                if m and int(m.group(1), 16) <= addr:
                    loops.append((int(m.group(1), 16), addr))
        for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
            body = [i for i in ins if lo <= i[0] <= hi]
            report(body, "loop 0x%x..0x%x" % (lo, hi), ccost, cn, cw)


def report(body, title, ccost, cn, cw):
    n, wide, cyc = show(title + " (inline part)", body, top=12)
    calls = collections.Counter(int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) for _, op, t, _ in body if op.startswith("CALL.REL"))
    tot_n = n + sum(cn.get(a, 0) * k for a, k in calls.items())
    tot_c = cyc + sum(ccost.get(a, 0) * k for a, k in calls.items())
    tot_w = wide + sum(cw.get(a, 0) * k for a, k in calls.items())
    print("   calls: %s -> with callees %d instructions, %d IMAD.WIDE (%.0f cycles), est. %.0f issue cycles; glue share %.1f %%"
          % (", ".join("0x%x x%d" % (a, k) for a, k in calls.items()), tot_n, tot_w, tot_w * 4.2, tot_c, 100.0 * cyc / max(tot_c, 1)))


if __name__ == "__main__":
    main()
