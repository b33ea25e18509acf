#!/usr/bin/env python3
"""ncu_source_stats.py <source-page csv> -- per kernel: executed warp instructions and stall samples by opcode, and by
"inside the out-of-line multiplier / outside" (the callees are the code after the kernel's last EXIT).  Input: the output of
`ncu -i report.ncu-rep --page source --csv` (tests/gpu_profile.sh).  Development tool."""
import collections
import csv
import re
import sys

STALLS = ["stall_wait", "stall_math", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_long_sb",
          "stall_short_sb", "stall_branch_resolving", "stall_lg", "stall_mio", "stall_misc", "stall_drain"]


def main():
    kernels, cur, hdr = [], None, None
    for row in csv.reader(open(sys.argv[1])):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
            hdr = None
            continue
        if row and row[0] == "Address":
            hdr = row
            continue
        if cur is not None and hdr and len(row) >= len(hdr) - 2 and row[0].startswith("0x"):
            cur["rows"].append(dict(zip(hdr, row)))
    for k in kernels:
        rows = k["rows"]
        ops = [re.sub(r"^@!?U?P\d+\s+", "", r["Source"].strip()).split()[0] for r in rows]
        ex = [int(r["Instructions Executed"]) for r in rows]
        tot = sum(ex)
        # the callees start after the last EXIT of the kernel body
        last_exit = max(i for i, o in enumerate(ops) if o == "EXIT")
        # ... more precisely after the last instruction that is reached by fallthrough; approximate with the first RET's function
        by_op = collections.Counter()
        st_by_op = collections.defaultdict(collections.Counter)
        region = collections.Counter()
        st_region = collections.defaultdict(collections.Counter)
        for i, (o, e, r) in enumerate(zip(ops, ex, rows)):
            key = "wide" if o.startswith("IMAD.WIDE") else o
            by_op[key] += e
            reg = "callee" if i > last_exit else "inline"
            region[reg] += e
            for s in STALLS:
                v = int(r.get(s, "0") or 0)
                st_by_op[key][s] += v
                st_region[reg][s] += v
        samples = sum(sum(c.values()) for c in st_by_op.values())
        print("==", k["name"], "executed warp instructions %.3e, stall samples %d" % (tot, samples))
        print("   inline %.1f %%, out-of-line multiplier bodies %.1f %%" % (100.0 * region["inline"] / tot, 100.0 * region["callee"] / tot))
        for reg in ("inline", "callee"):
            t = sum(st_region[reg].values())
            print("   samples in %-6s %5.1f %%: " % (reg, 100.0 * t / max(samples, 1)) +
                  ", ".join("%s %.1f" % (s[6:], 100.0 * v / max(samples, 1)) for s, v in st_region[reg].most_common(6)))
        print("   opcode: share of executed | share of samples (top stall reasons)")
        for o, e in by_op.most_common(16):
            t = sum(st_by_op[o].values())
            print("   %-18s %5.1f %% | %5.1f %%  %s" % (o, 100.0 * e / tot, 100.0 * t / max(samples, 1),
                                                      ", ".join("%s %.1f" % (s[6:], 100.0 * v / max(samples, 1)) for s, v in st_by_op[o].most_common(3))))


if __name__ == "__main__":
    main()
