#!/bin/bash
# gpu_sweep.sh -- development helper for a gpurun box: bench.py once per experiment library.
#   tests/gpu_sweep.sh <tag> [variant ...]      variant "" / "default" = the shipped libplume_b200.so
# Each run writes its JSON line to gpurun_out/sweep_<tag>_<variant>.json and a one-line summary to stdout.
# Extra bench.py flags: SWEEP_ARGS (default: --steps 5 --warmup 3 --no-cpu-baseline).
tag=$1; shift
mkdir -p gpurun_out
args=${SWEEP_ARGS:---steps 5 --warmup 3 --no-cpu-baseline}
for v in "$@"; do
    lib=""
    if [ "$v" != "default" ] && [ -n "$v" ]; then lib="$PWD/zk-nullifier-sig_b200/libplume_b200_$v.so"; fi
    out=gpurun_out/sweep_${tag}_${v}.json
    PLUME_B200_LIB=$lib timeout 600 python bench.py $args > $out 2> gpurun_out/sweep_${tag}_${v}.err
    python - "$out" "$v" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    st = {k: round(v["ms_total"] / v["launches"], 3) for k, v in d.get("stages", {}).items()}
    print(sys.argv[2], "ms_per_step %.2f" % d["ms_per_step"], "e2e %.3g" % d["e2e"]["value"], "checks", all(d["checks"].values()), st)
except Exception as e:
    print(sys.argv[2], "FAILED", e)
EOF
done
