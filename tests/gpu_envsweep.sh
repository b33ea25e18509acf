for kv in "PLUME_BINV_K=8" "PLUME_BINV_K=32" "PLUME_BINV_K=64" "PLUME_FIXED_WINDOW=22" "PLUME_FIXED_WINDOW=18" "PLUME_HOST_CHUNK_ITEMS=113664" "PLUME_HOST_CHUNK_ITEMS=340992"; do
  env $kv python bench.py --steps 10 --no-cpu-baseline --no-extra-configs > gpurun_out/r2i_$kv.json 2>/dev/null
  python - "$kv" <<PY
import json,sys
d=json.loads(open("gpurun_out/r2i_%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
st={k: round(v["ms_total"]/v["launches"],3) for k,v in d["stages"].items()}
print(sys.argv[1], round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), st["binv"], st["sign_fixed"], st["verify_mul_a"], all(d["checks"].values()))
PY
done
