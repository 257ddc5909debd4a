#!/bin/bash
# usage: quick_bench.sh name [lib]  -> one line: fps, stage ms, dense ms
name=$1; lib=$2
if [ -n "$lib" ]; then export TSDF_B200_LIB=$lib; fi
timeout 150 python bench.py --steps 100 --warmup 10 --no-cpu --no-color --no-mesh > /tmp/qb_$name.json 2>/dev/null
python - <<PY
import json
d=json.load(open("/tmp/qb_$name.json"))
print("%-8s fps %.0f e2e %.0f | prep %.1f track %.1f fuse %.1f us | dense %.3f ms (%.1f%% hbm)" % ("$name", d["value"], d["e2e"]["value"], d["stage_ms"]["prep"]*1e3, d["stage_ms"]["track"]*1e3, d["stage_ms"]["fuse"]*1e3, d["dense_fuse"]["ms_per_launch"], 100*d["dense_fuse"]["frac"]))
PY
