#!/bin/bash
# ncu captures for round 2 (run via gpurun, 1 GPU).  Numbers printed by runs under ncu are never bench values.
#   $1 = tag (default r02)
set -u
O=gpurun_out
T=${1:-r02}
B="python bench.py --steps 40 --warmup 10 --no-cpu --no-dense --no-color --no-mesh --no-sharded --no-k0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/${T}_launches.csv $B > $O/p1.log 2>&1
# --cache-control none: iterations 2..10 of a frame find the surface band in L2/L1 — that is the steady state
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_linearize -s 205 -c 1 -f -o $O/${T}_k_linearize $B > $O/p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fuse_cert|k_fuse_exact|k_fuse_plan" -s 60 -c 3 -f -o $O/${T}_k_fuse_traj $B > $O/p3.log 2>&1
ls -la $O/${T}_*
