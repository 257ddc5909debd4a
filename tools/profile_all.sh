#!/bin/bash
# One GPU-box pass that produces every ncu artefact summarised under profiles/ (run via gpurun, 1 GPU).
#   gpurun_out/r01_launches.csv          launch list of the bench command (gpu__time_duration.sum)
#   gpurun_out/r01_k_linearize.ncu-rep   --set full of one k_linearize launch (trajectory workload)
#   gpurun_out/r01_k_fuse_traj.ncu-rep   --set full of k_fuse_cert + k_fuse_exact on a trajectory frame
#   gpurun_out/r01_k_fuse_dense.ncu-rep  --set full of k_fuse_cert on the dense micro-benchmark
#   gpurun_out/r01_k_mesh.ncu-rep        --set full of the two mesher sweeps
# Numbers printed by runs under ncu are never bench values.
set -u
O=gpurun_out
B="python bench.py --steps 40 --warmup 10 --no-cpu --no-dense --no-color --no-mesh"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/r01_launches.csv $B > $O/p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linearize -s 205 -c 1 -f -o $O/r01_k_linearize $B > $O/p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fuse_cert|k_fuse_exact" -s 40 -c 2 -f -o $O/r01_k_fuse_traj $B > $O/p3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_cert -s 4 -c 1 -f -o $O/r01_k_fuse_dense python tools/dense_probe.py 512 4 > $O/p4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_sweep -s 2 -c 2 -f -o $O/r01_k_mesh python tools/mesh_probe.py 512 > $O/p5.log 2>&1
ls -la $O/r01_*
