#!/bin/bash
# One GPU-box pass that produces every ncu artefact summarised under profiles/ for round 2 (run via gpurun, 1 GPU).
#   gpurun_out/r02_launches.csv            launch list of the bench command (gpu__time_duration.sum)
#   gpurun_out/r02_k_linearize.ncu-rep     --set full of one k_linearize launch, --cache-control none (iterations 2..10 of a
#                                          frame find the surface band in L1/L2: that is the steady state)
#   gpurun_out/r02_k_fuse_traj.ncu-rep     --set full of k_fuse_plan + k_fuse_cert + k_fuse_exact on a trajectory frame
#   gpurun_out/r02_k_fuse_dense.ncu-rep    --set full of k_fuse_cert on the dense micro-benchmark
#   gpurun_out/r02_k_mesh.ncu-rep          --set full of the mesher (sweep + list emit)
#   gpurun_out/r02_k_color.ncu-rep         --set full of the dense colour pass (k_fuse_cert + k_fuse_exact<colour>)
#   gpurun_out/r02_k0_launches.csv         launch list of the K0 kernels on a noisy frame
# Numbers printed by runs under ncu are never bench values.
set -u
O=gpurun_out
T=${1:-r02}
B="python bench.py --steps 40 --warmup 10 --no-cpu --no-dense --no-color --no-mesh --no-sharded --no-k0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/${T}_launches.csv $B > $O/p1.log 2>&1
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_linearize -s 205 -c 1 -f -o $O/${T}_k_linearize $B > $O/p2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fuse_cert|k_fuse_exact|k_fuse_plan" -s 60 -c 3 -f -o $O/${T}_k_fuse_traj $B > $O/p3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_cert -s 4 -c 1 -f -o $O/${T}_k_fuse_dense python tools/dense_probe.py 512 4 > $O/p4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_mc_" -s 2 -c 2 -f -o $O/${T}_k_mesh python tools/mesh_probe.py 512 > $O/p5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fuse_cert|k_fuse_exact" -s 4 -c 2 -f -o $O/${T}_k_color python tools/color_probe.py > $O/p6.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k0_|k_prep|k_pyramid" -c 40 --csv --log-file $O/${T}_k0_launches.csv python tools/k0_probe.py > $O/p7.log 2>&1
ls -la $O/${T}_*
