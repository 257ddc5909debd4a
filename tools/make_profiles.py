"""Turn the artefacts of tools/profile_all.sh (gpurun_out/r01_*) into the committed summaries under profiles/."""
import collections, csv, json, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open(R + "gpurun_out/%s_launches.csv" % TAG)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", "")) / 1000.0
tot = sum(a[1] for a in agg.values())
with open(R + "profiles/%s_launch_list_summary.csv" % TAG, "w") as f:
    f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n")
    f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 python bench.py --steps 40 --warmup 10 --no-cpu --no-dense --no-color --no-mesh [--no-sharded --no-k0]   (tools/profile_all.sh / tools/profile_r02.sh)\n")
    f.write("kernel,launches,mean_us,total_us,share\n")
    for k, (n, t) in agg.items():
        f.write("%s,%d,%.2f,%.1f,%.3f\n" % (k, n, t / n, t, t / tot))
open(R + "profiles/%s_launches_raw.csv" % TAG, "w").write(open(R + "gpurun_out/%s_launches.csv" % TAG).read())
desc = {"k_linearize": "512^3 trajectory workload, one GN iteration" + (", --cache-control none (steady state of iterations 2..10)" if TAG != "r01" else ""),
        "k_fuse_traj": "512^3 trajectory frame: k_fuse_plan, k_fuse_cert, k_fuse_exact" if TAG != "r01" else "512^3 trajectory frame: k_fuse_cert then k_fuse_exact",
        "k_fuse_dense": "512^3 dense micro-benchmark: every voxel updated",
        "k_mesh": "512^3, volume fused from 10 trajectory frames: " + ("one sweep (count + surface-cell list) then the list emit" if TAG != "r01" else "count sweep then emit sweep")}
if TAG != "r01":
    desc["k_color"] = "512^3 dense colour pass (48 B per voxel): k_fuse_cert (rows certified free bypass the queue) then k_fuse_exact<colour>"
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for f in desc:
    rep = R + "gpurun_out/%s_%s.ncu-rep" % (TAG, f)
    summ = subprocess.run([sys.executable, R + "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
    src = "/tmp/src_%s.csv" % f
    open(src, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout)
    n = len(re.findall(r"^-- ", summ, flags=re.M))
    out = "# ncu --set full --clock-control none --import-source on  (%s_%s.ncu-rep, B200, %s; capture not committed: summary below; tools/profile_all.sh)\n" % (TAG, f, desc[f])
    out += summ + "\n# hottest source lines (warp-instructions executed / stall samples)\n"
    for w in range(n):
        out += subprocess.run([sys.executable, R + "tools/ncu_lines.py", src, "16", str(w)], capture_output=True, text=True).stdout
    open(R + "profiles/%s_%s_summary.txt" % (TAG, f), "w").write(out)
    traffic[f] = sum(float(a) * UNIT[b] for a, b in re.findall(r"dram__bytes_(?:read|write).sum\s+([0-9.]+) (\w+)", summ))
    print(f, n, "kernels, dram traffic %.1f MB" % (traffic[f] / 1e6), re.findall(r"gpu__time_duration.sum\s+([0-9.]+)", summ))
json.dump({"fuse_dense": traffic["k_fuse_dense"], "fuse_trajectory": traffic["k_fuse_traj"], "k_linearize": traffic["k_linearize"], "mesh": traffic["k_mesh"],
           "color_dense": traffic.get("k_color"),
           "source": "profiles/%s_k_fuse_dense_summary.txt, %s_k_fuse_traj_summary.txt, %s_k_linearize_summary.txt, %s_k_mesh_summary.txt (ncu --set full, "
                     "dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the kernels of the stage; cold cache)" % (TAG, TAG, TAG, TAG)},
          open(R + "profiles/ncu_traffic.json", "w"), indent=1)
for k, (n, t) in agg.items():
    print("%-28s %4d %8.2f %.3f" % (k, n, t / n, t / tot))
