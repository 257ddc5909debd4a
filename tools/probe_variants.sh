#!/bin/bash
# usage: probe_variants.sh v1 v2 ...   (names under tracking_sdf_b200/_lib/variants)
V=tracking_sdf_b200/_lib/variants
echo "== base"; python tools/lin_probe.py | tail -1
for v in "$@"; do echo "== $v"; TSDF_B200_LIB=$V/libtsdf_$v.so python tools/lin_probe.py | tail -1; done
