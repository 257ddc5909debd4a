#!/bin/bash
# bench every tuning variant under tracking_sdf_b200/_lib/variants (tools/build_variants.py) + the default build
# usage (gpurun): bash tools/variant_bench.sh "<extra bench args>"
set -u
ARGS="--steps 300 --warmup 10 --no-cpu --no-mesh --no-sharded --no-k0 $*"
for lib in default tracking_sdf_b200/_lib/variants/*.so; do
  if [ "$lib" = default ]; then unset TSDF_B200_LIB; else export TSDF_B200_LIB=$PWD/$lib; fi
  python bench.py $ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-28s fps %7.1f  track %.4f fuse %.4f ms | dense %.3f | colour dense %s' % ('$lib'.split('/')[-1], d['value'], d['stage_ms']['track'], d['stage_ms']['fuse'], d.get('dense_fuse',{}).get('frac',0), d.get('color_fuse',{}).get('frac')))"
done
