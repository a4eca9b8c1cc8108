#!/bin/bash
# fused vs unfused decoder tail on one box: bash tools/d78_ab.sh [steps...]
for steps in ${@:-20 60}; do for f in 1 0 1 0; do
  echo -n "steps=$steps NNIC_FUSE_D78=$f: "
  NNIC_FUSE_D78=$f timeout 200 python bench.py --steps $steps --no-cpu-baseline --no-strong-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['kernels'];print('step',d['ms_per_step'],'dconv7',k['dconv7']['ms_per_launch'],'dconv8',k['dconv8']['ms_per_launch'],'dconv6',k['dconv6']['ms_per_launch'],'conv2',k['conv2']['ms_per_launch'],'sm_mhz',d['clocks'].get('sm_mhz'))"
done; done
