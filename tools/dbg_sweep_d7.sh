#!/bin/bash
# Switch-off ablation of the fused dconv7 -> dconv8 tail (development build without timers: make BUILD=build_dev2 OUT=../libnnic_dev2.so EXTRA=-DNNIC_TC_DEVELOP)
export NNIC_LIB=${NNIC_LIB:-$PWD/neural_network_image_compression_b200/libnnic_dev2.so}
for d in ${DBG_LIST:-0 2 8 10 2048 4096}; do echo "dbg=$d"; NNIC_TC_DBG=$d timeout 120 python bench.py --steps 20 --no-cpu-baseline --no-strong-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['kernels'];print('   dconv7',k['dconv7']['ms_per_launch'],'dconv8',k['dconv8']['ms_per_launch'],'dconv6',k['dconv6']['ms_per_launch'],'total',d['value'])"; done
