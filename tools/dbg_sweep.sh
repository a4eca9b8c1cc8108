#!/bin/bash
# Switch-off ablation of k_tc_conv_patch (profiles/r2_component_ablation.log): DBG_LIST="0 2 32 1024" bash tools/dbg_sweep.sh
# Needs the development build (make -C neural_network_image_compression_b200/csrc dev); the product library has no switches.
export NNIC_LIB=${NNIC_LIB:-$PWD/neural_network_image_compression_b200/libnnic_dev.so}
for d in ${DBG_LIST:-0 15}; do echo "dbg=$d"; NNIC_TC_DBG=$d timeout 120 python bench.py --steps 20 --no-cpu-baseline --no-strong-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['kernels'];print('   conv3',k['conv3']['ms_per_launch'],'conv4',k['conv4']['ms_per_launch'],'dconv7',k['dconv7']['ms_per_launch'],'conv2',k['conv2']['ms_per_launch'],'total',d['value'])"; done
