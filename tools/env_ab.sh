#!/bin/bash
# A/B of an environment switch of libnnic on one box: bash tools/env_ab.sh NNIC_TC_PIN "20 200"  (values 1 / 0 interleaved)
VAR=$1; for steps in ${2:-20 200}; do for f in 1 0 1 0; do
  echo -n "steps=$steps $VAR=$f: "
  env $VAR=$f timeout 200 python bench.py --steps $steps --no-cpu-baseline --no-strong-c5 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['kernels'];print('step',d['ms_per_step'],' '.join('%s %.4f'%(n,k[n]['ms_per_launch']) for n in ('conv1','conv2','conv3','conv4','conv8','dconv1','dconv5','dconv6','dconv7','dconv8')),'sm_mhz',d['clocks'].get('sm_mhz'))"
done; done
