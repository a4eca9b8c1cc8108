for d in ${DBG_LIST:-0 15}; do echo "dbg=$d"; NNIC_TC_DBG=$d timeout 120 python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());k=d['kernels'];print('   conv3',k['conv3']['ms_per_launch'],'conv4',k['conv4']['ms_per_launch'],'dconv7',k['dconv7']['ms_per_launch'],'conv2',k['conv2']['ms_per_launch'],'total',d['value'])"; done
