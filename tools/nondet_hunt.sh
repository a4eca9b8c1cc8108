cd /root/repo
fail=0
for i in 1 2 3 4; do r=$(timeout 150 python tools/stress_decode.py 16 270 480 300 profile geometric 2>&1 | tail -2); echo "final build: $r"; case "$r" in *differs*|*Error*) fail=1;; esac; done
if [ $fail = 1 ]; then
  for i in 1 2 3 4; do echo -n "base lib: "; NNIC_LIB=$PWD/neural_network_image_compression_b200/libnnic_base.so timeout 150 python tools/stress_decode.py 16 270 480 300 profile geometric 2>&1 | tail -2; done
  for i in 1 2 3; do echo -n "PDL=0 FUSE=0 PIN=0: "; NNIC_PDL=0 NNIC_FUSE_D78=0 NNIC_TC_PIN=0 timeout 150 python tools/stress_decode.py 16 270 480 300 profile geometric 2>&1 | tail -2; done
fi
nvidia-smi --query-gpu=serial,uuid --format=csv,noheader
