#!/bin/bash
# A/B two builds of libnnic on the same box: tools/ab.sh <libA> <libB> [reps]  (sustained enc+rate+dec, c2 shape)
A=$1; B=$2; R=${3:-300}
for i in 1 2; do
  for L in $A $B; do NNIC_LIB=$PWD/$L timeout 300 python tools/gpu_time.py 24 512 768 tc_split $R; done
done
