#!/bin/bash
# A/B builds of libnnic on the same box: tools/ab.sh <reps> <libA> <libB> [...]  (sustained enc+rate+dec, c2 shape)
R=$1; shift
for i in 1 2; do
  for L in "$@"; do NNIC_LIB=$PWD/$L timeout 300 python tools/gpu_time.py 24 512 768 tc_split $R; done
done
