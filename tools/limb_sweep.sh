#!/bin/bash
# A/B of the limb-sweep kernel's launch shapes on one 2^20-row ECDSA-shaped proof (P2G_LIMB_MINB: 2 = 2x256 threads @126 regs,
# 4 = 3x192 @96, 5 = 5x128 @<=102, 3 = 3x256 @80 with spills) x strips (P2G_LIMB_PHASES)
for minb in ${MINBS:-2 3}; do for ph in ${PHASES:-4}; do
  P2G_LIMB_MINB=$minb P2G_LIMB_PHASES=$ph ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_quotient_limb --csv \
     --log-file /tmp/ql.csv python tools/one_proof.py 20 ecdsa 1 > /dev/null 2>&1
  echo "minb=$minb phases=$ph $(grep -o 'dram__bytes_read.sum.*\|gpu__time_duration.sum.*' /tmp/ql.csv | tr '\n' ' ')"
done; done
