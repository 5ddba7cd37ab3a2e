#!/bin/bash
# A/B of the limb-sweep kernel's knobs on one 2^20-row ECDSA-shaped proof: strips (P2G_LIMB_PHASES) x blocks per SM (P2G_LIMB_MINB)
for minb in 3 2; do for ph in 1 2 4; do
  P2G_LIMB_MINB=$minb P2G_LIMB_PHASES=$ph ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_quotient_limb --csv \
     --log-file /tmp/ql.csv python tools/one_proof.py 20 ecdsa 1 > /dev/null 2>&1
  echo "minb=$minb phases=$ph $(grep -o 'dram__bytes_read.sum.*\|gpu__time_duration.sum.*' /tmp/ql.csv | tr '\n' ' ')"
done; done
