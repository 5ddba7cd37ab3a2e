#!/bin/bash
# per-kernel time of one proof (ncu launch list) -- quotient kernels only
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv -k regex:k_quotient python tools/one_proof.py 20 ecdsa 1 > /dev/null 2>&1
python tools/launch_summary.py /tmp/l.csv 20
