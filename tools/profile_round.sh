#!/bin/bash
# Round profile: (1) ncu launch list of the bench command, (2) ncu --set full of the top kernels of one proof, exported to CSV on
# the box (the .ncu-rep files are too large to bring back).  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --inflight 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
python tools/launch_summary.py $out/${tag}_launches_bench.csv 50 > $out/${tag}_launches_bench_summary.txt
ncu --set full --clock-control none --import-source on -k regex:'k_pass_strided|k_pass_contig' -s 4 -c 4 -o /tmp/ntt \
    python tools/one_proof.py 20 ecdsa 1 > $out/${tag}_ncu_ntt.log 2>&1
ncu -i /tmp/ntt.ncu-rep --page raw --csv > $out/${tag}_ntt_ncu_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_leaf_keccak -s 1 -c 1 -o /tmp/keccak \
    python tools/one_proof.py 20 ecdsa 1 > $out/${tag}_ncu_keccak.log 2>&1
ncu -i /tmp/keccak.ncu-rep --page raw --csv > $out/${tag}_keccak_ncu_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'k_quotient' -c 17 -o /tmp/quot \
    python tools/one_proof.py 20 ecdsa 1 > $out/${tag}_ncu_quot.log 2>&1
ncu -i /tmp/quot.ncu-rep --page raw --csv > $out/${tag}_quotient_ncu_raw.csv 2>/dev/null
ls -la $out
