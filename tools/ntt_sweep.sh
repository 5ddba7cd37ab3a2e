#!/bin/bash
# sweep of the NTT tuning knobs on one proof (prints lde_ms / ntt_ms)
run() { env "$@" python tools/one_proof.py 20 ecdsa 2 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print('$*', {k:d[k] for k in ['ntt_ms','lde_ms','wires_commit_ms']})"; }
run P2G_NTT_R=3 P2G_NTT_LAST=10 P2G_NTT_TILE=12
run P2G_NTT_R=3 P2G_NTT_LAST=11 P2G_NTT_TILE=12
run P2G_NTT_R=3 P2G_NTT_LAST=11 P2G_NTT_TILE=12 P2G_NTT_TH=512
run P2G_NTT_R=3 P2G_NTT_LAST=11 P2G_NTT_TILE=12 P2G_NTT_TH=128
run P2G_NTT_R=3 P2G_NTT_LAST=10 P2G_NTT_TILE=13
run P2G_NTT_R=4 P2G_NTT_LAST=11 P2G_NTT_TILE=12
run P2G_NTT_R=4 P2G_NTT_LAST=11 P2G_NTT_TILE=12 P2G_NTT_TH=128
run P2G_NTT_R=3 P2G_NTT_LAST=9 P2G_NTT_TILE=12
