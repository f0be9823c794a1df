#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { env "$@" ITERS=14 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/diag_2rank.py 2>&1 | grep "^rank" | awk '{it=$4; f[it]+=$10; b[it]+=$12; ms[it]=$5} END {for (i in f) print i, "ms", ms[i], "relays f", f[i], "b", b[i]}' | sort -n | tr "\n" ";"; echo; }
echo "== no speculation"; run PMG_NO_SPECULATE=1
echo "== barrier M-step kernel"; run PMG_MSTEP_LAG=0
