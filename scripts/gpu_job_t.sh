#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for s in 1238 1239 1240 1241; do
  echo "== seed $s"
  SEED=$s ITERS=12 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29541 scripts/diag_2rank.py 2>&1 | grep "^rank" | awk '{printf "%s %s f%s b%s; ", $4, $5, $9, $11}'; echo
done
