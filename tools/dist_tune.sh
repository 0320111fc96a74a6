#!/bin/bash
R=$1
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $R --steps 100 --warmup 10 --quick 2>/dev/null | grep quick; }
run ABT_DIST_CE=0 ABT_COMM_MAX_CTAS=0 ABT_DIST_RESERVE_SMS=0
run ABT_DIST_CE=1 ABT_COMM_MAX_CTAS=0 ABT_DIST_RESERVE_SMS=0
run ABT_DIST_CE=1 ABT_COMM_MAX_CTAS=8 ABT_DIST_RESERVE_SMS=12
run ABT_DIST_CE=1 ABT_COMM_MAX_CTAS=16 ABT_DIST_RESERVE_SMS=20
run ABT_DIST_CE=1 ABT_COMM_MAX_CTAS=0 ABT_DIST_RESERVE_SMS=12
run ABT_DIST_CE=0 ABT_COMM_MAX_CTAS=8 ABT_DIST_RESERVE_SMS=12
