N=${1:-2}; CFG=${2:-cfg2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG --steps 30 --warmup 5 --no-cpu-baseline --only 2>&1 | python tools/bench_line.py "$1"; }
NEKO_DP_BACKEND=p2p run p2p_ce_nccl_tail
NEKO_DP_BACKEND=p2p NEKO_DP_NCCL_TAIL=0 run p2p_ce_only
