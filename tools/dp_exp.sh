# data-parallel experiments at N=2 (usage: bash tools/dp_exp.sh): all-reduce bucket size
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python tools/bench_line.py "$1"; }
NEKO_DP_BUCKET_MB=16 run bucket16
NEKO_DP_BUCKET_MB=32 run bucket32
NEKO_DP_BUCKET_MB=64 run bucket64
NEKO_DP_BUCKET_MB=160 run bucket160
