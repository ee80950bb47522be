# data-parallel experiments (usage: bash tools/dp_exp2.sh N [config]): overlap vs tail, fp32 vs bf16 gradients on the wire
N=${1:-2}; CFG=${2:-cfg2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG --steps 30 --warmup 5 --no-cpu-baseline --only 2>&1 | python tools/bench_line.py "$1"; }
python bench.py --config $CFG --steps 30 --warmup 5 --no-cpu-baseline --no-front-end --only 2>&1 | python tools/bench_line.py "n1"
NEKO_DP_MODE=overlap NEKO_DP_COMPRESS=none run overlap_f32
NEKO_DP_MODE=overlap NEKO_DP_COMPRESS=bf16 run overlap_bf16
NEKO_DP_MODE=tail NEKO_DP_COMPRESS=none run tail_f32
NEKO_DP_MODE=tail NEKO_DP_COMPRESS=bf16 run tail_bf16
NEKO_DP_BACKEND=p2p run p2p_overlap
NEKO_DP_BACKEND=p2p NEKO_DP_MODE=tail run p2p_tail
