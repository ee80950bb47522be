run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python tools/bench_line.py "$1"; }
run base
NCCL_MAX_CTAS=8 run nccl_max_ctas=8
NEKO_GEMM_SMS=132 run gemm_sms=132
NEKO_GEMM_SMS=132 NCCL_MAX_CTAS=16 run gemm_sms=132,nccl16
NEKO_GEMM_SMS=140 NCCL_MAX_CTAS=8 run gemm_sms=140,nccl8
