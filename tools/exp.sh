mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_n2.json | cut -c1-1200
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r01_n1_b.json | cut -c1-900
