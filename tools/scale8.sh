# weak-scaling check on 8 GPUs: the default bench shard (65 536 envs/GPU) and BASELINE config 5's shard (1 048 576 envs/GPU)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r01_bench_v8_n8.json 2> gpurun_out/r01_bench_v8_n8.err
cut -c1-260 gpurun_out/r01_bench_v8_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 3 --warmup 3 --envs-per-gpu 1048576 --rollout-steps 200 > gpurun_out/r01_bench_v8_config5_n8.json 2> gpurun_out/r01_bench_v8_config5_n8.err
cut -c1-260 gpurun_out/r01_bench_v8_config5_n8.json; tail -2 gpurun_out/r01_bench_v8_config5_n8.err
