# round 2, 8-GPU A/B of the sharded host entry: copy engines (default) vs the one-kernel form
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2w_bench_n8_dma.json 2> gpurun_out/r2w_bench_n8.err; tail -c 1000 gpurun_out/r2w_bench_n8_dma.json; tail -3 gpurun_out/r2w_bench_n8.err
BIPYMC_B200_SHARD_IN_KERNEL=1 timeout 600 $TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2w_bench_n8_kernel.json 2>> gpurun_out/r2w_bench_n8.err; tail -c 1000 gpurun_out/r2w_bench_n8_kernel.json
timeout 400 $TR --nproc-per-node 4 --master-port 29563 bench.py --gpus 4 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2w_bench_n4_dma.json 2>> gpurun_out/r2w_bench_n8.err; tail -c 700 gpurun_out/r2w_bench_n4_dma.json
