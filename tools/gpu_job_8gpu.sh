# round 2 (session 2), final 8-GPU job of the shipped tree: the 8-GPU weak-scaling point (sharded end-to-end entry, selfcheck) and
# C5 at full scale (10^7 x 1000-D, sub-population mode k = 10) with the block-wise propose / accept kernels
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2zi_bench_n8.json 2> gpurun_out/r2zi_bench_n8.err; tail -c 2000 gpurun_out/r2zi_bench_n8.json; tail -3 gpurun_out/r2zi_bench_n8.err
C5_GENS=10 timeout 500 $TR --nproc-per-node 8 --master-port 29562 tools/bench_configs.py c5full > gpurun_out/r2zi_c5full.txt 2>&1; grep "^{" gpurun_out/r2zi_c5full.txt
