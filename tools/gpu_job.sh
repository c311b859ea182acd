# round 2, 2-GPU job: sharded host entry on the copy engines (A/B against the one-kernel form) + sharded correctness
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2v_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2v_mg.log | tail -5
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2v_bench_n2_dma.json 2> gpurun_out/r2v_bench_n2.err; tail -c 900 gpurun_out/r2v_bench_n2_dma.json; tail -3 gpurun_out/r2v_bench_n2.err
BIPYMC_B200_SHARD_IN_KERNEL=1 timeout 600 $TR --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2v_bench_n2_kernel.json 2>> gpurun_out/r2v_bench_n2.err; tail -c 900 gpurun_out/r2v_bench_n2_kernel.json
