# round 2, final 8-GPU job: weak-scaling points of the shipped tree (8, 4 GPUs), sharded correctness incl. the serial sampler
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err; tail -c 2200 gpurun_out/r2u_bench_n8.json; tail -4 gpurun_out/r2u_bench_n8.err
timeout 500 $TR --nproc-per-node 4 --master-port 29564 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r2u_bench_n4.json 2> gpurun_out/r2u_bench_n4.err; tail -c 1200 gpurun_out/r2u_bench_n4.json
timeout 600 $TR --nproc-per-node 4 --master-port 29565 tools/multigpu_check.py > gpurun_out/r2u_mg4.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2u_mg4.log | tail -8
timeout 300 $TR --nproc-per-node 4 --master-port 29566 tools/bench_configs.py c4multi > gpurun_out/r2u_c4_n4.txt 2>&1; grep config gpurun_out/r2u_c4_n4.txt
