# round 2, second 8-GPU job: bulk-store peer exchange + on-the-fly inverse in the list packing; C5 with the NVLink-copy re-deal
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; tail -c 2200 gpurun_out/r2n_bench_n8.json; tail -4 gpurun_out/r2n_bench_n8.err
timeout 900 $TR --nproc-per-node 8 --master-port 29562 tools/bench_configs.py c5full > gpurun_out/r2n_c5full.txt 2>&1; grep config gpurun_out/r2n_c5full.txt | cut -c1-1800; tail -3 gpurun_out/r2n_c5full.txt | cut -c1-300
timeout 500 $TR --nproc-per-node 4 --master-port 29564 bench.py --gpus 4 --steps 50 --warmup 5 --no-stationary > gpurun_out/r2n_bench_n4.json 2> gpurun_out/r2n_bench_n4.err; tail -c 1200 gpurun_out/r2n_bench_n4.json
timeout 600 $TR --nproc-per-node 8 --master-port 29565 tools/multigpu_check.py > gpurun_out/r2n_mg8.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2n_mg8.log | tail -6
