# round 2, 8-GPU job: weak-scaling points (4, 8 GPUs) with the sharded end-to-end entry and the N > 1 selfcheck,
# the peer-store A/B, C5 at full scale (10^7 x 1000-D, sub-population mode), sharded correctness on 8 ranks, C4 on 4 GPUs
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2k_bench_n8.json 2> gpurun_out/r2k_bench_n8.err; tail -c 2600 gpurun_out/r2k_bench_n8.json; tail -4 gpurun_out/r2k_bench_n8.err
timeout 900 $TR --nproc-per-node 8 --master-port 29562 tools/bench_configs.py c5full > gpurun_out/r2k_c5full.txt 2>&1; grep config gpurun_out/r2k_c5full.txt; tail -4 gpurun_out/r2k_c5full.txt
BIPYMC_B200_NO_PEER_STORES=1 timeout 400 $TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e --no-stationary > gpurun_out/r2k_bench_n8_nopeerstores.json 2> gpurun_out/r2k_bench_n8_nopeerstores.err; tail -c 1200 gpurun_out/r2k_bench_n8_nopeerstores.json
timeout 500 $TR --nproc-per-node 4 --master-port 29564 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; tail -c 1500 gpurun_out/r2k_bench_n4.json
timeout 900 $TR --nproc-per-node 8 --master-port 29565 tools/multigpu_check.py > gpurun_out/r2k_mg8.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2k_mg8.log | tail -20
timeout 300 $TR --nproc-per-node 4 --master-port 29566 tools/bench_configs.py c4multi > gpurun_out/r2k_c4_n4.txt 2>&1; grep config gpurun_out/r2k_c4_n4.txt
