# round 2, GPU job 7 (2 GPUs): small-d persistent kernel re-measured; sharded host entry; 2-GPU bench / C4 / C5 plumbing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -6 gpurun_out/r2g_pytest.log
timeout 300 python tools/bench_configs.py c3 c4 > gpurun_out/r2g_secondary.txt 2>&1; cat gpurun_out/r2g_secondary.txt
timeout 300 python tools/bench_configs.py c4multi > gpurun_out/r2g_c4_n1.txt 2>&1; cat gpurun_out/r2g_c4_n1.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2g_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2g_mg.log | tail -22
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; tail -c 2400 gpurun_out/r2g_bench_n2.json; tail -5 gpurun_out/r2g_bench_n2.err
timeout 300 $TR --nproc-per-node 2 --master-port 29553 tools/bench_configs.py c4multi > gpurun_out/r2g_c4_n2.txt 2>&1; grep config gpurun_out/r2g_c4_n2.txt
C5_PER_GPU=40000 C5_GENS=12 C5_K=5 timeout 300 $TR --nproc-per-node 2 --master-port 29554 tools/bench_configs.py c5full > gpurun_out/r2g_c5small_n2.txt 2>&1; grep config gpurun_out/r2g_c5small_n2.txt; tail -3 gpurun_out/r2g_c5small_n2.txt
