# round 2 (session 2), 2 GPUs: sharded correctness + 2-GPU bench of the current tree
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2zg_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2zg_mg.log | tail -12
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 30 --warmup 5 --no-stationary > gpurun_out/r2zg_bench_n2.json 2> gpurun_out/r2zg_bench_n2.err; tail -c 1500 gpurun_out/r2zg_bench_n2.json; tail -3 gpurun_out/r2zg_bench_n2.err
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/r2zg_pytest_mg.log 2>&1; tail -3 gpurun_out/r2zg_pytest_mg.log
timeout 300 $TR --nproc-per-node 2 --master-port 29553 tools/bench_configs.py c4multi > gpurun_out/r2zg_c4_n2.txt 2>&1; grep "^{" gpurun_out/r2zg_c4_n2.txt
