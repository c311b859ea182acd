# round 2, 2-GPU job: side-stream prefetch of the next generation's shuffle -- parity, A/B on 1 and 2 GPUs, sharded check
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2q_bench_side.json 2> gpurun_out/r2q_bench.err; tail -c 700 gpurun_out/r2q_bench_side.json; tail -3 gpurun_out/r2q_bench.err
BIPYMC_B200_NO_SIDE=1 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2q_bench_noside.json 2>> gpurun_out/r2q_bench.err; tail -c 700 gpurun_out/r2q_bench_noside.json
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2q_secondary.txt 2>&1; cat gpurun_out/r2q_secondary.txt
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2q_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2q_mg.log | tail -6
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e --no-stationary > gpurun_out/r2q_bench_n2_side.json 2> gpurun_out/r2q_bench_n2.err; tail -c 700 gpurun_out/r2q_bench_n2_side.json
BIPYMC_B200_NO_SIDE=1 timeout 600 $TR --nproc-per-node 2 --master-port 29553 bench.py --gpus 2 --steps 50 --warmup 5 --no-e2e --no-stationary > gpurun_out/r2q_bench_n2_noside.json 2>> gpurun_out/r2q_bench_n2.err; tail -c 700 gpurun_out/r2q_bench_n2_noside.json
