# round 2, GPU job 2 (2 GPUs): sharded correctness over the peer-memory barrier, 2-GPU bench, host-entry tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -5 gpurun_out/r2b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2b_mg.log 2>&1; tail -25 gpurun_out/r2b_mg.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -c 2500 gpurun_out/r2b_bench_n2.json; tail -5 gpurun_out/r2b_bench_n2.err
BIPYMC_B200_HOST_PEER=1 timeout 300 python -m pytest tests -m gpu -q -k "host_buffer" > gpurun_out/r2b_hostpeer.log 2>&1; tail -5 gpurun_out/r2b_hostpeer.log
