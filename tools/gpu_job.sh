# round 2, 2-GPU job: TMA bulk stores of accepted rows (own + peer + host-mapped replicas), NVLink-copy re-deal
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -6 gpurun_out/r2m_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-stationary > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -c 1500 gpurun_out/r2m_bench.json; tail -3 gpurun_out/r2m_bench.err
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2m_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2m_mg.log | tail -18
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 50 --warmup 5 --no-stationary > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; tail -c 1800 gpurun_out/r2m_bench_n2.json; tail -5 gpurun_out/r2m_bench_n2.err
C5_PER_GPU=1250000 C5_GENS=10 C5_K=5 timeout 600 $TR --nproc-per-node 2 --master-port 29571 tools/bench_configs.py c5full > gpurun_out/r2m_c5_n2.txt 2>&1; grep config gpurun_out/r2m_c5_n2.txt | cut -c1-1700; tail -3 gpurun_out/r2m_c5_n2.txt | cut -c1-300
