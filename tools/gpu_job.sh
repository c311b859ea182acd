# round 2, final 2-GPU job: the shipped tree -- parity suite, smoke, full bench line + reference arm, secondary configs,
# checked build, ncu launch list + full capture of the dominant kernel, sharded correctness + 2-GPU bench
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -4 gpurun_out/r2p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -2 gpurun_out/r2p_smoke.log
timeout 600 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -c 2800 gpurun_out/r2p_bench.json; tail -3 gpurun_out/r2p_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; tail -c 700 gpurun_out/r2p_bench_ref.json
timeout 400 python tools/bench_configs.py c3 c4 demc100 c5shape > gpurun_out/r2p_secondary.txt 2>&1; cat gpurun_out/r2p_secondary.txt
BIPYMC_B200_LIB=$PWD/build_ab/lib_checks.so timeout 300 python tools/sanitize_case.py > gpurun_out/r2p_checked_build.log 2>&1; tail -5 gpurun_out/r2p_checked_build.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2p_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2p_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v4 -s 130 -c 1 -o gpurun_out/prof_r2p -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2p_ncu_full.log 2>&1; tail -2 gpurun_out/r2p_ncu_full.log
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tools/multigpu_check.py > gpurun_out/r2p_mg.log 2>&1; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2p_mg.log | tail -6
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err; tail -c 1500 gpurun_out/r2p_bench_n2.json
