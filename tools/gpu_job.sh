# round 2 (session 2), 1 GPU: ncu --set full of the rewritten split-path kernels at d = 1000 (one report: the merge-back limit is 64 MiB)
set -x
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:"propose_kernel|accept_kernel" -s 20 -c 2 -o gpurun_out/prof_r2zm_c5 -f python tools/bench_configs.py c5shape > gpurun_out/r2zm_ncu_c5.log 2>&1; tail -n 2 gpurun_out/r2zm_ncu_c5.log; ls -la gpurun_out
