# round 2, 2-GPU job: where does the C5 re-deal spend its time?  (8-GPU run: 778 ms per generation against 124 ms of kernels)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P C5_PER_GPU=1250000 C5_GENS=10 C5_K=5 timeout 600 $TR --nproc-per-node 2 --master-port 29571 tools/bench_configs.py c5full > gpurun_out/r2l_c5_n2.txt 2>&1; grep config gpurun_out/r2l_c5_n2.txt | cut -c1-1500; grep -i "via \|P2P\|SHM\|NVLS" gpurun_out/r2l_c5_n2.txt | head -12
