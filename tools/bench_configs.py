#!/usr/bin/env python
"""Secondary measurements for the BASELINE configs that are NOT the headline bench line
(bench.py measures configs[1]).  Prints one JSON line per configuration:
  c3  DREAM, bimodal 2-D Gaussian, 10^6 chains, CR adaptation + outlier reset on, no history
  c4  DREAM, line fit (3 parameters, 50 data points), 10^5 chains
  demc100  DE-MC on the 100-D Gaussian, 10^5 chains (32 d + 16 bytes per chain-step)
Device-resident timing with CUDA events, same rules as bench.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def run(name, make, N, d, bytes_per_step, gens=100, warm=20):
    from bipymc_b200 import _lib
    np.random.seed(1)
    s = make()
    k = 0
    s.run_mcmc(N * (warm + 1))
    k += warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s.run_mcmc(N * (gens + 1), _k_gen0=k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    v = N * gens / (ms * 1e-3)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # untimed extra pass with CUDA events around every launch: average launch time per kernel family
    import ctypes as C
    pg = max(4, gens // 5)
    _lib.check(s._libh.bpm_profile(s._handle, 1))
    s.run_mcmc(N * (pg + 1), _k_gen0=k + gens)
    torch.cuda.synchronize()
    ms_k, n_k = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(s._libh.bpm_profile_read(s._handle, ms_k, n_k))
    _lib.check(s._libh.bpm_profile(s._handle, 0))
    kinds = ["split", "propose", "likelihood", "accept", "fused_phase", "cr_reduce", "other6", "other7"]
    kms = dict((kinds[i], {"avg_ms": ms_k[i] / n_k[i], "per_generation": n_k[i] / pg}) for i in range(8) if n_k[i])
    print(json.dumps({"config": name, "n_chains": N, "dim": d, "chain_steps_per_s": v, "ms_per_generation": ms / gens,
                      "algorithmic_bytes_per_chain_step": bytes_per_step,
                      "algorithmic_GBps": v * bytes_per_step / 1e9, "frac_of_hbm_peak": v * bytes_per_step / 1e9 / peak,
                      "acceptance_fraction": s.acceptance_fraction, "launches_event_pass": kms}))


def main():
    from bipymc_b200 import DreamMpi, DeMcMpi, targets
    which = sys.argv[1:] or ["c3", "c4", "demc100", "c5shape"]
    if "c3" in which:
        N = 1000000
        t = targets.BimodeGauss_2D(log_of_pdf=False)
        # own row + 6 partners + write (8 d each) + lnL r/w, + moments r/w (32 d) during adaptation
        run("c3 bimodal DREAM 1e6 chains", lambda: DreamMpi(t.ln_like, [0.0, 0.0], n_chains=N, seed=3, varepsilon=1.0,
                                                              history="none", burnin_gen=10 ** 6, n_cr_gen=10,
                                                              outlier_gen=50), N, 2, 64 * 2 + 16 + 32 * 2 + 8 * 2)
    if "c4" in which:
        N = 100000
        t = targets.LineFit()
        run("c4 linefit DREAM 1e5 chains", lambda: DreamMpi(t.ln_like, [-0.8, 4.5, 0.2], n_chains=N, seed=3,
                                                              varepsilon=1e-2, history="none", burnin_gen=10 ** 6,
                                                              n_cr_gen=10), N, 3, 64 * 3 + 16 + 32 * 3 + 8 * 3)
    if "demc100" in which:
        N = 100000
        t = targets.Gauss_100D()
        run("DE-MC Gauss_100D 1e5 chains", lambda: DeMcMpi(t.ln_like, np.zeros(100), n_chains=N, seed=3,
                                                             varepsilon=np.arange(100) + 1.0, history="none"),
            N, 100, 32 * 100 + 16 + 32 * 100)


def c5shape():
    """configs[4] shape on one GPU: DREAM on the 1000-D correlated Gaussian (direct log-density),
    2 x 10^4 chains -- the generic split path (propose / tiled FP64 quadratic form / accept)."""
    from bipymc_b200 import DreamMpi, targets
    N, d = 20000, 1000
    t = targets.Gauss_100D(dim=d)
    run("c5-shape DREAM Gauss_1000D 2e4 chains", lambda: DreamMpi(t.ln_like, np.zeros(d), n_chains=N, seed=3,
                                                                    varepsilon=np.arange(d) + 1.0, history="none",
                                                                    burnin_gen=10 ** 6, n_cr_gen=10),
        N, d, 64 * d + 16 + 32 * d + 8 * d, gens=20, warm=12)


def c5multi():
    """configs[4] shape sharded over the GPUs of one box (launch under torchrun): DREAM on the
    1000-D Gaussian, 2.5 x 10^4 chains per GPU, once as ONE population (replicas + peer-memory
    exchange) and once in the stated sub-population mode (islands re-dealt every 10 generations)."""
    import torch.distributed as dist
    from bipymc_b200 import DreamMpi, targets
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    N, d, gens, warm = 25000 * world, 1000, 20, 10
    t = targets.Gauss_100D(dim=d)
    for mode, kw in (("one population, p2p exchange", {}), ("sub-population mode k=10", dict(subpop_k=10))):
        np.random.seed(1)
        s = DreamMpi(t.ln_like, np.zeros(d), n_chains=N, seed=3, varepsilon=np.arange(d) + 1.0, history="none",
                     burnin_gen=10 ** 6, n_cr_gen=5, device=local, **kw)
        s.run_mcmc(N * (warm + 1))
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run_mcmc(N * (gens + 1), _k_gen0=warm)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"config": "c5-shape DREAM Gauss_1000D, %d GPUs x 2.5e4 chains, %s" % (world, mode),
                              "n_chains": N, "chain_steps_per_s": N * gens / (float(ms.item()) * 1e-3),
                              "ms_per_generation": float(ms.item()) / gens,
                              "acceptance_fraction": s.acceptance_fraction}), flush=True)
        s.close()
        dist.barrier()
    dist.destroy_process_group()


def c5full():
    """BASELINE configs[4] AT SCALE (launch under torchrun, one rank per GPU): DREAM on the 1000-D correlated
    Gaussian (direct log-density: the reference's log(pdf) underflows at d = 1000) with
    C5_PER_GPU chains per GPU (default 1.25e6: 10^7 on 8 GPUs), in the stated sub-population mode -- every GPU
    steps its chains as one island (pairs from the local opposite half, no replica of the 80 GB population),
    islands re-dealt with one all-to-all every k generations (default 10).  History off, running moments and CR
    adaptation on.  Reports chain-steps/s (device events, max over ranks, re-deal included), the FP64-roof
    fraction of the quadratic-form kernel (2 d^2 flop per chain-step against the measured 36.5 TFLOP/s,
    profiles/r1_fp64_probe_b200.txt), acceptance and the R-hat trend from the streaming moments."""
    import ctypes as C
    import torch.distributed as dist
    from bipymc_b200 import DreamMpi, targets, _lib
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    per = int(float(os.environ.get("C5_PER_GPU", "1250000")))
    k = int(os.environ.get("C5_K", "10"))
    gens = int(os.environ.get("C5_GENS", "10"))
    N, d, warm = per * world, 1000, 3
    t = targets.Gauss_100D(dim=d)
    np.random.seed(1)
    s = DreamMpi(t.ln_like, np.zeros(d), n_chains=N, seed=3, varepsilon=np.arange(d) + 1.0, history="none",
                 burnin_gen=10 ** 6, n_cr_gen=2, device=local, subpop_k=k, device_init=True)
    done = 0
    trend = []

    def gens_run(n):
        nonlocal done
        s.run_mcmc(N * (n + 1), _k_gen0=done)
        done += n

    def note():
        rh = s.rhat()
        trend.append({"generation": done, "rhat_max": float(np.max(rh)), "rhat_median": float(np.median(rh))})

    gens_run(warm)
    note()
    _lib.check(s._libh.bpm_profile(s._handle, 1))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gens_run(gens)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_k, n_k = (C.c_double * 8)(), (C.c_int64 * 8)()
    _lib.check(s._libh.bpm_profile_read(s._handle, ms_k, n_k))
    _lib.check(s._libh.bpm_profile(s._handle, 0))
    acc = s.acceptance_fraction
    note()
    gens_run(gens)
    note()
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    if rank == 0:
        kinds = ["split", "propose", "likelihood", "accept", "fused_phase", "cr_reduce"]
        kms = dict((kinds[i], ms_k[i] / max(1, n_k[i])) for i in range(6) if n_k[i])
        chains_per_launch = per / 2.0
        out = {"config": "C5: DREAM Gauss_1000D, %d GPUs x %d chains, sub-population mode k=%d" % (world, per, k),
               "n_chains": N, "dim": d, "generations_timed": gens,
               "chain_steps_per_s": N * gens / (float(ms.item()) * 1e-3), "ms_per_generation": float(ms.item()) / gens,
               "acceptance_fraction": acc, "rhat_trend": trend, "avg_launch_ms_by_kernel": kms,
               "redeals": getattr(s, "n_redeals", 0), "redeal_seconds_total": getattr(s, "redeal_seconds", 0.0),
               "torch_peak_alloc_GiB_per_gpu": mem}
        if "likelihood" in kms:
            tf = 2.0 * d * d * chains_per_launch / (kms["likelihood"] * 1e-3) / 1e12
            out["quadratic_form_fp64_tflops"] = tf
            out["frac_of_fp64_peak_36.5"] = tf / 36.5
            tot = sum(ms_k[i] for i in range(6))
            out["quadratic_form_share_of_kernel_time"] = ms_k[2] / tot
        print(json.dumps(out), flush=True)
    s.close()
    dist.barrier()
    dist.destroy_process_group()


def c4multi():
    """BASELINE configs[3] (launch under torchrun, 1 / 2 / 4 ranks): DREAM on the examples/ex_para_fit.py line
    fit (3 parameters, 50 data points, likelihood evaluated in-kernel), 10^5 chains in TOTAL sharded over the
    GPUs (the config as stated: strong scaling) and 10^5 chains PER GPU (weak scaling).  One rank: every
    generation of the call runs inside one persistent cooperative kernel; several ranks: per-phase kernels with
    accepted rows stored into the peer replicas and the peer-memory barrier / CR exchange between them."""
    import torch.distributed as dist
    from bipymc_b200 import DreamMpi, targets
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = targets.LineFit()
    gens, warm = 200, 20
    for label, N in (("1e5 chains in total", 100000 // world * world), ("1e5 chains per GPU", 100000 * world)):
        np.random.seed(1)
        s = DreamMpi(t.ln_like, [-0.8, 4.5, 0.2], n_chains=N, seed=3, varepsilon=1e-2, history="none",
                     burnin_gen=10 ** 6, n_cr_gen=10, device=local)
        s.run_mcmc(N * (warm + 1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run_mcmc(N * (gens + 1), _k_gen0=warm)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"config": "C4 linefit DREAM, %d GPU(s), %s" % (world, label), "n_gpus": world,
                              "n_chains": N, "chain_steps_per_s": N * gens / (float(ms.item()) * 1e-3),
                              "us_per_generation": 1e3 * float(ms.item()) / gens,
                              "acceptance_fraction": s.acceptance_fraction,
                              "path": "persistent cooperative kernel" if world == 1 else
                                      "per-phase kernels + peer-memory barrier (%s)" % ("on" if s._sync_on else "off")}),
                  flush=True)
        s.close()
        if world > 1:
            dist.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    if sys.argv[1:] == ["c4multi"]:
        c4multi()
        sys.exit(0)
    if sys.argv[1:] == ["c5full"]:
        c5full()
        sys.exit(0)
    if sys.argv[1:] == ["c5multi"]:
        c5multi()
        sys.exit(0)
    if "c5shape" in (sys.argv[1:] or ["c5shape"]):
        c5shape()
        if sys.argv[1:] == ["c5shape"]:
            sys.exit(0)
    main()
