"""world_size-2 checks of the host-side multi-rank logic over gloo (no GPU needed):
shard ownership (demc.py:39), the generation count of the reference's while-loop
(demc.py:78-79), the torch.distributed communicator facade and the population /
per-chain all-gather bookkeeping used between half-phases (demc.py:93,116)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Shim(object):
    """Just enough of a sampler for the host-side methods under test."""
    def __init__(self, n_chains, comm):
        self.n_chains, self.comm = n_chains, comm
        self.rank_chain_ids = np.array_split(np.array(range(n_chains)), comm.size)[comm.rank]


def _worker(rank, world, port, n_chains, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bipymc_b200.demc import DeMcMpi, _default_comm
        comm = _default_comm()
        assert comm.size == world and comm.rank == rank and comm.Get_size() == world
        comm.Barrier()
        s = _Shim(n_chains, comm)
        lo, hi = int(s.rank_chain_ids[0]), int(s.rank_chain_ids[-1]) + 1
        # population replica: every rank fills only its own rows, the gather completes it
        d = 5
        X = torch.full((n_chains, d), -1.0, dtype=torch.float64)
        X[lo:hi] = torch.arange(lo, hi, dtype=torch.float64)[:, None] * 10 + torch.arange(d, dtype=torch.float64)
        DeMcMpi._allgather_rows(s, X, lo, hi)
        want = torch.arange(n_chains, dtype=torch.float64)[:, None] * 10 + torch.arange(d, dtype=torch.float64)
        assert torch.equal(X, want)
        v = torch.zeros(n_chains, dtype=torch.float64)
        v[lo:hi] = torch.arange(lo, hi, dtype=torch.float64) + 0.5
        DeMcMpi._allgather_rows(s, v, lo, hi)
        assert torch.equal(v, torch.arange(n_chains, dtype=torch.float64) + 0.5)
        # generation count == the reference's loop: j += n_local while j < int((n - N) / size)
        for n in (0, n_chains, n_chains + 1, 7 * n_chains, 7 * n_chains + 3, 50 * n_chains):
            j, gens, limit = 0, 0, int((n - n_chains) / world)
            while j < limit:
                j += len(s.rank_chain_ids)
                gens += 1
            assert DeMcMpi._n_generations(s, n) == gens, (n, gens)
        # uneven shards are refused up front (the reference hangs there: per-rank j counters, demc.py:79,93) --
        # the check sits before any device work, so it runs on this CPU-only box
        if n_chains % world != 0:
            try:
                DeMcMpi(lambda th: 0.0, np.zeros(2), n_chains=n_chains)
                raise AssertionError("uneven shards were accepted")
            except ValueError as e:
                assert "divisible" in str(e)
        # an mpi4py-style communicator must describe the torch.distributed world
        class _Fake(object):
            def __init__(self, rank, size):
                self.rank, self.size = rank, size

            def Get_size(self):
                return self.size

            def Get_rank(self):
                return self.rank
        from bipymc_b200.demc import _resolve_comm
        assert _resolve_comm(_Fake(rank, world)).size == world
        try:
            _resolve_comm(_Fake(rank, world + 1))
            raise AssertionError("mismatching communicator was accepted")
        except RuntimeError as e:
            assert "disagree" in str(e)
        # ownership lookup (demc.py:327-338)
        for c in (0, lo, hi - 1, n_chains - 1):
            r = DeMcMpi.get_chain_rank(s, c)
            assert c in np.array_split(np.array(range(n_chains)), world)[r]
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [16, 17])
def test_world2_gloo_host_logic(n_chains):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chains, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
