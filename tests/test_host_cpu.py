"""CPU-side checks: the C-ABI library loads without a GPU and exports every symbol the
header declares; the host mirrors of the device RNG pass known-answer tests; the host
logic of the drop-in classes (targets, util, McmcChain, history store) matches the
oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bipymc_b200 import _lib, targets, util
from bipymc_b200.chain import McmcChain
from oracle import targets as otargets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "bipymc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bpm_[a-z0-9_]+)\s*\(", txt)) - {"bpm_lnl_fn"})


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.bpm_version() >= 100


def test_struct_layouts_match_header():
    # sizes the C compiler would produce for the header's structs (LP64)
    assert C.sizeof(_lib.Config) == 12 * 4 + 5 * 8 + 8
    assert C.sizeof(_lib.State) == 8 * 8
    assert C.sizeof(_lib.Replay) == 8 + 9 * 8
    assert C.sizeof(_lib.TraceOut) == 3 * 8


def philox(ctr, key):
    lib = _lib.load()
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    assert lib.bpm_test_philox(c, k, o) == 0
    return list(o)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert philox([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("n", [4, 5, 10, 11, 64, 1000, 4097, 100000])
def test_shuffle_permutation_is_a_permutation(n):
    lib = _lib.load()
    seen = []
    for g in range(3):
        out = (C.c_int32 * n)()
        assert lib.bpm_test_permutation(42, g, n, out) == 0
        a = np.array(out[:])
        assert np.array_equal(np.sort(a), np.arange(n))
        seen.append(a)
    if n >= 64:
        assert not np.array_equal(seen[0], seen[1])      # a new shuffle every generation


def test_shuffle_split_is_balanced_and_mixing():
    """Over many generations every chain should land in half 'a' about half the time and
    any two chains should share a half about half the time (np.random.shuffle's law)."""
    lib = _lib.load()
    n, G = 50, 4000
    in_a = np.zeros((G, n), dtype=bool)
    for g in range(G):
        out = (C.c_int32 * n)()
        lib.bpm_test_permutation(7, g, n, out)
        in_a[g, np.array(out[:n // 2])] = True
    frac = in_a.mean(axis=0)
    assert np.all(np.abs(frac - 0.5) < 0.05)
    same = (in_a[:, :1] == in_a).mean(axis=0)[1:]
    assert np.all(np.abs(same - (n / 2 - 1) / (n - 1)) < 0.06)


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from bipymc_b200 import DreamMpi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DreamMpi(targets.Banana_2D().ln_like, [0.0, 0.0], n_chains=8)


def test_host_targets_match_oracle_targets():
    rs = np.random.RandomState(0)
    b, ob = targets.Banana_2D(), otargets.Banana2D()
    g, og = targets.BimodeGauss_2D(), otargets.BimodeGauss2D()
    for _ in range(50):
        y = rs.randn(2) * 2
        np.testing.assert_allclose(b.ln_like(y), ob.ln_like(y), rtol=0, atol=1e-12)
        y = rs.rand(2) * 3 - 0.5
        np.testing.assert_allclose(g.ln_like(y), og.ln_like(y), rtol=0, atol=1e-11)
    for dim in (7, 100):
        t, ot = targets.Gauss_100D(dim=dim), otargets.GaussND(dim=dim)
        for _ in range(10):
            y = rs.randn(dim)
            np.testing.assert_allclose(t.ln_like(y), ot.ln_like(y), rtol=0, atol=1e-10)
    t, ot = targets.Gauss_100D(dim=1000), otargets.GaussND(dim=1000, use_logpdf=True)
    assert not t.log_of_pdf
    np.testing.assert_allclose(t.ln_like(np.zeros(1000)), -3531.8834, atol=1e-3)
    np.testing.assert_allclose(t.ln_like(np.zeros(1000)), ot.ln_like(np.zeros(1000)), atol=1e-8)
    lf, olf = targets.LineFit(), otargets.LineFit()
    for th in ([-0.8, 4.5, 0.2], [-1.0, 4.0, -0.5], [0.6, 4.0, 0.0]):
        assert lf.ln_like(th) == olf.ln_like(th)
    np.testing.assert_allclose(targets.Gauss_100D().ln_like(np.zeros(100)), -241.4137423, atol=1e-6)
    ef, oef = targets.ExpFit(), otargets.ExpFit()
    assert np.array_equal(ef.t, oef.t) and np.array_equal(ef.y, oef.y)
    for th in ([12.0, 1.5, 0.6, 1e-3, 2e-3], [20.0, 1.0, -0.3, 0.5, 0.5], [0.5, 1.0, 0.1, 0.0, 0.1],
               [10.0, 1.0, 0.1, 0.0, 1.5]):
        assert ef.ln_like(th) == oef.ln_like(th)
    assert ef.ln_like([0.5, 1.0, 0.1, 0.0, 0.1]) == -np.inf       # tau outside its prior box


def test_var_ball_matches_numpy_stream():
    for v in (1e-6, 0.3):
        np.random.seed(3)
        a = np.array([np.random.multivariate_normal(np.zeros(4), np.eye(4) * v, size=1)[0]
                      for _ in range(5)])
        np.random.seed(3)
        assert np.array_equal(a, util.var_ball_batch(v, 4, 5))
        np.random.seed(3)
        assert np.array_equal(a, np.array([util.var_ball(v, 4) for _ in range(5)]))
    v = np.array([1e-2, 3.0, 0.5, 1e-6])
    np.random.seed(4)
    a = np.array([np.random.multivariate_normal(np.zeros(4), np.eye(4) * v, size=1)[0] for _ in range(5)])
    np.random.seed(4)
    assert np.array_equal(a, util.var_ball_batch(v, 4, 5))
    st = np.random.get_state()[1].copy()
    assert util.var_ball(0.0, 3) == 0.0 and util.var_box(0.0, 3) == 0.0     # util.py: no RNG consumed
    assert np.array_equal(st, np.random.get_state()[1])


def test_free_standing_mcmc_chain_behaves_like_reference():
    np.random.seed(5)
    ch = McmcChain([1.0, 2.0], varepsilon=1e-6, global_id=3)
    np.random.seed(5)
    want = np.array([1.0, 2.0]) + np.sqrt(1e-6) * np.random.standard_normal(2)
    assert np.array_equal(ch.current_pos, want)
    assert ch.chain.shape == (1, 2) and ch.dim == 2 and ch.chain_len == 1 and ch.global_id == 3
    ch.append_sample([3.0, 4.0])
    assert ch.chain_len == 2 and np.array_equal(ch[-1], [3.0, 4.0]) and ch[0:1].shape == (1, 2)
    ch.pop_sample()
    assert ch.chain_len == 1
    with pytest.raises(AssertionError):
        ch.chain = np.zeros((3, 5))
    ch.set_t_kernel(np.array([[0.0, 1.0], [1.0, 0.0]]))            # chain.py:31-43
    w, _ = ch.t_kernel_eig()
    assert sorted(np.round(w.real, 12)) == [-1.0, 1.0]
    with pytest.raises(AssertionError):
        ch.set_t_kernel(np.zeros((3, 3)))
    assert ch.auto_corr(1) is None


def test_history_store_chunks_and_super_chain_layout():
    import torch
    from bipymc_b200.demc import HistoryStore
    hs = HistoryStore(n_local=3, dim=2, ld=2, device=torch.device("cpu"), chunk_bytes=3 * 2 * 8 * 4)
    x0 = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    hs.set_initial(x0)
    t = 1
    while t < 11:
        base, avail = hs.reserve(11 - t)
        take = min(avail, 11 - t)
        # emulate the kernel: row index `hs.length + r` of the flat array
        for r in range(take):
            addr = base + (hs.length + r) * hs.row_bytes
            last = hs.chunks[-1]
            off = (addr - last.data_ptr()) // hs.row_bytes
            assert 0 <= off < last.shape[0]
            last[off] = x0 + (t + r)
        hs.advance(take)
        t += take
    full = hs.tensor()
    assert full.shape == (11, 3, 2)
    for r in range(11):
        assert torch.equal(full[r], x0 + r)
    # flattened (T*N, d) is the reference's interlaced super chain: row t*N + i
    sc = full.reshape(-1, 2).numpy()
    assert np.array_equal(sc[1::3], full[:, 1, :].numpy())
    hs2 = HistoryStore(3, 2, 2, torch.device("cpu"), policy="none")
    hs2.set_initial(x0)
    assert hs2.reserve(5) == (None, 5)
    hs2.advance(5)
    assert hs2.length == 6 and hs2.stored == 1


def test_history_store_pending_row_and_contiguous_reserve():
    """Host bookkeeping of the lazy protocol (bpm_state.pending): the newest row of a generation is written
    one generation later (or by bpm_flush) at flat_base() + (length - 1) rows -- which must stay inside the LAST
    chunk, so the owner flushes before reserve() opens a new one (will_grow()); the before-read hook runs
    whenever the stored rows are read; replay steps get every stored row in ONE block (reserve_contiguous)."""
    import torch
    from bipymc_b200.demc import HistoryStore
    hs = HistoryStore(n_local=2, dim=2, ld=2, device=torch.device("cpu"), chunk_bytes=2 * 2 * 8 * 3)
    x0 = torch.zeros((2, 2), dtype=torch.float64)
    hs.set_initial(x0)
    calls = []
    hs._before_read = lambda: calls.append(hs.length)
    pending = None                      # row index a lazy generation left unwritten
    for g in range(1, 9):
        if pending is not None and hs.will_grow():
            # flush: the pending row lives in the chunk that is full now
            base = hs.flat_base()
            off = (base + pending * hs.row_bytes - hs.chunks[-1].data_ptr()) // hs.row_bytes
            assert 0 <= off < hs.chunks[-1].shape[0]
            hs.chunks[-1][off] = x0 + pending
            pending = None
        base, avail = hs.reserve(1)
        assert avail >= 1
        if pending is not None:         # the next generation's proposal stage writes the pending row
            off = (base + pending * hs.row_bytes - hs.chunks[-1].data_ptr()) // hs.row_bytes
            assert 0 <= off < hs.chunks[-1].shape[0], "pending row must sit in the chunk reserve() returned"
            hs.chunks[-1][off] = x0 + pending
        pending = hs.length             # this generation's row stays pending
        hs.advance(1)
    base = hs.flat_base()
    off = (base + pending * hs.row_bytes - hs.chunks[-1].data_ptr()) // hs.row_bytes
    hs.chunks[-1][off] = x0 + pending
    full = hs.tensor()
    assert calls and full.shape[0] == 9
    for r in range(9):
        assert torch.equal(full[r], x0 + r)
    # one block for replay steps: base is a real array base and old rows are preserved
    hs3 = HistoryStore(2, 2, 2, torch.device("cpu"), chunk_bytes=2 * 2 * 8 * 2)
    hs3.set_initial(x0 + 7)
    for g in range(5):
        base, avail = hs3.reserve_contiguous(1)
        assert len(hs3.chunks) == 1 and base == hs3.chunks[0].data_ptr() and avail >= 1
        hs3.chunks[0][hs3.length] = x0 + 7 + hs3.length
        hs3.advance(1)
    t3 = hs3.tensor()
    assert t3.shape[0] == 6 and all(torch.equal(t3[r], x0 + 7 + r) for r in range(6))
