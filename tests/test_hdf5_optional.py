"""HDF5 checkpoint layout of the reference (bipymc/chain.py:59-93, demc.py:198-233): one gzip dataset
/chains/chain_id_<id> of shape (T, dim) float64 per chain.  h5py is NOT in this image, so these tests skip
here and the HDF5 branch of save_state / load_state / McmcChain.write_chain_h5 has never executed in this
repo (DESIGN.md section 8 says so); everywhere h5py imports they run and pin the layout."""
import os

import numpy as np
import pytest

h5py = pytest.importorskip("h5py")


def test_free_standing_chain_h5_round_trip(tmp_path):
    from bipymc_b200.chain import McmcChain
    np.random.seed(0)
    c = McmcChain(np.zeros(3), varepsilon=1e-2, global_id=7)
    for _ in range(5):
        c.append_sample(np.random.randn(3))
    f = str(tmp_path / "one.h5")
    c.write_chain_h5(f)
    with h5py.File(f, "r") as h:
        ds = h["/chains/chain_id_7"]
        assert ds.shape == (6, 3) and ds.dtype == np.float64 and ds.compression == "gzip"
        assert np.array_equal(ds[:], c.chain)
    d = McmcChain(np.zeros(3), varepsilon=0.0, global_id=7)
    d.read_chain_h5(f)
    assert np.array_equal(d.chain, c.chain)


@pytest.mark.gpu
def test_sampler_checkpoint_is_the_reference_layout(tmp_path):
    from bipymc_b200 import DreamMpi, targets
    np.random.seed(1)
    a = DreamMpi(targets.Banana_2D().ln_like, [0.0, 0.0], n_chains=12, seed=3, n_cr_gen=2, burnin_gen=10)
    a.run_mcmc(12 * 9)
    f = str(tmp_path / "ckpt.h5")
    a.save_state(f)
    with h5py.File(f, "r") as h:
        assert sorted(h["/chains"].keys()) == sorted("chain_id_%d" % i for i in range(12))
        for i in range(12):
            ds = h["/chains/chain_id_%d" % i]
            assert ds.shape == (9, 2) and ds.compression == "gzip"
            assert np.array_equal(ds[:], a.am_chains[i].chain)
    np.random.seed(1)
    b = DreamMpi(targets.Banana_2D().ln_like, [0.0, 0.0], n_chains=12, seed=99, warm_start=True, h5_file=f, dim=2,
                 n_cr_gen=2, burnin_gen=10)
    assert np.array_equal(b.super_chain, a.super_chain)
    a.run_mcmc(12 * 5, _k_gen0=8)
    b.run_mcmc(12 * 5, _k_gen0=8)
    assert np.array_equal(b.super_chain, a.super_chain)      # Philox seed and CR state travelled with the file
