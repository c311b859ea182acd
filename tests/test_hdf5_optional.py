"""Cross-checks of the package's own HDF5 writer / reader (bipymc_b200/h5lite.py) against h5py / libhdf5, for the
reference's checkpoint layout (bipymc/chain.py:59-93, demc.py:198-233): one gzip dataset /chains/chain_id_<id> of
shape (T, dim) float64 per chain.  h5py is NOT in the build image, so these tests skip there -- the format is
pinned byte by byte in tests/test_h5lite_cpu.py instead, and the checkpoint tests (tests/test_native_gpu.py) run
through h5lite; everywhere h5py imports, these run and prove both directions."""
import numpy as np
import pytest

h5py = pytest.importorskip("h5py")
if not hasattr(h5py, "version"):            # the oracle's empty import shim, not the library
    pytest.skip("h5py is a stub", allow_module_level=True)

from bipymc_b200 import h5lite  # noqa: E402


def _chains(n, seed=0):
    rng = np.random.RandomState(seed)
    return [rng.randn(7 + i % 4, 3) for i in range(n)]


def test_h5py_reads_what_h5lite_wrote(tmp_path):
    f = str(tmp_path / "lite.h5")
    chains = _chains(300)
    big = np.random.RandomState(1).randn(7000, 40)
    with h5lite.File(f, "w") as h:
        for i, c in enumerate(chains):
            h.create_dataset("/chains/chain_id_%d" % i, data=c, compression="gzip")
        h["/chains"].attrs["b200_seed"] = np.uint64(31)
        h["/chains"].attrs["b200_p_cr"] = np.array([0.2, 0.3, 0.5])
        h.create_dataset("big", data=big, compression="gzip", chunks=(50, 40), shuffle=True)
        h.create_dataset("plain", data=np.arange(12, dtype=np.int32).reshape(3, 4))
        h.create_dataset("scalar", data=np.float64(3.5))
    with h5py.File(f, "r") as h:
        assert sorted(h["/chains"].keys()) == sorted("chain_id_%d" % i for i in range(300))
        for i, c in enumerate(chains):
            ds = h["/chains/chain_id_%d" % i]
            assert ds.shape == c.shape and ds.dtype == np.float64 and ds.compression == "gzip"
            assert np.array_equal(ds[:], c)
        assert int(h["/chains"].attrs["b200_seed"]) == 31
        assert np.array_equal(h["/chains"].attrs["b200_p_cr"], [0.2, 0.3, 0.5])
        assert np.array_equal(h["big"][:], big) and h["big"].chunks == (50, 40)
        assert np.array_equal(h["plain"][:], np.arange(12).reshape(3, 4))
        assert h["scalar"][()] == 3.5
    # and h5py can extend the file h5lite wrote
    with h5py.File(f, "a") as h:
        h.create_dataset("/chains/extra", data=np.ones(3))
    with h5lite.File(f, "r") as h:
        assert np.array_equal(h["/chains/extra"][:], np.ones(3)) and len(h["chains"]) == 301


def test_h5lite_reads_what_h5py_wrote(tmp_path):
    f = str(tmp_path / "py.h5")
    chains = _chains(300, seed=2)
    big = np.random.RandomState(3).randn(5000, 16)
    with h5py.File(f, "w") as h:
        for i, c in enumerate(chains):
            h.create_dataset("/chains/chain_id_%d" % i, data=c, compression="gzip")
        h["/chains"].attrs["b200_seed"] = np.uint64(77)
        h["/chains"].attrs["b200_p_cr"] = np.array([0.1, 0.9])
        h.create_dataset("big", data=big, compression="gzip", shuffle=True, fletcher32=True)
        h.create_dataset("contig", data=np.arange(6.0))
    with h5lite.File(f, "r") as h:
        assert h["chains"].keys() == sorted("chain_id_%d" % i for i in range(300))
        for i, c in enumerate(chains):
            ds = h["/chains/chain_id_%d" % i]
            assert ds.shape == c.shape and ds.dtype == np.float64 and ds.compression == "gzip"
            assert np.array_equal(ds[:], c)
        assert int(h["/chains"].attrs["b200_seed"]) == 77
        assert np.array_equal(h["/chains"].attrs["b200_p_cr"], [0.1, 0.9])
        assert np.array_equal(h["big"][:], big)
        assert np.array_equal(h["contig"][:], np.arange(6.0))


def test_free_standing_chain_h5_round_trip_across_implementations(tmp_path, monkeypatch):
    from bipymc_b200.chain import McmcChain
    np.random.seed(0)
    c = McmcChain(np.zeros(3), varepsilon=1e-2, global_id=7)
    for _ in range(5):
        c.append_sample(np.random.randn(3))
    for writer, reader in (("1", "0"), ("0", "1")):
        f = str(tmp_path / ("one_%s.h5" % writer))
        monkeypatch.setenv("BIPYMC_B200_H5LITE", writer)
        c.write_chain_h5(f)
        monkeypatch.setenv("BIPYMC_B200_H5LITE", reader)
        d = McmcChain(np.zeros(3), varepsilon=0.0, global_id=7)
        d.read_chain_h5(f)
        assert np.array_equal(d.chain, c.chain)
