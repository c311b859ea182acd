"""McmcChain -- drop-in surface of bipymc/chain.py:9-132 over device-resident storage.

In the reference every chain owns a numpy array re-allocated by ``np.vstack`` on each
append (chain.py:51-54).  Here the whole population's history lives in one device tensor
``[T][n_local][ld]`` (see ``HistoryStore``); a ``McmcChain`` handed out by a sampler's
``am_chains`` is a lazy view of column ``i`` of that tensor, materialised on the host
only when ``.chain`` is read.  A free-standing ``McmcChain(theta_0, ...)`` (no sampler)
behaves exactly like the reference's host object, so user code that builds chains by
hand keeps working.
"""
import numpy as np

from .util import var_ball


class McmcChain(object):
    def __init__(self, theta_0, varepsilon=1e-6, global_id=0, mpi_comm=None, mpi_rank=None,
                 _store=None, _local_index=None):
        assert isinstance(global_id, (int, np.integer))
        assert global_id >= 0
        self.global_id = int(global_id)
        self._store = _store
        self._li = _local_index
        if _store is not None:
            self._dim = _store.dim
            self._chain = None
            return
        assert np.all(np.asarray(varepsilon) >= 0.0)
        theta_0_flat = np.asarray(theta_0, dtype=float).flatten()
        self._dim = len(theta_0_flat)
        theta_0_flat = theta_0_flat + var_ball(varepsilon, self._dim)
        self._chain = np.array([theta_0_flat])

    # -- storage -------------------------------------------------------------------
    @property
    def chain(self):
        """(T, dim) float64 history of this chain (host copy when device-backed)."""
        if self._store is not None:
            return self._store.chain_host(self._li)
        return self._chain

    @chain.setter
    def chain(self, input_chain):
        input_chain = np.asarray(input_chain, dtype=float)
        assert input_chain.shape[1] == self._dim
        if self._store is not None:
            self._store.set_chain_host(self._li, input_chain)
        else:
            self._chain = input_chain

    @property
    def current_pos(self):
        if self._store is not None:
            return self._store.current_host(self._li)
        return self.chain[-1, :]

    @property
    def chain_len(self):
        if self._store is not None:
            return self._store.length
        return self.chain.shape[0]

    @property
    def dim(self):
        return self._dim

    def append_sample(self, theta_new):
        theta_new = np.asarray(theta_new)
        assert theta_new.shape[0] == self._dim
        if self._store is not None:
            raise RuntimeError("device-backed chains are appended by the sampler's generation "
                               "kernels; append_sample is only valid on a free-standing McmcChain")
        self._chain = np.vstack((self._chain, theta_new))

    def pop_sample(self):
        if self._store is not None:
            raise RuntimeError("pop_sample is only valid on a free-standing McmcChain")
        self._chain = self._chain[:-1, :]

    # -- HDF5 (chain.py:59-93): dataset /chains/chain_id_<global_id>, gzip, (T, dim) f64 ----
    def write_chain_h5(self, h5_file):
        from . import h5lite       # h5py where it imports, else the package's own HDF5 writer / reader
        h5py = h5lite.get_h5()
        name = "/chains/chain_id_" + str(self.global_id)
        if isinstance(h5_file, str):
            with h5py.File(h5_file, "w") as h5f:
                h5f.create_dataset(name, data=self.chain, compression="gzip")
        elif h5lite.is_file(h5_file):
            if name in h5_file:
                del h5_file[name]
            h5_file.create_dataset(name, data=self.chain, compression="gzip")
        else:
            raise RuntimeError

    def read_chain_h5(self, h5_file, c_id=None):
        from . import h5lite
        h5py = h5lite.get_h5()
        name = "/chains/chain_id_" + str(self.global_id)
        if isinstance(h5_file, str):
            with h5py.File(h5_file, "r") as h5f:
                self.load_chain_state(h5f[name][:])
        elif h5lite.is_file(h5_file):
            self.load_chain_state(h5_file[name][:])
        else:
            raise RuntimeError

    def load_chain_state(self, chain_state):
        self.chain = chain_state

    # chain.py:31-49: transition-kernel helpers nothing on the sampler path calls; kept so the
    # McmcChain surface is complete
    def set_t_kernel(self, t_kernel):
        t_kernel = np.asarray(t_kernel)
        assert t_kernel.shape[1] == self.chain.shape[1]
        assert t_kernel.shape[1] == t_kernel.shape[0]          # must be square
        self.t_kernel = t_kernel

    def t_kernel_eig(self):
        return np.linalg.eig(self.t_kernel)

    def apply_t_kernel(self, apply_new_state=True):
        new_state = np.dot(self.t_kernel, self.chain[:-1])
        if apply_new_state:
            self.append_sample(new_state)
        return new_state

    def auto_corr(self, lag):
        pass

    def __getitem__(self, get_index):
        if isinstance(get_index, slice):
            return self.chain[get_index]
        return self.chain[get_index, :]
