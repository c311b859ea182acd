"""Fake single-rank communicator: just enough of mpi4py.MPI for bipymc's
DeMcMpi/DreamMpi (demc.py:93,116,135,145-151) to run with comm.size == 1."""
import numpy as np

DOUBLE = "DOUBLE"
INT = "INT"
ANY_SOURCE = -1
ANY_TAG = -1


class Status(object):
    pass


class _Comm(object):
    size = 1
    rank = 0

    def Get_size(self):
        return 1

    def Get_rank(self):
        return 0

    def Barrier(self):
        pass

    def Allgather(self, send, recv):
        src = np.asarray(send[0])
        dst = recv[0]
        dst[...] = src.reshape(dst.shape)

    def send(self, *a, **k):
        raise RuntimeError("single-rank shim: send is never reached")

    def recv(self, *a, **k):
        raise RuntimeError("single-rank shim: recv is never reached")


COMM_WORLD = _Comm()


class ShmComm(object):
    """P-rank communicator over multiprocessing shared memory, for timing the UNMODIFIED reference on
    all host cores (mpi4py / mpirun are not in the image).  Implements what
    DeMcMpi / DreamMpi call on the sampling path: Allgather of equal-sized float64 / int blocks
    (demc.py:93,116,145-148) and Barrier (demc.py:135,151).  Created by oracle/ref_runner.py in each
    forked rank and passed as mpi_comm."""
    def __init__(self, rank, size, shared, barrier):
        self.rank, self.size = rank, size
        self._shared, self._barrier = shared, barrier

    def Get_size(self):
        return self.size

    def Get_rank(self):
        return self.rank

    def Barrier(self):
        self._barrier.wait()

    def Allgather(self, send, recv):
        src = np.ascontiguousarray(send[0]).reshape(-1)
        dst = recv[0]
        n = src.size
        buf = np.frombuffer(self._shared, dtype=np.float64)
        if n * self.size > buf.size:
            raise RuntimeError("ShmComm: shared buffer too small")
        buf[self.rank * n:(self.rank + 1) * n] = src.astype(np.float64)
        self._barrier.wait()
        dst.reshape(-1)[...] = buf[:n * self.size].astype(dst.dtype)
        self._barrier.wait()

    def send(self, *a, **k):
        raise RuntimeError("ShmComm: point-to-point is not on the sampling path")

    def recv(self, *a, **k):
        raise RuntimeError("ShmComm: point-to-point is not on the sampling path")
