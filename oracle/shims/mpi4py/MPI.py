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
