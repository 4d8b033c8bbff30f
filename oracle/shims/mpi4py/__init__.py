"""TEST INFRASTRUCTURE ONLY. Single-process stand-in for `mpi4py`, which the
reference imports unconditionally (lattice/dispatch.py:6) although the
elemental path never communicates."""


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, obj, root=0):
        return obj

    def Barrier(self):
        pass


class _MPI:
    COMM_WORLD = _Comm()


MPI = _MPI()
