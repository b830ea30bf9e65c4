"""File-based stand-in for mpi4py (TEST INFRASTRUCTURE): lets the reference's MPSCoefParallel run as several plain
Python processes in a container without MPI.  Rank / size / mailbox directory come from FAKE_MPI_RANK,
FAKE_MPI_SIZE, FAKE_MPI_DIR.  Only what pytdscf uses is implemented: pickled send/recv with tags, barrier, bcast,
scatter, gather, allgather, allreduce(LOR), Abort."""
from . import MPI  # noqa: F401
