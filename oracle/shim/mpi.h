// oracle/shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
//
// Single-process stand-in for the MPI declarations that the reference's XMPI.h (S/preloop/utilities/XMPI.h) uses inline,
// so that reference sources which ask XMPI::rank()/nproc() (Connectivity.cpp:125-181) can be run "as rank r of n" inside
// one process: MPI_Comm_rank / MPI_Comm_size return the two globals below, which the harness sets.  No message ever
// moves: the data-moving calls fail loudly.  MPI is not in this image.
#pragma once
#include <cstdio>
#include <cstdlib>

extern int ax_mpi_rank, ax_mpi_size;      // defined by the harness (oracle/ref_connectivity.cpp)

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef struct { int dummy; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_FLOAT 3
#define MPI_DOUBLE 4
#define MPI_C_FLOAT_COMPLEX 5
#define MPI_C_DOUBLE_COMPLEX 6
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_SUCCESS 0

inline int ax_mpi_unsupported(const char *what) {
    std::fprintf(stderr, "oracle/shim/mpi.h: %s is not available in the single-process stand-in\n", what);
    std::abort();
    return 1;
}
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = ax_mpi_rank; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = ax_mpi_size; return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::exit(code); return MPI_SUCCESS; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return ax_mpi_unsupported("MPI_Bcast"); }
inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return ax_mpi_unsupported("MPI_Isend"); }
inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return ax_mpi_unsupported("MPI_Irecv"); }
inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return ax_mpi_unsupported("MPI_Waitall"); }
inline int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) { return ax_mpi_unsupported("MPI_Allreduce"); }
inline int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm) { return ax_mpi_unsupported("MPI_Gather"); }
inline int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm) { return ax_mpi_unsupported("MPI_Gatherv"); }
inline int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm) { return ax_mpi_unsupported("MPI_Allgather"); }
inline int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm) { return ax_mpi_unsupported("MPI_Allgatherv"); }
