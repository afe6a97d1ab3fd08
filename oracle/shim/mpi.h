// oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// In-process stand-in for the MPI subset that the reference
// (/root/reference/src/*.cpp + deps/2Decomp_C) uses, so that the UNMODIFIED reference sources can be
// compiled and run in this image, which has no MPI library.  Each "rank" is a thread of one process
// (MIF_SHIM_NP environment variable, default 1); messages are eager-buffered copies through an
// in-memory mailbox, collectives are built on those.  Semantics follow the MPI standard for exactly
// the calls listed here; anything else is absent on purpose so that new uses fail at compile time.
//
// Call sites in the reference this covers:
//   halos            src/StaggeredTensor.cpp:60-165   (MPI_Isend / MPI_Recv / MPI_Wait, contiguous types)
//   pencil transposes deps/2Decomp_C/Transpose*.cpp   (MPI_Alltoallv on Cartesian sub-communicators)
//   topology         deps/2Decomp_C/C2Decomp.cpp:34-62,106-117 (MPI_Cart_create/sub/coords/shift)
//   reductions       src/Norms.cpp:122-148, src/PressureEquation.cpp:299-333 (MPI_Send / MPI_Recv of scalars)
//   output           src/VTKDatExport.cpp (MPI_File_*, MPI_Allgather, MPI_Gather(v)), src/main.cpp (MPI_Bcast)
#ifndef MIF_SHIM_MPI_H
#define MIF_SHIM_MPI_H

#include <cstddef>

struct mifshim_comm;
typedef mifshim_comm *MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_Op;
typedef long long MPI_Offset;
struct mifshim_file;
typedef mifshim_file *MPI_File;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_ERR_OTHER 15
#define MPI_PROC_NULL (-2)
#define MPI_ANY_TAG (-1)
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

MPI_Comm mifshim_comm_world();
MPI_Comm mifshim_comm_self();
#define MPI_COMM_WORLD (mifshim_comm_world())
#define MPI_COMM_SELF (mifshim_comm_self())

// Predefined datatypes: the handle is an index into the shim's size table.
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_RDWR 8

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize();
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime();
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);
int MPI_Type_size(MPI_Datatype type, int *size);

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *status);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Waitall(int count, MPI_Request *reqs, MPI_Status *statuses);

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, int rcount, MPI_Datatype rtype,
               int root, MPI_Comm comm);
int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, const int *rcounts,
                const int *displs, MPI_Datatype rtype, int root, MPI_Comm comm);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype stype, void *rbuf, int rcount,
                  MPI_Datatype rtype, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype, void *rbuf,
                  const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Ialltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype, void *rbuf,
                   const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm,
                   MPI_Request *req);

int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder,
                    MPI_Comm *newcomm);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_sub(MPI_Comm comm, const int *remain_dims, MPI_Comm *newcomm);
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest);

int MPI_File_open(MPI_Comm comm, const char *filename, int amode, MPI_Info info, MPI_File *fh);
int MPI_File_close(MPI_File *fh);
int MPI_File_delete(const char *filename, MPI_Info info);
int MPI_File_write_at(MPI_File fh, MPI_Offset offset, const void *buf, int count, MPI_Datatype type,
                      MPI_Status *status);

// Launcher hook: run `fn(argc, argv)` once per rank on `np` threads; returns the maximum return code.
// Used by oracle/shim/mpi_launcher.cpp (tests' `main` is renamed to mifshim_user_main with objcopy).
int mifshim_run(int np, int (*fn)(int, char **), int argc, char **argv);

#endif  // MIF_SHIM_MPI_H
