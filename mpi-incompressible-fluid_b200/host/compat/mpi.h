// mpi.h -- the handful of MPI calls the reference's drivers make (src/main.cpp:50-52,80-86,107-176,
// test/full_test.cpp:26-28,116-138,187, test/pressure_test_mixed.cpp), served by the host layer so that those sources
// compile unchanged without an MPI installation.  One process per GPU is started by scripts/mifrun (or mpirun / srun /
// torchrun); rank and size come from the launcher's environment (Launch.h); the data exchange of the solver itself happens
// inside libmifgpu over NCCL.  File existence checks and deletions are plain POSIX.  Collectives that carry data between
// processes are only provided in the forms the drivers use: a broadcast of values every rank already has (all ranks parse
// the same input file unless MIF_NO_DISTRIBUTED_FS is defined, which this stand-in does not support) and barriers around
// diagnostic output.
#ifndef MIF_COMPAT_MPI_H
#define MIF_COMPAT_MPI_H

#include <chrono>
#include <cstdio>

#include "Launch.h"

typedef int MPI_Comm;
typedef int MPI_Info;
typedef int MPI_Datatype;
typedef struct mif_mpi_file_ { std::FILE *handle; } *MPI_File;
enum { MPI_COMM_WORLD = 0, MPI_COMM_SELF = 1 };
enum { MPI_SUCCESS = 0, MPI_ERR_NO_SUCH_FILE = 37 };
enum { MPI_INFO_NULL = 0 };
enum { MPI_MODE_RDONLY = 2 };
enum { MPI_DOUBLE = 8 };

inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm comm, int *rank) {
  *rank = (comm == MPI_COMM_SELF) ? 0 : mif::launch_rank();
  return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (comm == MPI_COMM_SELF) ? 1 : mif::launch_size();
  return MPI_SUCCESS;
}
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }  // only orders diagnostic output in the drivers
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }  // every rank holds the values already
inline int MPI_File_open(MPI_Comm, const char *name, int, MPI_Info, MPI_File *file) {
  std::FILE *f = std::fopen(name, "rb");
  if (!f) return MPI_ERR_NO_SUCH_FILE;
  *file = new mif_mpi_file_{f};
  return MPI_SUCCESS;
}
inline int MPI_File_close(MPI_File *file) {
  if (file && *file) {
    std::fclose((*file)->handle);
    delete *file;
    *file = nullptr;
  }
  return MPI_SUCCESS;
}
inline int MPI_File_delete(const char *name, MPI_Info) { return std::remove(name) == 0 ? MPI_SUCCESS : MPI_ERR_NO_SUCH_FILE; }

#endif  // MIF_COMPAT_MPI_H
