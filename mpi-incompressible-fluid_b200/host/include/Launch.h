// Launch.h -- what MPI_Init / MPI_Comm_rank / MPI_Comm_size / MPI_Finalize give the reference's drivers
// (src/main.cpp:50-56,176; test/full_test.cpp:24-28), without linking an MPI: one process per GPU is started by
// scripts/mifrun (or by any mpirun / srun / torchrun -- their environment variables are understood), and the ranks
// find each other through libmifgpu's communicator id, which rank 0 publishes in a rendezvous file.
#ifndef MIF_LAUNCH_H
#define MIF_LAUNCH_H

namespace mif {

// Rank and number of processes of this job: MIF_RANK / MIF_WORLD_SIZE (scripts/mifrun), else the variables of Open MPI
// (OMPI_COMM_WORLD_*), MPICH / PMI (PMI_*), Slurm (SLURM_PROCID / SLURM_NTASKS) or torchrun (RANK / WORLD_SIZE);
// 0 and 1 when none is set.
int launch_rank();
int launch_size();
// Rank among the processes of this node (selects the GPU when MIFGPU_DEVICE is not set).
int launch_local_rank();

}  // namespace mif

#endif  // MIF_LAUNCH_H
