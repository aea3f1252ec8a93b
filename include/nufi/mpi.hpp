// nufi/mpi.hpp -- nufi::mpi::{programme, comm_size, comm_rank, allreduce_add, allgatherv} with the reference's
// interface (nufi/mpi.hpp:33-171), so that bin/test_nufi_gpu_{2,3}d.cpp compile against this include tree alone.
//
// Two flavours, chosen at compile time:
//   * <mpi.h> is on the include path (and NUFI_B200_NO_MPI is not defined): thin throwing wrappers over the real MPI.
//   * otherwise: a SINGLE-RANK stand-in.  The target of libnufi_b200 is one 8 x B200 NVSwitch box driven either by one
//     process (cuda_scheduler owns every visible GPU, the reference's own intra-process mode) or by one process per GPU
//     through torch.distributed/NCCL (numericalflowiteration_b200/distributed.py); neither needs MPI.  With one rank
//     MPI_Allreduce(IN_PLACE, SUM) is the identity, which is what the stand-in implements; only the handful of MPI
//     names the reference drivers use are defined (bin/test_nufi_gpu_3d.cpp:37-47, 74-75, 158, 195, 232).
#ifndef NUFI_B200_NUFI_MPI_HPP
#define NUFI_B200_NUFI_MPI_HPP

#include <cstring>
#include <exception>
#include <iostream>
#include <stdexcept>
#include <string>

#if !defined(NUFI_B200_NO_MPI) && defined(__has_include)
#if __has_include(<mpi.h>)
#define NUFI_B200_HAVE_MPI 1
#endif
#endif

#ifdef NUFI_B200_HAVE_MPI
#include <mpi.h>
#else
// ---- single-rank stand-in for the MPI names the drivers touch
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 64
#define MPI_IN_PLACE (reinterpret_cast<void *>(1))
inline int MPI_Error_string(int, char *buf, int *len)
{
    static const char msg[] = "single-rank MPI stand-in";
    std::memcpy(buf, msg, sizeof(msg));
    *len = static_cast<int>(sizeof(msg)) - 1;
    return MPI_SUCCESS;
}
#endif

namespace nufi
{

namespace mpi
{

#ifdef NUFI_B200_HAVE_MPI

namespace detail
{
inline void guard(int errcode, const char *who)
{
    if (errcode == MPI_SUCCESS) return;
    char text[MPI_MAX_ERROR_STRING + 1];
    int len = 0;
    MPI_Error_string(errcode, text, &len);
    text[len] = '\0';
    throw std::runtime_error(std::string(who) + ": " + text);
}
} // namespace detail

// RAII around MPI_Init / MPI_Finalize (nufi/mpi.hpp:33-59)
struct programme
{
    programme() = delete;
    programme(const programme &) = delete;
    programme &operator=(const programme &) = delete;
    programme(int *argc, char ***argv)
    {
        if (MPI_Init(argc, argv) != MPI_SUCCESS) throw std::runtime_error("nufi::mpi::programme: Error initialising MPI.");
    }
    ~programme()
    {
        if (MPI_Finalize() != MPI_SUCCESS) {
            std::cerr << "nufi::mpi::programme: Error finalizing MPI. Terminating." << std::flush;
            std::terminate();
        }
    }
};

inline void comm_size(MPI_Comm comm, int *size) { detail::guard(MPI_Comm_size(comm, size), "nufi::mpi::comm_size"); }
inline void comm_rank(MPI_Comm comm, int *rank) { detail::guard(MPI_Comm_rank(comm, rank), "nufi::mpi::comm_rank"); }

// sendbuf is void* so that MPI_IN_PLACE is accepted (nufi/mpi.hpp:135-171)
inline void allreduce_add(const void *sendbuf, float *recvbuf, int count, MPI_Comm comm)
{
    detail::guard(MPI_Allreduce(sendbuf, recvbuf, count, MPI_FLOAT, MPI_SUM, comm), "nufi::mpi::allreduce_add(float)");
}
inline void allreduce_add(const void *sendbuf, double *recvbuf, int count, MPI_Comm comm)
{
    detail::guard(MPI_Allreduce(sendbuf, recvbuf, count, MPI_DOUBLE, MPI_SUM, comm), "nufi::mpi::allreduce_add(double)");
}
inline void allgatherv(const void *sendbuf, int sendcount, float *recvbuf, const int recvcounts[], const int displs[], MPI_Comm comm)
{
    detail::guard(MPI_Allgatherv(sendbuf, sendcount, MPI_FLOAT, recvbuf, recvcounts, displs, MPI_FLOAT, comm), "nufi::mpi::allgatherv(float)");
}
inline void allgatherv(const void *sendbuf, int sendcount, double *recvbuf, const int recvcounts[], const int displs[], MPI_Comm comm)
{
    detail::guard(MPI_Allgatherv(sendbuf, sendcount, MPI_DOUBLE, recvbuf, recvcounts, displs, MPI_DOUBLE, comm), "nufi::mpi::allgatherv(double)");
}

#else // ---- single rank

struct programme
{
    programme() = delete;
    programme(const programme &) = delete;
    programme &operator=(const programme &) = delete;
    programme(int *, char ***) {}
};

inline void comm_size(MPI_Comm, int *size) { *size = 1; }
inline void comm_rank(MPI_Comm, int *rank) { *rank = 0; }

namespace detail
{
// one rank: the sum over ranks of sendbuf is sendbuf itself
template <typename real> void copy_unless_in_place(const void *sendbuf, real *recvbuf, int count)
{
    if (sendbuf != MPI_IN_PLACE && sendbuf != static_cast<const void *>(recvbuf))
        std::memcpy(recvbuf, sendbuf, sizeof(real) * static_cast<size_t>(count));
}
} // namespace detail

inline void allreduce_add(const void *sendbuf, float *recvbuf, int count, MPI_Comm) { detail::copy_unless_in_place(sendbuf, recvbuf, count); }
inline void allreduce_add(const void *sendbuf, double *recvbuf, int count, MPI_Comm) { detail::copy_unless_in_place(sendbuf, recvbuf, count); }
inline void allgatherv(const void *sendbuf, int sendcount, float *recvbuf, const int[], const int displs[], MPI_Comm)
{
    if (sendbuf != MPI_IN_PLACE) std::memcpy(recvbuf + displs[0], sendbuf, sizeof(float) * static_cast<size_t>(sendcount));
}
inline void allgatherv(const void *sendbuf, int sendcount, double *recvbuf, const int[], const int displs[], MPI_Comm)
{
    if (sendbuf != MPI_IN_PLACE) std::memcpy(recvbuf + displs[0], sendbuf, sizeof(double) * static_cast<size_t>(sendcount));
}

#endif

} // namespace mpi

} // namespace nufi

#endif
