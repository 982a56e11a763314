/*
 * Single-rank MPI shim — TEST INFRASTRUCTURE ONLY (oracle/ is never on the product path).
 *
 * The reference's umbrella header LAM.hpp:6 pulls ConjugateGradient_CPU_MPI_OMP.hpp, which
 * includes <mpi.h> unconditionally, and this image ships no MPI.  This header lets the
 * UNMODIFIED reference sources compile as a 1-rank job: every collective degenerates to a
 * local copy, MPI-IO maps onto stdio.  Only the entry points the reference CPU path calls
 * are provided (ConjugateGradient_CPU_MPI_OMP.hpp:73-74,325-406,464,505 and
 * test_CG_CPU_MPI_OMP.cpp:207-209,289).
 *
 * Timing aid: with LAMCG_SHIM_SIZE=P in the environment MPI_Comm_size reports P while this
 * process stays rank 0, so the reference allocates, generates and multiplies only rank 0's
 * n/P-row block of the n-column system.  bench.py uses that to time a BOUNDED SAMPLE of a system
 * too large for a quick CPU run (one rank's share of the work; the numbers it computes are then
 * not a solution and are never used as an oracle).
 */
#ifndef LAMCG_ORACLE_MPI_SHIM_H
#define LAMCG_ORACLE_MPI_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int MPI_Comm;
typedef int MPI_Datatype; /* value == element size in bytes */
typedef int MPI_Op;
typedef int MPI_Info;
typedef FILE *MPI_File;
typedef long long MPI_Offset;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_BYTE 1
#define MPI_C_BOOL 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG 8
#define MPI_DATATYPE_NULL 0
#define MPI_SUM 1
#define MPI_INFO_NULL 0
#define MPI_MODE_RDONLY 1
#define MPI_SEEK_SET 0
#define MPI_SEEK_CUR 1
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *size)
{
    (void)c;
    const char *e = getenv("LAMCG_SHIM_SIZE");
    int p = e ? atoi(e) : 1;
    *size = p > 0 ? p : 1;
    return MPI_SUCCESS;
}
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }

static inline int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c)
{ (void)buf; (void)count; (void)t; (void)root; (void)c; return MPI_SUCCESS; }

static inline int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{
    (void)op; (void)c;
    if (send != MPI_IN_PLACE) memcpy(recv, send, (size_t)count * (size_t)t);
    return MPI_SUCCESS;
}

static inline int MPI_Allgatherv(const void *send, int sendcount, MPI_Datatype st, void *recv,
                                 const int *recvcounts, const int *displs, MPI_Datatype rt, MPI_Comm c)
{
    (void)recvcounts; (void)rt; (void)c;
    if (send != MPI_IN_PLACE)
        memcpy((char *)recv + (size_t)displs[0] * (size_t)st, send, (size_t)sendcount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Allgather(const void *send, int sendcount, MPI_Datatype st, void *recv,
                                int recvcount, MPI_Datatype rt, MPI_Comm c)
{
    (void)recvcount; (void)rt; (void)c;
    if (send != MPI_IN_PLACE) memcpy(recv, send, (size_t)sendcount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_Gatherv(const void *send, int sendcount, MPI_Datatype st, void *recv,
                              const int *recvcounts, const int *displs, MPI_Datatype rt, int root, MPI_Comm c)
{
    (void)recvcounts; (void)rt; (void)root; (void)c;
    if (send != MPI_IN_PLACE)
        memcpy((char *)recv + (size_t)displs[0] * (size_t)st, send, (size_t)sendcount * (size_t)st);
    return MPI_SUCCESS;
}

static inline int MPI_File_open(MPI_Comm c, const char *name, int mode, MPI_Info info, MPI_File *fh)
{
    (void)c; (void)mode; (void)info;
    *fh = fopen(name, "rb");
    return *fh ? MPI_SUCCESS : 1;
}

/* NB: `count` is int on purpose — that is the real MPI signature and the reason the
 * reference fails for blocks above 2^31 elements (TESTS/BEST_RESULTS:114). */
static inline int MPI_File_read(MPI_File fh, void *buf, int count, MPI_Datatype t, MPI_Status *s)
{
    (void)s;
    if (count < 0) return 1;
    size_t got = fread(buf, (size_t)t, (size_t)count, fh);
    return got == (size_t)count ? MPI_SUCCESS : 1;
}

static inline int MPI_File_seek(MPI_File fh, MPI_Offset off, int whence)
{ return fseeko(fh, (off_t)off, whence == MPI_SEEK_CUR ? SEEK_CUR : SEEK_SET) == 0 ? MPI_SUCCESS : 1; }

static inline int MPI_File_close(MPI_File *fh)
{ int r = fclose(*fh); *fh = NULL; return r == 0 ? MPI_SUCCESS : 1; }

#endif /* LAMCG_ORACLE_MPI_SHIM_H */
