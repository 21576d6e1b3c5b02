/* oracle/refshim/include/mpi.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Single-rank stand-in for the MPI subset the reference sources touch
 * (reference call sites: src/main.c:80-82,187,200; src/cn.c:152; src/log.c:124-249).
 * Rank is always 0 of 1, point-to-point calls are inert, and MPI_Abort is routed to
 * refshim_abort() so a fatal reference error unwinds back into the test harness
 * instead of killing the Python process that loaded the shim.
 */
#ifndef REFSHIM_MPI_H
#define REFSHIM_MPI_H

typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD   0
#define MPI_REQUEST_NULL (-1)
#define MPI_INT          1
#define MPI_ANY_SOURCE   (-1)
#define MPI_ANY_TAG      (-1)
#define MPI_SUCCESS      0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

void refshim_abort(int code);

static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; refshim_abort(code); return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq)
{ (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; *rq = 0; return MPI_SUCCESS; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; *rq = 0; return MPI_SUCCESS; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c)
{ (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request *rq, int *flag, MPI_Status *st) { (void)rq; (void)st; *flag = 0; return MPI_SUCCESS; }
static inline int MPI_Cancel(MPI_Request *rq) { (void)rq; return MPI_SUCCESS; }
static inline int MPI_Request_free(MPI_Request *rq) { *rq = MPI_REQUEST_NULL; return MPI_SUCCESS; }

#endif
