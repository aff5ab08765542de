/* Single-rank stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY.
 *
 * Lets the unmodified reference headers (/root/reference/include, which
 * include <mpi.h> from GMDP/gmdp.h:36) compile and run as exactly one rank.
 * Nothing here is linked into the product library.  Collectives over one rank
 * are copies; point-to-point is a FIFO mailbox to self (only reached by the
 * reference's ingest helpers when nrank > 1, which never happens here).
 */
#ifndef GM_ORACLE_STUB_MPI_H_
#define GM_ORACLE_STUB_MPI_H_
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <sys/time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype; /* datatype == element size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG 8
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_LAND 4
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

#ifdef __cplusplus
#define GM_STUB_INLINE inline
#else
#define GM_STUB_INLINE static inline
#endif

GM_STUB_INLINE int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return 0; }
GM_STUB_INLINE int MPI_Finalize(void) { return 0; }
GM_STUB_INLINE int MPI_Comm_size(MPI_Comm c, int* n) { (void)c; *n = 1; return 0; }
GM_STUB_INLINE int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
GM_STUB_INLINE int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
GM_STUB_INLINE double MPI_Wtime(void) {
  struct timeval tv; gettimeofday(&tv, 0);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}
GM_STUB_INLINE int MPI_Allreduce(const void* s, void* r, int cnt, MPI_Datatype dt, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c;
  if (s != MPI_IN_PLACE) memcpy(r, s, (size_t)cnt * (size_t)dt);
  return 0;
}
GM_STUB_INLINE int MPI_Bcast(void* b, int cnt, MPI_Datatype dt, int root, MPI_Comm c) {
  (void)b; (void)cnt; (void)dt; (void)root; (void)c; return 0;
}
GM_STUB_INLINE int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype* out) { *out = n * old; return 0; }
GM_STUB_INLINE int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }

/* FIFO mailbox to self. */
typedef struct gm_stub_msg { void* data; size_t bytes; int tag; struct gm_stub_msg* next; } gm_stub_msg;
GM_STUB_INLINE gm_stub_msg** gm_stub_box(void) { static gm_stub_msg* head = 0; return &head; }
GM_STUB_INLINE int MPI_Send(const void* buf, int cnt, MPI_Datatype dt, int dst, int tag, MPI_Comm c) {
  (void)dst; (void)c;
  gm_stub_msg* m = (gm_stub_msg*)malloc(sizeof(gm_stub_msg));
  m->bytes = (size_t)cnt * (size_t)dt; m->data = malloc(m->bytes ? m->bytes : 1);
  memcpy(m->data, buf, m->bytes); m->tag = tag; m->next = 0;
  gm_stub_msg** p = gm_stub_box(); while (*p) p = &(*p)->next; *p = m;
  return 0;
}
GM_STUB_INLINE int MPI_Recv(void* buf, int cnt, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status* st) {
  (void)src; (void)c; (void)st; (void)cnt; (void)dt;
  gm_stub_msg** p = gm_stub_box();
  while (*p && (*p)->tag != tag) p = &(*p)->next;
  if (!*p) abort(); /* receive before send: impossible with one rank */
  gm_stub_msg* m = *p; *p = m->next;
  memcpy(buf, m->data, m->bytes); free(m->data); free(m);
  return 0;
}
GM_STUB_INLINE int MPI_Isend(const void* buf, int cnt, MPI_Datatype dt, int dst, int tag, MPI_Comm c, MPI_Request* rq) {
  *rq = 0; return MPI_Send(buf, cnt, dt, dst, tag, c);
}
/* Irecv is deferred to Waitall in real MPI; with one rank every matching Isend
 * in the reference is posted before the Irecv, so receive immediately. */
GM_STUB_INLINE int MPI_Irecv(void* buf, int cnt, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request* rq) {
  *rq = 0; return MPI_Recv(buf, cnt, dt, src, tag, c, 0);
}
GM_STUB_INLINE int MPI_Waitall(int n, MPI_Request* rq, MPI_Status* st) { (void)n; (void)rq; (void)st; return 0; }
GM_STUB_INLINE int MPI_Wait(MPI_Request* rq, MPI_Status* st) { (void)rq; (void)st; return 0; }
#endif
