// One-process stand-in for the few MPI names src/eQ.h and diffuclass.cpp mention.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <ctime>
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 0
inline double MPI_Wtime() { return (double)clock() / CLOCKS_PER_SEC; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
inline int MPI_Comm_split(MPI_Comm c, int, int, MPI_Comm *o) { *o = c; return 0; }
inline int MPI_Barrier(MPI_Comm) { return 0; }
// The rest of the names src/eQmpi.h and src/simulation.cpp mention, so that the reference's controller code can be
// COMPILED against the drop-in class (scripts/check_dropin_compiles.sh).  Inert: a one-process stand-in cannot carry
// the controller <-> HSL-rank exchange, and nothing here is ever run for it.
typedef int MPI_Request;
typedef int MPI_Group;
typedef int MPI_Datatype;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
#define MPI_COMM_NULL (-1)
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_LONG 3
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
inline double MPI_Wtick() { return 1.0 / CLOCKS_PER_SEC; }
inline int MPI_Get_processor_name(char *name, int *len) { name[0] = 'x'; name[1] = 0; *len = 1; return 0; }
inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
inline int MPI_Comm_group(MPI_Comm, MPI_Group *g) { *g = 0; return 0; }
inline int MPI_Group_incl(MPI_Group, int, const int *, MPI_Group *g) { *g = 0; return 0; }
inline int MPI_Group_free(MPI_Group *) { return 0; }
inline int MPI_Comm_create(MPI_Comm c, MPI_Group, MPI_Comm *o) { *o = c; return 0; }
inline int MPI_Comm_free(MPI_Comm *) { return 0; }
inline int MPI_Group_excl(MPI_Group, int, const int *, MPI_Group *g) { *g = 0; return 0; }
