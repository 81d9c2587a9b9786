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
