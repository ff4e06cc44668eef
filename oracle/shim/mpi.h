#pragma once
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
inline int MPI_Abort(MPI_Comm, int) { return 0; }
