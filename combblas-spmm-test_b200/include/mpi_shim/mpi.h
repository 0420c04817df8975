// <mpi.h> for reference drivers compiled UNMODIFIED against the B200 host layer on a box without MPI: the handful of MPI
// names the CombBLAS surface needs (CombBLAS/cb_mpi.h).  Put this directory on the include path only when no real MPI is
// present; with a real one, compile with -DCB_HAVE_MPI and leave it out.
#ifndef CB_MPI_SHIM_H
#define CB_MPI_SHIM_H
#include "../CombBLAS/cb_mpi.h"
#endif
