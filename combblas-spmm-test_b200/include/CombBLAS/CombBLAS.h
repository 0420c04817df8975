// Umbrella include of the B200 host layer (the role of reference include/CombBLAS/CombBLAS.h:119-137).
#ifndef CB_COMBBLAS_H
#define CB_COMBBLAS_H
#include "cb_mpi.h"
#include "SpDefs.h"
#include "promote.h"
#include "Semirings.h"
#include "CommGrid.h"
#include "SpTuples.h"
#include "FullyDistVec.h"
#include "SpParMat.h"
#include "DenseParMat.h"
#include "ParFriends.h"
#endif
