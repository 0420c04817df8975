// K2 instantiations: SelectMaxSRing<bool,T> (reference include/CombBLAS/Semirings.h:191-210) and
// PlusTimesSRing<bool,bool> = OR/AND (promote.h:78).
#include "cb_spmm_dispatch.cuh"
using namespace cbk;
int cb_launch_select_max(int dtype, const LaunchParams& p) {
    switch (dtype) {
        case CB_F32: return launch_op<SelectMax<float>>(p);
        case CB_F64: return launch_op<SelectMax<double>>(p);
        case CB_I32: return launch_op<SelectMax<int32_t>>(p);
        case CB_I64: return launch_op<SelectMax<int64_t>>(p);
    }
    return CB_ERR_UNSUPPORTED;
}
int cb_launch_or_and(int akind, const LaunchParams& p) {
    if (akind == A_PATTERN) return launch_op<OrAnd<A_PATTERN>>(p);
    return launch_op<OrAnd<A_BOOL>>(p);
}
