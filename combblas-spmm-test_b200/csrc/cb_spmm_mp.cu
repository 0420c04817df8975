// K2 instantiations: MinPlusSRing (reference include/CombBLAS/Semirings.h:235-255, inf_plus :40-47).
#include "cb_spmm_dispatch.cuh"
using namespace cbk;
int cb_launch_min_plus(int dtype, const LaunchParams& p) {
    switch (dtype) {
        case CB_F32: return launch_op<MinPlus<float>>(p);
        case CB_F64: return launch_op<MinPlus<double>>(p);
        case CB_I32: return launch_op<MinPlus<int32_t>>(p);
        case CB_I64: return launch_op<MinPlus<int64_t>>(p);
    }
    return CB_ERR_UNSUPPORTED;
}
