// K2 instantiations: PlusTimesSRing over int32/int64 (reference include/CombBLAS/Semirings.h:212-232).
#include "cb_spmm_dispatch.cuh"
using namespace cbk;
int cb_launch_plus_times_i(int dtype, int akind, const LaunchParams& p) {
    if (dtype == CB_I32) {
        if (akind == A_SAME) return launch_op<PlusTimes<int32_t, A_SAME>>(p);
        if (akind == A_PATTERN) return launch_op<PlusTimes<int32_t, A_PATTERN>>(p);
        return launch_op<PlusTimes<int32_t, A_BOOL>>(p);
    }
    if (akind == A_SAME) return launch_op<PlusTimes<int64_t, A_SAME>>(p);
    if (akind == A_PATTERN) return launch_op<PlusTimes<int64_t, A_PATTERN>>(p);
    return launch_op<PlusTimes<int64_t, A_BOOL>>(p);
}
