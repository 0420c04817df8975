// K2 instantiations: PlusTimesSRing over float/double (reference include/CombBLAS/Semirings.h:212-232).
#include "cb_spmm_dispatch.cuh"
using namespace cbk;
int cb_launch_plus_times_f(int dtype, int akind, const LaunchParams& p) {
    if (dtype == CB_F32) {
        if (akind == A_SAME) return launch_op<PlusTimes<float, A_SAME>>(p);
        if (akind == A_PATTERN) return launch_op<PlusTimes<float, A_PATTERN>>(p);
        return launch_op<PlusTimes<float, A_BOOL>>(p);
    }
    if (akind == A_SAME) return launch_op<PlusTimes<double, A_SAME>>(p);
    if (akind == A_PATTERN) return launch_op<PlusTimes<double, A_PATTERN>>(p);
    return launch_op<PlusTimes<double, A_BOOL>>(p);
}
