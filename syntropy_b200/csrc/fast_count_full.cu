// Single-subspace windowed count pass: instantiations (DP = 2, 4, 6, 8) and launcher.
#include "fast_count_launch.cuh"
#include "fast_count_group.cuh"

namespace ksg {

int launch_count_full(int dim_padded, const CountF32Args& a, const CountUnitsArgs& w) {
    switch (dim_padded) {
        case 2: return launch_count_full_dp<2>(a, w);
        case 4: return launch_count_full_dp<4>(a, w);
        case 6: return launch_count_full_dp<6>(a, w);
        case 8: return launch_count_full_dp<8>(a, w);
        default: return KSG_EUNSUPPORTED;
    }
}

}  // namespace ksg
