// K4 instantiation: DT=8 NET=0 A2=1 GEN=0 (one translation unit per variant so that they build in parallel)
#include "fgp_mll.cuh"
namespace fgp {
int mll_lat_x_a2_d8(const MllArgs& a, const PassGeom& g, int B, cudaStream_t st) { return launch_mll<8, false, true, false>(a, g, B, st); }
}  // namespace fgp
