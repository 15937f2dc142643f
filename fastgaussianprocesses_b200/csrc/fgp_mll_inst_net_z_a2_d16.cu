// K4 instantiation: DT=16 NET=1 A2=1 GEN=1 (one translation unit per variant so that they build in parallel)
#include "fgp_mll.cuh"
namespace fgp {
int mll_net_z_a2_d16(const MllArgs& a, const PassGeom& g, int B, cudaStream_t st) { return launch_mll<16, true, true, true>(a, g, B, st); }
}  // namespace fgp
