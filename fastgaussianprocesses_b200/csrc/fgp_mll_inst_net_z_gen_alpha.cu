// K4 instantiation: DT=0 NET=1 A2=0 GEN=1 (one translation unit per variant so that they build in parallel)
#include "fgp_mll.cuh"
namespace fgp {
int mll_net_z_gen_alpha(const MllArgs& a, const PassGeom& g, int B, cudaStream_t st) { return launch_mll<0, true, false, true>(a, g, B, st); }
}  // namespace fgp
