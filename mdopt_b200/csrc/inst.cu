// One capacity of the kernels: nvcc -DMDOPT_CHI=8|16|32|64 -c inst.cu (the four objects build in parallel).
#include "kernels.cuh"
#include "launchers.h"

#ifndef MDOPT_CHI
#error "compile with -DMDOPT_CHI=<capacity>"
#endif
#define MD_CAT2(a, b) a##b
#define MD_CAT(a, b) MD_CAT2(a, b)

namespace mdopt {
namespace {
size_t ws_doubles_this(int n_sites) { return ws_doubles_per_cta(MDOPT_CHI, n_sites); }
const ChiOps kOps = {MDOPT_CHI,
                     &num_ctas_impl<MDOPT_CHI>,
                     &ws_doubles_this,
                     &launch_decode<MDOPT_CHI>,
                     &launch_svd<MDOPT_CHI>,
                     &launch_site_apply<MDOPT_CHI>,
                     &launch_mps_program<MDOPT_CHI>};
}  // namespace
const ChiOps* MD_CAT(chi_ops_, MDOPT_CHI)() { return &kOps; }
}  // namespace mdopt
