// One capacity of the monomial warp engine: nvcc -DMDOPT_CHI=8|...|512 -c mono_inst.cu (objects build in parallel).
#include "mono_kernels.cuh"
#include "launchers.h"

#ifndef MDOPT_CHI
#error "compile with -DMDOPT_CHI=<capacity>"
#endif
#define MD_CAT2(a, b) a##b
#define MD_CAT(a, b) MD_CAT2(a, b)

namespace mdopt {
namespace {
size_t mono_ws_this(int n_sites) { return mono_ws_bytes_per_warp(MDOPT_CHI, n_sites); }
const MonoOps kMonoOps = {MDOPT_CHI, &mono_num_warps_impl<MDOPT_CHI>, &mono_ws_this, &launch_mono_decode<MDOPT_CHI>};
}  // namespace
const MonoOps* MD_CAT(mono_ops_, MDOPT_CHI)() { return &kMonoOps; }
}  // namespace mdopt
