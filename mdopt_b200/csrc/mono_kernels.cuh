// Kernel and launcher of the monomial warp engine (mono_core.cuh); included by mono_inst.cu once per capacity.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/mdopt_b200.h"
#include "mono_core.cuh"

namespace mdopt {
namespace {

// doubles of output per syndrome: the 2^k class marginals, or (more than 12 logical sites: Dephasing-DMRG read-out)
// the k bits of the main component
__host__ __device__ inline int mono_out_per_syndrome(int n_logical) { return n_logical > 12 ? n_logical : (1 << n_logical); }

// bytes of global workspace one warp needs: its MPS in monomial form + the bond dimensions (+ the environment
// vectors of the Dephasing-DMRG read-out, (n_logical + 1) x chi doubles, sized for up to MONO_MAX_LOGICAL sites)
constexpr int MONO_MAX_LOGICAL = 64;
__host__ __device__ inline size_t mono_ws_bytes_per_warp(int chi, int n_sites) {
  size_t b = (size_t)n_sites * ((size_t)2 * chi * 10) + ((size_t)n_sites + 1) * sizeof(int);
  b = (b + 255) & ~(size_t)255;
  b += (size_t)(MONO_MAX_LOGICAL + 1) * chi * sizeof(double);
  return (b + 255) & ~(size_t)255;
}
__host__ __device__ inline size_t mono_env_offset(int chi, int n_sites) {
  size_t b = (size_t)n_sites * ((size_t)2 * chi * 10) + ((size_t)n_sites + 1) * sizeof(int);
  return (b + 255) & ~(size_t)255;
}

// bytes at the end of the workspace: the work counter (256 B) and the per-launch tensor tables
__host__ __device__ inline size_t mono_ws_tail_bytes(int n_tensors) {
  return 256 + (((size_t)n_tensors + 1) * sizeof(MonoTensor) + 255) / 256 * 256;
}
constexpr int MONO_MAX_TENSORS = 62;   // tail reserved by mdopt_mono_workspace_bytes: 256 + 63 * 160 bytes, rounded

// Per-launch set-up: the MPO tensor table in the form the engine reads (one thread per tensor; slot 0 = identity).
__global__ void mono_prepare_kernel(const double* __restrict__ wtab, const int* __restrict__ wdim, int n_tensors,
                                    MonoTensor* tabs, int* work_counter) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) { mono_build_identity(tabs); *work_counter = 0; }
  if (t < n_tensors) mono_build_tensor(wtab + (size_t)t * 16, wdim + (size_t)t * 4, tabs + 1 + t);
}

// -------------------------------------------------------------------------------------------------
// One warp per CTA, one syndrome per warp at a time; persistent warps pull syndromes from a global counter.
// A warp's whole state is its ~8 KB of shared memory (chi = 64) plus its MPS slice in the workspace, so an SM
// keeps 24 syndromes in flight and the site records stream through L2 / HBM.
// -------------------------------------------------------------------------------------------------
// resident one-warp CTAs per SM the register budget is tuned for (shared memory allows 28 at chi = 64)
#ifndef MONO_CTAS_64
#define MONO_CTAS_64 28
#endif
constexpr int mono_min_ctas(int chi) { return chi <= 64 ? MONO_CTAS_64 : (chi <= 128 ? 12 : (chi <= 256 ? 6 : 3)); }
template <int CHI>
__global__ void __launch_bounds__(32, mono_min_ctas(CHI))
mono_decode_kernel(DecodeParams P, const int* __restrict__ prog, const MonoTensor* __restrict__ tabs,
                   const uint8_t* __restrict__ syms, int batch, double* dist, int* success, int* status,
                   double* trace, double* counters, unsigned char* ws, int* work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MonoSmem<CHI>& s = *reinterpret_cast<MonoSmem<CHI>*>(smem_raw);
  static_assert(sizeof(MonoSite<CHI>) == (size_t)2 * CHI * 10, "site record layout");
  const int lane = threadIdx.x;
  if (lane == 0) {
    unsigned char* base = ws + (size_t)blockIdx.x * mono_ws_bytes_per_warp(CHI, P.n_sites);
    MonoState<CHI>& st = s.state;
    st.sites = reinterpret_cast<MonoSite<CHI>*>(base);
    st.bd = reinterpret_cast<int*>(base + (size_t)P.n_sites * sizeof(MonoSite<CHI>));
    st.env = reinterpret_cast<double*>(base + mono_env_offset(CHI, P.n_sites));
    st.tabs = tabs; st.P = P; st.trace = nullptr;
    st.cnt = MonoCounters{};
  }
  __syncwarp();
  const int ncls = mono_out_per_syndrome(P.n_logical);
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= batch) break;
    if (lane == 0) s.state.trace = trace ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    __syncwarp();
    Mono<CHI>::run(s, prog, syms + (size_t)b * P.n_sites, dist + (size_t)b * ncls, success + b);
    __syncwarp();
    const int st = s.state.status;
    if (st == ST_CAPACITY || st == ST_NOT_MONOMIAL) {
      for (int i = lane; i < ncls; i += 32) dist[(size_t)b * ncls + i] = nan("");
      if (lane == 0) { success[b] = 0; s.state.cnt.n_fallback += 1.0; }
    }
    if (lane == 0) status[b] = st;
    __syncwarp();
  }
  if (lane == 0 && counters) {
    atomicAdd(counters + 3, s.state.cnt.bytes_hbm);
    atomicAdd(counters + 4, s.state.cnt.n_steps);
    atomicAdd(counters + 5, s.state.cnt.n_trunc);
    atomicAdd(counters + 7, s.state.cnt.n_fallback);
  }
}

// -------------------------------------------------------------------------------------------------
// Lock-step variant: ONE CTA per SM, WPC warps, one syndrome per warp, `bar.sync` after every two-site step so that
// the warps of the SM walk the program together and share the instruction caches (Mono::run_lockstep).
// -------------------------------------------------------------------------------------------------
template <int CHI, int WPC>
__global__ void __launch_bounds__(32 * WPC, 1)
mono_decode_lockstep_kernel(DecodeParams P, const int* __restrict__ prog, const MonoTensor* __restrict__ tabs,
                            const uint8_t* __restrict__ syms, int batch, double* dist, int* success, int* status,
                            double* trace, double* counters, unsigned char* ws, int* work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  MonoSmem<CHI>& s = reinterpret_cast<MonoSmem<CHI>*>(smem_raw)[warp];
  if (lane == 0) {
    unsigned char* base = ws + (size_t)(blockIdx.x * WPC + warp) * mono_ws_bytes_per_warp(CHI, P.n_sites);
    MonoState<CHI>& st = s.state;
    st.sites = reinterpret_cast<MonoSite<CHI>*>(base);
    st.bd = reinterpret_cast<int*>(base + (size_t)P.n_sites * sizeof(MonoSite<CHI>));
    st.env = reinterpret_cast<double*>(base + mono_env_offset(CHI, P.n_sites));
    st.tabs = tabs; st.P = P; st.trace = nullptr;
    st.cnt = MonoCounters{};
    st.status = ST_OK;
  }
  __syncwarp();
  const int ncls = mono_out_per_syndrome(P.n_logical);
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    const bool active = b < batch;
    if (!__syncthreads_or(active ? 1 : 0)) break;
    if (lane == 0) s.state.trace = (active && trace) ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    __syncwarp();
    Mono<CHI>::run_lockstep(s, prog, syms + (size_t)(active ? b : 0) * P.n_sites, dist + (size_t)(active ? b : 0) * ncls,
                            success + (active ? b : 0), active);
    __syncwarp();
    if (active) {
      const int st = s.state.status;
      if (st == ST_CAPACITY || st == ST_NOT_MONOMIAL) {
        for (int i = lane; i < ncls; i += 32) dist[(size_t)b * ncls + i] = nan("");
        if (lane == 0) { success[b] = 0; s.state.cnt.n_fallback += 1.0; }
      }
      if (lane == 0) status[b] = st;
    }
    __syncwarp();
  }
  if (lane == 0 && counters) {
    atomicAdd(counters + 3, s.state.cnt.bytes_hbm);
    atomicAdd(counters + 4, s.state.cnt.n_steps);
    atomicAdd(counters + 5, s.state.cnt.n_trunc);
    atomicAdd(counters + 7, s.state.cnt.n_fallback);
  }
}

template <int CHI>
int mono_num_warps_impl() {
  int dev = 0, sms = 0, per = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaFuncSetAttribute((const void*)mono_decode_kernel<CHI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(MonoSmem<CHI>)) != cudaSuccess)
    return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, mono_decode_kernel<CHI>, 32, sizeof(MonoSmem<CHI>)) !=
      cudaSuccess)
    return 0;
  return sms * per;
}

template <int CHI>
int launch_mono_decode(const mdopt_decode_desc* d, int n_warps, const int32_t* program, const double* wtab,
                       const int32_t* wdim, int n_tensors, const uint8_t* symbols, int batch, double* dist,
                       int32_t* success, int32_t* status, double* trace, double* counters, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  const size_t per = mono_ws_bytes_per_warp(CHI, d->n_sites);
  if (n_tensors > MONO_MAX_TENSORS) return MDOPT_E_CAPACITY;
  if (workspace_bytes < per * (size_t)n_warps + mono_ws_tail_bytes(MONO_MAX_TENSORS)) return MDOPT_E_BADARG;
  cudaError_t e = cudaFuncSetAttribute((const void*)mono_decode_kernel<CHI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MonoSmem<CHI>));
  if (e != cudaSuccess) return (int)e;
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : (d->n_logical >= 32 ? 0xffffffffu : ((1u << d->n_logical) - 1u));
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = n_tensors;
  int* work_counter = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + per * (size_t)n_warps);
  MonoTensor* tabs = reinterpret_cast<MonoTensor*>(reinterpret_cast<char*>(work_counter) + 256);
  mono_prepare_kernel<<<(n_tensors + 63) / 64, 64, 0, st>>>(wtab, wdim, n_tensors, tabs, work_counter);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (counters) {
    e = cudaMemsetAsync(counters, 0, 16 * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
  }
  if constexpr (CHI == 64) {
    // default at chi = 64: the lock-step kernel, one CTA of 28 warps per SM (+15 % over 28 free-running one-warp
    // CTAs: 227 k vs 198 k non-trivial syndromes/s).  MDOPT_MONO_LOCKSTEP=0 selects the free-running kernel,
    // 16 / 24 other CTA sizes (measurement switches).
    const char* ls = getenv("MDOPT_MONO_LOCKSTEP");
    const int wpc = ls ? atoi(ls) : 28;
    if (wpc == 16 || wpc == 24 || wpc == 28) {
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int ctas = (n_warps / wpc) < sms ? (n_warps / wpc) : sms;
      if (ctas < 1) return MDOPT_E_BADARG;
      const size_t smem = (size_t)wpc * sizeof(MonoSmem<CHI>);
#define MD_LAUNCH_LS(W)                                                                                          \
  {                                                                                                              \
    e = cudaFuncSetAttribute((const void*)mono_decode_lockstep_kernel<CHI, W>,                                   \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                            \
    if (e != cudaSuccess) return (int)e;                                                                         \
    mono_decode_lockstep_kernel<CHI, W><<<ctas, 32 * W, smem, st>>>(                                             \
        P, program, tabs, symbols, batch, dist, success, status, d->trace_stride > 0 ? trace : nullptr, counters, \
        reinterpret_cast<unsigned char*>(workspace), work_counter);                                              \
  }
      if (wpc == 16) MD_LAUNCH_LS(16) else if (wpc == 24) MD_LAUNCH_LS(24) else MD_LAUNCH_LS(28)
#undef MD_LAUNCH_LS
      return (int)cudaGetLastError();
    }
  }
  const int grid = batch < n_warps ? batch : n_warps;
  mono_decode_kernel<CHI><<<grid, 32, sizeof(MonoSmem<CHI>), st>>>(
      P, program, tabs, symbols, batch, dist, success, status, d->trace_stride > 0 ? trace : nullptr, counters,
      reinterpret_cast<unsigned char*>(workspace), work_counter);
  return (int)cudaGetLastError();
}

}  // namespace
}  // namespace mdopt
