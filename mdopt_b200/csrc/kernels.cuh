// Templated kernels and launchers of libmdopt_b200.so (included by inst.cu once per capacity).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/mdopt_b200.h"
#include "mps_core.cuh"
#include "site_apply_mma.cuh"

namespace mdopt {
namespace {


template <int CHI>
struct Cfg {
  static constexpr int NT = (CHI * 8 < 64) ? 64 : (CHI * 8 > 1024 ? 1024 : CHI * 8);  // threads per CTA
};

__host__ __device__ inline size_t ws_doubles_per_cta(int chi, int n_sites) {
  size_t d = (size_t)n_sites * chi * 2 * chi + 2 * (size_t)(2 * chi) * (2 * chi) + mat_store_doubles(chi);
  size_t ints = ((size_t)n_sites + 1 + 1) / 2;  // bond dims, packed two per double slot
  d += ints;
  return (d + 31) & ~(size_t)31;
}

// -------------------------------------------------------------------------------------------------
// Fused decode kernel: persistent CTAs pull syndromes from a global counter; each CTA runs the whole
// program (bias, moves, MPO sweeps, marginal, read-out) for its syndrome out of shared memory, with the
// MPS itself in the CTA's private slice of the workspace (L2-resident at small bond dimension).
// -------------------------------------------------------------------------------------------------
template <int CHI>
__global__ void __launch_bounds__(Cfg<CHI>::NT, 1)
decode_kernel(DecodeParams P, const int* __restrict__ prog, const double* __restrict__ wtab,
              const int* __restrict__ wdim, const uint8_t* __restrict__ syms, int batch, double* dist,
              int* success, int* status, double* trace, double* counters, double* ws, int* work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CHI>& s = *reinterpret_cast<Smem<CHI>*>(smem_raw);
  __shared__ int next_item;
  Engine<CHI> E(s);
  if (threadIdx.x == 0) {
    const size_t stride = ws_doubles_per_cta(CHI, P.n_sites);
    double* base = ws + (size_t)blockIdx.x * stride;
    EngineState& st = s.st;
    st.sites = base;
    st.scr0 = base + (size_t)P.n_sites * Engine<CHI>::SITE_STRIDE;
    st.scr1 = st.scr0 + (size_t)(2 * CHI) * (2 * CHI);
    double* mats = st.scr1 + (size_t)(2 * CHI) * (2 * CHI);
    s.bind_mats(mats);
    st.bd = (P.n_sites + 1 <= Smem<CHI>::BOND_SMEM) ? s.bd_s : reinterpret_cast<int*>(mats + mat_store_doubles(CHI));
    const bool cache = P.n_tensors > 0 && P.n_tensors <= Smem<CHI>::TABLE_SMEM;
    st.wtab = cache ? s.wt_s : wtab;
    st.wdim = cache ? s.wd_s : wdim;
    st.P = P;
    s.cnt = Counters{};
  }
  if (P.n_tensors > 0 && P.n_tensors <= Smem<CHI>::TABLE_SMEM) {
    for (int i = threadIdx.x; i < P.n_tensors * 16; i += blockDim.x) s.wt_s[i] = wtab[i];
    for (int i = threadIdx.x; i < P.n_tensors * 4; i += blockDim.x) s.wd_s[i] = wdim[i];
  }
  __syncthreads();
  const int ncls = 1 << P.n_logical;
  while (true) {
    if (threadIdx.x == 0) next_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int b = next_item;
    __syncthreads();
    if (b >= batch) break;
    if (threadIdx.x == 0) s.st.trace = trace ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    __syncthreads();
    E.run(prog, syms + (size_t)b * P.n_sites, dist + (size_t)b * ncls, success + b);
    __syncthreads();
    if (threadIdx.x == 0) {
      status[b] = s.st.status;
      if (s.st.status == ST_CAPACITY) {
        for (int i = 0; i < ncls; ++i) dist[(size_t)b * ncls + i] = nan("");
        success[b] = 0;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && counters) {
    atomicAdd(counters + 0, s.cnt.flops_qr);
    atomicAdd(counters + 1, s.cnt.flops_jacobi);
    atomicAdd(counters + 2, s.cnt.flops_apply);
    atomicAdd(counters + 3, s.cnt.bytes_hbm);
    atomicAdd(counters + 4, s.cnt.n_steps);
    atomicAdd(counters + 5, s.cnt.n_trunc);
    atomicAdd(counters + 6, s.cnt.n_jacobi_sweeps);
    atomicAdd(counters + 7, s.cnt.n_qr_reflectors);
    atomicAdd(counters + 8, s.cnt.cyc_qr);
    atomicAdd(counters + 9, s.cnt.cyc_formq);
    atomicAdd(counters + 10, s.cnt.cyc_jacobi);
    atomicAdd(counters + 11, s.cnt.cyc_apply);
    atomicAdd(counters + 12, s.cnt.cyc_io);
    atomicAdd(counters + 13, s.cnt.cyc_coef);
    atomicAdd(counters + 14, s.cnt.cyc_total);
    atomicAdd(counters + 15, s.cnt.cyc_pad);
  }
}

// -------------------------------------------------------------------------------------------------
// Program on caller-owned MPS buffers (single-state API: mps_mpo_contract, apply_constraints,
// move_orth_centre, marginal).  One CTA per state; the state is updated in place.
// -------------------------------------------------------------------------------------------------
template <int CHI>
__global__ void __launch_bounds__(Cfg<CHI>::NT, 1)
mps_program_kernel(DecodeParams P, const int* __restrict__ prog, const double* __restrict__ wtab,
                   const int* __restrict__ wdim, int batch, double* sites_io, int* bonds_io, int* centre_io,
                   double* dist, int* success, int* status, double* trace, double* ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CHI>& s = *reinterpret_cast<Smem<CHI>*>(smem_raw);
  Engine<CHI> E(s);
  const int ncls = 1 << P.n_logical;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      EngineState& st = s.st;
      st.sites = sites_io + (size_t)b * P.n_sites * Engine<CHI>::SITE_STRIDE;
      st.bd = (P.n_sites + 1 <= Smem<CHI>::BOND_SMEM) ? s.bd_s : bonds_io + (size_t)b * (P.n_sites + 1);
      st.scr0 = ws + (size_t)blockIdx.x * (2 * (size_t)(2 * CHI) * (2 * CHI) + mat_store_doubles(CHI));
      st.scr1 = st.scr0 + (size_t)(2 * CHI) * (2 * CHI);
      s.bind_mats(st.scr1 + (size_t)(2 * CHI) * (2 * CHI));
      st.wtab = wtab; st.wdim = wdim; st.P = P;
      s.cnt = Counters{};
      st.centre = centre_io[b]; st.mirrored = 0; st.status = ST_OK; st.n_traced = 0;
      st.trace = trace ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    }
    const bool bd_cached = (P.n_sites + 1 <= Smem<CHI>::BOND_SMEM);
    if (bd_cached)
      for (int i = threadIdx.x; i <= P.n_sites; i += blockDim.x) s.bd_s[i] = bonds_io[(size_t)b * (P.n_sites + 1) + i];
    __syncthreads();
    E.run_ops(prog, dist ? dist + (size_t)b * ncls : nullptr, success ? success + b : nullptr);
    __syncthreads();
    if (bd_cached)
      for (int i = threadIdx.x; i <= P.n_sites; i += blockDim.x) bonds_io[(size_t)b * (P.n_sites + 1) + i] = s.bd_s[i];
    if (threadIdx.x == 0) { status[b] = s.st.status; centre_io[b] = s.st.centre; }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// Batched truncated SVD: one CTA per matrix, one-sided Jacobi in shared memory.
// -------------------------------------------------------------------------------------------------
template <int CHI>
__global__ void __launch_bounds__(Cfg<CHI>::NT, 1)
svd_kernel(int batch, int m, int n, const double* __restrict__ a, int chi_max, double cut, int renormalise,
           double* u, double* sv, double* vh, int* kept, double* trunc, int* status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CHI>& s = *reinterpret_cast<Smem<CHI>*>(smem_raw);
  constexpr int LD = Smem<CHI>::LD;
  const int kmax = min(m, n);
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double* ab = a + (size_t)b * m * n;
    MD_PAR_FOR(t, m * n) { int i = t / n, c = t - i * n; s.X[i * LD + c] = ab[t]; }
    MD_SYNC();
    jacobi_columns_dev<CHI>(s, m, n, 60);   // plain sweeps: the stand-alone SVD serves dense inputs (the probe-mode shortcuts
                                            // cost 30 % there, measured at 64 x 64); the decode's splits go through Engine::factor
    truncation_rule<CHI>(s, n, min(chi_max, kmax), cut);
    const int k = s.iscr[0];
    const double scale = (renormalise && s.dscr[3] > 0.0) ? 1.0 / s.dscr[3] : 1.0;
    double* ub = u + (size_t)b * m * kmax;
    double* sb = sv + (size_t)b * kmax;
    double* vb = vh + (size_t)b * kmax * n;
    MD_PAR_FOR(t, m * kmax) {
      int i = t / kmax, j = t - i * kmax;
      ub[t] = (j < k) ? s.X[i * LD + s.order[j]] / s.sig[s.order[j]] : 0.0;
    }
    MD_PAR_FOR(j, kmax) { sb[j] = (j < k) ? s.sig[s.order[j]] * scale : 0.0; }
    // vh[j, c] = (1/sigma_j^2) sum_i X[i, order[j]] a[i, c]
    MD_PAR_FOR(t, kmax * n) {
      int j = t / n, c = t - j * n;
      double acc = 0.0;
      if (j < k) {
        int col = s.order[j];
        for (int i = 0; i < m; ++i) acc += s.X[i * LD + col] * ab[(size_t)i * n + c];
        double sg = s.sig[col];
        acc = acc / sg / sg;
      }
      vb[t] = acc;
    }
    MD_LEADER {
      kept[b] = k;
      trunc[b] = s.dscr[2];
      status[b] = s.iscr[5] ? ST_NOT_CONVERGED : ST_OK;
    }
    MD_SYNC();
  }
}

// -------------------------------------------------------------------------------------------------
// Batched one-site contraction theta = M . (A x W): one CTA per item, operands staged in shared memory.
// -------------------------------------------------------------------------------------------------
template <int CHI>
__global__ void __launch_bounds__(Cfg<CHI>::NT, 1)
site_apply_kernel(int batch, int rows, int vl, int vr, int mr, int bn, const double* __restrict__ m_in,
                  const double* __restrict__ a_site, const double* __restrict__ w, double* theta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<CHI>& s = *reinterpret_cast<Smem<CHI>*>(smem_raw);
  constexpr int LD = Smem<CHI>::LD;
  Engine<CHI> E(s);
  const int cin = vl * mr;
  MD_PAR_FOR(i, 16) { s.W[i] = (i < vl * vr * 4) ? w[i] : 0.0; }
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double* mb = m_in + (size_t)b * rows * cin;
    const double* ab = a_site + (size_t)b * mr * 2 * bn;
    MD_SYNC();
    MD_PAR_FOR(t, rows * cin) { int i = t / cin, c = t - i * cin; s.X[i * LD + c] = mb[t]; }
    MD_PAR_FOR(t, mr * 2 * bn) { s.P2[t] = ab[t]; }
    MD_SYNC();
    E.apply_site(s.X, LD, rows, vl, vr, mr, bn, s.P2, theta + (size_t)b * rows * 2 * vr * bn);
  }
}

template <int CHI>
cudaError_t set_smem_attr(const void* fn) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<CHI>));
}

template <int CHI>
int num_ctas_impl() {
  int dev = 0, sms = 0, per = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (set_smem_attr<CHI>((const void*)decode_kernel<CHI>) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, decode_kernel<CHI>, Cfg<CHI>::NT, sizeof(Smem<CHI>)) !=
      cudaSuccess)
    return 0;
  return sms * per;
}

template <int CHI>
int launch_decode(const mdopt_decode_desc* d, int n_ctas, const int32_t* program, const double* wtab,
                  const int32_t* wdim, int n_tensors, const uint8_t* symbols, int batch, double* dist, int32_t* success,
                  int32_t* status, double* trace, double* counters, void* workspace, size_t workspace_bytes,
                  cudaStream_t st) {
  {  // the dense read-out pages hold 2^(k-1) x chi (k <= 6, one-sided) or 2^ceil(k/2) x chi (meeting in the middle)
    const int k = d->n_logical;
    const int need = (k <= 6) ? (1 << (k - 1)) : (1 << ((k + 1) / 2));
    if (need > CHI) return MDOPT_E_CAPACITY;
  }
  const size_t per = ws_doubles_per_cta(CHI, d->n_sites) * sizeof(double);
  const size_t need = per * (size_t)n_ctas + 256;
  if (workspace_bytes < need) return MDOPT_E_BADARG;
  cudaError_t e = set_smem_attr<CHI>((const void*)decode_kernel<CHI>);
  if (e != cudaSuccess) return (int)e;
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : ((1u << d->n_logical) - 1u);
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = n_tensors;
  // the work counter lives in the last 256 bytes of the workspace
  int* work_counter = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + per * (size_t)n_ctas);
  e = cudaMemsetAsync(work_counter, 0, 256, st);
  if (e != cudaSuccess) return (int)e;
  if (counters) {
    e = cudaMemsetAsync(counters, 0, 16 * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = batch < n_ctas ? batch : n_ctas;
  decode_kernel<CHI><<<grid, Cfg<CHI>::NT, sizeof(Smem<CHI>), st>>>(
      P, program, wtab, wdim, symbols, batch, dist, success, status, d->trace_stride > 0 ? trace : nullptr,
      counters, reinterpret_cast<double*>(workspace), work_counter);
  return (int)cudaGetLastError();
}

template <int C_>
int launch_svd(int sms, int batch, int m, int n, const double* a, int chi_max, double cut, int renormalise,
               double* u, double* s, double* vh, int32_t* kept, double* trunc, int32_t* status, cudaStream_t st) {
  // the stand-alone kernels work out of shared memory only; the large-chi tier goes through the MPS program
  if constexpr (C_ > SMEM_CHI_MAX) {
    return MDOPT_E_CAPACITY;
  } else {
    cudaError_t e = set_smem_attr<C_>((const void*)svd_kernel<C_>);
    if (e != cudaSuccess) return (int)e;
    int per = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, svd_kernel<C_>, Cfg<C_>::NT, sizeof(Smem<C_>));
    if (per < 1) per = 1;
    int grid = batch < sms * per ? batch : sms * per;
    svd_kernel<C_><<<grid, Cfg<C_>::NT, sizeof(Smem<C_>), st>>>(batch, m, n, a, chi_max, cut, renormalise, u, s, vh,
                                                                 kept, trunc, status);
    return (int)cudaGetLastError();
  }
}

template <int C_>
int launch_site_apply(int sms, int batch, int rows, int vl, int vr, int mr, int bn, const double* m_in,
                      const double* a_site, const double* w, double* theta, cudaStream_t st) {
  if constexpr (C_ > SMEM_CHI_MAX) {
    return MDOPT_E_CAPACITY;
  } else {
    // TMA-staged FP64-MMA kernel for tile-aligned shapes (every full-chi step); the scalar kernel otherwise
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(m_in) | reinterpret_cast<uintptr_t>(a_site) |
                             reinterpret_cast<uintptr_t>(theta)) & 15u) == 0;
    if (C_ >= 16 && aligned16 && site_apply_mma_ok(rows, vl, vr, mr, bn)) {
      const size_t smem = sizeof(SiteApplyMmaSmem<C_>);
      cudaError_t e = cudaFuncSetAttribute((const void*)site_apply_mma_kernel<C_>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      const int threads = (rows / 16) * (bn / 32) * 2 * vr >= 16 ? 512 : 256;
      int grid = batch < sms ? batch : sms;
      site_apply_mma_kernel<C_><<<grid, threads, smem, st>>>(batch, rows, vl, vr, mr, bn, m_in, a_site, w, theta);
      return (int)cudaGetLastError();
    }
    cudaError_t e = set_smem_attr<C_>((const void*)site_apply_kernel<C_>);
    if (e != cudaSuccess) return (int)e;
    int per = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, site_apply_kernel<C_>, Cfg<C_>::NT, sizeof(Smem<C_>));
    if (per < 1) per = 1;
    int grid = batch < sms * per ? batch : sms * per;
    site_apply_kernel<C_><<<grid, Cfg<C_>::NT, sizeof(Smem<C_>), st>>>(batch, rows, vl, vr, mr, bn, m_in, a_site, w,
                                                                        theta);
    return (int)cudaGetLastError();
  }
}

template <int C_>
int launch_mps_program(const mdopt_decode_desc* d, const int32_t* program, const double* wtab, const int32_t* wdim,
                       int n_tensors, int batch, double* sites_io, int32_t* bonds_io, int32_t* centre_io, double* dist,
                       int32_t* success, int32_t* status, double* trace, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
  const size_t per = (2 * (size_t)(2 * C_) * (2 * C_) + mat_store_doubles(C_)) * sizeof(double);
  if (workspace_bytes < per * (size_t)batch) return MDOPT_E_BADARG;
  cudaError_t e = set_smem_attr<C_>((const void*)mps_program_kernel<C_>);
  if (e != cudaSuccess) return (int)e;
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : ((1u << d->n_logical) - 1u);
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = n_tensors;
  mps_program_kernel<C_><<<batch, Cfg<C_>::NT, sizeof(Smem<C_>), st>>>(
      P, program, wtab, wdim, batch, sites_io, bonds_io, centre_io, dist, success, status,
      d->trace_stride > 0 ? trace : nullptr, reinterpret_cast<double*>(workspace));
  return (int)cudaGetLastError();
}


}  // namespace
}  // namespace mdopt
