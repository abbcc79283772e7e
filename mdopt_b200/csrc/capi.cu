// C ABI of libmdopt_b200.so (include/mdopt_b200.h): argument checks and dispatch on the compiled capacity.
#include <cuda_runtime.h>

#include "../../include/mdopt_b200.h"
#include "launchers.h"

using mdopt::ChiOps;

namespace {

const ChiOps* ops_for(int chi_cap) {
  switch (chi_cap) {
    case 8: return mdopt::chi_ops_8();
    case 16: return mdopt::chi_ops_16();
    case 32: return mdopt::chi_ops_32();
    case 64: return mdopt::chi_ops_64();
    case 128: return mdopt::chi_ops_128();
    case 256: return mdopt::chi_ops_256();
    default: return nullptr;
  }
}

const mdopt::MonoOps* mono_ops_for(int chi_cap) {
  switch (chi_cap) {
    case 8: return mdopt::mono_ops_8();
    case 16: return mdopt::mono_ops_16();
    case 32: return mdopt::mono_ops_32();
    case 64: return mdopt::mono_ops_64();
    case 128: return mdopt::mono_ops_128();
    case 256: return mdopt::mono_ops_256();
    case 512: return mdopt::mono_ops_512();
    default: return nullptr;
  }
}

// Fused read-out of a batch of dense logical vectors (decoding.py:1362-1379).
__global__ void readout_kernel(int batch, int ncls, const double* __restrict__ v, int renormalise, double* dist,
                               int* success) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* vb = v + (size_t)b * ncls;
  double nrm = 0.0;
  for (int i = 0; i < ncls; ++i) nrm += vb[i] * vb[i];
  nrm = sqrt(nrm);
  double inv = (renormalise && nrm > 1e-17) ? 1.0 / nrm : 1.0;
  double top = 0.0;
  int nan_seen = 0;
  for (int i = 0; i < ncls; ++i) {
    double x = fabs(vb[i]) * inv;
    dist[(size_t)b * ncls + i] = x;
    if (x != x) nan_seen = 1;
    if (x > top) top = x;
  }
  double eps = fmax(1e-9 * top, 1e-12);
  success[b] = (!nan_seen && dist[(size_t)b * ncls] >= top - eps) ? 1 : 0;
}

}  // namespace

extern "C" {

int mdopt_b200_abi_version(void) { return 1; }

int mdopt_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

int mdopt_chi_capacity(int chi) {
  if (chi <= 0) return 0;
  if (chi <= 8) return 8;
  if (chi <= 16) return 16;
  if (chi <= 32) return 32;
  if (chi <= 64) return 64;
  if (chi <= 128) return 128;  // large-chi tier: work matrices in global memory
  if (chi <= 256) return 256;
  return 0;
}

int mdopt_decode_num_ctas(int chi_cap) {
  const ChiOps* ops = ops_for(chi_cap);
  return ops ? ops->num_ctas() : 0;
}

size_t mdopt_decode_workspace_bytes(int chi_cap, int n_sites, int n_ctas) {
  const ChiOps* ops = ops_for(chi_cap);
  if (!ops || n_sites <= 0 || n_ctas <= 0) return 0;
  return ops->ws_doubles_per_cta(n_sites) * sizeof(double) * (size_t)n_ctas + 256;
}

int mdopt_decode_batch_device(const mdopt_decode_desc* desc, int chi_cap, int n_ctas, const int32_t* program,
                              const double* wtab, const int32_t* wdim, int n_tensors, const uint8_t* symbols,
                              int batch, double* dist, int32_t* success, int32_t* status, double* trace,
                              double* counters, void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !program || !wtab || !wdim || !symbols || !dist || !success || !status || !workspace)
    return MDOPT_E_BADARG;
  if (batch <= 0 || n_ctas <= 0 || n_tensors <= 0) return MDOPT_E_BADARG;
  if (desc->n_logical < 1 || desc->n_logical > 12 || desc->n_sites <= desc->n_logical) return MDOPT_E_BADARG;
  const ChiOps* ops = ops_for(chi_cap);
  if (!ops || desc->chi_max < 1 || desc->chi_max > chi_cap) return MDOPT_E_CAPACITY;
  return ops->decode(desc, n_ctas, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters,
                     workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int mdopt_decode_batch_host(const mdopt_decode_desc* desc, int chi_cap, int n_ctas, const int32_t* d_program,
                            const double* d_wtab, const int32_t* d_wdim, int n_tensors, const uint8_t* h_symbols,
                            int batch, double* h_dist, int32_t* h_success, int32_t* h_status, uint8_t* d_symbols,
                            double* d_dist, int32_t* d_success, int32_t* d_status, double* d_counters,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !h_symbols || !h_dist || !h_success || !h_status || !d_symbols || !d_dist || !d_success || !d_status)
    return MDOPT_E_BADARG;
  if (batch <= 0) return MDOPT_E_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t ncls = (size_t)1 << desc->n_logical;
  cudaError_t e = cudaMemcpyAsync(d_symbols, h_symbols, (size_t)batch * desc->n_sites, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  mdopt_decode_desc d2 = *desc;
  d2.trace_stride = 0;
  int rc = mdopt_decode_batch_device(&d2, chi_cap, n_ctas, d_program, d_wtab, d_wdim, n_tensors, d_symbols, batch,
                                     d_dist, d_success, d_status, nullptr, d_counters, workspace, workspace_bytes,
                                     stream);
  if (rc != 0) return rc;
  e = cudaMemcpyAsync(h_dist, d_dist, (size_t)batch * ncls * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(h_success, d_success, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(h_status, d_status, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(st);
}

int mdopt_mono_chi_capacity(int chi) {
  if (chi <= 0) return 0;
  for (int cap = 8; cap <= 512; cap *= 2)
    if (chi <= cap) return cap;
  return 0;
}

int mdopt_mono_num_warps(int chi_cap) {
  const mdopt::MonoOps* ops = mono_ops_for(chi_cap);
  return ops ? ops->num_warps() : 0;
}

size_t mdopt_mono_workspace_bytes(int chi_cap, int n_sites, int n_warps) {
  const mdopt::MonoOps* ops = mono_ops_for(chi_cap);
  if (!ops || n_sites <= 0 || n_warps <= 0) return 0;
  return ops->ws_bytes_per_warp(n_sites) * (size_t)n_warps + 256 + 64 * 160;   // + work counter + tensor tables
}

int mdopt_mono_decode_batch_device(const mdopt_decode_desc* desc, int chi_cap, int n_warps, const int32_t* program,
                                   const double* wtab, const int32_t* wdim, int n_tensors, const uint8_t* symbols,
                                   int batch, double* dist, int32_t* success, int32_t* status, double* trace,
                                   double* counters, void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !program || !wtab || !wdim || !symbols || !dist || !success || !status || !workspace)
    return MDOPT_E_BADARG;
  if (batch <= 0 || n_warps <= 0 || n_tensors <= 0) return MDOPT_E_BADARG;
  // up to 12 logical sites: dense read-out (2^k doubles per syndrome); 13..64: Dephasing-DMRG read-out (k doubles)
  if (desc->n_logical < 1 || desc->n_logical > 64 || desc->n_sites <= desc->n_logical) return MDOPT_E_BADARG;
  if (desc->n_logical > 2 * chi_cap) return MDOPT_E_CAPACITY;
  const mdopt::MonoOps* ops = mono_ops_for(chi_cap);
  if (!ops || desc->chi_max < 1 || desc->chi_max > chi_cap) return MDOPT_E_CAPACITY;
  return ops->decode(desc, n_warps, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace,
                     counters, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int mdopt_mono_decode_batch_host(const mdopt_decode_desc* desc, int chi_cap, int n_warps, const int32_t* d_program,
                                 const double* d_wtab, const int32_t* d_wdim, int n_tensors, const uint8_t* h_symbols,
                                 int batch, double* h_dist, int32_t* h_success, int32_t* h_status, uint8_t* d_symbols,
                                 double* d_dist, int32_t* d_success, int32_t* d_status, double* d_counters,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !h_symbols || !h_dist || !h_success || !h_status || !d_symbols || !d_dist || !d_success || !d_status)
    return MDOPT_E_BADARG;
  if (batch <= 0) return MDOPT_E_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t ncls = desc->n_logical > 12 ? (size_t)desc->n_logical : ((size_t)1 << desc->n_logical);
  cudaError_t e = cudaMemcpyAsync(d_symbols, h_symbols, (size_t)batch * desc->n_sites, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  mdopt_decode_desc d2 = *desc;
  d2.trace_stride = 0;
  int rc = mdopt_mono_decode_batch_device(&d2, chi_cap, n_warps, d_program, d_wtab, d_wdim, n_tensors, d_symbols,
                                          batch, d_dist, d_success, d_status, nullptr, d_counters, workspace,
                                          workspace_bytes, stream);
  if (rc != 0) return rc;
  e = cudaMemcpyAsync(h_dist, d_dist, (size_t)batch * ncls * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(h_success, d_success, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyAsync(h_status, d_status, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(st);
}

size_t mdopt_mps_program_workspace_bytes(int chi_cap, int batch) {
  const size_t mats = chi_cap <= 64 ? 0 : (size_t)3 * chi_cap * (2 * chi_cap + 1);  // mat_store_doubles()
  return (2 * (size_t)(2 * chi_cap) * (2 * chi_cap) + mats) * sizeof(double) * (size_t)batch;
}

size_t mdopt_site_stride(int chi_cap) { return (size_t)chi_cap * 2 * chi_cap; }

int mdopt_mps_program_device(const mdopt_decode_desc* desc, int chi_cap, const int32_t* program, const double* wtab,
                             const int32_t* wdim, int n_tensors, int batch, double* sites_io, int32_t* bonds_io,
                             int32_t* centre_io, double* dist, int32_t* success, int32_t* status, double* trace,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc || !program || !wtab || !wdim || !sites_io || !bonds_io || !centre_io || !status || !workspace)
    return MDOPT_E_BADARG;
  if (batch <= 0 || n_tensors <= 0 || desc->n_sites < 2 || desc->chi_max < 1) return MDOPT_E_BADARG;
  if (desc->n_logical < 1 || desc->n_logical > 12) return MDOPT_E_BADARG;
  const ChiOps* ops = ops_for(chi_cap);
  if (!ops) return MDOPT_E_CAPACITY;
  return ops->mps_program(desc, program, wtab, wdim, n_tensors, batch, sites_io, bonds_io, centre_io, dist, success, status,
                          trace, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int mdopt_svd_batch_device(int chi_cap, int batch, int m, int n, const double* a, int chi_max, double cut,
                           int renormalise, double* u, double* s, double* vh, int32_t* kept, double* trunc,
                           int32_t* status, void* stream) {
  if (!a || !u || !s || !vh || !kept || !trunc || !status || batch <= 0 || m <= 0 || n <= 0) return MDOPT_E_BADARG;
  const ChiOps* ops = ops_for(chi_cap);
  if (!ops || m > 2 * chi_cap || n > 2 * chi_cap) return MDOPT_E_CAPACITY;
  int sms = mdopt_device_sm_count();
  if (sms <= 0) return MDOPT_E_NOGPU;
  return ops->svd(sms, batch, m, n, a, chi_max, cut, renormalise, u, s, vh, kept, trunc, status,
                  reinterpret_cast<cudaStream_t>(stream));
}

int mdopt_site_apply_batch_device(int chi_cap, int batch, int rows, int vl, int vr, int mr, int bn,
                                  const double* m_in, const double* a_site, const double* w, double* theta,
                                  void* stream) {
  if (!m_in || !a_site || !w || !theta || batch <= 0) return MDOPT_E_BADARG;
  if (vl < 1 || vl > 2 || vr < 1 || vr > 2 || rows < 1 || mr < 1 || bn < 1) return MDOPT_E_BADARG;
  const ChiOps* ops = ops_for(chi_cap);
  if (!ops || rows > 2 * chi_cap || vl * mr > 2 * chi_cap || mr > chi_cap || bn > chi_cap) return MDOPT_E_CAPACITY;
  int sms = mdopt_device_sm_count();
  if (sms <= 0) return MDOPT_E_NOGPU;
  return ops->site_apply(sms, batch, rows, vl, vr, mr, bn, m_in, a_site, w, theta,
                         reinterpret_cast<cudaStream_t>(stream));
}

int mdopt_readout_batch_device(int batch, int n_classes, const double* v, int renormalise, double* dist,
                               int32_t* success, void* stream) {
  if (!v || !dist || !success || batch <= 0 || n_classes <= 0) return MDOPT_E_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  readout_kernel<<<(batch + 255) / 256, 256, 0, st>>>(batch, n_classes, v, renormalise, dist, success);
  return (int)cudaGetLastError();
}

}  // extern "C"
