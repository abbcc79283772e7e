// Per-capacity launcher table: inst.cu is compiled once per MDOPT_CHI and exports one of these.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mdopt_b200.h"

namespace mdopt {

struct ChiOps {
  int chi;
  int (*num_ctas)();
  size_t (*ws_doubles_per_cta)(int n_sites);
  int (*decode)(const mdopt_decode_desc* d, int n_ctas, const int32_t* program, const double* wtab,
                const int32_t* wdim, int n_tensors, const uint8_t* symbols, int batch, double* dist, int32_t* success,
                int32_t* status, double* trace, double* counters, void* workspace, size_t workspace_bytes,
                cudaStream_t st);
  int (*svd)(int sms, int batch, int m, int n, const double* a, int chi_max, double cut, int renormalise, double* u,
             double* s, double* vh, int32_t* kept, double* trunc, int32_t* status, cudaStream_t st);
  int (*site_apply)(int sms, int batch, int rows, int vl, int vr, int mr, int bn, const double* m_in,
                    const double* a_site, const double* w, double* theta, cudaStream_t st);
  int (*mps_program)(const mdopt_decode_desc* d, const int32_t* program, const double* wtab, const int32_t* wdim,
                     int n_tensors, int batch, double* sites_io, int32_t* bonds_io, int32_t* centre_io, double* dist,
                     int32_t* success, int32_t* status, double* trace, void* workspace, size_t workspace_bytes,
                     cudaStream_t st);
};

// the monomial warp engine (mono_core.cuh): compiled capacities 8 ... 512
struct MonoOps {
  int chi;
  int (*num_warps)();
  size_t (*ws_bytes_per_warp)(int n_sites);
  int (*decode)(const mdopt_decode_desc* d, int n_warps, const int32_t* program, const double* wtab,
                const int32_t* wdim, int n_tensors, const uint8_t* symbols, int batch, double* dist, int32_t* success,
                int32_t* status, double* trace, double* counters, void* workspace, size_t workspace_bytes,
                cudaStream_t st);
};
const MonoOps* mono_ops_8();
const MonoOps* mono_ops_16();
const MonoOps* mono_ops_32();
const MonoOps* mono_ops_64();
const MonoOps* mono_ops_128();
const MonoOps* mono_ops_256();
const MonoOps* mono_ops_512();

const ChiOps* chi_ops_8();
const ChiOps* chi_ops_16();
const ChiOps* chi_ops_32();
const ChiOps* chi_ops_64();
const ChiOps* chi_ops_128();
const ChiOps* chi_ops_256();

}  // namespace mdopt
