// Host emulation of the block-level engine (mdopt_b200/csrc/*.cuh compiled with -DMDOPT_EMU).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only test tier exercise the exact index logic, pivoting,
// truncation rules and program interpreter the CUDA kernels run.  Never linked into libmdopt_b200.so.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/mdopt_b200.h"
#include "../../mdopt_b200/csrc/mps_core.cuh"
#include "../../mdopt_b200/csrc/mono_core.cuh"

using namespace mdopt;

namespace {
size_t ws_doubles(int chi, int n_sites) {
  return (size_t)n_sites * chi * 2 * chi + 2 * (size_t)(2 * chi) * (2 * chi) + (n_sites + 2) / 2 + 8;
}

template <int CHI>
int decode_impl(const mdopt_decode_desc* d, const int32_t* program, const double* wtab, const int32_t* wdim,
                const uint8_t* symbols, int batch, double* dist, int32_t* success, int32_t* status, double* trace,
                double* counters) {
  Smem<CHI>* sm = new Smem<CHI>();
  std::vector<double> ws(ws_doubles(CHI, d->n_sites));
  std::vector<double> mats(mat_store_doubles(CHI) + 1);
  sm->bind_mats(mats.data());
  Engine<CHI> E(*sm);
  E.st.sites = ws.data();
  E.st.scr0 = E.st.sites + (size_t)d->n_sites * Engine<CHI>::SITE_STRIDE;
  E.st.scr1 = E.st.scr0 + (size_t)(2 * CHI) * (2 * CHI);
  E.st.bd = reinterpret_cast<int*>(E.st.scr1 + (size_t)(2 * CHI) * (2 * CHI));
  E.st.wtab = wtab;
  E.st.wdim = wdim;
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : ((1u << d->n_logical) - 1u);
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = 0;
  E.st.P = P;
  sm->cnt = Counters{};
  const int ncls = 1 << d->n_logical;
  for (int b = 0; b < batch; ++b) {
    E.st.trace = (trace && d->trace_stride > 0) ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    E.run(program, symbols + (size_t)b * d->n_sites, dist + (size_t)b * ncls, success + b);
    status[b] = E.st.status;
  }
  if (counters) {
    counters[0] = sm->cnt.flops_qr; counters[1] = sm->cnt.flops_jacobi; counters[2] = sm->cnt.flops_apply;
    counters[3] = sm->cnt.bytes_hbm; counters[4] = sm->cnt.n_steps; counters[5] = sm->cnt.n_trunc;
    counters[6] = sm->cnt.n_jacobi_sweeps; counters[7] = 0;
  }
  delete sm;
  return 0;
}

// the monomial (one warp per syndrome) engine, lane loops run sequentially
template <int CHI>
int mono_impl(const mdopt_decode_desc* d, const int32_t* program, const double* wtab, const int32_t* wdim,
              int n_tensors, const uint8_t* symbols, int batch, double* dist, int32_t* success, int32_t* status,
              double* trace, double* counters) {
  MonoSmem<CHI>* sm = new MonoSmem<CHI>();
  std::vector<MonoSite<CHI>> sites(d->n_sites);
  std::vector<int> bd(d->n_sites + 1);
  std::vector<MonoTensor> tabs(n_tensors + 1);
  std::vector<double> env((size_t)(d->n_logical + 2) * CHI);
  mono_build_identity(&tabs[0]);
  for (int t = 0; t < n_tensors; ++t) mono_build_tensor(wtab + (size_t)t * 16, wdim + (size_t)t * 4, &tabs[1 + t]);
  MonoState<CHI>& st = sm->state;
  st.sites = sites.data(); st.bd = bd.data(); st.tabs = tabs.data(); st.env = env.data();
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : (d->n_logical >= 32 ? 0xffffffffu : ((1u << d->n_logical) - 1u));
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = n_tensors;
  st.P = P;
  st.cnt = MonoCounters{};
  const int ncls = d->n_logical > 12 ? d->n_logical : (1 << d->n_logical);
  for (int b = 0; b < batch; ++b) {
    st.trace = (trace && d->trace_stride > 0) ? trace + (size_t)b * P.trace_max * P.trace_stride : nullptr;
    Mono<CHI>::run(*sm, program, symbols + (size_t)b * d->n_sites, dist + (size_t)b * ncls, success + b);
    status[b] = st.status;
    if (st.status == ST_CAPACITY || st.status == ST_NOT_MONOMIAL) {
      for (int i = 0; i < ncls; ++i) dist[(size_t)b * ncls + i] = NAN;
      success[b] = 0;
    }
  }
  if (counters) { counters[3] = st.cnt.bytes_hbm; counters[4] = st.cnt.n_steps; counters[5] = st.cnt.n_trunc; }
  delete sm;
  return 0;
}

template <int CHI>
int svd_impl(int m, int n, const double* a, int chi_max, double cut, double* u, double* sv, double* vh, int* kept) {
  Smem<CHI>* sm = new Smem<CHI>();
  Smem<CHI>& s = *sm;
  constexpr int LD = Smem<CHI>::LD;
  for (int i = 0; i < m; ++i) for (int c = 0; c < n; ++c) s.X[i * LD + c] = a[i * n + c];
  jacobi_columns<CHI>(s, m, n);
  const int kmax = m < n ? m : n;
  truncation_rule<CHI>(s, n, chi_max < kmax ? chi_max : kmax, cut);
  int k = s.iscr[0];
  *kept = k;
  for (int j = 0; j < kmax; ++j) {
    int col = s.order[j];
    sv[j] = j < k ? s.sig[col] : 0.0;
    for (int i = 0; i < m; ++i) u[i * kmax + j] = j < k ? s.X[i * LD + col] / s.sig[col] : 0.0;
    for (int c = 0; c < n; ++c) {
      double acc = 0.0;
      if (j < k) { for (int i = 0; i < m; ++i) acc += s.X[i * LD + col] * a[i * n + c]; acc = acc / s.sig[col] / s.sig[col]; }
      vh[j * n + c] = acc;
    }
  }
  int bad = s.iscr[5];
  delete sm;
  return bad;
}

// The split of the SVD modes on the portable tier, as Engine::factor runs it: orthogonality probe; one round of
// disjoint pair rotations; greedy rounds, each followed by the probe; cyclic sweeps.  Returns the route taken
// (0 orthogonal, 1 single pair round, 2 greedy rounds, 3 cyclic sweeps), X J in xout (m x n) and the ranking.
template <int CHI>
int split_impl(int m, int n, const double* a, double* xout, int* order, double* sig) {
  Smem<CHI>* sm = new Smem<CHI>();
  Smem<CHI>& s = *sm;
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  std::vector<double> table((size_t)N2 * N2);
  for (int i = 0; i < m; ++i) for (int c = 0; c < n; ++c) s.X[i * LD + c] = a[i * n + c];
  int route = 0;
  if (!ortho_probe<CHI>(s, m, n)) {
    if (pair_fix<CHI>(s, m, n, table.data())) route = 1;
    else {
      int settled = 0;
      for (int it = 0; it < 12 && !settled; ++it) {
        if (!pair_fix<CHI>(s, m, n, table.data(), 1)) break;
        settled = ortho_probe<CHI>(s, m, n);
      }
      route = settled ? 2 : 3;
      if (!settled) jacobi_columns<CHI>(s, m, n);
    }
  }
  for (int i = 0; i < m; ++i) for (int c = 0; c < n; ++c) xout[i * n + c] = s.X[i * LD + c];
  for (int c = 0; c < n; ++c) { order[c] = s.order[c]; sig[c] = s.sig[c]; }
  if (s.iscr[5]) route = -1;
  delete sm;
  return route;
}

// QRCP + form Q: returns K; q (m x K, ld kcap), coeff (K x n, ld n)
template <int CHI>
int qr_impl(int m, int n, const double* a, int kmax, double rel_tol, double* q, double* coeff, int* more) {
  Smem<CHI>* sm = new Smem<CHI>();
  Smem<CHI>& s = *sm;
  constexpr int LD = Smem<CHI>::LD;
  for (int i = 0; i < m; ++i) for (int c = 0; c < n; ++c) s.X[i * LD + c] = a[i * n + c];
  qrcp<CHI>(s, m, n, kmax, rel_tol);
  int K = s.iscr[0];
  *more = s.iscr[1];
  extract_coeff<CHI>(s, K, n);
  form_q<CHI>(s, m, K);
  for (int i = 0; i < m; ++i) for (int k = 0; k < K; ++k) q[i * K + k] = s.X[i * LD + k];
  for (int k = 0; k < K; ++k) for (int c = 0; c < n; ++c) coeff[k * n + c] = s.P2[k * LD + c];
  delete sm;
  return K;
}

template <int CHI>
int program_impl(const mdopt_decode_desc* d, const int32_t* program, const double* wtab, const int32_t* wdim,
                 double* sites_io, int32_t* bonds_io, int32_t* centre_io, double* dist, int32_t* success,
                 int32_t* status, double* trace) {
  Smem<CHI>* sm = new Smem<CHI>();
  std::vector<double> ws(2 * (size_t)(2 * CHI) * (2 * CHI));
  std::vector<double> mats(mat_store_doubles(CHI) + 1);
  sm->bind_mats(mats.data());
  Engine<CHI> E(*sm);
  E.st.sites = sites_io; E.st.bd = bonds_io;
  E.st.scr0 = ws.data(); E.st.scr1 = E.st.scr0 + (size_t)(2 * CHI) * (2 * CHI);
  E.st.wtab = wtab; E.st.wdim = wdim;
  DecodeParams P;
  P.n_sites = d->n_sites; P.n_logical = d->n_logical; P.chi_max = d->chi_max; P.mode = d->mode;
  P.renormalise = d->renormalise; P.trace_stride = d->trace_stride; P.trace_max = d->trace_max;
  P.n_ops = d->n_ops; P.cut_sweep = d->cut_sweep; P.cut_move = d->cut_move; P.rank_tol = d->rank_tol;
  P.bias_p = d->bias_p;
  P.n_head = d->n_head > 0 ? d->n_head : d->n_logical;
  P.kept_mask = d->kept_mask ? d->kept_mask : ((1u << d->n_logical) - 1u);
  P.readout_rel = d->readout_rel; P.readout_abs = d->readout_abs;
  P.n_tensors = 0;
  E.st.P = P;
  sm->cnt = Counters{};
  E.st.centre = *centre_io; E.st.mirrored = 0; E.st.status = ST_OK; E.st.n_traced = 0;
  E.st.trace = (trace && d->trace_stride > 0) ? trace : nullptr;
  E.run_ops(program, dist, success);
  *status = E.st.status; *centre_io = E.st.centre;
  delete sm;
  return 0;
}
}  // namespace

extern "C" {
// test switch: 1 = skip the pair rotations after a failed probe (every rotating split runs the cyclic sweeps);
// 2 = skip only the single-round shortcut (every rotating split runs the greedy rounds, each followed by the probe)
void emu_set_pair_fix_off(int off) { mdopt::emu_pair_fix_off() = off; }
int emu_decode_batch(const mdopt_decode_desc* d, int chi_cap, const int32_t* program, const double* wtab,
                     const int32_t* wdim, const uint8_t* symbols, int batch, double* dist, int32_t* success,
                     int32_t* status, double* trace, double* counters) {
  switch (chi_cap) {
    case 8: return decode_impl<8>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
    case 16: return decode_impl<16>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
    case 32: return decode_impl<32>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
    case 64: return decode_impl<64>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
    case 128: return decode_impl<128>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
    case 256: return decode_impl<256>(d, program, wtab, wdim, symbols, batch, dist, success, status, trace, counters);
  }
  return -2;
}
int emu_mono_decode_batch(const mdopt_decode_desc* d, int chi_cap, const int32_t* program, const double* wtab,
                          const int32_t* wdim, int n_tensors, const uint8_t* symbols, int batch, double* dist,
                          int32_t* success, int32_t* status, double* trace, double* counters) {
  switch (chi_cap) {
    case 8: return mono_impl<8>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 16: return mono_impl<16>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 32: return mono_impl<32>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 64: return mono_impl<64>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 128: return mono_impl<128>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 256: return mono_impl<256>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
    case 512: return mono_impl<512>(d, program, wtab, wdim, n_tensors, symbols, batch, dist, success, status, trace, counters);
  }
  return -2;
}
int emu_mps_program(const mdopt_decode_desc* d, int chi_cap, const int32_t* program, const double* wtab,
                    const int32_t* wdim, double* sites_io, int32_t* bonds_io, int32_t* centre_io, double* dist,
                    int32_t* success, int32_t* status, double* trace) {
  switch (chi_cap) {
    case 8: return program_impl<8>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
    case 16: return program_impl<16>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
    case 32: return program_impl<32>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
    case 64: return program_impl<64>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
    case 128: return program_impl<128>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
    case 256: return program_impl<256>(d, program, wtab, wdim, sites_io, bonds_io, centre_io, dist, success, status, trace);
  }
  return -2;
}
int emu_svd(int chi_cap, int m, int n, const double* a, int chi_max, double cut, double* u, double* sv, double* vh,
            int* kept) {
  switch (chi_cap) {
    case 8: return svd_impl<8>(m, n, a, chi_max, cut, u, sv, vh, kept);
    case 16: return svd_impl<16>(m, n, a, chi_max, cut, u, sv, vh, kept);
    case 32: return svd_impl<32>(m, n, a, chi_max, cut, u, sv, vh, kept);
    case 64: return svd_impl<64>(m, n, a, chi_max, cut, u, sv, vh, kept);
  }
  return -2;
}
int emu_split(int chi_cap, int m, int n, const double* a, double* xout, int* order, double* sig) {
  switch (chi_cap) {
    case 8: return split_impl<8>(m, n, a, xout, order, sig);
    case 16: return split_impl<16>(m, n, a, xout, order, sig);
    case 32: return split_impl<32>(m, n, a, xout, order, sig);
    case 64: return split_impl<64>(m, n, a, xout, order, sig);
    default: return -2;
  }
}
int emu_qr(int chi_cap, int m, int n, const double* a, int kmax, double rel_tol, double* q, double* coeff, int* more) {
  switch (chi_cap) {
    case 8: return qr_impl<8>(m, n, a, kmax, rel_tol, q, coeff, more);
    case 16: return qr_impl<16>(m, n, a, kmax, rel_tol, q, coeff, more);
    case 32: return qr_impl<32>(m, n, a, kmax, rel_tol, q, coeff, more);
    case 64: return qr_impl<64>(m, n, a, kmax, rel_tol, q, coeff, more);
  }
  return -2;
}
}
