// Block-cooperative FP64 building blocks for the MPS-MPO decode path (sm_100a).
//
// One CTA owns one matrix / one syndrome.  All routines are written as sequences of "parallel-for"
// phases separated by CTA barriers, with every reduction going through shared memory in a fixed order.
// That keeps the arithmetic order independent of the thread count, and lets the *same source* be
// compiled by g++ with -DMDOPT_EMU (MD_PAR_FOR becomes a plain loop) so the index logic can be unit
// tested on a CPU-only box.  The emulation build is test infrastructure (tests/emu); the product library
// is built by nvcc only and has no host compute path.
//
// Matrix convention: row-major, leading dimension LD = 2*CHI+1 doubles (odd, so that column walks do not
// hit a single shared-memory bank).  "X" is the 2*CHI x 2*CHI work matrix, "P2" a CHI x 2*CHI page.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef MDOPT_EMU
#include <algorithm>
#define MD_DEV inline
#define MD_PHASE inline
#define MD_HD inline
#define MD_SYNC() ((void)0)
#define MD_PAR_FOR(i, n) for (int i = 0; i < (n); ++i)
#define MD_LEADER if (true)
#else
#define MD_DEV __device__ __forceinline__
// The engine's big phases (sweep, gate, product state, read-out) are force-inlined at every call site on the
// shared-memory capacities (the per-site specialisation is worth ~4 %); the large-chi objects are compiled with
// -DMDOPT_NOINLINE_PHASES, which keeps one copy of each and cuts their compile time several times.
#ifdef MDOPT_NOINLINE_PHASES
#define MD_PHASE __device__ __noinline__
#else
#define MD_PHASE __device__ __forceinline__
#endif
#define MD_HD __host__ __device__ inline
#define MD_SYNC() __syncthreads()
#define MD_PAR_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define MD_LEADER if (threadIdx.x == 0)
#endif

namespace mdopt {

#ifdef MDOPT_EMU
using std::max;
using std::min;
struct double2 { double x, y; };  // host stand-in for the CUDA vector type
#endif

// program opcodes, factorisation modes and per-syndrome status codes (include/mdopt_b200.h), shared by both engines
enum { OP_END = 0, OP_SWEEP = 1, OP_MOVE = 2, OP_MARGINAL = 3, OP_GATE2 = 4, OP_COMPRESS = 5, OP_MARGINAL_DMRG = 6 };
enum { MODE_FAST = 0, MODE_SVD = 1, MODE_SVD_GRAM = 2, MODE_MONO = 3 };
enum { ST_OK = 0, ST_NOT_CONVERGED = 1, ST_CAPACITY = 2, ST_ZERO_NORM = 3, ST_NOT_MONOMIAL = 4 };

constexpr int NCH = 8;            // row chunks used by every column-wise partial reduction
constexpr double EPS = 2.220446049250313e-16;
constexpr double TOL3Z = 1.4901161193847656e-08;  // sqrt(eps), LAPACK dlaqp2's norm-recompute trigger

// Work counters a CTA accumulates (leader thread only); reported per launch for the roofline.
struct Counters {
  double flops_qr;      // Householder QRCP + forming Q
  double flops_jacobi;  // one-sided Jacobi rotations + dot products
  double flops_apply;   // site-apply / theta-build contractions and U^T X products
  double bytes_hbm;     // global-memory bytes moved by the CTA (site tensors + carried-matrix scratch)
  double n_steps;       // two-site steps (sweep + move)
  double n_trunc;       // steps that needed the truncating SVD
  double n_jacobi_sweeps;
  double n_qr_reflectors;
  // leader-thread clock64() deltas per phase (device build only)
  double cyc_qr, cyc_formq, cyc_jacobi, cyc_apply, cyc_io, cyc_coef, cyc_total, cyc_pad;
};

#ifdef MDOPT_EMU
#define MD_CLOCK() 0ll
#else
#define MD_CLOCK() clock64()
#endif

// 8-byte asynchronous global -> shared copy (cp.async, completes in the background; md_async_wait() before the
// data is read or the destination reused).  The host build copies synchronously.
#ifdef MDOPT_EMU
inline void md_async_copy8(double* dst_shared, const double* src_global) { *dst_shared = *src_global; }
inline void md_async_wait() {}
#else
__device__ __forceinline__ void md_async_copy8(double* dst_shared, const double* src_global) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src_global) : "memory");
}
// one FP64 tensor-core tile product, D(8x8) += A(8x4, row) B(4x8, col)  (SASS: DMMA.8x8x4)
__device__ __forceinline__ void md_dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void md_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
#endif

struct DecodeParams {
  int n_sites;        // MPS length N
  int n_logical;      // number of logical sites kL (they are sites 0..kL-1)
  int chi_max;        // user chi_max for sweeps
  int mode;           // MODE_FAST / MODE_SVD
  int renormalise;    // decode_css(renormalise=...)
  int trace_stride;   // doubles per traced step (0 = no trace)
  int trace_max;      // max traced steps per syndrome
  int n_ops;          // length of the program in ints
  double cut_sweep;   // user cut for sweep splits
  double cut_move;    // 1e-12 (reference default of split_two_site_tensor inside move_orth_centre)
  double rank_tol;    // MODE_FAST relative rank threshold
  double bias_p;      // bit-flip bias probability; < 0 disables the bias
  int n_head;         // read-out: number of leading sites in the dense stage (>= n_logical)
  unsigned kept_mask; // read-out: which of them are kept (the others are traced out)
  double readout_rel; // tolerant arg-max: eps = max(readout_rel * max v, readout_abs)
  double readout_abs;
  int n_tensors;      // entries of the MPO tensor table (<= TABLE_SMEM of them are cached in shared memory)
};

// Persistent per-CTA engine state (shared memory; see Engine in mps_core.cuh).
struct EngineState {
  double* sites;       // [N][SITE_STRIDE], packed (vL, p, vR) row-major per site
  double* scr0;        // carried matrix, flat (global, L2-resident)
  double* scr1;        // theta scratch for non-isometric steps
  int* bd;             // [N+1] bond dimensions (global, per CTA)
  const double* wtab;  // [n_tensors][16]
  const int* wdim;     // [n_tensors][4] = vL, vR, iso, pad
  double* trace;       // per-syndrome trace block or nullptr
  DecodeParams P;
  int n_traced;
  int status;
  int centre;
  int K, Wq, Mr;       // carried-matrix dimensions: M(K, 2, Wq, Mr)
  int mirrored;        // frame: 0 normal, 1 mirrored (moving the centre left)
  int in_x;            // 1: the carried matrix lives in s.X (ld LD) instead of scr0 (orthogonal fast path)
};

// Capacities up to SMEM_CHI_MAX keep the two work matrices in shared memory; larger ones (the large-chi
// tier) keep the same layout in a per-CTA slice of the global workspace (L2-resident at chi = 128) and only
// the O(chi) vectors in shared memory.
constexpr int SMEM_CHI_MAX = 64;
template <int CHI, bool IN_SMEM>
struct MatStore {
  double X[2 * CHI * (2 * CHI + 1)];   // work matrix (up to 2CHI x 2CHI)
  double P2[CHI * (2 * CHI + 1)];      // coefficient page (up to CHI x 2CHI)
};
template <int CHI>
struct MatStore<CHI, false> {
  double* X;
  double* P2;
};
// doubles of global scratch a CTA needs for X and P2 (0 when they live in shared memory)
MD_HD constexpr size_t mat_store_doubles(int chi) {
  return chi <= SMEM_CHI_MAX ? 0 : (size_t)3 * chi * (2 * chi + 1);
}

template <int CHI>
struct Smem : MatStore<CHI, (CHI <= SMEM_CHI_MAX)> {
  static constexpr int LD = 2 * CHI + 1;
  static constexpr int N2 = 2 * CHI;
  static constexpr bool MATS_IN_SMEM = (CHI <= SMEM_CHI_MAX);
  // bind X / P2 to a slice of global memory (no-op for the shared-memory capacities)
  MD_DEV void bind_mats(double* g) {
    if constexpr (!MATS_IN_SMEM) { this->X = g; this->P2 = g + (size_t)N2 * LD; }
  }
  static constexpr int NPART = (NCH * N2 * 3 / 2 > 256) ? NCH * N2 * 3 / 2 : 256;
  double part[NPART];
  double vn1[N2], vn2[N2], tau[N2], sig[N2], wv[N2];
  alignas(16) double rot[N2 * 2];
  double W[16];         // current MPO site tensor (vL, vR, pU, pD), vL,vR <= 2
  double dscr[16];
  static constexpr int NFLAG = (N2 > 32) ? N2 : 32;
  int perm[N2], order[N2], flag[NFLAG];
  int iscr[16];
  // per-launch constants cached next to the engine: MPO tensor table (when it has <= TABLE_SMEM entries) and the
  // bond dimensions of the chain (when it has < BOND_SMEM sites); both are read several times per step
  static constexpr int TABLE_SMEM = 16, BOND_SMEM = 1024;
  double wt_s[TABLE_SMEM * 16];
  int wd_s[TABLE_SMEM * 4];
  int bd_s[BOND_SMEM];
  Counters cnt;   // leader-only work counters of this CTA
  EngineState st;
};

// Column swizzle of the flat R x K matrix Q that form_q_reg leaves in s.P2: element (r, c) sits at
// r*K + (c ^ q_swizzle(r, K)).  With K a multiple of 8 the 8 lanes of a column group (rows 2 apart) and the
// transposed reads of store_u land on 8 different banks instead of one.
MD_DEV int q_swizzle(int r, int K) { return (K & 7) ? 0 : ((r >> 1) & 7); }

// -------------------------------------------------------------------------------------------------
// Householder QR with column pivoting, in place (LAPACK dlaqp2 data flow).
//   X (R x C) -> R factor in the upper trapezoid, reflector tails below the diagonal, tau[], perm[].
//   Stops at step j when the largest remaining column norm is <= rel_tol * (largest initial column norm)
//   or when j == kmax.  s.iscr[0] = number of reflectors K, s.iscr[1] = 1 if it stopped at kmax with a
//   residual above the threshold (i.e. a genuine truncation is required).
// -------------------------------------------------------------------------------------------------
template <int CHI>
MD_DEV void qrcp(Smem<CHI>& s, int R, int C, int kmax, double rel_tol) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  double* X = s.X;
  const int rows_per = (R + NCH - 1) / NCH;
  MD_PAR_FOR(t, NCH * C) {
    int ch = t / C, c = t - ch * C;
    int i0 = ch * rows_per, i1 = min(R, i0 + rows_per);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) { double v = X[i * LD + c]; acc += v * v; }
    s.part[ch * N2 + c] = acc;
  }
  MD_SYNC();
  MD_PAR_FOR(c, C) {
    double acc = 0.0;
    for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
    acc = sqrt(acc);
    s.vn1[c] = acc; s.vn2[c] = acc; s.perm[c] = c;
  }
  MD_SYNC();
  MD_LEADER {
    double mx = 0.0;
    for (int c = 0; c < C; ++c) mx = fmax(mx, s.vn1[c]);
    s.dscr[0] = mx * rel_tol;  // absolute stop threshold
    s.iscr[0] = 0; s.iscr[1] = 0;
    s.cnt.flops_qr += 2.0 * R * C;
  }
  MD_SYNC();
  const double thresh = s.dscr[0];
  const int jmax = min(R, C);
  int K = 0;
  for (int j = 0; j < jmax; ++j) {
    // ---- pivot search over trailing columns (two-level argmax through shared memory) ----
    MD_PAR_FOR(g, 32) {
      int span = (C - j + 31) / 32;
      int c0 = j + g * span, c1 = min(C, c0 + span);
      double best = -1.0; int bi = -1;
      for (int c = c0; c < c1; ++c) if (s.vn1[c] > best) { best = s.vn1[c]; bi = c; }
      s.part[g] = best; s.flag[g] = bi;
    }
    MD_SYNC();
    MD_LEADER {
      double best = -1.0; int bi = j;
      for (int g = 0; g < 32; ++g) if (s.part[g] > best) { best = s.part[g]; bi = s.flag[g]; }
      int stop = 0;
      if (!(best > thresh)) stop = 1;               // numerically rank deficient: done
      else if (j >= kmax) { stop = 1; s.iscr[1] = 1; }  // more weight than kmax columns can hold
      s.iscr[2] = stop; s.iscr[3] = bi;
    }
    MD_SYNC();
    if (s.iscr[2]) break;
    const int pv = s.iscr[3];
    if (pv != j) {
      MD_PAR_FOR(i, R) { double a = X[i * LD + j]; X[i * LD + j] = X[i * LD + pv]; X[i * LD + pv] = a; }
      MD_LEADER {
        int tp = s.perm[j]; s.perm[j] = s.perm[pv]; s.perm[pv] = tp;
        s.vn1[pv] = s.vn1[j]; s.vn2[pv] = s.vn2[j];
      }
    }
    MD_SYNC();
    // ---- Householder vector from X[j:R, j] ----
    MD_PAR_FOR(ch, NCH) {
      int len = R - (j + 1);
      int per = (len + NCH - 1) / NCH;
      int i0 = j + 1 + ch * per, i1 = min(R, i0 + per);
      double acc = 0.0;
      for (int i = i0; i < i1; ++i) { double v = X[i * LD + j]; acc += v * v; }
      s.part[ch] = acc;
    }
    MD_SYNC();
    MD_LEADER {
      double ss = 0.0;
      for (int ch = 0; ch < NCH; ++ch) ss += s.part[ch];
      double alpha = X[j * LD + j];
      double xnorm = sqrt(ss);
      double tau = 0.0, scal = 0.0, beta = alpha;
      if (xnorm != 0.0) {
        beta = -copysign(sqrt(alpha * alpha + ss), alpha);
        tau = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
      }
      s.tau[j] = tau; s.dscr[1] = scal;
      X[j * LD + j] = beta;
      s.cnt.flops_qr += 4.0 * (R - j) * (C - j - 1) + 3.0 * (R - j);
    }
    MD_SYNC();
    const double tau = s.tau[j];
    const double scal = s.dscr[1];
    MD_PAR_FOR(i, R - j - 1) { X[(j + 1 + i) * LD + j] *= scal; }
    MD_SYNC();
    const int nt = C - j - 1;  // trailing columns
    if (nt > 0 && tau != 0.0) {
      const int len = R - j;
      const int per = (len + NCH - 1) / NCH;
      MD_PAR_FOR(t, NCH * nt) {
        int ch = t / nt, c = j + 1 + (t - ch * nt);
        int i0 = j + ch * per, i1 = min(R, i0 + per);
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) {
          double v = (i == j) ? 1.0 : X[i * LD + j];
          acc += v * X[i * LD + c];
        }
        s.part[ch * N2 + c] = acc;
      }
      MD_SYNC();
      MD_PAR_FOR(t, nt) {
        int c = j + 1 + t;
        double acc = 0.0;
        for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
        s.wv[c] = tau * acc;
      }
      MD_SYNC();
      MD_PAR_FOR(t, len * nt) {
        int ii = t / nt, c = j + 1 + (t - ii * nt);
        int i = j + ii;
        double v = (i == j) ? 1.0 : X[i * LD + j];
        X[i * LD + c] -= v * s.wv[c];
      }
      MD_SYNC();
    }
    // ---- column-norm downdate with LAPACK's cancellation safeguard ----
    MD_PAR_FOR(t, nt) {
      int c = j + 1 + t;
      int redo = 0;
      if (s.vn1[c] != 0.0) {
        double temp = fabs(X[j * LD + c]) / s.vn1[c];
        temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
        double r = s.vn1[c] / s.vn2[c];
        double temp2 = temp * r * r;
        if (temp2 <= TOL3Z) redo = 1;
        else s.vn1[c] *= sqrt(temp);
      }
      s.flag[c] = redo;
    }
    MD_SYNC();
    MD_PAR_FOR(t, nt) {
      int c = j + 1 + t;
      if (s.flag[c]) {
        double acc = 0.0;
        for (int i = j + 1; i < R; ++i) { double v = X[i * LD + c]; acc += v * v; }
        acc = sqrt(acc);
        s.vn1[c] = acc; s.vn2[c] = acc;
      }
    }
    MD_SYNC();
    K = j + 1;
  }
  MD_LEADER { s.iscr[0] = K; }
  MD_SYNC();
}

// Coefficients of the original columns in the new basis: P2[k, perm[c]] = R[k, c] (K x C, ld LD).
template <int CHI>
MD_DEV void extract_coeff(Smem<CHI>& s, int K, int C) {
  constexpr int LD = Smem<CHI>::LD;
  MD_PAR_FOR(t, K * C) {
    int k = t / C, c = t - k * C;
    s.P2[k * LD + s.perm[c]] = (c >= k) ? s.X[k * LD + c] : 0.0;
  }
  MD_SYNC();
}

// Form the first K columns of Q in place over the reflectors (LAPACK dorg2r data flow).
template <int CHI>
MD_DEV void form_q(Smem<CHI>& s, int R, int K) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  double* X = s.X;
  for (int j = K - 1; j >= 0; --j) {
    const double tau = s.tau[j];
    const int nt = K - 1 - j;
    const int len = R - j;
    if (nt > 0 && tau != 0.0) {
      const int per = (len + NCH - 1) / NCH;
      MD_PAR_FOR(t, NCH * nt) {
        int ch = t / nt, c = j + 1 + (t - ch * nt);
        int i0 = j + ch * per, i1 = min(R, i0 + per);
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) {
          double v = (i == j) ? 1.0 : X[i * LD + j];
          acc += v * X[i * LD + c];
        }
        s.part[ch * N2 + c] = acc;
      }
      MD_SYNC();
      MD_PAR_FOR(t, nt) {
        int c = j + 1 + t;
        double acc = 0.0;
        for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
        s.wv[c] = tau * acc;
      }
      MD_SYNC();
      MD_PAR_FOR(t, len * nt) {
        int ii = t / nt, c = j + 1 + (t - ii * nt);
        int i = j + ii;
        double v = (i == j) ? 1.0 : X[i * LD + j];
        X[i * LD + c] -= v * s.wv[c];
      }
      MD_SYNC();
    }
    MD_PAR_FOR(i, R) {
      double v;
      if (i < j) v = 0.0;
      else if (i == j) v = 1.0 - tau;
      else v = -tau * X[i * LD + j];
      X[i * LD + j] = v;
    }
    MD_LEADER { s.cnt.flops_qr += 4.0 * len * nt + 2.0 * len; }
    MD_SYNC();
  }
}

// -------------------------------------------------------------------------------------------------
// Canonical order of the columns by singular value (the tie-break that makes truncating decodes well-posed; the
// monomial engine and oracle/structured_svd.py state the same rule): sigma descending; values that agree to
// ORDER_TIE_REL relative are tied, and ties go to the column whose first non-zero row comes first, then to the lower
// column index.  first_row[c] must hold the first row with a non-zero entry in column c (R if none).
//   before(o, c) == true  <=>  column o is ranked ahead of column c
// -------------------------------------------------------------------------------------------------
constexpr double ORDER_TIE_REL = 1e-10;
MD_HD bool canon_before(double w, int fo, int o, double v, int fc, int c) {
  if (w > v * (1.0 + ORDER_TIE_REL)) return true;
  if (v > w * (1.0 + ORDER_TIE_REL)) return false;
  return (fo < fc) || (fo == fc && o < c);
}
// first non-zero row of every column of X (R x C, ld LD) -> out[c]
template <int CHI>
MD_DEV void first_rows(const Smem<CHI>& s, int R, int C, int* out) {
  constexpr int LD = Smem<CHI>::LD;
#ifdef MDOPT_EMU
  MD_PAR_FOR(c, C) {
    int f = R;
    for (int i = 0; i < R; ++i) { if (s.X[i * LD + c] != 0.0) { f = i; break; } }
    out[c] = f;
  }
#else
  // four adjacent lanes per column, each scanning every fourth row without an early exit (independent loads
  // instead of a chain of dependent ones); the minimum meets by shuffle
  for (int t4 = threadIdx.x; t4 < 4 * (((C + 7) >> 3) << 3); t4 += blockDim.x) {
    const int c = t4 >> 2, q = t4 & 3;
    int f = R;
    if (c < C) {
#pragma unroll 8
      for (int i = R - 4 + q; i >= 0; i -= 4) { if (s.X[i * LD + c] != 0.0) f = i; }
    }
    f = min(f, __shfl_xor_sync(0xffffffffu, f, 1));
    f = min(f, __shfl_xor_sync(0xffffffffu, f, 2));
    if (c < C && q == 0) out[c] = f;
  }
#endif
  MD_SYNC();
}

// -------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on the columns of X (R x C), in place: X <- X J = U diag(sigma).
//   Round-robin pair schedule, partial dot products through shared memory, rotation applied by all
//   threads.  On exit: s.sig[c] = ||X[:,c]||, s.order[] = columns sorted by sigma descending (ties by
//   index).  Returns the number of sweeps in s.iscr[4]; s.iscr[5] = 1 when it did not converge.
// -------------------------------------------------------------------------------------------------
template <int CHI>
MD_DEV void jacobi_columns(Smem<CHI>& s, int R, int C, int max_sweeps = 60) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  double* X = s.X;
  const int n = (C + 1) & ~1;     // even number of players (last one may be a phantom)
  const int npair = n / 2;
  const double tol = EPS * sqrt((double)R);
  const int per = (R + NCH - 1) / NCH;
  // columns whose norm is below 4 eps * (largest column norm) are numerically zero: pairs involving
  // one are never rotated (the angle is rounding noise and would never settle) and truncation_rule
  // never keeps them, so U stays an isometry.
  MD_PAR_FOR(t, NCH * C) {
    int ch = t / C, c = t - ch * C;
    int i0 = ch * per, i1 = min(R, i0 + per);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) { double v = X[i * LD + c]; acc += v * v; }
    s.part[ch * N2 + c] = acc;
  }
  MD_SYNC();
  MD_LEADER {
    double mx = 0.0;
    for (int c = 0; c < C; ++c) {
      double acc = 0.0;
      for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
      mx = fmax(mx, acc);
    }
    s.dscr[6] = mx * (4.0 * EPS) * (4.0 * EPS);  // squared-norm floor
  }
  MD_SYNC();
  const double floor2 = s.dscr[6];
  int sweeps = 0;
  int converged = (C < 2);
  while (!converged && sweeps < max_sweeps) {
    MD_LEADER { s.iscr[6] = 0; }
    MD_SYNC();
    for (int rnd = 0; rnd < n - 1; ++rnd) {
      // pair k of this round: circle method with player n-1 fixed
      MD_PAR_FOR(t, npair * NCH) {
        int k = t / NCH, ch = t - k * NCH;
        int a, b;
        if (k == 0) { a = n - 1; b = rnd; }
        else { a = (rnd + k) % (n - 1); b = (rnd - k + (n - 1)) % (n - 1); }
        double saa = 0.0, sbb = 0.0, sab = 0.0;
        if (a < C && b < C) {
          int i0 = ch * per, i1 = min(R, i0 + per);
          for (int i = i0; i < i1; ++i) {
            double xa = X[i * LD + a], xb = X[i * LD + b];
            saa += xa * xa; sbb += xb * xb; sab += xa * xb;
          }
        }
        double* p = s.part + (k * NCH + ch) * 3;
        p[0] = saa; p[1] = sbb; p[2] = sab;
      }
      MD_SYNC();
      MD_PAR_FOR(k, npair) {
        double saa = 0.0, sbb = 0.0, sab = 0.0;
        for (int ch = 0; ch < NCH; ++ch) {
          const double* p = s.part + (k * NCH + ch) * 3;
          saa += p[0]; sbb += p[1]; sab += p[2];
        }
        double c = 1.0, sn = 0.0;
        int did = 0;
        if (sab != 0.0 && fabs(sab) > tol * sqrt(saa) * sqrt(sbb) && saa > floor2 && sbb > floor2) {
          double zeta = (sbb - saa) / (2.0 * sab);
          double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = 1.0 / sqrt(1.0 + t * t);
          sn = c * t;
          did = 1;
        }
        s.rot[2 * k] = c; s.rot[2 * k + 1] = sn;
        s.flag[k] = did;
      }
      MD_SYNC();
      MD_PAR_FOR(t, npair * R) {
        int k = t / R, i = t - k * R;
        if (s.flag[k]) {
          int a, b;
          if (k == 0) { a = n - 1; b = rnd; }
          else { a = (rnd + k) % (n - 1); b = (rnd - k + (n - 1)) % (n - 1); }
          double c = s.rot[2 * k], sn = s.rot[2 * k + 1];
          double xa = X[i * LD + a], xb = X[i * LD + b];
          X[i * LD + a] = c * xa - sn * xb;
          X[i * LD + b] = sn * xa + c * xb;
        }
      }
      MD_LEADER {
        int nrot = 0;
        for (int k = 0; k < npair; ++k) nrot += s.flag[k];
        s.iscr[6] += nrot;
        s.cnt.flops_jacobi += 6.0 * R * npair + 6.0 * R * nrot;
      }
      MD_SYNC();
    }
    ++sweeps;
    converged = (s.iscr[6] == 0);
    MD_SYNC();
  }
  // singular values = column norms; order by value (descending), ties by column index
  MD_PAR_FOR(t, NCH * C) {
    int ch = t / C, c = t - ch * C;
    int i0 = ch * per, i1 = min(R, i0 + per);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) { double v = X[i * LD + c]; acc += v * v; }
    s.part[ch * N2 + c] = acc;
  }
  MD_SYNC();
  MD_PAR_FOR(c, C) {
    double acc = 0.0;
    for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
    s.sig[c] = sqrt(acc);
  }
  MD_SYNC();
  first_rows<CHI>(s, R, C, s.flag);
  MD_PAR_FOR(c, C) {
    double v = s.sig[c];
    int pos = 0;
    for (int o = 0; o < C; ++o) pos += canon_before(s.sig[o], s.flag[o], o, v, s.flag[c], c) ? 1 : 0;
    s.order[pos] = c;
  }
  MD_LEADER {
    s.iscr[4] = sweeps; s.iscr[5] = converged ? 0 : 1;
    s.cnt.n_jacobi_sweeps += sweeps;
    s.cnt.flops_jacobi += 2.0 * R * C;
  }
  MD_SYNC();
}

// -------------------------------------------------------------------------------------------------
// Orthogonality probe, portable version (host emulation and the large-chi tier; the shared-memory capacities use
// the fused copy inside jacobi_columns_dev).  Freivalds test with two fixed pseudo-random vectors z:
// X^T (X z) = diag(|x_c|^2) z holds for every z exactly when the columns of X are mutually orthogonal.
// Returns 1 when every component agrees to 1e-12 |x_c| |X z|; then s.sig = column norms, s.order = their ranking
// (descending, ties by column index), s.dscr[6] = the zero-column floor, and X is untouched.
// -------------------------------------------------------------------------------------------------
#ifdef MDOPT_EMU
// host emulation only: lets a test run the same decode with and without the pair rotations (tests/emu/emu.cpp)
inline int& emu_pair_fix_off() { static int off = 0; return off; }
#endif
// s.vn2 (squared column norms) -> s.sig, s.order (the canonical ranking); s.iscr[4] = s.iscr[5] = 0
template <int CHI>
MD_DEV void rank_columns(Smem<CHI>& s, int R, int C) {
  MD_PAR_FOR(c, C) s.sig[c] = sqrt(s.vn2[c]);
  MD_SYNC();
  first_rows<CHI>(s, R, C, s.flag);
  // four tasks per column, each ranking it against a quarter of the others; the partial ranks meet in s.part (ints)
  constexpr int N2 = Smem<CHI>::N2;
  int* ipart = reinterpret_cast<int*>(s.part);
  MD_PAR_FOR(t, 4 * C) {
    const int c = t % C, q = t / C;
    const double v = s.sig[c];
    const int fc = s.flag[c];
    int pos = 0;
    for (int o = q; o < C; o += 4) pos += canon_before(s.sig[o], s.flag[o], o, v, fc, c) ? 1 : 0;
    ipart[q * N2 + c] = pos;
  }
  MD_SYNC();
  MD_PAR_FOR(c, C) s.order[(ipart[c] + ipart[N2 + c]) + (ipart[2 * N2 + c] + ipart[3 * N2 + c])] = c;
  MD_LEADER { s.iscr[4] = 0; s.iscr[5] = 0; }
  MD_SYNC();
}

MD_HD double probe_z(int c, unsigned salt) {
  unsigned h = (unsigned)c * 2654435761u + salt;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  return 1.0 + (double)(h & 0xFFFFFu) * (1.0 / 1048576.0);   // in [1, 2)
}

template <int CHI>
MD_DEV int ortho_probe(Smem<CHI>& s, int R, int C) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  const double* X = s.X;
  double* z0 = s.vn1;
  double* z1 = s.sig;
  double* y0 = s.wv;
  double* y1 = s.tau;
  double* pr = s.part;   // [3][4][N2] partial sums
  MD_PAR_FOR(c, C) { z0[c] = probe_z(c, 0x9E3779B9u); z1[c] = probe_z(c, 0x7F4A7C15u); }
  MD_SYNC();
  // pass 1: y = X z; task t = (row t % N2, quarter t / N2 of the columns -- a contiguous quarter, so that a task
  // walking its row re-uses the 32-byte sectors it fetched when X lives in global memory)
  const int cq = (C + 3) / 4;
  MD_PAR_FOR(t, 4 * N2) {
    const int idx = t % N2, grp = t / N2;
    double a0 = 0.0, a1 = 0.0;
    if (idx < R) {
      const int c1 = (grp + 1) * cq < C ? (grp + 1) * cq : C;
      for (int c = grp * cq; c < c1; ++c) { const double x = X[idx * LD + c]; a0 += x * z0[c]; a1 += x * z1[c]; }
    }
    pr[(0 * 4 + grp) * N2 + idx] = a0;
    pr[(1 * 4 + grp) * N2 + idx] = a1;
  }
  MD_SYNC();
  MD_PAR_FOR(r, R) {
    y0[r] = (pr[r] + pr[N2 + r]) + (pr[2 * N2 + r] + pr[3 * N2 + r]);
    y1[r] = (pr[4 * N2 + r] + pr[5 * N2 + r]) + (pr[6 * N2 + r] + pr[7 * N2 + r]);
  }
  MD_SYNC();
  MD_PAR_FOR(k, 16) {   // |y|^2 in 16 interleaved partial sums (s.rot holds 2 N2 >= 32 doubles) (s.rot), then the leader
    double n0 = 0.0, n1 = 0.0;
    for (int r = k; r < R; r += 16) { n0 += y0[r] * y0[r]; n1 += y1[r] * y1[r]; }
    s.rot[k] = n0; s.rot[16 + k] = n1;
  }
  MD_SYNC();
  MD_LEADER {
    double n0 = 0.0, n1 = 0.0;
    for (int k = 0; k < 16; ++k) { n0 += s.rot[k]; n1 += s.rot[16 + k]; }
    s.dscr[8] = n0; s.dscr[9] = n1;
  }
  // pass 2: w = X^T y and the squared column norms; task t = (column t % N2, quarter t / N2 of the rows)
  MD_PAR_FOR(t, 4 * N2) {
    const int idx = t % N2, grp = t / N2;
    double w0 = 0.0, w1 = 0.0, d = 0.0;
    if (idx < C) {
      for (int r = grp; r < R; r += 4) { const double x = X[r * LD + idx]; w0 += x * y0[r]; w1 += x * y1[r]; d += x * x; }
    }
    pr[(0 * 4 + grp) * N2 + idx] = w0;
    pr[(1 * 4 + grp) * N2 + idx] = w1;
    pr[(2 * 4 + grp) * N2 + idx] = d;
  }
  MD_SYNC();
  const double ny0 = s.dscr[8], ny1 = s.dscr[9];
  MD_PAR_FOR(c, C) {
    const double w0 = (pr[c] + pr[N2 + c]) + (pr[2 * N2 + c] + pr[3 * N2 + c]);
    const double w1 = (pr[4 * N2 + c] + pr[5 * N2 + c]) + (pr[6 * N2 + c] + pr[7 * N2 + c]);
    const double d = (pr[8 * N2 + c] + pr[9 * N2 + c]) + (pr[10 * N2 + c] + pr[11 * N2 + c]);
    const double e0 = w0 - d * z0[c], e1 = w1 - d * z1[c];
    s.vn2[c] = d;
    s.flag[c] = (e0 * e0 > 1e-24 * d * ny0 || e1 * e1 > 1e-24 * d * ny1) ? 1 : 0;
  }
  MD_SYNC();
  MD_PAR_FOR(k, 16) {
    double mx = 0.0;
    for (int c = k; c < C; c += 16) mx = fmax(mx, s.vn2[c]);
    s.rot[k] = mx;
  }
  MD_SYNC();
  MD_LEADER {
    double mx = 0.0;
    for (int k = 0; k < 16; ++k) mx = fmax(mx, s.rot[k]);
    s.dscr[6] = mx * (4.0 * EPS) * (4.0 * EPS);
    s.cnt.flops_jacobi += 10.0 * R * C;
  }
  MD_SYNC();
  // a numerically zero column (rounding residue of earlier rotations, below the floor the rotations and the
  // truncation rule ignore) has a random direction and would fail the test for ever: it does not count
  MD_PAR_FOR(k, 16) {
    int bad = 0;
    for (int c = k; c < C; c += 16) {
      if (s.flag[c] && !(s.vn2[c] > s.dscr[6])) s.flag[c] = 0;
      bad |= s.flag[c];
    }
    s.rot[16 + k] = bad ? 1.0 : 0.0;
  }
  MD_SYNC();
  MD_LEADER {
    int bad = 0;
    for (int k = 0; k < 16; ++k) bad |= (s.rot[16 + k] != 0.0);
    s.iscr[11] = bad ? 0 : 1;
  }
  MD_SYNC();
  if (!s.iscr[11]) return 0;
  rank_columns<CHI>(s, R, C);
  return 1;
}

#ifndef MDOPT_EMU
// Device form of the dot-product tables of pair_fix: 8 x 8 tiles on the FP64 tensor pipe (mma.m8n8k4.f64), one tile per
// warp task, rows la[0..na) x columns lb[0..nb) (a null list is the identity).  Lane l feeds A[m = l/4][k = l%4] =
// X[r0 + k][la[m]] and B[k][n = l/4] = X[r0 + k][lb[n]] and receives the entries (m = l/4, n = 2 (l%4) + {0, 1}).
// A warp load touches 4 rows x 8 columns, so a tile of 64 dot products costs R/2 load instructions instead of the
// 128 R of 64 scalar column walks -- what matters when X is a row-major matrix in global memory.
// row_wanted(m) (list position) lets the caller skip tiles none of whose rows is of interest; visit(m, n, value)
// receives list positions.
template <typename Want, typename Visit>
__device__ __forceinline__ void md_tile_dots(const double* X, int LD, int R, const int* la, int na, const int* lb,
                                             int nb, Want&& row_wanted, Visit&& visit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (int)(blockDim.x >> 5);
  const int nta = (na + 7) >> 3, ntb = (nb + 7) >> 3;
  for (int t = warp; t < nta * ntb; t += nwarps) {
    const int ti = t / ntb, tj = t - ti * ntb;
    const int ma = 8 * ti + (lane >> 2);
    if (!__any_sync(0xffffffffu, ma < na && row_wanted(ma))) continue;
    const int ia = min(ma, na - 1), ib = min(8 * tj + (lane >> 2), nb - 1);
    const int ca = la ? la[ia] : ia, cb = lb ? lb[ib] : ib;
    const double* pa = X + (size_t)(lane & 3) * LD + ca;
    const double* pb = X + (size_t)(lane & 3) * LD + cb;
    double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
    const int R8 = R & ~7;
#pragma unroll 4
    for (int r0 = 0; r0 < R8; r0 += 8) {
      const double a0 = pa[(size_t)r0 * LD], b0 = pb[(size_t)r0 * LD];
      const double a1 = pa[(size_t)(r0 + 4) * LD], b1 = pb[(size_t)(r0 + 4) * LD];
      md_dmma_8x8x4(acc0, a0, b0);
      md_dmma_8x8x4(acc1, a1, b1);
    }
    for (int r0 = R8; r0 < R; r0 += 4) {
      const bool ok = r0 + (lane & 3) < R;
      md_dmma_8x8x4(acc0, ok ? pa[(size_t)r0 * LD] : 0.0, ok ? pb[(size_t)r0 * LD] : 0.0);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ob = 8 * tj + 2 * (lane & 3) + e;
      if (ma < na && ob < nb) visit(ma, ob, acc0[e] + acc1[e]);
    }
  }
}
#endif

// dot product of two strided vectors of length R: four independent partial sums, so that the loads of a matrix in
// global memory are in flight together instead of one per dependent FMA
MD_HD double dot4(const double* pa, int sa, const double* pc, int sc, int R) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = 0;
  for (; r + 3 < R; r += 4) {
    s0 += pa[(r + 0) * sa] * pc[(r + 0) * sc];
    s1 += pa[(r + 1) * sa] * pc[(r + 1) * sc];
    s2 += pa[(r + 2) * sa] * pc[(r + 2) * sc];
    s3 += pa[(r + 3) * sa] * pc[(r + 3) * sc];
  }
  for (; r < R; ++r) s0 += pa[r * sa] * pc[r * sc];
  return (s0 + s1) + (s2 + s3);
}

// -------------------------------------------------------------------------------------------------
// Portable form of the probe-mode pair rotations of jacobi_columns_dev (same rule, same outputs), asked after a failed
// ortho_probe: the flagged columns (s.flag) are the candidates; their mutual dot products, then those of the still
// unmatched ones with the unflagged columns, are tested exactly, and when EVERY flagged column has exactly one
// partner the disjoint pairs are rotated in one round.  Returns 1 then: X <- X J, s.sig / s.order as after the
// sweeps, and the 2 x 2 blocks of J in s.vn1 (J[c, c]), s.wv (J[partner, c]), s.perm (partner, -1 = untouched) for
// the two-entries-per-row coefficient matrix (layout 4 of Engine::step).  Returns 0 with X untouched otherwise.
// greedy = 1 (the step before the cyclic sweeps, X already saved by the caller): columns with several partners take
// the one with the largest cosine, whatever disjoint pairs result are rotated, and the caller asks the probe again
// -- classical Jacobi restricted to the offending pairs instead of C - 1 rounds over all of them.
// No atomics: candidates are tabulated in `table` and matched by the leader, so host emulation and device agree.
// -------------------------------------------------------------------------------------------------
template <int CHI>
MD_DEV int pair_fix(Smem<CHI>& s, int R, int C, double* table, int greedy = 0) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  double* X = s.X;
  const double tol = EPS * sqrt((double)R);
  const double tol2 = tol * tol;
  const double floor2 = s.dscr[6];
  // cand[-1 - i]: partner of flagged column i (-1 none yet, -2 several), in the tail of s.part viewed as ints (the
  // probe's partial sums are dead by now)
  static_assert(2 * Smem<CHI>::NPART >= N2, "one int per column fits s.part");
  int* cand = reinterpret_cast<int*>(s.part) + N2;
  MD_LEADER {
    int nf = 0;
    for (int c = 0; c < C; ++c) if (s.flag[c]) s.order[nf++] = c;
    s.iscr[13] = nf;
  }
  MD_PAR_FOR(c, C) s.perm[c] = -1;
  MD_SYNC();
  const int nf = s.iscr[13];
  if (nf < 1) return 0;
  // tables of offending dot products (0 elsewhere) in `table` (N2 * N2 doubles: the flat scratch matrix the split
  // does not read); G1 is dead once its rows are scanned
  double* G1 = table;   // nf x nf: flagged x flagged
  double* G2 = table;   // nf x C : rows of the flagged columns still without a partner x unflagged columns
#ifndef MDOPT_EMU
  md_tile_dots(X, LD, R, s.order, nf, s.order, nf, [](int) { return true; }, [&](int i, int j, double sab) {
    double g = 0.0;
    if (i != j) {
      const double saa = s.vn2[s.order[i]], scc = s.vn2[s.order[j]];
      if (sab != 0.0 && sab * sab > tol2 * saa * scc && saa > floor2 && scc > floor2) g = sab;
    }
    G1[i * nf + j] = g;
  });
#else
  MD_PAR_FOR(t, nf * nf) {
    const int i = t / nf, j = t - i * nf;
    double g = 0.0;
    if (i != j) {
      const int a = s.order[i], c = s.order[j];
      const double sab = dot4(X + a, LD, X + c, LD, R);
      const double saa = s.vn2[a], scc = s.vn2[c];
      if (sab != 0.0 && sab * sab > tol2 * saa * scc && saa > floor2 && scc > floor2) g = sab;
    }
    G1[t] = g;
  }
#endif
  MD_LEADER { s.cnt.flops_jacobi += 2.0 * R * nf * nf; }
  MD_SYNC();
  MD_PAR_FOR(i, nf) {
    int cnt = 0, b = -1;
    double best = 0.0;
    for (int j = 0; j < nf; ++j) {
      const double g = G1[i * nf + j];
      if (g != 0.0) {
        ++cnt;
        const double w = g * g / s.vn2[s.order[j]];   // greedy: the partner with the largest cosine
        if (!greedy || w > best) { best = w; b = j; }
      }
    }
    cand[-1 - i] = (cnt == 0) ? -1 : ((cnt == 1 || greedy) ? s.order[b] : -2);
    if (cnt == 1 || (greedy && cnt > 0)) s.tau[s.order[i]] = G1[i * nf + b];   // the pair's dot product, kept for the rotation
  }
  MD_SYNC();
  // Flagged columns without a partner among the flagged: the probe's criterion for a column is relative to that
  // column's own norm, so the larger column of a very unequal pair can pass while the smaller one is flagged.  Their
  // partners are looked for among the unflagged columns; consecutive tasks walk consecutive columns c (coalesced when
  // X lives in global memory), the flagged column is the same for a run of tasks (broadcast).
#ifndef MDOPT_EMU
  md_tile_dots(X, LD, R, s.order, nf, nullptr, C, [&](int i) { return cand[-1 - i] == -1; },
               [&](int i, int c, double sab) {
    if (cand[-1 - i] == -1) {
      double g = 0.0;
      if (!s.flag[c]) {
        const double saa = s.vn2[s.order[i]], scc = s.vn2[c];
        if (sab != 0.0 && sab * sab > tol2 * saa * scc && saa > floor2 && scc > floor2) g = sab;
      }
      G2[i * C + c] = g;
    }
  });
#else
  MD_PAR_FOR(t, nf * C) {
    const int i = t / C, c = t - i * C;
    if (cand[-1 - i] == -1) {
      double g = 0.0;
      if (!s.flag[c]) {
        const int a = s.order[i];
        const double sab = dot4(X + a, LD, X + c, LD, R);
        const double saa = s.vn2[a], scc = s.vn2[c];
        if (sab != 0.0 && sab * sab > tol2 * saa * scc && saa > floor2 && scc > floor2) g = sab;
      }
      G2[t] = g;
    }
  }
#endif
  MD_SYNC();
  MD_PAR_FOR(i, nf) {
    if (cand[-1 - i] == -1) {
      int cnt = 0, b = -1;
      double best = 0.0;
      for (int c = 0; c < C; ++c) {
        const double g = G2[i * C + c];
        if (g != 0.0) {
          ++cnt;
          const double w = g * g / s.vn2[c];
          if (!greedy || w > best) { best = w; b = c; }
        }
      }
      cand[-1 - i] = (cnt == 0) ? -1 : ((cnt == 1 || greedy) ? b : -2);
      if (cnt == 1 || (greedy && cnt > 0)) s.tau[s.order[i]] = G2[i * C + b];
    }
  }
  MD_SYNC();
  // every flagged column must have exactly one partner, and the pairs must be disjoint
  // (greedy: whatever disjoint pairs the flagged columns' best partners form, at least one)
  MD_LEADER {
    int ok = 1, np = 0;
    for (int i = 0; i < nf && ok; ++i) {
      const int a = s.order[i], b = cand[-1 - i];
      if (b < 0) ok = greedy;
      else if (s.perm[a] == -1 && s.perm[b] == -1) {
        s.perm[a] = b; s.perm[b] = a; s.rot[2 * (a < b ? a : b)] = s.tau[a];
        ++np;
      } else if (!(s.perm[a] == b && s.perm[b] == a)) ok = greedy;
    }
    s.iscr[12] = (ok && np > 0) ? 0 : 1;
  }
  MD_SYNC();
  if (s.iscr[12]) return 0;
  // every flagged column has exactly one partner: rotate the pairs (a < b = s.perm[a]); s.rot[2a] = their dot product
  MD_PAR_FOR(a, C) {
    const int b = s.perm[a];
    if (b > a) {
      const double saa = s.vn2[a], sbb = s.vn2[b], sab = s.rot[2 * a];
      const double alpha = sbb - saa, beta = 2.0 * sab;
      const double t = copysign(beta, alpha * beta) / (fabs(alpha) + sqrt(alpha * alpha + beta * beta));
      const double c = 1.0 / sqrt(1.0 + t * t);
      const double sn = c * t;
      s.vn1[a] = c; s.wv[a] = -sn;
      s.vn1[b] = c; s.wv[b] = sn;
    }
  }
  MD_LEADER {
    int np = 0;
    for (int a = 0; a < C; ++a) if (s.perm[a] > a) s.order[np++] = a;
    s.iscr[13] = np;
    s.cnt.flops_jacobi += 12.0 * R * np;
  }
  MD_SYNC();
  const int np = s.iscr[13];
  MD_PAR_FOR(t, np * R) {
    const int q = t / R, r = t - q * R;
    const int a = s.order[q], b = s.perm[a];
    const double c = s.vn1[a], sn = s.wv[b];
    const double xa = X[r * LD + a], xb = X[r * LD + b];
    X[r * LD + a] = c * xa - sn * xb;
    X[r * LD + b] = sn * xa + c * xb;
  }
  MD_SYNC();
  // fresh squared norms of the rotated columns: NCH row chunks per column, partial sums in s.part (behind the ints)
  double* pn = s.part + N2;
  const int rows_per = (R + NCH - 1) / NCH;
  MD_PAR_FOR(t, 2 * np * NCH) {
    const int ch = t % NCH, q = t / NCH;
    const int a = s.order[q >> 1];
    const int col = (q & 1) ? s.perm[a] : a;
    const int r0 = ch * rows_per, r1 = (r0 + rows_per < R) ? r0 + rows_per : R;
    pn[ch * N2 + col] = (r0 < r1) ? dot4(X + (size_t)r0 * LD + col, LD, X + (size_t)r0 * LD + col, LD, r1 - r0) : 0.0;
  }
  MD_SYNC();
  MD_PAR_FOR(t, 2 * np) {
    const int a = s.order[t >> 1];
    const int col = (t & 1) ? s.perm[a] : a;
    double acc = 0.0;
    for (int ch = 0; ch < NCH; ++ch) acc += pn[ch * N2 + col];
    s.vn2[col] = acc;
  }
  MD_SYNC();
  MD_LEADER { s.cnt.n_jacobi_sweeps += 1.0; }
  if (greedy) { MD_SYNC(); return 1; }   // the caller asks the probe again
  rank_columns<CHI>(s, R, C);
  MD_LEADER { s.iscr[4] = 1; }
  MD_SYNC();
  return 1;
}

// Reference truncation rule (mdopt/utils/utils.py:119): keep min(chi_max, #(sigma > cut)) values.
// Requires s.sig / s.order from jacobi_columns.  Result in s.iscr[0]; s.dscr[2] = truncation error
// sum_{i>=k} sigma_i^2, s.dscr[3] = ||sigma_kept||_2.
template <int CHI>
MD_DEV void truncation_rule(Smem<CHI>& s, int C, int chi_lim, double cut) {
  MD_LEADER {
    int k = 0;
    const double sig_floor = sqrt(s.dscr[6]);
    for (int i = 0; i < C; ++i) { double v = s.sig[s.order[i]]; k += (v > cut && v > sig_floor); }
    if (k > chi_lim) {   // chi_max cuts into the values the cut would have kept: a truncating step
      k = chi_lim;
      s.cnt.n_trunc += 1.0;
    }
    double err = 0.0, kept = 0.0;
    for (int i = 0; i < C; ++i) {
      double v = s.sig[s.order[i]];
      if (i >= k) err += v * v; else kept += v * v;
    }
    s.iscr[0] = k; s.dscr[2] = err; s.dscr[3] = sqrt(kept);
  }
  MD_SYNC();
}


#ifndef MDOPT_EMU
// =================================================================================================
// Device-only fast path: the same Householder QR with column pivoting, with the matrix held in
// REGISTERS (4 threads per column, rows interleaved mod 4) and warp shuffles for every reduction.
// Per reflector: one 4-lane serial section (pivot column -> v, tau), two CTA barriers, and an
// embarrassingly parallel rank-1 update.  Reflectors are kept in shared memory (s.P2 viewed as
// Vall[K][N2]) for the Q formation, which runs in registers with no barriers at all.
//   in : orig (R x C, row-major ld C, global)
//   out: s.X[k*LD + c] = coefficient of original column c on basis vector k  (k < K)
//        s.P2 = Vall, s.tau;  s.iscr[0] = K, s.iscr[1] = genuine-truncation flag
// Pivot rule and stop rule are identical to qrcp<> above (largest remaining column norm, ties to the
// lowest column index; stop at rel_tol * largest initial norm or at kmax).
// =================================================================================================
template <int CHI, int TPC, int CPT>
__device__ __noinline__ void qrcp_reg(Smem<CHI>& s, const double* orig, int R, int C, int kmax, double rel_tol) {
  // TPC threads share a column (rows interleaved in pairs), each thread group owns CPT columns
  // (col, col + NCOL, ...).  <8,1> covers C <= CHI (every centre move), <8,2> covers C <= 2*CHI.
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int NT = (CHI * 8 < 64) ? 64 : CHI * 8;
  constexpr int NCOL = NT / TPC;            // columns covered by one pass of the thread groups
  constexpr int RPT = N2 / TPC;             // register rows per thread and column
  static_assert(RPT <= 16, "the pivot-slot switch has 16 cases");
  constexpr int CH = RPT < 8 ? RPT : 8;     // register rows per chunk
  constexpr int NCHK = RPT / CH;            // chunk c covers matrix rows [TPC*CH*c, TPC*CH*(c+1))
  constexpr int NWARP = NT / 32;
  // register slot idx = 2*i2 + e holds matrix row 2*TPC*i2 + 2*part + e (pairs of rows -> 16-byte LDS of v)
  const int tid = threadIdx.x;
  const int col0 = tid / TPC, part = tid % TPC, lane = tid & 31;
  double* Vall = s.P2;
  unsigned* keys = reinterpret_cast<unsigned*>(s.vn1);  // one key per warp
  // key = upper 24 bits of the SQUARED norm (sign, exponent, 12 mantissa bits) | (255 - column): an unsigned max
  // picks the largest norm to 2.4e-4 relative, ties towards the lowest column index; 0 = closed / padded column.
  // 32-bit keys reduce with one REDUX per warp instead of a shuffle tree.
  auto norm_key = [](double nn, int c) -> unsigned {
    return ((unsigned)__double2hiint(nn) & ~0xFFu) | (unsigned)(0xFF - c);
  };
  auto warp_key_publish = [&](unsigned k) {
    k = __reduce_max_sync(0xffffffffu, k);
    if (lane == 0) keys[tid >> 5] = k;
  };
  double x[CPT][RPT];
  double nn_me[CPT];   // SQUARED norm of the not-yet-eliminated part of each owned column
  bool open[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int col = col0 + k * NCOL;
    open[k] = col < C;
    // closing record of a column: pivot row (N2 while open) and diagonal R entry
    if (part == 0 && col < N2) { s.perm[col] = N2; s.sig[col] = 0.0; }
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) {
      int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
      x[k][idx] = (open[k] && r < R) ? orig[(size_t)r * C + col] : 0.0;
    }
  }
  {
    unsigned kmax_key = 0u;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      double nn = 0.0;
#pragma unroll
      for (int idx = 0; idx < RPT; ++idx) nn += x[k][idx] * x[k][idx];
#pragma unroll
      for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
      nn_me[k] = nn;
      const unsigned kk = open[k] ? norm_key(nn, col0 + k * NCOL) : 0u;
      kmax_key = kk > kmax_key ? kk : kmax_key;
    }
    warp_key_publish(kmax_key);
  }
  __syncthreads();
  double thresh = 0.0;
  const int jmax = min(R, C);
  int K = 0, more = 0;
  for (int j = 0; j < jmax; ++j) {
    // ---- pivot: max over the per-warp keys (norm bits | inverted column index), redundantly per warp ----
    unsigned key = (lane < NWARP) ? keys[lane] : 0u;
    key = __reduce_max_sync(0xffffffffu, key);
    const double best = key ? __hiloint2double((int)(key & ~0xFFu), 0) : -1.0;  // rounded down by < 2.5e-4
    const int bi = 0xFF - (int)(key & 0xFFu);
    if (j == 0) thresh = best * rel_tol * rel_tol;  // keys carry squared norms
    if (!(best > thresh)) break;
    if (j >= kmax) { more = 1; break; }
    const int ko = (bi >= NCOL) ? bi / NCOL : 0;      // which of my columns could be the pivot
    const bool owner = (col0 + ko * NCOL == bi);
    // pivot row entry a = x[row j] of column bi: every lane picks its candidate (register slot jslot is
    // warp-uniform, one indexed jump) and the value is broadcast with a full-mask shuffle OUTSIDE the
    // owner branch -- a sub-warp mask inside divergent code costs a MATCH + WARPSYNC sequence
    const int jslot = ((j / (2 * TPC)) << 1) | (j & 1), jpart = (j >> 1) % TPC;
    double apiv;
    {
      double a[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) a[k] = 0.0;
#define MD_QR_SLOT(I)                                                                 \
  case I:                                                                             \
    if constexpr ((I) < RPT) {                                                        \
      _Pragma("unroll") for (int k = 0; k < CPT; ++k) a[k] = x[k][(I) < RPT ? (I) : 0]; \
    }                                                                                 \
    break;
      switch (jslot) {
        MD_QR_SLOT(0) MD_QR_SLOT(1) MD_QR_SLOT(2) MD_QR_SLOT(3) MD_QR_SLOT(4) MD_QR_SLOT(5) MD_QR_SLOT(6)
        MD_QR_SLOT(7) MD_QR_SLOT(8) MD_QR_SLOT(9) MD_QR_SLOT(10) MD_QR_SLOT(11) MD_QR_SLOT(12)
        MD_QR_SLOT(13) MD_QR_SLOT(14) MD_QR_SLOT(15)
        default: break;
      }
#undef MD_QR_SLOT
      double asel = a[0];
#pragma unroll
      for (int k = 1; k < CPT; ++k) asel = (k == ko) ? a[k] : asel;
      apiv = __shfl_sync(0xffffffffu, asel, (lane & ~(TPC - 1)) | jpart);
    }
    if (owner) {
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        if (k == ko) {
          double a = apiv;
          // The reflector is published UNSCALED: v = (0, .., 0, a - beta, x[j+1:]) with tau = -1/(beta (a - beta)),
          // i.e. H_j = I - tau v v^T with the 1/(a - beta)^2 of the textbook form folded into tau.  The column
          // goes out first (it does not depend on the square root / reciprocal chain); only v[j] and tau
          // follow.  The column's registers are left as they are: a closed column gets w = 0 in every later
          // update and is rebuilt from (registers, pivot row, beta) at the final write-out.
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            double2 v;
            v.x = (r0 >= j) ? x[k][2 * i2] : 0.0;
            v.y = (r0 + 1 >= j) ? x[k][2 * i2 + 1] : 0.0;
            *reinterpret_cast<double2*>(Vall + j * N2 + r0) = v;
          }
          const double beta = -copysign(sqrt(nn_me[k]), a);
          const double d = a - beta;
          const double tau = -1.0 / (beta * d);
          if (part == jpart) Vall[j * N2 + j] = d;
          if (part == 0) { s.perm[bi] = j; s.sig[bi] = beta; }
          open[k] = false;
          if (part == 0) s.tau[j] = tau;
        }
      }
    }
    __syncthreads();
    bool any_open = false;
#pragma unroll
    for (int k = 0; k < CPT; ++k) any_open |= open[k];
    if (__any_sync(0xffffffffu, any_open)) {
      // ---- apply H_j to the open columns (closed ones get w = 0 and stay as they are).  Row chunks entirely above
      //      the pivot row or below the matrix are skipped (warp-uniform).  v is loaded once per chunk and
      //      used for all CPT columns of the thread group. ----
      const double tau = s.tau[j];
      const double2* vj = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
      double w[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) w[k] = 0.0;
#pragma unroll
      for (int c = 0; c < NCHK; ++c) {
        if (TPC * CH * (c + 1) > j && TPC * CH * c < R) {
          double w0[CPT], w1[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) { w0[k] = 0.0; w1[k] = 0.0; }
#pragma unroll
          for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
            const double2 v = vj[TPC * i2];
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              w0[k] += v.x * x[k][2 * i2];
              w1[k] += v.y * x[k][2 * i2 + 1];
            }
          }
#pragma unroll
          for (int k = 0; k < CPT; ++k) w[k] += w0[k] + w1[k];
        }
      }
#pragma unroll
      for (int off = 1; off < TPC; off <<= 1) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) w[k] += __shfl_xor_sync(0xffffffffu, w[k], off);
      }
#pragma unroll
      for (int k = 0; k < CPT; ++k) w[k] = open[k] ? w[k] * tau : 0.0;  // closed columns are final
      double nn[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) nn[k] = 0.0;
#pragma unroll
      for (int c = 0; c < NCHK; ++c) {
        if (TPC * CH * (c + 1) > j && TPC * CH * c < R) {
          double s0[CPT], s1[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) { s0[k] = 0.0; s1[k] = 0.0; }
          if (TPC * CH * c > j) {  // every row of the chunk is below the pivot row
#pragma unroll
            for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
              const double2 v = vj[TPC * i2];
#pragma unroll
              for (int k = 0; k < CPT; ++k) {
                const double xa = x[k][2 * i2] - w[k] * v.x, xb = x[k][2 * i2 + 1] - w[k] * v.y;
                x[k][2 * i2] = xa; x[k][2 * i2 + 1] = xb;
                s0[k] += xa * xa; s1[k] += xb * xb;
              }
            }
          } else {
#pragma unroll
            for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
              const int r0 = 2 * TPC * i2 + 2 * part;
              const double2 v = vj[TPC * i2];
#pragma unroll
              for (int k = 0; k < CPT; ++k) {
                const double xa = x[k][2 * i2] - w[k] * v.x, xb = x[k][2 * i2 + 1] - w[k] * v.y;
                x[k][2 * i2] = xa; x[k][2 * i2 + 1] = xb;
                if (r0 > j) s0[k] += xa * xa;
                if (r0 + 1 > j) s1[k] += xb * xb;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < CPT; ++k) nn[k] += s0[k] + s1[k];
        }
      }
#pragma unroll
      for (int off = 1; off < TPC; off <<= 1) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) nn[k] += __shfl_xor_sync(0xffffffffu, nn[k], off);
      }
#pragma unroll
      for (int k = 0; k < CPT; ++k) nn_me[k] = nn[k];
    }
    // per-warp maximum of the open columns' keys; written after the first barrier of the iteration, so no
    // warp's pivot search can race it
    {
      unsigned kmax_key = 0u;
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const unsigned kk = open[k] ? norm_key(nn_me[k], col0 + k * NCOL) : 0u;
        kmax_key = kk > kmax_key ? kk : kmax_key;
      }
      warp_key_publish(kmax_key);
    }
    __syncthreads();
    K = j + 1;
  }
  if (!more) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int col = col0 + k * NCOL;
      const int jc = (col < C) ? s.perm[col] : N2;
      const double bet = (col < C) ? s.sig[col] : 0.0;
#pragma unroll
      for (int idx = 0; idx < RPT; ++idx) {
        int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
        // closed column: R entries above its pivot row, beta on it, zero below; open column (negligible
        // remainder): its rows < K
        if (col < C && r < K) s.X[r * LD + col] = (r < jc) ? x[k][idx] : ((r == jc) ? bet : 0.0);
      }
    }
  }
  if (tid == 0) {
    s.iscr[0] = K; s.iscr[1] = more;
    double f = 2.0 * R * C;
    for (int j = 0; j < K; ++j) f += 4.0 * (R - j) * (C - j - 1) + 3.0 * (R - j);
    s.cnt.flops_qr += f;
    s.cnt.n_qr_reflectors += K;
  }
  __syncthreads();
}

// Q = H_0 ... H_{K-1} [I_K; 0] in registers (8 threads per column, rows interleaved mod 8), written to
// s.P2 as a flat R x K matrix (ld K) after every thread has finished reading the reflectors from it.
template <int CHI>
__device__ __noinline__ void form_q_reg(Smem<CHI>& s, int R, int K) {
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int RPT = N2 / 8;               // register rows per thread
  constexpr int CH = RPT < 8 ? RPT : 8;
  constexpr int NCHK = RPT / CH;            // chunk c covers matrix rows [8*CH*c, 8*CH*(c+1))
  // register slot idx = 2*i2 + e holds matrix row 16*i2 + 2*part + e
  const int tid = threadIdx.x;
  const int col = tid >> 3, part = tid & 7;
  const int warp_col0 = (tid >> 5) * 4;     // first of the 4 columns this warp owns
  const double* Vall = s.P2;
  double q[RPT];
#pragma unroll
  for (int idx = 0; idx < RPT; ++idx) q[idx] = (16 * (idx >> 1) + 2 * part + (idx & 1) == col) ? 1.0 : 0.0;
  // c_j = v_j^T v_{j-1} for every consecutive pair (group g computes c_{g+1}): lets two reflectors be applied
  // per shuffle round below, H_{j-1} H_j q = q - wa v_j - wb v_{j-1} with wa = tau_j v_j^T q and
  // wb = tau_{j-1} (v_{j-1}^T q - c_j wa)
  {
    const int j = col + 1;
    double c = 0.0;
    if (j < K) {
      const double2* va = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
      const double2* vb = reinterpret_cast<const double2*>(Vall + (j - 1) * N2 + 2 * part);
#pragma unroll
      for (int i2 = 0; i2 < RPT / 2; ++i2) {
        const double2 x = va[8 * i2], y = vb[8 * i2];
        c += x.x * y.x + x.y * y.y;
      }
    }
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    if (part == 0 && j < N2) s.wv[j] = c;
  }
  __syncthreads();
  // H_j leaves e_c untouched for j > c, so a warp starts at its largest column index
  const int jstart = (warp_col0 < K) ? min(K - 1, warp_col0 + 3) : -1;
  int j = jstart;
  if (j >= 0 && ((j + 1) & 1)) {  // odd number of reflectors: apply the first one alone
    const double tau = s.tau[j];
    const double2* vj = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
    double w = 0.0;
#pragma unroll
    for (int c = 0; c < NCHK; ++c) {
      if (8 * CH * (c + 1) > j && 8 * CH * c < R) {
        double w0 = 0.0, w1 = 0.0;
#pragma unroll
        for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
          const double2 v = vj[8 * i2];
          w0 += v.x * q[2 * i2];
          w1 += v.y * q[2 * i2 + 1];
        }
        w += w0 + w1;
      }
    }
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 4);
    w *= tau;
#pragma unroll
    for (int c = 0; c < NCHK; ++c) {
      if (8 * CH * (c + 1) > j && 8 * CH * c < R) {
#pragma unroll
        for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
          const double2 v = vj[8 * i2];
          q[2 * i2] -= w * v.x;
          q[2 * i2 + 1] -= w * v.y;
        }
      }
    }
    --j;
  }
  for (; j >= 1; j -= 2) {  // pairs (j, j-1), H_j applied first
    const double tau_a = s.tau[j], tau_b = s.tau[j - 1], cab = s.wv[j];
    const double2* va = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
    const double2* vb = reinterpret_cast<const double2*>(Vall + (j - 1) * N2 + 2 * part);
    double wa = 0.0, wb = 0.0;
#pragma unroll
    for (int c = 0; c < NCHK; ++c) {
      if (8 * CH * (c + 1) > j - 1 && 8 * CH * c < R) {
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
          const double2 x = va[8 * i2], y = vb[8 * i2];
          a0 += x.x * q[2 * i2];
          a1 += x.y * q[2 * i2 + 1];
          b0 += y.x * q[2 * i2];
          b1 += y.y * q[2 * i2 + 1];
        }
        wa += a0 + a1;
        wb += b0 + b1;
      }
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
      const double ta = __shfl_xor_sync(0xffffffffu, wa, off), tb = __shfl_xor_sync(0xffffffffu, wb, off);
      wa += ta; wb += tb;
    }
    wa *= tau_a;
    wb = tau_b * (wb - cab * wa);
#pragma unroll
    for (int c = 0; c < NCHK; ++c) {
      if (8 * CH * (c + 1) > j - 1 && 8 * CH * c < R) {
#pragma unroll
        for (int i2 = c * CH / 2; i2 < (c + 1) * CH / 2; ++i2) {
          const double2 x = va[8 * i2], y = vb[8 * i2];
          q[2 * i2] -= wa * x.x + wb * y.x;
          q[2 * i2 + 1] -= wa * x.y + wb * y.y;
        }
      }
    }
  }
  __syncthreads();
  if (col < K) {
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) {
      int r = 16 * (idx >> 1) + 2 * part + (idx & 1);
      if (r < R) s.P2[r * K + (col ^ q_swizzle(r, K))] = q[idx];
    }
  }
  if (tid == 0) {
    double f = 0.0;
    for (int j = 0; j < K; ++j) f += 4.0 * (R - j) * (K - 1 - j) + 2.0 * (R - j);
    s.cnt.flops_qr += f;
  }
  __syncthreads();
}

// truncation_rule with block-wide reductions instead of the leader's serial loops (same outputs; the sums are
// accumulated in a different order, i.e. equal to rounding).  Needs blockDim.x >= C.
template <int CHI>
__device__ __forceinline__ void truncation_rule_dev(Smem<CHI>& s, int C, int chi_lim, double cut) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double sig_floor = sqrt(s.dscr[6]);
  double v = 0.0;
  int pred = 0;
  if (tid < C) { v = s.sig[s.order[tid]]; pred = (v > cut && v > sig_floor); }
  int k = __syncthreads_count(pred);
  if (k > chi_lim) {   // chi_max cuts into the values the cut would have kept: a truncating step
    k = chi_lim;
    if (tid == 0) s.cnt.n_trunc += 1.0;
  }
  double err = (tid < C && tid >= k) ? v * v : 0.0;
  double kept = (tid < C && tid < k) ? v * v : 0.0;
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    err += __shfl_xor_sync(0xffffffffu, err, off);
    kept += __shfl_xor_sync(0xffffffffu, kept, off);
  }
  if (lane == 0) { s.rot[warp] = err; s.rot[16 + warp] = kept; }
  __syncthreads();
  if (tid == 0) {
    double e = 0.0, kp = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { e += s.rot[w]; kp += s.rot[16 + w]; }
    s.iscr[0] = k; s.dscr[2] = e; s.dscr[3] = sqrt(kp);
  }
  __syncthreads();
}

// One-sided Jacobi, device version: 8 lanes per column pair, both columns of the pair in registers for
// the round (load, 3 dot products reduced by shuffles, rotation, store), one CTA barrier per round.
// Same round-robin order, rotation formula, thresholds and outputs as jacobi_columns<> above.
template <int CHI>
__device__ __noinline__ void jacobi_columns_dev(Smem<CHI>& s, int R, int C, int max_sweeps, int try_ortho = 0,
                                                double* spill = nullptr) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int RPT = N2 / 8;  // register rows per lane; slot idx holds matrix row 8*idx + sub
  double* X = s.X;
  const int tid = threadIdx.x;
  // The two 8-lane groups of a half-warp take pair slots 8 apart (bits 0 and 3 of the group index swapped):
  // their columns then sit 8 apart, so the two groups' 64-bit accesses fall on disjoint halves of the
  // shared-memory banks instead of colliding 2-way (the row stride LD is odd).
  constexpr int NGRP = ((CHI * 8 < 64) ? 64 : CHI * 8) / 8;
  const int grp = tid >> 3, sub = tid & 7;
  const int slot = (NGRP >= 16) ? ((grp & ~9) | ((grp & 1) << 3) | ((grp >> 3) & 1)) : grp;
  const int n = (C + 1) & ~1;
  const int npair = n / 2;
  const double tol = EPS * sqrt((double)R);
  const double tol2 = tol * tol;
  const int per = (R + NCH - 1) / NCH;
  // squared-norm floor below which a column counts as numerically zero (the Gram pass below provides the
  // norms itself, so the try_ortho path skips this extra pass over X)
  if (!(try_ortho && C >= 2)) {
  MD_PAR_FOR(t, NCH * C) {
    int ch = t / C, c = t - ch * C;
    int i0 = ch * per, i1 = min(R, i0 + per);
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) { double v = X[i * LD + c]; acc += v * v; }
    s.part[ch * N2 + c] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0;
    for (int c = 0; c < C; ++c) {
      double acc = 0.0;
      for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
      mx = fmax(mx, acc);
    }
    s.dscr[6] = mx * (4.0 * EPS) * (4.0 * EPS);
  }
  __syncthreads();
  }  // !try_ortho
  double floor2 = s.dscr[6];
  int sweeps = 0;
  int converged = (C < 2);
  int ortho_done = 0;
  // ---- Probe check (Freivalds), see below; a lambda because it is asked again after a pair round ----
  auto probe_check = [&]() -> bool {
    // ---- Probe check (Freivalds): for two fixed pseudo-random vectors z the identity X^T (X z) = D z with
    // D = diag(|x_c|^2) holds for every z exactly when the columns are mutually orthogonal; a deviation above
    // 1e-12 |x_c| |X z| in any component sends the call to the deterministic machinery below (Gram check /
    // sweeps).  Three passes over X (8 flops per entry) instead of the R C^2 of the full Gram matrix.  What a
    // pass can miss is non-orthogonality below 1e-12 relative, which leaves the factorisation X = (X/sigma) sigma
    // exact and only perturbs the gauge at that level. ----
    const int lane = tid & 31, warp = tid >> 5;
    auto zval = [](int c, unsigned salt) -> double {
      unsigned h = (unsigned)c * 2654435761u + salt;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      return 1.0 + (double)(h & 0xFFFFFu) * (1.0 / 1048576.0);   // in [1, 2)
    };
    double* z0 = s.vn1;       // probe vectors (C entries)
    double* z1 = s.sig;
    double* y0 = s.wv;        // X z0   (R entries)
    double* y1 = s.tau;       // X z1
    for (int c = tid; c < C; c += blockDim.x) { z0[c] = zval(c, 0x9E3779B9u); z1[c] = zval(c, 0x7F4A7C15u); }
    __syncthreads();
    // Thread t works on index t % N2 (a row in pass 1, a column in pass 2) and on the quarter t / N2 of the
    // other dimension; consecutive lanes touch consecutive rows / columns, which is conflict-free with the odd
    // leading dimension.  The four partial sums meet in shared memory (s.part viewed as [3][4][N2]).
    const int idx = tid % N2, qtr = tid / N2;      // blockDim.x == 4 * N2 for every capacity with CHI >= 8
    double* pr = s.part;
    // pass 1: y = X z
    {
      double a0s = 0.0, a1s = 0.0;
      if (idx < R) {
#pragma unroll 4
        for (int c = qtr; c < C; c += 4) { const double x = X[idx * LD + c]; a0s += x * z0[c]; a1s += x * z1[c]; }
      }
      pr[(0 * 4 + qtr) * N2 + idx] = a0s;
      pr[(1 * 4 + qtr) * N2 + idx] = a1s;
    }
    __syncthreads();
    double ny0 = 0.0, ny1 = 0.0;
    if (qtr == 0 && idx < R) {
      const double a0s = (pr[idx] + pr[N2 + idx]) + (pr[2 * N2 + idx] + pr[3 * N2 + idx]);
      const double a1s = (pr[4 * N2 + idx] + pr[5 * N2 + idx]) + (pr[6 * N2 + idx] + pr[7 * N2 + idx]);
      y0[idx] = a0s; y1[idx] = a1s; ny0 = a0s * a0s; ny1 = a1s * a1s;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      ny0 += __shfl_xor_sync(0xffffffffu, ny0, off);
      ny1 += __shfl_xor_sync(0xffffffffu, ny1, off);
    }
    if (lane == 0) { s.rot[warp] = ny0; s.rot[16 + warp] = ny1; }
    __syncthreads();
    ny0 = 0.0; ny1 = 0.0;
    {  // 16-byte loads of the per-warp partials (blockDim.x / 32 is even for every capacity)
      const double2* r2 = reinterpret_cast<const double2*>(s.rot);
      const int nw2 = ((int)(blockDim.x >> 5) + 1) >> 1;
#pragma unroll 4
      for (int w = 0; w < nw2; ++w) { const double2 a = r2[w], b = r2[8 + w]; ny0 += a.x + a.y; ny1 += b.x + b.y; }
    }
    // pass 2: w = X^T y and the squared column norms
    {
      double w0 = 0.0, w1 = 0.0, d = 0.0;
      if (idx < C) {
#pragma unroll 4
        for (int r = qtr; r < R; r += 4) { const double x = X[r * LD + idx]; w0 += x * y0[r]; w1 += x * y1[r]; d += x * x; }
      }
      pr[(0 * 4 + qtr) * N2 + idx] = w0;
      pr[(1 * 4 + qtr) * N2 + idx] = w1;
      pr[(2 * 4 + qtr) * N2 + idx] = d;
    }
    __syncthreads();
    int bad = 0;
    double mxd = 0.0;
    if (qtr == 0 && idx < C) {
      const double w0 = (pr[idx] + pr[N2 + idx]) + (pr[2 * N2 + idx] + pr[3 * N2 + idx]);
      const double w1 = (pr[4 * N2 + idx] + pr[5 * N2 + idx]) + (pr[6 * N2 + idx] + pr[7 * N2 + idx]);
      const double d = (pr[8 * N2 + idx] + pr[9 * N2 + idx]) + (pr[10 * N2 + idx] + pr[11 * N2 + idx]);
      s.vn2[idx] = d;
      mxd = d;
      const double e0 = w0 - d * z0[idx], e1 = w1 - d * z1[idx];
      if (e0 * e0 > 1e-24 * d * ny0 || e1 * e1 > 1e-24 * d * ny1) bad = 1;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) mxd = fmax(mxd, __shfl_xor_sync(0xffffffffu, mxd, off));
    if (lane == 0) s.rot[warp] = mxd;
    __syncthreads();
    {
      double mx = 0.0;
      const double2* r2 = reinterpret_cast<const double2*>(s.rot);
      const int nw2 = ((int)(blockDim.x >> 5) + 1) >> 1;
#pragma unroll 4
      for (int w = 0; w < nw2; ++w) { const double2 a = r2[w]; mx = fmax(mx, fmax(a.x, a.y)); }
      floor2 = mx * (4.0 * EPS) * (4.0 * EPS);
      if (tid == 0) s.dscr[6] = floor2;
    }
    // a numerically zero column (rounding residue of earlier rotations, below the floor the rotations and the
    // truncation rule ignore) has a random direction and would fail the test for ever: it does not count
    if (qtr == 0 && idx < C) {
      bad = bad && (s.vn2[idx] > floor2);
      s.flag[idx] = bad;   // columns whose component of X^T X z deviates: they take part in a non-orthogonal pair
    }
    const int any_bad = __syncthreads_or(bad);
    if (tid == 0) s.cnt.flops_jacobi += 10.0 * R * C;
    __syncthreads();
    return !any_bad;
    };
  int probe_failed = 0;
  if (!converged && try_ortho == 1) {
    converged = probe_check();
    ortho_done = converged;
    probe_failed = !converged;
  }
  // ---- Gram check (also the convergence test after every sweep: one register-tiled pass over X^T X and three
  //      barriers instead of a whole checking sweep of C-1 barrier-separated rounds) ----
  auto gram_check = [&](bool record) -> bool {
    // ---- Gram check: are the columns already mutually orthogonal (to the rotation threshold)?  In the decode
    // with bit-flip bias every bond stays in its Schmidt basis under the XOR / SWAP constraint tensors, so the
    // matrix handed to an SVD-mode split needs no rotation at all; one register-tiled pass over X^T X (warp w
    // owns columns 8w..8w+7, lane l pairs them with columns l, l+32, ..) replaces the round-robin sweep that
    // would only have confirmed it.  Any pair above the threshold sends the call to the sweeps below. ----
    constexpr int NBJ = (N2 + 31) / 32;     // column groups per lane
    const int warp = tid >> 5, lane = tid & 31;
    const int a0 = 8 * warp;
    double g[8][NBJ];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int v = 0; v < NBJ; ++v) g[u][v] = 0.0;
    // the Gram matrix is symmetric: a warp skips the 32-column groups that lie entirely left of its own
    // columns (those pairs belong to the warps owning the other column); the first group it needs is
    // warp-uniform, so each case below is a fully unrolled loop over compile-time column groups
    bool use[NBJ];
#pragma unroll
    for (int v = 0; v < NBJ; ++v) use[v] = (32 * v + 31 >= a0);
    if (a0 < C) {
#define MD_GRAM_ROWS(V0)                                                                     \
  _Pragma("unroll 2") for (int r = 0; r < R; ++r) {                                          \
    const double* row = X + r * LD;                                                          \
    double xa[8];                                                                            \
    _Pragma("unroll") for (int u = 0; u < 8; ++u) xa[u] = (a0 + u < C) ? row[a0 + u] : 0.0;  \
    _Pragma("unroll") for (int v = (V0); v < NBJ; ++v) {                                     \
      const double xb = (lane + 32 * v < C) ? row[lane + 32 * v] : 0.0;                      \
      _Pragma("unroll") for (int u = 0; u < 8; ++u) g[u][v] += xa[u] * xb;                   \
    }                                                                                        \
  }
      switch (a0 / 32) {
        case 0: MD_GRAM_ROWS(0) break;
        case 1: if constexpr (NBJ > 1) { MD_GRAM_ROWS(1) } break;
        case 2: if constexpr (NBJ > 2) { MD_GRAM_ROWS(2) } break;
        default: if constexpr (NBJ > 3) { MD_GRAM_ROWS(3) } break;
      }
#undef MD_GRAM_ROWS
      // publish the diagonal (squared column norms)
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < NBJ; ++v)
          if (a0 + u == lane + 32 * v && a0 + u < C) s.vn2[a0 + u] = g[u][v];
    }
    __syncthreads();
    {  // floor = (4 eps)^2 * largest squared column norm: block maximum of the diagonal
      double mx = 0.0;
      for (int c = tid; c < C; c += blockDim.x) mx = fmax(mx, s.vn2[c]);
#pragma unroll
      for (int off = 16; off; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if (lane == 0) s.rot[warp] = mx;
    }
    __syncthreads();
    {
      double mx = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, s.rot[w]);
      floor2 = mx * (4.0 * EPS) * (4.0 * EPS);
      if (tid == 0) s.dscr[6] = floor2;
    }
    int bad = 0;
    if (a0 < C) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < NBJ; ++v) {
          const int ca = a0 + u, cb = lane + 32 * v;
          if (use[v] && ca < C && cb < C && ca != cb) {
            const double saa = s.vn2[ca], sbb = s.vn2[cb], sab = g[u][v];
            if (sab != 0.0 && sab * sab > tol2 * saa * sbb && saa > floor2 && sbb > floor2) {
              bad = 1;
              // remember the offending pair; a column with two different partners breaks the matching -- and once it
              // is broken nothing more needs recording (a dense matrix would queue thousands of atomics here)
              if (record && *(volatile int*)&s.iscr[12] == 0) {
                const int o1 = atomicCAS(&s.perm[ca], -1, cb);
                const int o2 = atomicCAS(&s.perm[cb], -1, ca);
                if ((o1 != -1 && o1 != cb) || (o2 != -1 && o2 != ca)) s.iscr[12] = 1;
              }
            }
          }
        }
    }
    const bool gram_ok = !__syncthreads_or(bad);
    if (tid == 0) s.cnt.flops_jacobi += 1.0 * R * C * (C + 1);   // R (C^2 + C) / 2 multiply-adds
      return gram_ok;
  };
  int norms_ready = ortho_done;   // s.vn2 holds the squared column norms of the final X (Gram diagonal / the probe's d)
  int pair_round_done = 0;
  int gram_recorded = 0;
  int spilled = 0;
  int want_probe_pairs = 0;
  if (probe_failed) {
    // ---- Probe mode, failed probe: the columns the probe flagged are the only candidates for non-orthogonal pairs.
    // With few of them, the pairs AMONG them are tested exactly (8 lanes per pair), and if the offending ones form a
    // matching they are rotated in one round and the probe is asked again -- no Gram matrix, no sweep. ----
    if (tid < 32) {   // warp 0 compacts the flagged columns into s.order (ascending), 32 columns per ballot
      int nf = 0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + tid;
        const int f = (c < C) ? s.flag[c] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) s.order[nf + __popc(m & ((1u << tid) - 1u))] = c;
        nf += __popc(m);
      }
      if (tid == 0) { s.iscr[13] = nf; s.iscr[12] = 0; s.iscr[14] = 0; s.iscr[15] = 0; }
    }
    MD_PAR_FOR(c, C) s.perm[c] = -1;
    __syncthreads();
    const int nf = s.iscr[13];
    want_probe_pairs = (nf >= 1);   // any number of flagged columns: the dense-looking splits are perfect matchings too
  }
  // Dot products between two lists of columns (la[0..na) x lb[0..nb); a null list is the identity) on the FP64
  // tensor pipe: a warp task is one 8 x 8 tile, accumulated over the rows in steps of 4 with mma.m8n8k4.f64 (DMMA);
  // lane l feeds A[m = l/4][k = l%4] = X[r0 + k][la[m]] and B[k][n = l/4] = X[r0 + k][lb[n]] and receives the
  // entries (m = l/4, n = 2 (l%4) + {0, 1}).  `same` lists: only the tiles / entries above the diagonal.
  auto tile_dots = [&](const int* la, int na, const int* lb, int nb, bool same, auto&& visit) {
    const int lane = tid & 31, warp = tid >> 5, nwarps = (int)(blockDim.x >> 5);
    const int nta = (na + 7) >> 3, ntb = (nb + 7) >> 3;
    for (int t = warp; t < nta * ntb; t += nwarps) {
      const int ti = t / ntb, tj = t - ti * ntb;
      if (same && tj < ti) continue;   // warp-uniform
      const int ia = min(8 * ti + (lane >> 2), na - 1), ib = min(8 * tj + (lane >> 2), nb - 1);
      const int ca = la ? la[ia] : ia, cb = lb ? lb[ib] : ib;
      const double* pa = X + (lane & 3) * LD + ca;
      const double* pb = X + (lane & 3) * LD + cb;
      double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
      const int R8 = R & ~7;
      for (int r0 = 0; r0 < R8; r0 += 8) {
        const double a0 = pa[r0 * LD], b0 = pb[r0 * LD], a1 = pa[(r0 + 4) * LD], b1 = pb[(r0 + 4) * LD];
        md_dmma_8x8x4(acc0, a0, b0);
        md_dmma_8x8x4(acc1, a1, b1);
      }
      for (int r0 = R8; r0 < R; r0 += 4) {
        const bool ok = r0 + (lane & 3) < R;
        md_dmma_8x8x4(acc0, ok ? pa[r0 * LD] : 0.0, ok ? pb[r0 * LD] : 0.0);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int oa = 8 * ti + (lane >> 2), ob = 8 * tj + 2 * (lane & 3) + e;
        if (oa < na && ob < nb && (!same || ob > oa)) visit(ca, lb ? lb[ob] : ob, acc0[e] + acc1[e]);
      }
    }
  };
  // one 8-lane group: dot products of columns a, b (all lanes of the warp take part in the shuffles)
  auto pair_dots = [&](int a, int b, bool valid, double (&xa)[RPT], double (&xb)[RPT], double& saa, double& sbb,
                       double& sab) {
    saa = 0.0; sbb = 0.0; sab = 0.0;
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) {
      const int r = 8 * idx + sub;
      const bool ok = valid && r < R;
      xa[idx] = ok ? X[r * LD + a] : 0.0;
      xb[idx] = ok ? X[r * LD + b] : 0.0;
      saa += xa[idx] * xa[idx]; sbb += xb[idx] * xb[idx]; sab += xa[idx] * xb[idx];
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
      saa += __shfl_xor_sync(0xffffffffu, saa, off);
      sbb += __shfl_xor_sync(0xffffffffu, sbb, off);
      sab += __shfl_xor_sync(0xffffffffu, sab, off);
    }
  };
  // ---- Pair round.  Measured on depolarising-bias decodes (4 012 rotating splits): the columns are mutually
  // orthogonal except for DISJOINT pairs -- the first cyclic sweep never rotates a column twice (median 6 rotations,
  // at most C/2) and the second sweep only confirms.  So when the offending pairs the Gram check recorded form a
  // matching, all of them are rotated in ONE barrier-separated round (8 lanes per pair, fresh dot products) and the
  // Gram check is asked again; a cyclic sweep is C-1 rounds of ~1 500 cycles each.  Anything else falls through to
  // the sweeps below. ----
  // rotation (cos, sin) of the pairs this thread's group rotated (a group sees at most two pairs: C <= 2 NGRP)
  double keep_c[2] = {1.0, 1.0}, keep_s[2] = {0.0, 0.0};
  int keep_a[2] = {-1, -1};
  int pair_rounds = 0;
  static_assert(2 * NGRP >= N2, "a pair round visits every column in two passes of the groups");
  auto pair_round = [&](bool save_orig, bool write_norms) {
    int nrot_pairs = 0;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (it * NGRP >= C) continue;   // uniform
      const int a = grp + it * NGRP;
      const int b = (a < C) ? s.perm[a] : -1;
      const bool valid = (b > a);
      if (!__any_sync(0xffffffffu, valid)) continue;   // none of the warp's four groups holds a pair
      double xa[RPT], xb[RPT];
      double saa, sbb, sab;
      pair_dots(a, b, valid, xa, xb, saa, sbb, sab);
      if (valid && sab != 0.0 && sab * sab > tol2 * saa * sbb && saa > floor2 && sbb > floor2) {
        const double alpha = sbb - saa, beta = 2.0 * sab;
        const double t = copysign(beta, alpha * beta) / (fabs(alpha) + sqrt(alpha * alpha + beta * beta));
        const double c = rsqrt(1.0 + t * t);
        const double sn = c * t;
        double na = 0.0, nb = 0.0;
#pragma unroll
        for (int idx = 0; idx < RPT; ++idx) {
          const int r = 8 * idx + sub;
          if (r < R) {
            const double ya = c * xa[idx] - sn * xb[idx], yb = sn * xa[idx] + c * xb[idx];
            X[r * LD + a] = ya;
            X[r * LD + b] = yb;
            na += ya * ya; nb += yb * yb;
            if (save_orig) { spill[(size_t)r * C + a] = xa[idx]; spill[(size_t)r * C + b] = xb[idx]; }
          }
        }
        keep_c[it] = c; keep_s[it] = sn; keep_a[it] = a;
        if (write_norms) {   // fresh squared norms of the two new columns (the 8 lanes of the group are all here)
          const unsigned gm = 0xffu << ((tid & 31) & ~7);
#pragma unroll
          for (int off = 1; off < 8; off <<= 1) {
            na += __shfl_xor_sync(gm, na, off);
            nb += __shfl_xor_sync(gm, nb, off);
          }
          if (sub == 0) { s.vn2[a] = na; s.vn2[b] = nb; }
        }
        if (sub == 0) ++nrot_pairs;
      } else if (valid && sub == 0) {
        // below the threshold on the fresh dot product: the pair stays as it is and counts as untouched
        s.perm[a] = -1; s.perm[b] = -1;
      }
    }
    pair_round_done = 1;
    ++pair_rounds;
    if (nrot_pairs) atomicAdd(&s.iscr[15], nrot_pairs);   // (a double atomicAdd on shared memory is a CAS loop)
    __syncthreads();
    if (tid == 0) { s.cnt.flops_jacobi += 12.0 * R * s.iscr[15]; s.iscr[15] = 0; }
  };
  if (want_probe_pairs) {
    const int nf = s.iscr[13];
    int found = 0;
    // every pair of flagged columns, one 8 x 8 tile of dot products per warp task on the FP64 tensor pipe
    tile_dots(s.order, nf, s.order, nf, true, [&](int a, int b, double sab) {
      const double saa = s.vn2[a], sbb = s.vn2[b];   // squared norms from the probe's second pass
      if (sab != 0.0 && sab * sab > tol2 * saa * sbb && saa > floor2 && sbb > floor2) {
        found = 1;
        if (*(volatile int*)&s.iscr[12]) return;   // the matching is already broken: nothing more to record
        const int o1 = atomicCAS(&s.perm[a], -1, b);
        const int o2 = atomicCAS(&s.perm[b], -1, a);
        if ((o1 != -1 && o1 != b) || (o2 != -1 && o2 != a)) s.iscr[12] = 1;
        atomicAdd(&s.iscr[14], 2);   // columns that found a partner
      }
    });
    int any_found = __syncthreads_or(found);
    if (s.iscr[12] == 0 && s.iscr[14] < nf) {
      // Flagged columns without a partner among the flagged: the probe's criterion for a column is relative to that
      // column's own norm, so the larger column of a very unequal pair can pass while the smaller one is flagged.
      // Their partners are looked for among the unflagged columns (exact dot products again).
      if (tid < 32) {   // warp 0: the unmatched flagged columns, compacted in place at the head of s.order
        int nu = 0;
        for (int i0 = 0; i0 < nf; i0 += 32) {
          const int i = i0 + tid;
          const int c = (i < nf) ? s.order[i] : 0;
          const int f = (i < nf) && (s.perm[c] < 0);
          const unsigned m = __ballot_sync(0xffffffffu, f);
          __syncwarp();
          if (f) s.order[nu + __popc(m & ((1u << tid) - 1u))] = c;
          nu += __popc(m);
          __syncwarp();
        }
        if (tid == 0) s.iscr[13] = nu;
      }
      __syncthreads();
      const int nu = s.iscr[13];
      int found2 = 0;
      tile_dots(s.order, nu, nullptr, C, false, [&](int a, int b, double sab) {
        if (s.flag[b]) return;   // flagged x flagged pairs were all tested above
        const double saa = s.vn2[a], sbb = s.vn2[b];
        if (sab != 0.0 && sab * sab > tol2 * saa * sbb && saa > floor2 && sbb > floor2) {
          found2 = 1;
          if (*(volatile int*)&s.iscr[12]) return;
          const int o1 = atomicCAS(&s.perm[a], -1, b);
          const int o2 = atomicCAS(&s.perm[b], -1, a);
          if ((o1 != -1 && o1 != b) || (o2 != -1 && o2 != a)) s.iscr[12] = 1;
          atomicAdd(&s.iscr[14], 1);   // one more flagged column with a partner
        }
      });
      any_found |= __syncthreads_or(found2);
    }
    if (any_found && s.iscr[12] == 0) {
      // the rotations change X: the round saves the original of the columns it rotates (the coefficient product
      // reads it if sweeps follow); when the round settles the split the coefficients come from the rotations
      // themselves and nothing else is copied
      //
      // Perfect matching (every flagged column has exactly one partner): the probe's verdict on the
      // other columns stands -- an unflagged column is orthogonal to ALL columns within the probe's tolerance, a
      // property rotations among other columns preserve -- and the rotated pairs are orthogonal by construction,
      // so the round settles the split without asking the probe again; the round writes the new squared norms.
      const bool perfect = (s.iscr[14] == nf);
      pair_round(spill != nullptr, perfect);
      converged = perfect ? 1 : probe_check();
      norms_ready = converged;
      if (!converged && spill != nullptr) {   // complete the copy: untouched columns still hold the original
        MD_PAR_FOR(t, R * C) { int r = t / C, c = t - r * C; if (s.perm[c] < 0) spill[t] = X[r * LD + c]; }
        spilled = 1;
      }
      __syncthreads();
    }
  }
  if (!converged && try_ortho) {
    if (pair_rounds) pair_rounds = 2;      // s.perm is about to be reused: the earlier round no longer counts as the only one
    MD_PAR_FOR(c, C) s.perm[c] = -1;       // partner of a column in an offending pair (s.perm is free in the SVD path)
    if (tid == 0) { s.iscr[12] = 0; s.iscr[15] = 0; }   // [12] = 1: the offending pairs do not form a matching; [15]: rotations
    __syncthreads();
    converged = gram_check(true);
    gram_recorded = 1;
    if (!pair_round_done) ortho_done = converged;
    norms_ready = converged;
  }
  if (!converged && !spilled && spill != nullptr) {
    // the rotations are about to change X: keep the original for the coefficient product (flat, ld C)
    MD_PAR_FOR(t, R * C) { int r = t / C, c = t - r * C; spill[t] = X[r * LD + c]; }
    __syncthreads();
    spilled = 1;
  }
  if (!converged && try_ortho && gram_recorded && s.iscr[12] == 0) {
    pair_round(false, false);
    converged = gram_check(false);
    norms_ready = converged;
    __syncthreads();
  }
  while (!converged && sweeps < max_sweeps) {
    if (tid == 0) s.iscr[6] = 0;
    __syncthreads();
    for (int rnd = 0; rnd < n - 1; ++rnd) {
      // every lane of a warp with a valid pair runs the load / dot / shuffle sequence (invalid slots contribute
      // zeros), so the three reductions are plain full-mask shuffles; only the rotation itself is conditional
      int a = 0, b = 0;
      bool valid = false;
      if (slot < npair) {
        if (slot == 0) { a = n - 1; b = rnd; }
        else { a = (rnd + slot) % (n - 1); b = (rnd - slot + (n - 1)) % (n - 1); }
        valid = (a < C && b < C);  // uniform over the 8 lanes of the slot
      }
      // A round is issue-bound (~350 instructions per warp): warps none of whose four pair slots is valid (small C)
      // go straight to the barrier.  (Stopping the unrolled row-slot loops at the matrix height with uniform breaks
      // was tried and lost 2x: the loads no longer issue as one batch.)
      if (__any_sync(0xffffffffu, valid)) {
        double xa[RPT], xb[RPT];
        double saa = 0.0, sbb = 0.0, sab = 0.0;
#pragma unroll
        for (int idx = 0; idx < RPT; ++idx) {
          const int r = 8 * idx + sub;
          const bool ok = valid && r < R;
          xa[idx] = ok ? X[r * LD + a] : 0.0;
          xb[idx] = ok ? X[r * LD + b] : 0.0;
          saa += xa[idx] * xa[idx]; sbb += xb[idx] * xb[idx]; sab += xa[idx] * xb[idx];
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
          saa += __shfl_xor_sync(0xffffffffu, saa, off);
          sbb += __shfl_xor_sync(0xffffffffu, sbb, off);
          sab += __shfl_xor_sync(0xffffffffu, sab, off);
        }
        if (valid && sab != 0.0 && sab * sab > tol2 * saa * sbb && saa > floor2 && sbb > floor2) {
          // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = alpha / beta, written with one division:
          // t = sign(alpha) beta / (|alpha| + sqrt(alpha^2 + beta^2))   (the round's latency is a chain of these)
          const double alpha = sbb - saa, beta = 2.0 * sab;
          const double t = copysign(beta, alpha * beta) / (fabs(alpha) + sqrt(alpha * alpha + beta * beta));
          const double c = rsqrt(1.0 + t * t);
          const double sn = c * t;
#pragma unroll
          for (int idx = 0; idx < RPT; ++idx) {
            const int r = 8 * idx + sub;
            if (r < R) {
              X[r * LD + a] = c * xa[idx] - sn * xb[idx];
              X[r * LD + b] = sn * xa[idx] + c * xb[idx];
            }
          }
          if (sub == 0) atomicAdd(&s.iscr[6], 1);
        }
      }
      __syncthreads();
    }
    ++sweeps;
    const int nrot = s.iscr[6];
    converged = (nrot == 0);
    if (tid == 0) s.cnt.flops_jacobi += 6.0 * R * npair * (n - 1) + 6.0 * R * nrot;
    __syncthreads();
    // did this sweep finish the job?  (same pair threshold as the rotations.)  Only asked when the sweep rotated
    // less than a quarter of its pairs: a sweep that still rotates most pairs is far from done and the check would be
    // wasted (random dense matrices: ~10 sweeps, only the last ones qualify)
    if (!converged && 4 * nrot <= npair * (n - 1)) {
      converged = gram_check(false);
      norms_ready = converged;
      __syncthreads();
    }
  }
  // One pair round and nothing else: X' = X J with J the identity plus 2x2 rotation blocks, so the coefficient
  // matrix sigma J^T has one or two entries per row.  s.vn1[c] = J[c, c], s.wv[c] = J[partner, c] (s.perm[c] = the
  // partner, -1 for untouched columns); step() reads them instead of forming U^T X (layout 4).
  const int pairs_only = (converged && sweeps == 0 && pair_rounds == 1);
  if (pairs_only) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      if (keep_a[it] >= 0 && sub == 0) {
        const int a = keep_a[it], b = s.perm[a];
        s.vn1[a] = keep_c[it]; s.wv[a] = -keep_s[it];
        s.vn1[b] = keep_c[it]; s.wv[b] = keep_s[it];
      }
    }
  }
  // singular values = column norms; order by value (descending), ties by column index
  if (norms_ready) {
    MD_PAR_FOR(c, C) s.sig[c] = sqrt(s.vn2[c]);   // the Gram diagonal
  } else {
    MD_PAR_FOR(t, NCH * C) {
      int ch = t / C, c = t - ch * C;
      int i0 = ch * per, i1 = min(R, i0 + per);
      double acc = 0.0;
      for (int i = i0; i < i1; ++i) { double v = X[i * LD + c]; acc += v * v; }
      s.part[ch * N2 + c] = acc;
    }
    __syncthreads();
    MD_PAR_FOR(c, C) {
      double acc = 0.0;
      for (int ch = 0; ch < NCH; ++ch) acc += s.part[ch * N2 + c];
      s.sig[c] = sqrt(acc);
    }
  }
  __syncthreads();
  // rank of every column in the canonical order (sigma descending, ties by first non-zero row, then column index):
  // four adjacent lanes per column, each counting a quarter of the others (blockDim.x == 4 * N2)
  first_rows<CHI>(s, R, C, s.flag);
  for (int t4 = tid; t4 < 4 * N2; t4 += blockDim.x) {
    const int c = t4 >> 2, q = t4 & 3;
    int pos = 0;
    if (c < C) {
      const double v = s.sig[c];
      const int fc = s.flag[c];
      for (int o = q; o < C; o += 4) pos += canon_before(s.sig[o], s.flag[o], o, v, fc, c) ? 1 : 0;
    }
    pos += __shfl_xor_sync(0xffffffffu, pos, 1);
    pos += __shfl_xor_sync(0xffffffffu, pos, 2);
    if (c < C && q == 0) s.order[pos] = c;
  }
  if (tid == 0) {
    s.iscr[4] = sweeps + pair_round_done; s.iscr[5] = converged ? 0 : 1;
    // 1: columns were orthogonal on entry, X is untouched; 2: one round of disjoint pair rotations settled it
    s.iscr[9] = (sweeps == 0 && !pair_round_done && C >= 2) ? 1 : (pairs_only ? 2 : 0);
    s.cnt.n_jacobi_sweeps += sweeps + pair_round_done;
    s.cnt.flops_jacobi += 2.0 * R * C;
  }
  __syncthreads();
}
#endif  // !MDOPT_EMU

}  // namespace mdopt
