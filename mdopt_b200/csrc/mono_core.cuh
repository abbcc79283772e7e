// Monomial ("trellis") MPS engine: one WARP decodes one syndrome, every site tensor in sparse form.
//
// Same path as mps_core.cuh (mdopt/contractor/contractor.py:177-378 mps_mpo_contract, mdopt/mps/canonical.py:345-420
// move_orth_centre, mdopt/optimiser/utils.py:284-345, canonical.py:906-991 marginal + :241-272 dense,
// examples/decoding/decoding.py:1362-1379 read-out), same program, same formulation (factor the
// carried matrix, one-site centre moves, the reference's truncation rule k = min(chi_max, #(s > cut))) -- specialised to
// the structure a decode with product-state inputs and 0/1 constraint tensors has and *verifies* at every step:
//
//   every site tensor A[a, p, b] is p-MONOMIAL: for fixed p, each row a has at most one non-zero b and each column b
//   at most one non-zero a (a partial permutation with weights).
//
// Reason: the bond basis never leaves the trellis states of the parity constraints.  For such tensors the carried
// matrix X (rows (k, p), columns (w', b)) has ONE non-zero per row, so its columns have disjoint supports: they are
// exactly orthogonal, the SVD is X = (X / sigma) diag(sigma) with sigma the column norms, no rotation is ever needed,
// and U is again p-monomial.  A site tensor is 2*chi (value, target) pairs instead of 2*chi^2 doubles, a two-site
// step is O(chi) work plus one sort of <= 2*chi column norms, and a whole step fits one warp.  The structure is
// checked, not assumed: a row that would need two different targets, or two rows of equal p meeting in one column,
// ends the syndrome with MDOPT_ST_NOT_MONOMIAL and the caller re-runs it on the dense engine (decode_kernel).
//
// Ordering of the new bond (the tie-break that makes truncating decodes well-posed; oracle/structured_svd.py states the
// same rule for the reference's SVD hook): sigma descending, values within 1e-10 relative (chained) form one cluster,
// inside a cluster the column whose first supporting row comes first goes first.
//
// The source compiles for the host with -DMDOPT_EMU (lane loops become plain loops) so the CPU test tier runs the
// same index logic; the product is the CUDA build only.
#pragma once
#include "block_ops.cuh"

#ifdef MDOPT_EMU
#include <algorithm>
#include <string.h>
#define MW_FOR(i, n) for (int i = 0; i < (n); ++i)
#define MW_SYNC() ((void)0)
#define MW_ANY(x) (x)
#define MW_LANE0 if (true)
#define MW_PHASE inline
#else
// Engine phases are NOT inlined: inside a function the compiler knows the warp entered converged, so the shuffles
// of the reductions are plain SHFLs; inlined into the program interpreter's data-dependent control flow every one
// of them carried a WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair.
#define MW_PHASE __device__ __noinline__
// lane loops are NOT unrolled: trip counts are <= 2*chi/32, and the unrolled copies (4x + remainder per loop) made the
// hot code of a step larger than the 32 KB L1.5 instruction cache
#ifdef MW_UNROLL_LANE_LOOPS
#define MW_FOR(i, n) for (int i = (int)(threadIdx.x & 31u); i < (n); i += 32)
#else
#define MW_FOR(i, n) _Pragma("unroll 1") for (int i = (int)(threadIdx.x & 31u); i < (n); i += 32)
#endif
#define MW_SYNC() __syncwarp()
#define MW_ANY(x) __any_sync(0xffffffffu, (x))
#define MW_LANE0 if ((threadIdx.x & 31u) == 0)
#endif

namespace mdopt {

constexpr double MONO_TIE_REL2 = 2e-10;   // cluster width on sigma^2 (= 1e-10 relative on sigma)
constexpr int MONO_ROW_BITS = 10;         // low key bits that carry the first supporting row (2*chi <= 1024)
constexpr unsigned long long MONO_ROW_MASK = (1ull << MONO_ROW_BITS) - 1ull;

// One site tensor in HBM: rows r = a*2 + p -> (target b or -1, value).  Only the first 2*dim_left rows are live.
template <int CHI>
struct MonoSite {
  double val[2 * CHI];
  int16_t tgt[2 * CHI];
};

// One MPO site tensor (vL, vR, pU, pD), vL, vR <= 2, as the engine reads it: its non-zeros per (incoming bond q,
// outgoing physical pd), idx = q*2 + pd -- at most two (w', pU, value) entries (one for XOR / SWAP / IDENTITY, two
// for COPY_LEFT); en = 3 marks a tensor the monomial form cannot take.  Built once per launch (mono_build_tables).
struct MonoTensor {
  double ew[4][2];
  int en[4];
  int ewo[4][2], epu[4][2];
  int vl, vr, iso, pad;
};

MD_HD void mono_build_tensor(const double* w16, const int* dim4, MonoTensor* out) {
  const int vl = dim4[0], vr = dim4[1];
  out->vl = vl; out->vr = vr; out->iso = dim4[2]; out->pad = 0;
  for (int idx = 0; idx < 4; ++idx) {
    const int q = idx >> 1, pd = idx & 1;
    int n = 0;
    for (int e = 0; e < 2; ++e) { out->ew[idx][e] = 0.0; out->ewo[idx][e] = 0; out->epu[idx][e] = 0; }
    if (q < vl && vl <= 2 && vr <= 2 && dim4[3] == 0) {
      for (int wo = 0; wo < vr; ++wo)
        for (int pu = 0; pu < 2; ++pu) {
          const double w = w16[((q * vr + wo) * 2 + pu) * 2 + pd];
          if (w == 0.0) continue;
          if (n < 2) { out->ew[idx][n] = w; out->ewo[idx][n] = wo; out->epu[idx][n] = pu; }
          ++n;
        }
    } else if (q < vl) {
      n = 3;   // two-site gate entries (dim4[3] != 0) or oversized bonds: not a monomial MPO tensor
    }
    out->en[idx] = n > 2 ? 3 : n;
  }
}
// slot 0 = the identity MPO tensor of a centre move, slot 1 + t = tensor t of the program's table
MD_HD void mono_build_identity(MonoTensor* out) {
  const double w[16] = {1.0, 0.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const int dim[4] = {1, 1, 1, 0};
  mono_build_tensor(w, dim, out);
}

struct MonoCounters {
  double bytes_hbm, n_steps, n_trunc, n_fallback;
};

// Warp-uniform engine state.  It lives in shared memory (like the dense engine's EngineState) so that the phases can
// be separate functions without a per-thread stack object.
template <int CHI>
struct MonoState {
  MonoSite<CHI>* sites;     // [n_sites] this warp's MPS (global)
  int* bd;                  // [n_sites + 1] bond dimensions (global)
  const MonoTensor* tabs;   // [1 + n_tensors]
  double* trace;
  double* env;              // [(n_head + 1) * CHI] Dephasing-DMRG environments (global), or nullptr
  DecodeParams P;
  MonoCounters cnt;
  int n_traced, status, centre, K, Wq, Mr, mirrored;
};

// Per-warp shared memory.
template <int CHI>
struct MonoSmem {
  static constexpr int N2 = 2 * CHI;
  // carried matrix, row-singleton form: row r = (k, p') has its only entry cv[r] in column cc[r] (-1: empty row)
  double cv[N2];
  // per column: squared entry of the row with p = 0 / p = 1 that hits it; acc[1] is reused for the 64-bit sort keys,
  // acc[0] for sigma^2 and then 1 / sigma of the kept columns
  double acc[2][N2];
  double sig[N2];        // sigma per column
  double sv[N2];         // staged next site, frame view: rows (m, p) -> (st, sv)
  int16_t cc[N2];
  int16_t own[2][N2];    // k of the row with p = 0 / 1 that hits the column, -1 none; then the cluster flags
  int16_t rank[N2];      // new bond index of a column, -1 dropped
  int16_t order[N2];     // column at sorted position
  int16_t st[N2];
  // the outgoing record of a mirrored-frame store is assembled in acc[1] (values) / own[1] (targets): both are dead
  // once the columns are ordered, and the 1.25 KB this saves per warp is what lets 28 warps share an SM
  MD_DEV double* uv_buf() { return acc[1]; }
  MD_DEV int16_t* ut_buf() { return own[1]; }
  MonoTensor tw;         // the MPO tensor of the current step (copied from the per-launch table)
  MonoState<CHI> state;
};

#ifndef MDOPT_EMU
__device__ __forceinline__ int mw_clz(unsigned w) { return __clz((int)w); }
__device__ __forceinline__ int mw_ctz(unsigned w) { return __ffs((int)w) - 1; }
#else
inline int mw_clz(unsigned w) { return __builtin_clz(w); }
inline int mw_ctz(unsigned w) { return __builtin_ctz(w); }
#endif
#ifndef MDOPT_EMU
static __device__ __noinline__ double mw_sum(double v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
static __device__ __noinline__ double mw_max(double v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ int mw_isum(int v) { return __reduce_add_sync(0xffffffffu, v); }
#else
inline double mw_sum(double v) { return v; }
inline double mw_max(double v) { return v; }
inline int mw_isum(int v) { return v; }
#endif

// -------------------------------------------------------------------------------------------------
// Sort n 64-bit keys (shared memory, n <= NMAX = 32*EPL) in DESCENDING order, in place.
// Device: bitonic network with EPL keys per lane in registers (element e = lane*EPL + j), written as LOOPS over the
// merge sizes and strides: the merge of size S first pairs e with its mirror e ^ (S-1), then with e ^ t for
// t = S/4 ... 1, so every comparator points the same way (the lower element keeps the larger key).  Strides below
// EPL stay inside a lane, the others are one 64-bit shuffle per key.  The code of one variant is ~2 KB: with 24
// independent warps per SM the instruction caches (6 KB L0 per scheduler, 32 KB L1.5) decide the speed of this
// kernel, and the fully unrolled network (16 KB per variant) made `no_instruction` the top stall.
// Not inlined: inside the function the warp is known to be converged, so every shuffle is a plain SHFL.
// Host: std::sort.  Keys are unique or zero.
// -------------------------------------------------------------------------------------------------
#ifndef MDOPT_EMU
template <int EPL>
__device__ __forceinline__ void mw_sort_inlane(unsigned long long (&k)[EPL], int tmax) {
  // comparators e ^ t for t = tmax, tmax/2, ..., 1 (all < EPL): pairs (j, j | t), the lower index keeps the maximum
#pragma unroll
  for (int t = EPL >> 1; t > 0; t >>= 1) {
    if (t <= tmax) {
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        if ((j & t) == 0) {
          const unsigned long long a = k[j], b = k[j | t];
          const bool sw = a < b;
          k[j] = sw ? b : a; k[j | t] = sw ? a : b;
        }
      }
    }
  }
}

template <int EPL>
__device__ __noinline__ void mw_sort_desc_impl(unsigned long long* keys, int n) {
  const int lane = threadIdx.x & 31;
  unsigned long long k[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) { const int e = lane * EPL + j; k[j] = (e < n) ? keys[e] : 0ull; }
  // merges that stay inside a lane (S <= EPL): mirror stage j ^ (S-1), then strides S/4 ... 1
#pragma unroll
  for (int S = 2; S <= EPL; S <<= 1) {
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int o = j ^ (S - 1);
      if (j < o) {
        const unsigned long long a = k[j], b = k[o];
        const bool sw = a < b;
        k[j] = sw ? b : a; k[o] = sw ? a : b;
      }
    }
    mw_sort_inlane<EPL>(k, S >> 2);
  }
  // merges across lanes: S = 2*EPL ... 32*EPL
#pragma unroll 1
  for (int sl = 2; sl <= 32; sl <<= 1) {       // sl = S / EPL
    {  // mirror stage: partner lane = lane ^ (sl - 1), partner key index = EPL-1-j
      const bool keep_max = (lane & (sl >> 1)) == 0;
      unsigned long long o[EPL];
#pragma unroll
      for (int j = 0; j < EPL; ++j) o[j] = __shfl_xor_sync(0xffffffffu, k[EPL - 1 - j], sl - 1);
#pragma unroll
      for (int j = 0; j < EPL; ++j) { const bool gt = k[j] > o[j]; k[j] = (gt == keep_max) ? k[j] : o[j]; }
    }
#pragma unroll 1
    for (int lm = sl >> 2; lm > 0; lm >>= 1) {   // strides t = lm * EPL >= EPL
      const bool keep_max = (lane & lm) == 0;
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[j], lm);
        const bool gt = k[j] > o;
        k[j] = (gt == keep_max) ? k[j] : o;
      }
    }
    mw_sort_inlane<EPL>(k, EPL >> 1);
  }
#pragma unroll
  for (int j = 0; j < EPL; ++j) { const int e = lane * EPL + j; if (e < n) keys[e] = k[j]; }
  __syncwarp();
}
template <int NMAX>
__device__ __forceinline__ void mw_sort_desc(unsigned long long* keys, int n) {
  __syncwarp();
  if constexpr (NMAX >= 1024) { if (n > 512) { mw_sort_desc_impl<32>(keys, n); return; } }
  if constexpr (NMAX >= 512) { if (n > 256) { mw_sort_desc_impl<16>(keys, n); return; } }
  if constexpr (NMAX >= 256) { if (n > 128) { mw_sort_desc_impl<8>(keys, n); return; } }
  if constexpr (NMAX >= 128) { if (n > 64) { mw_sort_desc_impl<4>(keys, n); return; } }
  if constexpr (NMAX >= 64) { if (n > 32) { mw_sort_desc_impl<2>(keys, n); return; } }
  mw_sort_desc_impl<1>(keys, n);
}
#else
template <int NMAX>
inline void mw_sort_desc(unsigned long long* keys, int n) {
  std::sort(keys, keys + n, [](unsigned long long a, unsigned long long b) { return a > b; });
}
#endif

// -------------------------------------------------------------------------------------------------
// The per-warp engine: free functions over the warp's shared memory.
// -------------------------------------------------------------------------------------------------
template <int CHI>
struct Mono {
  static constexpr int N2 = 2 * CHI;
  using SM = MonoSmem<CHI>;

  // ---- frame helpers: logical site j of the current frame ----
  static MD_DEV int phys(const SM& s, int j) { return s.state.mirrored ? (s.state.P.n_sites - 1 - j) : j; }
  static MD_DEV int dim_left(const SM& s, int j) { const int p = phys(s, j); return s.state.mirrored ? s.state.bd[p + 1] : s.state.bd[p]; }
  static MD_DEV int dim_right(const SM& s, int j) { const int p = phys(s, j); return s.state.mirrored ? s.state.bd[p] : s.state.bd[p + 1]; }

  // ---- copy the step's tensor entry (160 bytes) from the per-launch table into shared memory ----
  static MD_DEV void load_tensor(SM& s, int tid) {
    static_assert(sizeof(MonoTensor) == 160, "MonoTensor layout");
    const int* src = reinterpret_cast<const int*>(s.state.tabs + (tid + 1));
    int* dst = reinterpret_cast<int*>(&s.tw);
    MW_FOR(i, (int)(sizeof(MonoTensor) / sizeof(int))) dst[i] = src[i];
    MW_SYNC();
  }

  // ---- ask for site j's record ahead of its use: one 128-byte line per lane into L1 (no registers, no staging
  //      buffer; load_site() a step later hits the cache instead of waiting for L2 / HBM) ----
  static MD_DEV void prefetch_site(const SM& s, int j) {
#if !defined(MDOPT_EMU) && !defined(MONO_NO_PREFETCH)
    const char* base = reinterpret_cast<const char*>(&s.state.sites[phys(s, j)]);
    const int A = s.state.mirrored ? dim_right(s, j) : dim_left(s, j);   // stored left dimension
    const int lane = threadIdx.x & 31;
    const int nv = (2 * A * 8 + 127) >> 7, nt = (2 * A * 2 + 127) >> 7;   // lines of the value / target arrays
    if (lane < nv) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + lane * 128));
    else if (lane < nv + nt) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + 2 * CHI * 8 + (lane - nv) * 128));
#endif
  }

  // ---- stage site j in the frame view: rows (a, p) -> (s.st, s.sv), a < dim_left(j) ----
  static MD_DEV void load_site(SM& s, int j) {
    const MonoSite<CHI>& src = s.state.sites[phys(s, j)];
    const int A = dim_left(s, j), B = dim_right(s, j), mir = s.state.mirrored;
    if (!mir) {
      MW_FOR(r, 2 * A) { s.sv[r] = src.val[r]; s.st[r] = src.tgt[r]; }
    } else {
      // the stored tensor has dims (B, 2, A): rows (b, p) -> a.  Invert per p (unique writers: p-monomial).
      MW_FOR(r, 2 * A) { s.st[r] = -1; s.sv[r] = 0.0; }
      MW_SYNC();
      MW_FOR(r, 2 * B) {
        const int t = src.tgt[r];
        if (t >= 0) { const int o = t * 2 + (r & 1); s.st[o] = (int16_t)(r >> 1); s.sv[o] = src.val[r]; }
      }
    }
    MW_LANE0 { s.state.cnt.bytes_hbm += 10.0 * 2 * (mir ? B : A); }
    MW_SYNC();
  }

  // ---- new carried row from entry (q, m) of the coefficient side, output physical index pd, factor f:
  //      sum over (w', pu) of f * W(q, w', pu, pd) * A[m, pu, .] must live in a single column ----
  static MD_DEV bool gather_row(const SM& s, const MonoTensor& T, int Bn, int q, int m, int pd, double f,
                                int* col_out, double* val_out) {
    const int idx = q * 2 + pd;
    const int n = T.en[idx];
    int col = -1; double val = 0.0; bool ok = (n <= 2);
    for (int e = 0; e < n && e < 2; ++e) {
      const int r = m * 2 + T.epu[idx][e];
      const int t = s.st[r];
      if (t < 0) continue;
      const int c = T.ewo[idx][e] * Bn + t;
      if (col >= 0 && c != col) ok = false;
      col = c;
      val += (f * T.ew[idx][e]) * s.sv[r];
    }
    *col_out = col; *val_out = val;
    return ok;
  }

  // ---- carried <- site j with the first MPO tensor applied (contractor.py:276-291, first half) ----
  static MW_PHASE void carried_from_site(SM& s, int j, int tid, int jnext) {
    load_tensor(s, tid);
    const MonoTensor& T = s.tw;
    const int A = dim_left(s, j), B = dim_right(s, j), vr = T.vr;
    if (jnext >= 0) prefetch_site(s, jnext);
    load_site(s, j);
    int bad = (T.vl != 1);
    // rows (k, p') of the carried matrix; columns (q, m) = q*B + m
    MW_FOR(r, 2 * A) {
      int col; double val;
      if (!gather_row(s, T, B, 0, r >> 1, r & 1, 1.0, &col, &val)) bad = 1;
      s.cv[r] = val; s.cc[r] = (int16_t)col;
    }
    bad = MW_ANY(bad);
    s.state.K = A; s.state.Wq = vr; s.state.Mr = B;
    if (bad) s.state.status = ST_NOT_MONOMIAL;
    MW_SYNC();
  }

  // ---- one two-site step: absorb site jn (frame index) with MPO tensor tid ----
  static MW_PHASE void step(SM& s, int jn, int tid, int chi_lim, double cut, int is_move, int renorm_sv, int jnext) {
    load_tensor(s, tid);
    const MonoTensor& T = s.tw;
    const int vl = T.vl, vr = T.vr, iso = T.iso;
    const int K = s.state.K, Wq = s.state.Wq, Mr = s.state.Mr, mir = s.state.mirrored;
    const int Bn = dim_right(s, jn);
    const int R = K * 2;
    int C = Wq * Mr;
    if (vl != Wq || dim_left(s, jn) != Mr) { s.state.status = ST_CAPACITY; MW_SYNC(); return; }
    if (jnext >= 0) prefetch_site(s, jnext);
    load_site(s, jn);
    int bad = 0;
    if (!iso) {
      // theta = Mmat . T formed explicitly (contractor.py:328-342): row (k, r) keeps one entry, now in column
      // (p', w', m') = (p'*vr + w')*Bn + m'
      if (2 * vr * Bn > N2) { s.state.status = ST_CAPACITY; MW_SYNC(); return; }
      MW_FOR(r, R) {
        const int c = s.cc[r];
        if (c < 0) continue;
        const int q = (c >= Mr) ? 1 : 0, m = c - q * Mr;   // MPO bonds are <= 2
        const double v = s.cv[r];
        int col = -1; double val = 0.0;
        for (int pd = 0; pd < 2; ++pd) {
          int c1; double v1;
          if (!gather_row(s, T, Bn, q, m, pd, v, &c1, &v1)) bad = 1;
          if (c1 < 0) continue;
          c1 += pd * vr * Bn;
          if (col >= 0 && c1 != col) bad = 1;
          col = c1; val += v1;
        }
        s.cc[r] = (int16_t)col; s.cv[r] = val;
      }
      C = 2 * vr * Bn;
      MW_SYNC();
    }
    // ---- column sums: at most one row per p may hit a column ----
    MW_FOR(c, C) { s.acc[0][c] = 0.0; s.acc[1][c] = 0.0; s.own[0][c] = -1; s.own[1][c] = -1; }
    MW_SYNC();
    MW_FOR(r, R) {
      const int c = s.cc[r];
      if (c >= 0) { const double v = s.cv[r]; s.own[r & 1][c] = (int16_t)(r >> 1); s.acc[r & 1][c] = v * v; }
    }
    MW_SYNC();
    MW_FOR(r, R) { const int c = s.cc[r]; if (c >= 0 && s.own[r & 1][c] != (r >> 1)) bad = 1; }
    if (MW_ANY(bad)) { s.state.status = ST_NOT_MONOMIAL; MW_SYNC(); return; }
    // sigma^2, sigma, first supporting row, largest sigma^2; the sort keys go into acc[1] (viewed as 64-bit integers)
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(s.acc[1]);
    double mx = 0.0;
    MW_FOR(c, C) {
      const double n2 = s.acc[0][c] + s.acc[1][c];
      const int o0 = s.own[0][c], o1 = s.own[1][c];
      const int first = (o0 >= 0 && (o1 < 0 || 2 * o0 < 2 * o1 + 1)) ? 2 * o0 : ((o1 >= 0) ? 2 * o1 + 1 : (int)MONO_ROW_MASK);
      s.acc[0][c] = n2;       // sigma^2 (each column is handled by one lane: no hazard with the read above)
      s.sig[c] = sqrt(n2);
      mx = fmax(mx, n2);
      unsigned long long kb;
#ifdef MDOPT_EMU
      memcpy(&kb, &n2, 8);
#else
      kb = (unsigned long long)__double_as_longlong(n2);
#endif
      keys[c] = (n2 > 0.0) ? ((kb & ~MONO_ROW_MASK) | (MONO_ROW_MASK - (unsigned long long)first)) : 0ull;
    }
    mx = mw_max(mx);
    const double floor2 = mx * (4.0 * EPS) * (4.0 * EPS);
    MW_SYNC();
    // truncation count (utils.py:119): #(sigma > cut) among the numerically non-zero columns
    int cnt_keep = 0, cnt_nz = 0;
    MW_FOR(c, C) {
      const double n2 = s.acc[0][c];
      cnt_keep += (s.sig[c] > cut && n2 > floor2) ? 1 : 0;
      cnt_nz += (n2 > 0.0) ? 1 : 0;
    }
    cnt_keep = mw_isum(cnt_keep);
    const int nnz = mw_isum(cnt_nz);
    int Knew = cnt_keep < chi_lim ? cnt_keep : chi_lim;
    // ---- canonical order: sort by (sigma^2, first row); values within the tie width of their predecessor (chained)
    //      form one cluster; inside a cluster the order is by first supporting row ----
    mw_sort_desc<N2>(keys, C);
    // cluster flags: position p opens a cluster.  Device: one ballot word per 32 sorted positions (position
    // p = lane + 32*i is bit `lane` of word i, positions >= nnz count as starts), kept in shared memory, so the
    // bounds of p's cluster are bit scans instead of walks.  The loops stay rolled: the instruction caches, not the
    // instruction count, limit this kernel.
    unsigned* fw = reinterpret_cast<unsigned*>(s.own[0]);   // own[] is free after the structure check (NW <= N2/2 ints)
    int multi = 0;
#ifndef MDOPT_EMU
    {
      const int lane = threadIdx.x & 31;
      const int nw = (nnz + 31) >> 5;
#pragma unroll 1
      for (int i = 0; i < nw; ++i) {
        const int p = lane + 32 * i;
        int flag = 1;
        if (p < nnz && p > 0) {
          const double a = __longlong_as_double((long long)(keys[p - 1] & ~MONO_ROW_MASK));
          const double b = __longlong_as_double((long long)(keys[p] & ~MONO_ROW_MASK));
          flag = (b < a * (1.0 - MONO_TIE_REL2)) ? 1 : 0;
        }
        const unsigned w = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) fw[i] = w;
        multi |= (w != 0xffffffffu);
      }
    }
#else
    {
      const int nw = (nnz + 31) >> 5;
      for (int i = 0; i < nw; ++i) fw[i] = 0u;
      for (int p = 0; p < 32 * nw; ++p) {
        int flag = 1;
        if (p < nnz && p > 0) {
          double a, b;
          const unsigned long long ka = keys[p - 1] & ~MONO_ROW_MASK, kb = keys[p] & ~MONO_ROW_MASK;
          memcpy(&a, &ka, 8); memcpy(&b, &kb, 8);
          flag = (b < a * (1.0 - MONO_TIE_REL2)) ? 1 : 0;
        }
        if (flag) fw[p >> 5] |= 1u << (p & 31); else multi = 1;
      }
    }
#endif
    MW_FOR(c, C) s.rank[c] = -1;
    MW_SYNC();
    double kept2 = 0.0, err2 = 0.0;
    MW_FOR(p, nnz) {
      const int first = (int)(MONO_ROW_MASK - (keys[p] & MONO_ROW_MASK));
      int pos = p;
      if (multi) {   // rank inside the cluster = members with an earlier first row
        const int lane = p & 31, nw = (nnz + 31) >> 5;
        // lo = last cluster start at or before p; hi = first cluster start after p (or nnz)
        int wi = p >> 5;
        unsigned w = fw[wi] & (0xffffffffu >> (31 - lane));
        while (w == 0u) w = fw[--wi];            // position 0 always opens a cluster
        const int lo = 32 * wi + (31 - mw_clz(w));
        int vi = p >> 5;
        unsigned v = (lane == 31) ? 0u : (fw[vi] & (0xffffffffu << (lane + 1)));
        while (v == 0u && vi + 1 < nw) v = fw[++vi];
        int hi = nnz;
        if (v != 0u) { const int h = 32 * vi + mw_ctz(v); hi = h < nnz ? h : nnz; }
        if (hi - lo > 1) {
          pos = lo;
          for (int q = lo; q < hi; ++q) pos += ((int)(MONO_ROW_MASK - (keys[q] & MONO_ROW_MASK)) < first) ? 1 : 0;
        }
      }
      const int c = s.cc[first];
      s.order[pos] = (int16_t)c;
      const double n2 = s.acc[0][c];
      // a kept column leaves 1 / sigma in acc[0] for the U store below (one division per kept column, not per row)
      if (pos < Knew) { s.rank[c] = (int16_t)pos; kept2 += n2; s.acc[0][c] = 1.0 / s.sig[c]; } else err2 += n2;
    }
    if (renorm_sv) kept2 = mw_sum(kept2);
    err2 = mw_sum(err2);
    MW_SYNC();
    if (s.state.trace != nullptr) {
      const int nt = s.state.n_traced;
      if (nt < s.state.P.trace_max) {
        const int stride = s.state.P.trace_stride;
        double* tr = s.state.trace + (size_t)nt * stride;
        MW_LANE0 { tr[0] = R; tr[1] = C; tr[2] = Knew; tr[3] = is_move; }
        MW_FOR(c, N2) { if (4 + c < stride) tr[4 + c] = (c < nnz) ? s.sig[s.order[c]] : 0.0; }
      }
      MW_SYNC();
      MW_LANE0 { s.state.n_traced = nt + 1; }
    }
    if (Knew > CHI) { s.state.status = ST_CAPACITY; MW_SYNC(); return; }
    const double dscale = (renorm_sv && kept2 > 0.0) ? 1.0 / sqrt(kept2) : 1.0;
    const int zero = (Knew == 0);
    if (zero) Knew = 1;   // zero matrix: keep a single zero direction (the state has zero norm)
    // ---- site jn-1 <- U: row (k, p) -> (rank of its column, value / sigma) ----
    {
      MonoSite<CHI>& dst = s.state.sites[phys(s, jn - 1)];
      if (!mir) {
        MW_FOR(r, R) {
          const int c = s.cc[r];
          const int t = (c >= 0) ? s.rank[c] : -1;
          double v = 0.0;
          if (t >= 0) v = s.cv[r] * s.acc[0][c];
          int tt = t;
          if (zero) { tt = (r == 0) ? 0 : -1; v = (r == 0) ? 1.0 : 0.0; }
          dst.tgt[r] = (int16_t)tt; dst.val[r] = v;
        }
      } else {   // stored dims (Knew, 2, K): rows (kn, p) -> k
        MW_FOR(r, 2 * Knew) { s.ut_buf()[r] = -1; s.uv_buf()[r] = 0.0; }
        MW_SYNC();
        MW_FOR(r, R) {
          const int c = s.cc[r];
          const int t = (c >= 0) ? s.rank[c] : -1;
          if (t >= 0 && !zero) { const int o = t * 2 + (r & 1); s.ut_buf()[o] = (int16_t)(r >> 1); s.uv_buf()[o] = s.cv[r] * s.acc[0][c]; }
        }
        if (zero) { MW_LANE0 { s.ut_buf()[0] = 0; s.uv_buf()[0] = 1.0; } }
        MW_SYNC();
        MW_FOR(r, 2 * Knew) { dst.tgt[r] = s.ut_buf()[r]; dst.val[r] = s.uv_buf()[r]; }
      }
    }
    MW_SYNC();
    // ---- next carried matrix (the old rows were last read by the U store above; the new ones depend on order / sig /
    //      the staged site only) ----
    bad = 0;
    if (iso) {
      MW_FOR(rn, 2 * Knew) {
        int col = -1; double val = 0.0;
        if (!zero) {
          const int c = s.order[rn >> 1];
          const int q = (c >= Mr) ? 1 : 0, m = c - q * Mr;
          if (!gather_row(s, T, Bn, q, m, rn & 1, dscale * s.sig[c], &col, &val)) bad = 1;
        }
        s.cv[rn] = val; s.cc[rn] = (int16_t)col;
      }
    } else {
      // coefficients diag(sigma) over columns (p', w', m'): carried(Knew, 2, vr, Bn), row (kn, p') -> (w', m')
      MW_FOR(rn, 2 * Knew) {
        int col = -1; double val = 0.0;
        if (!zero) {
          const int c = s.order[rn >> 1];
          const int pd = (c >= vr * Bn) ? 1 : 0;
          if (pd == (rn & 1)) { col = c - pd * vr * Bn; val = dscale * s.sig[c]; }
        }
        s.cv[rn] = val; s.cc[rn] = (int16_t)col;
      }
    }
    bad = MW_ANY(bad);
    MW_LANE0 {
      MonoState<CHI>& st = s.state;
      if (bad) st.status = ST_NOT_MONOMIAL;
      const int p = phys(s, jn - 1);
      if (mir) st.bd[p] = Knew; else st.bd[p + 1] = Knew;   // set_dim_right(jn - 1)
      st.K = Knew; st.Wq = vr; st.Mr = Bn;
      st.cnt.n_steps += 1.0;
      st.cnt.bytes_hbm += 10.0 * 2 * (mir ? Knew : K);
      if (cnt_keep > chi_lim && err2 > 0.0) st.cnt.n_trunc += 1.0;
    }
    MW_SYNC();
  }

  // ---- carried (K, 2, 1, Mr) -> site j, optionally divided by its Frobenius norm (optimiser/utils.py:340-345) ----
  static MW_PHASE void carried_to_site(SM& s, int j, int renorm) {
    const int K = s.state.K, Mr = s.state.Mr, mir = s.state.mirrored;
    const int n = K * 2;
    double scale = 1.0;
    if (renorm) {
      double acc = 0.0;
      MW_FOR(r, n) { if (s.cc[r] >= 0) { const double v = s.cv[r]; acc += v * v; } }
      acc = mw_sum(acc);
      const double nrm = sqrt(acc);
      scale = (nrm > 0.0) ? 1.0 / nrm : 1.0;
    }
    MonoSite<CHI>& dst = s.state.sites[phys(s, j)];
    if (!mir) {
      MW_FOR(r, n) { const int c = s.cc[r]; dst.tgt[r] = (int16_t)c; dst.val[r] = (c >= 0) ? s.cv[r] * scale : 0.0; }
    } else {   // stored dims (Mr, 2, K): rows (m, p) -> k
      MW_FOR(r, 2 * Mr) { s.ut_buf()[r] = -1; s.uv_buf()[r] = 0.0; }
      MW_SYNC();
      MW_FOR(r, n) {
        const int c = s.cc[r];
        if (c >= 0) { const int o = c * 2 + (r & 1); s.ut_buf()[o] = (int16_t)(r >> 1); s.uv_buf()[o] = s.cv[r] * scale; }
      }
      MW_SYNC();
      MW_FOR(r, 2 * Mr) { dst.tgt[r] = s.ut_buf()[r]; dst.val[r] = s.uv_buf()[r]; }
    }
    MW_LANE0 { s.state.cnt.bytes_hbm += 10.0 * 2 * (mir ? Mr : K); }
    MW_SYNC();
  }

  // ---- sweep of an MPO over frame sites start..start+len-1 (tids == nullptr: identity MPO = centre move) ----
  static MD_DEV void sweep(SM& s, int start, int len, const int* tids, int chi_lim, double cut, int renorm_after,
                           int is_move, int renorm_sv) {
    carried_from_site(s, start, tids ? tids[0] : -1, len > 1 ? start + 1 : -1);
    if (s.state.status == ST_NOT_MONOMIAL) return;
    for (int t = 1; t < len; ++t) {
      step(s, start + t, tids ? tids[t] : -1, chi_lim, cut, is_move, renorm_sv, t + 1 < len ? start + t + 1 : -1);
      const int st = s.state.status;
      if (st == ST_CAPACITY || st == ST_NOT_MONOMIAL) return;
    }
    if (s.state.Wq != 1) { s.state.status = ST_CAPACITY; MW_SYNC(); return; }
    carried_to_site(s, start + len - 1, renorm_after);
  }

  // ---- product state + bit-flip bias (mps/utils.py:439-512, decoding.py:140-190) ----
  static MW_PHASE void init_product(SM& s, const uint8_t* sym) {
    const DecodeParams& P = s.state.P;
    const int N = P.n_sites, kL = P.n_logical;
    const double bp = P.bias_p;
    const double d = (bp >= 0.0) ? sqrt(1.0 - bp) : 1.0;
    const double o = (bp >= 0.0) ? sqrt(bp) : 0.0;
    MonoSite<CHI>* sites = s.state.sites;
    int* bd = s.state.bd;
    MW_FOR(j, N) {
      double a0, a1;
      const int c = sym[j];
      if (c == 0) { a0 = 1.0; a1 = 0.0; }
      else if (c == 1) { a0 = 0.0; a1 = 1.0; }
      else { a0 = 1.0 / sqrt(2.0); a1 = a0; }
      if (j >= kL && bp >= 0.0) {
        const double b0 = a0 * d + a1 * o, b1 = a0 * o + a1 * d;
        a0 = b0; a1 = b1;
      }
      sites[j].val[0] = a0; sites[j].val[1] = a1;
      sites[j].tgt[0] = 0; sites[j].tgt[1] = 0;
    }
    MW_FOR(j, N + 1) bd[j] = 1;
    MW_LANE0 { s.state.centre = 0; s.state.mirrored = 0; s.state.status = ST_OK; s.state.n_traced = 0; }
    MW_SYNC();
  }

  // ---- Dephasing-DMRG read-out for more than 12 logical sites (decoding.py:1382-1400 ->
  //      mdopt/optimiser/dephasing_dmrg.py:147-374, started from |+...+>, mode "LA").
  // The reference's search MPS is snapped to one computational-basis configuration after every bond update
  // (:236-249), so it keeps bond dimension 1 and every site is '+', 0 or 1; its environments are then outer products
  // of a vector with itself and the effective two-site operator (_EINSUM, :47) is diagonal in the computational
  // basis with entries |l . T_i[a] . T_{i+1}[b] . r|^2.  eigsh(k=1) returns the basis vector of the largest entry
  // (first maximal index on exact ties).  Here the target tensors are monomial, so l / r are vectors of <= chi
  // entries updated by gathers / scatters, and a bond update is four warp reductions.
  // The engine works on the REVERSED logical chain (decoding.py:1351-1356): reversed site i = logical site
  // kL-1-i.  out[i] = bit of reversed site i (what engine.mps holds); *success = overlap with |0...0>. ----
  static MW_PHASE void dmrg_readout(SM& s, double* out, int* success, int iters) {
    const DecodeParams& P = s.state.P;
    const int kL = P.n_head;
    const MonoSite<CHI>* sites = s.state.sites;
    const int* bd = s.state.bd;
    double* env = s.state.env;                 // [kL + 1][CHI] far-side environments of the current half sweep (global)
    int16_t* bit = s.rank;                     // per logical site: -1 = '+', else the bit (kL <= 2*CHI)
    double* near = s.sig;                      // near-side environment, updated as the sweep moves
    double* tmp = s.sv;
    const double isq2 = 1.0 / sqrt(2.0);
    if (kL > N2 || env == nullptr) { s.state.status = ST_CAPACITY; MW_SYNC(); return; }
    MW_FOR(j, kL) bit[j] = -1;
    MW_SYNC();
    // weight of physical value p on site j under the current search state
    auto wgt = [&](int j, int p) -> double { const int b = bit[j]; return b < 0 ? isq2 : (b == p ? 1.0 : 0.0); };
    auto dimL = [&](int j) -> int { return bd[j]; };
    // far-side environments: forward[j] = contraction of sites < j (left bond vector of site j)
    auto build_forward = [&]() {
      MW_LANE0 { env[0] = 1.0; }
      MW_SYNC();
      for (int j = 0; j + 1 < kL; ++j) {
        const double* cur = env + (size_t)j * CHI;
        double* nxt = env + (size_t)(j + 1) * CHI;
        const int A = dimL(j), B = dimL(j + 1);
        MW_FOR(b, B) nxt[b] = 0.0;
        MW_SYNC();
        for (int p = 0; p < 2; ++p) {     // for a fixed p every target has one writer (p-monomial tensors)
          const double w = wgt(j, p);
          if (w != 0.0) {
            MW_FOR(a, A) {
              const int t = sites[j].tgt[a * 2 + p];
              if (t >= 0) nxt[t] += cur[a] * sites[j].val[a * 2 + p] * w;
            }
          }
          MW_SYNC();
        }
      }
    };
    // backward[j] = contraction of sites >= j (left bond vector of site j), backward[kL] = [1]
    auto build_backward = [&]() {
      MW_LANE0 { env[(size_t)kL * CHI] = 1.0; }
      MW_SYNC();
      for (int j = kL - 1; j >= 2; --j) {
        const double* nxt = env + (size_t)(j + 1) * CHI;
        double* cur = env + (size_t)j * CHI;
        const int A = dimL(j);
        MW_FOR(a, A) {
          double acc = 0.0;
          for (int p = 0; p < 2; ++p) {
            const int t = sites[j].tgt[a * 2 + p];
            if (t >= 0) acc += wgt(j, p) * sites[j].val[a * 2 + p] * nxt[t];
          }
          cur[a] = acc;
        }
        MW_SYNC();
      }
    };
    // bond update on logical sites (j, j+1): lv = left vector of site j, rv = left vector of site j+2
    auto update = [&](int j, const double* lv, const double* rv) {
      const int A = dimL(j);
      double w[4] = {0.0, 0.0, 0.0, 0.0};    // index a*2 + b: a = bit of logical site j+1 (reversed site i), b = site j
      MW_FOR(bl, A) {
        const double l = lv[bl];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int bm = sites[j].tgt[bl * 2 + b];
          if (bm < 0) continue;
          const double lb = l * sites[j].val[bl * 2 + b];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int br = sites[j + 1].tgt[bm * 2 + a];
            if (br >= 0) w[a * 2 + b] += lb * sites[j + 1].val[bm * 2 + a] * rv[br];
          }
        }
      }
      // first maximal index, like the snap's argmax (:245); weights within 1e-9 relative of the largest count as
      // tied (exact ties -- unconstrained logical sites -- differ by rounding noise between implementations)
      double v2[4], top = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) { const double v = mw_sum(w[q]); v2[q] = v * v; top = fmax(top, v2[q]); }
      int best = 3;
#pragma unroll
      for (int q = 3; q >= 0; --q) { if (v2[q] >= top * (1.0 - 1e-9)) best = q; }
      MW_SYNC();
      MW_LANE0 { bit[j + 1] = (int16_t)(best >> 1); bit[j] = (int16_t)(best & 1); }
      MW_SYNC();
    };
    for (int it = 0; it < iters; ++it) {
      // reversed bonds 0 .. kL-2 = logical bonds (kL-2, kL-1) down to (0, 1): far side = forward vectors
      build_forward();
      MW_LANE0 { near[0] = 1.0; }            // left vector of site kL (nothing to the right yet)
      MW_SYNC();
      for (int j = kL - 2; j >= 0; --j) {
        update(j, env + (size_t)j * CHI, near);
        // near <- left vector of site j+1 = site j+1 (now fixed) applied to the old near
        const int A = dimL(j + 1);
        MW_FOR(a, A) {
          const int p = bit[j + 1];
          const int t = sites[j + 1].tgt[a * 2 + p];
          tmp[a] = (t >= 0) ? sites[j + 1].val[a * 2 + p] * near[t] : 0.0;
        }
        MW_SYNC();
        MW_FOR(a, A) near[a] = tmp[a];
        MW_SYNC();
      }
      // back: logical bonds (0, 1) up to (kL-2, kL-1): far side = backward vectors, near = forward from the left
      build_backward();
      MW_LANE0 { near[0] = 1.0; }
      MW_SYNC();
      for (int j = 0; j + 1 < kL; ++j) {
        update(j, near, env + (size_t)(j + 2) * CHI);
        const int A = dimL(j), B = dimL(j + 1);
        MW_FOR(b, B) tmp[b] = 0.0;
        MW_SYNC();
        MW_FOR(a, A) {
          const int p = bit[j];
          const int t = sites[j].tgt[a * 2 + p];
          if (t >= 0) tmp[t] = near[a] * sites[j].val[a * 2 + p];
        }
        MW_SYNC();
        MW_FOR(b, B) near[b] = tmp[b];
        MW_SYNC();
      }
    }
    int nonzero = 0;
    MW_FOR(i, kL) { const int b = bit[kL - 1 - i]; out[i] = (double)b; nonzero |= (b != 0); }
    nonzero = MW_ANY(nonzero);
    MW_LANE0 { *success = nonzero ? 0 : 1; }
    MW_SYNC();
  }

  // ---- marginal over sites kL..N-1, reverse, dense, |.|, L2 renorm, tolerant arg-max ----
  //   canonical.py:956-989, :150-161, :241-272; decoding.py:1362-1379
  static MW_PHASE void marginal_readout(SM& s, double* out, int* success, int dmrg_iters) {
    const DecodeParams& P = s.state.P;
    const int N = P.n_sites, kL = P.n_head, ren = P.renormalise;
    // (erasures of qubits whose index is a logical-site index change the kept sites: dense engine; not with DMRG)
    if (dmrg_iters == 0 && P.kept_mask != ((kL >= 32) ? 0xffffffffu : ((1u << kL) - 1u))) {
      s.state.status = ST_NOT_MONOMIAL; MW_SYNC(); return;
    }
    MonoSite<CHI>* sites = s.state.sites;
    int* bd = s.state.bd;
    double* T = s.cv;     // current last tensor (L x 2)
    double* w = s.sig;    // traced vector
    int L = bd[N - 1];
    MW_FOR(t, L * 2) T[t] = (sites[N - 1].tgt[t] >= 0) ? sites[N - 1].val[t] : 0.0;
    MW_SYNC();
    double bytes = 0.0;
    for (int site = N - 1; site >= kL; --site) {
      MW_FOR(l, L) w[l] = (T[l * 2] + T[l * 2 + 1]) * (1.0 / sqrt(2.0));
      MW_SYNC();
      const int L2 = bd[site - 1];
      const MonoSite<CHI>& A = sites[site - 1];   // (L2, 2, L)
      double acc = 0.0;
      MW_FOR(t, L2 * 2) {
        const int b = A.tgt[t];
        const double v = (b >= 0) ? A.val[t] * w[b] : 0.0;
        T[t] = v; acc += v * v;
      }
      bytes += 10.0 * 2 * L2;
      if (ren) {
        acc = mw_sum(acc);
        const double nrm = sqrt(acc);
        MW_SYNC();
        MW_FOR(t, L2 * 2) T[t] = T[t] / nrm;   // unguarded like the reference: 0/0 -> NaN
      }
      MW_SYNC();
      L = L2;
    }
    // the merged last logical tensor goes back into the MPS (what CanonicalMPS.marginal returns)
    MW_FOR(t, L * 2) { sites[kL - 1].val[t] = T[t]; sites[kL - 1].tgt[t] = 0; }
    MW_LANE0 { bd[kL] = 1; s.state.cnt.bytes_hbm += bytes; }
    MW_SYNC();
    if (dmrg_iters > 0) { dmrg_readout(s, out, success, dmrg_iters); return; }
    // dense stage over the kL logical sites: every class follows one path through the monomial tensors
    const int ncls = 1 << kL;
    double nrm2 = 0.0, top = 0.0; int badv = 0;
    MW_FOR(i, ncls) {
      int b = 0; double amp = 1.0;
      for (int j = 0; j < kL; ++j) {
        const int r = b * 2 + ((i >> j) & 1);
        if (j == kL - 1) { amp *= T[r]; break; }
        const int t = sites[j].tgt[r];
        if (t < 0) { amp = 0.0; break; }
        amp *= sites[j].val[r]; b = t;
      }
      const double v = fabs(amp);
      out[i] = v; nrm2 += v * v; top = fmax(top, v); badv |= (v != v);
    }
    nrm2 = mw_sum(nrm2); top = mw_max(top); badv = MW_ANY(badv);
    const double nrm = sqrt(nrm2);
    const double inv = (ren && nrm > 1e-17) ? 1.0 / nrm : 1.0;
    const int bad = (badv || nrm != nrm) ? 1 : 0;
    MW_SYNC();
    MW_FOR(i, ncls) out[i] = out[i] * inv;
    MW_SYNC();
    MW_LANE0 {
      if (bad) *success = 0;
      else {
        const double t2 = top * inv;
        const double eps = fmax(P.readout_rel * t2, P.readout_abs);
        *success = (out[0] >= t2 - eps) ? 1 : 0;
      }
      if (bad && s.state.status == ST_OK) s.state.status = ST_ZERO_NORM;
    }
    MW_SYNC();
  }

  // ---- run the program.  Every sweep -- the centre move in front of an op (canonical.py:345-420; moving left runs in
  //      the mirrored frame like the reference's reverse()) and the MPO sweep itself -- goes through ONE call site ----
  static MD_DEV void run(SM& s, const int* prog, const uint8_t* sym, double* out, int* success) {
    init_product(s, sym);
    const int n_ops = s.state.P.n_ops, n_sites = s.state.P.n_sites;
    int pc = 0;
    while (pc < n_ops) {
      const int op = prog[pc];
      if (op == OP_END) break;
      if (op == OP_MARGINAL) {
        marginal_readout(s, out, success, 0);
        pc += 1;
      } else if (op == OP_MARGINAL_DMRG) {   // [6, sweeps]: marginal chain, then the Dephasing-DMRG read-out
        marginal_readout(s, out, success, prog[pc + 1] > 0 ? prog[pc + 1] : 1);
        pc += 2;
      } else if (op == OP_SWEEP || op == OP_MOVE) {
        const int target = prog[pc + 1];
        for (int phase = (target == s.state.centre) ? 1 : 0; phase < 2; ++phase) {
          if (phase == 1 && op == OP_MOVE) break;
          int start, len, chi_lim, ren = 0, rsv = 0, mir = 0;
          double cut;
          const int* tids = nullptr;
          const int centre = s.state.centre;
          if (phase == 0) {   // move the centre to `target`
            mir = (target < centre) ? 1 : 0;
            start = mir ? (n_sites - 1 - centre) : centre;
            len = (mir ? (centre - target) : (target - centre)) + 1;
            chi_lim = CHI; cut = s.state.P.cut_move;
          } else {
            start = target; len = prog[pc + 2]; ren = prog[pc + 3]; rsv = prog[pc + 4];
            tids = prog + pc + 5; chi_lim = s.state.P.chi_max; cut = s.state.P.cut_sweep;
          }
          MW_SYNC();
          MW_LANE0 { s.state.mirrored = mir; }
          MW_SYNC();
          sweep(s, start, len, tids, chi_lim, cut, ren, phase == 0, rsv);
          MW_SYNC();
          MW_LANE0 { s.state.mirrored = 0; }
          const int st = s.state.status;
          if (st == ST_CAPACITY || st == ST_NOT_MONOMIAL) break;
          MW_LANE0 { s.state.centre = (phase == 0) ? target : (target + len - 1); }
          MW_SYNC();
        }
        pc += (op == OP_MOVE) ? 2 : (5 + prog[pc + 2]);
      } else {
        MW_LANE0 { s.state.status = ST_NOT_MONOMIAL; }   // two-site gates / compression: dense engine
        MW_SYNC();
      }
      const int st = s.state.status;
      if (st == ST_CAPACITY || st == ST_NOT_MONOMIAL) break;
    }
  }

#ifndef MDOPT_EMU
  // ---- the same program with all warps of the CTA in LOCK-STEP: one `bar.sync` after every two-site step.  The
  //      program is error-independent, so its control flow is the same for every syndrome; a warp without work, or
  //      whose syndrome has failed, keeps walking the program and only takes the barriers.  Purpose: the warps of an
  //      SM then run the same ~2 K instructions of a step at the same time and share the L0 / L1.5 instruction
  //      caches instead of evicting each other (no_instruction is the largest stall of the free-running kernel). ----
  static __device__ __forceinline__ void run_lockstep(SM& s, const int* prog, const uint8_t* sym, double* out,
                                                      int* success, bool active) {
    if (active) init_product(s, sym);
    const int n_ops = s.state.P.n_ops, n_sites = s.state.P.n_sites;
    int pc = 0, centre = 0;     // CTA-uniform: follows the program, not the data
    auto ok = [&]() -> bool {
      if (!active) return false;
      const int st = s.state.status;
      return st != ST_CAPACITY && st != ST_NOT_MONOMIAL;
    };
    while (pc < n_ops) {
      const int op = prog[pc];
      if (op == OP_END) break;
      if (op == OP_MARGINAL || op == OP_MARGINAL_DMRG) {
        const int iters = (op == OP_MARGINAL) ? 0 : (prog[pc + 1] > 0 ? prog[pc + 1] : 1);
        if (ok()) marginal_readout(s, out, success, iters);
        pc += (op == OP_MARGINAL) ? 1 : 2;
      } else if (op == OP_SWEEP || op == OP_MOVE) {
        const int target = prog[pc + 1];
        for (int phase = (target == centre) ? 1 : 0; phase < 2; ++phase) {
          if (phase == 1 && op == OP_MOVE) break;
          int start, len, chi_lim, ren = 0, rsv = 0, mir = 0;
          double cut;
          const int* tids = nullptr;
          if (phase == 0) {
            mir = (target < centre) ? 1 : 0;
            start = mir ? (n_sites - 1 - centre) : centre;
            len = (mir ? (centre - target) : (target - centre)) + 1;
            chi_lim = CHI; cut = s.state.P.cut_move;
          } else {
            start = target; len = prog[pc + 2]; ren = prog[pc + 3]; rsv = prog[pc + 4];
            tids = prog + pc + 5; chi_lim = s.state.P.chi_max; cut = s.state.P.cut_sweep;
          }
          MW_SYNC();
          MW_LANE0 { s.state.mirrored = mir; }
          MW_SYNC();
          if (ok()) carried_from_site(s, start, tids ? tids[0] : -1, len > 1 ? start + 1 : -1);
          __syncthreads();
          for (int t = 1; t < len; ++t) {
            if (ok()) step(s, start + t, tids ? tids[t] : -1, chi_lim, cut, phase == 0, rsv, t + 1 < len ? start + t + 1 : -1);
            __syncthreads();
          }
          if (ok()) {
            if (s.state.Wq != 1) { MW_LANE0 { s.state.status = ST_CAPACITY; } }
            else carried_to_site(s, start + len - 1, ren);
          }
          MW_SYNC();
          MW_LANE0 { s.state.mirrored = 0; }
          centre = (phase == 0) ? target : (target + len - 1);
          MW_SYNC();
        }
        pc += (op == OP_MOVE) ? 2 : (5 + prog[pc + 2]);
      } else {
        MW_LANE0 { s.state.status = ST_NOT_MONOMIAL; }   // two-site gates / compression: dense engine
        MW_SYNC();
        pc = n_ops;    // uniform: the program itself is outside the engine's domain
      }
    }
  }
#endif
};

}  // namespace mdopt
