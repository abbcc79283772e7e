// EXPERIMENT (not built): Householder QR without column exchanges (row-echelon R) with one barrier per column
// and look-ahead for the next column's owner.  Correct (the whole GPU test tier passes with it), but not
// adopted: R's diagonal staying above the threshold does not bound the smallest singular value of the accepted
// columns, so on the decode workload it accepts more than chi_max columns 6.5x more often than the pivoted
// factorisation (517 k instead of 79 k truncating steps per 2 072 syndromes) and every such step pays a Jacobi
// SVD: 375 syndromes/s instead of 946.  Restricted to centre moves (where chi_max cannot bind) it measured 879/s,
// i.e. one barrier per column with look-ahead is not faster per reflector than the two-barrier pivoted loop either.
#pragma once
#include "../block_ops.cuh"
namespace mdopt {
#ifndef MDOPT_EMU
// -------------------------------------------------------------------------------------------------
// qr_echelon_reg: rank-revealing Householder QR WITHOUT column exchanges, matrix in registers (same layout
// and outputs as qrcp_reg).  Columns are visited left to right; a column whose remainder below the current
// rank K is <= rel_tol * (largest initial column norm) lies in the span of the accepted ones and is skipped,
// otherwise it becomes reflector K (row-echelon R).  Because the next candidate is known in advance, its
// owner applies the newest reflector to it first and goes straight on to building the next one while the
// other warps are still updating their columns: ONE barrier per column, no arg-max, no norm bookkeeping.
// Householder QR is backward stable without pivoting; the pivoting of qrcp_reg only served the rank
// decision, which the skip rule takes over (the span test is the same threshold).
//   out: s.X[k*LD + c] = R[k][c] (k < K), s.P2 = Vall, s.tau, s.iscr[0] = K, s.iscr[1] = 1 when more than
//        kmax columns are independent (the caller then runs the truncating SVD)
// -------------------------------------------------------------------------------------------------
template <int CHI, int TPC, int CPT>
__device__ __noinline__ void qr_echelon_reg(Smem<CHI>& s, const double* orig, int R, int C, int kmax, double rel_tol) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int NT = (CHI * 8 < 64) ? 64 : CHI * 8;
  constexpr int NCOL = NT / TPC;            // columns covered by one pass of the thread groups
  constexpr int RPT = N2 / TPC;             // register rows per thread and column
  static_assert(RPT <= 16, "the row-slot switch has 16 cases");
  constexpr int CH = RPT < 8 ? RPT : 8;     // register rows per chunk
  constexpr int NCHK = RPT / CH;            // chunk c covers matrix rows [TPC*CH*c, TPC*CH*(c+1))
  constexpr int NWARP = NT / 32;
  const int tid = threadIdx.x;
  const int col0 = tid / TPC, part = tid % TPC, lane = tid & 31;
  const int gbase = lane & ~(TPC - 1);
  const unsigned gmask = ((1u << TPC) - 1u) << gbase;
  double* Vall = s.P2;
  unsigned* keys = reinterpret_cast<unsigned*>(s.vn1);
  double x[CPT][RPT];
  {  // clear the reflector rows this call can produce (v_K is zero above row K)
    const int nclr = min(min(R, C), kmax) * (N2 / 2);
    double2* vz = reinterpret_cast<double2*>(Vall);
    for (int t = tid; t < nclr; t += NT) vz[t] = make_double2(0.0, 0.0);
  }
  {
    unsigned kmax_key = 0u;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int col = col0 + k * NCOL;
      double nn = 0.0;
#pragma unroll
      for (int idx = 0; idx < RPT; ++idx) {
        int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
        x[k][idx] = (col < C && r < R) ? orig[(size_t)r * C + col] : 0.0;
        nn += x[k][idx] * x[k][idx];
      }
#pragma unroll
      for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
      const unsigned kk = (unsigned)__double2hiint(nn);   // non-negative doubles order like their upper words
      kmax_key = kk > kmax_key ? kk : kmax_key;
      // closing record of the column: row where it closed (N2: never visited), accepted flag, diagonal entry
      if (part == 0 && col < N2) { s.perm[col] = N2; s.order[col] = 0; s.sig[col] = 0.0; }
    }
    kmax_key = __reduce_max_sync(0xffffffffu, kmax_key);
    if (lane == 0) keys[tid >> 5] = kmax_key;
  }
  __syncthreads();
  double thr2;
  {
    unsigned key = (lane < NWARP) ? keys[lane] : 0u;
    key = __reduce_max_sync(0xffffffffu, key);
    thr2 = __hiloint2double((int)key, 0) * rel_tol * rel_tol;   // largest squared norm, rounded down < 1e-6
  }
  int K = 0, more = 0;
  double fl = 2.0 * R * C;
  for (int c = 0; c < C; ++c) {
    const int kc = c / NCOL;
    if (col0 == c - kc * NCOL) {
      // ---- owner of column c (its registers already carry H_0 .. H_{K-1}): accept or skip ----
      const int jslot = ((K / (2 * TPC)) << 1) | (K & 1), jpart = (K >> 1) % TPC;
      const int i2p = K / (2 * TPC);
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        if (k == kc) {
          double n0 = 0.0, n1 = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            if (r0 >= K) n0 += x[k][2 * i2] * x[k][2 * i2];
            if (r0 + 1 >= K) n1 += x[k][2 * i2 + 1] * x[k][2 * i2 + 1];
          }
          double nn = n0 + n1;
#pragma unroll
          for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(gmask, nn, off);
          int verdict = 0;
          if (nn > thr2) {
            if (K >= kmax) {
              verdict = 2;
            } else {
              verdict = 1;
              double a = 0.0;
#define MD_QE_SLOT(I) case I: if constexpr ((I) < RPT) a = x[k][(I) < RPT ? (I) : 0]; break;
              switch (jslot) {
                MD_QE_SLOT(0) MD_QE_SLOT(1) MD_QE_SLOT(2) MD_QE_SLOT(3) MD_QE_SLOT(4) MD_QE_SLOT(5) MD_QE_SLOT(6)
                MD_QE_SLOT(7) MD_QE_SLOT(8) MD_QE_SLOT(9) MD_QE_SLOT(10) MD_QE_SLOT(11) MD_QE_SLOT(12)
                MD_QE_SLOT(13) MD_QE_SLOT(14) MD_QE_SLOT(15)
                default: break;
              }
#undef MD_QE_SLOT
              a = __shfl_sync(gmask, a, gbase | jpart);
              const double beta = -copysign(sqrt(nn), a);
              const double d = a - beta;
              const double inv = 1.0 / (beta * d);
              const double tau = -d * d * inv;   // (beta - a) / beta
              const double scal = beta * inv;    // 1 / (a - beta)
#pragma unroll
              for (int i2 = 0; i2 < RPT / 2; ++i2) {
                if (i2 >= i2p) {
                  const int r0 = 2 * TPC * i2 + 2 * part;
                  double2 v;
                  v.x = x[k][2 * i2] * scal;
                  v.y = x[k][2 * i2 + 1] * scal;
                  if (i2 == i2p) {
                    v.x = (r0 > K) ? v.x : ((r0 == K) ? 1.0 : 0.0);
                    v.y = (r0 + 1 > K) ? v.y : ((r0 + 1 == K) ? 1.0 : 0.0);
                  }
                  *reinterpret_cast<double2*>(Vall + K * N2 + r0) = v;
                }
              }
              if (part == 0) { s.tau[K] = tau; s.sig[c] = beta; s.order[c] = 1; }
            }
          }
          if (part == 0) { s.perm[c] = K; s.flag[c] = verdict; }
        }
      }
    }
    __syncthreads();
    const int verdict = s.flag[c];
    if (verdict == 2) { more = 1; break; }
    if (verdict == 0) continue;
    // ---- apply H_K to the columns right of c ----
    bool act[CPT];
    bool any_act = false;
#pragma unroll
    for (int k = 0; k < CPT; ++k) { act[k] = (col0 + k * NCOL > c) && (col0 + k * NCOL < C); any_act |= act[k]; }
    if (__any_sync(0xffffffffu, any_act)) {
      const double tau = s.tau[K];
      const double2* vj = reinterpret_cast<const double2*>(Vall + K * N2 + 2 * part);
      double w[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) w[k] = 0.0;
#pragma unroll
      for (int ch = 0; ch < NCHK; ++ch) {
        if (TPC * CH * (ch + 1) > K && TPC * CH * ch < R) {
          double w0[CPT], w1[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) { w0[k] = 0.0; w1[k] = 0.0; }
#pragma unroll
          for (int i2 = ch * CH / 2; i2 < (ch + 1) * CH / 2; ++i2) {
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
      for (int k = 0; k < CPT; ++k) w[k] = act[k] ? w[k] * tau : 0.0;   // visited columns are final
#pragma unroll
      for (int ch = 0; ch < NCHK; ++ch) {
        if (TPC * CH * (ch + 1) > K && TPC * CH * ch < R) {
#pragma unroll
          for (int i2 = ch * CH / 2; i2 < (ch + 1) * CH / 2; ++i2) {
            const double2 v = vj[TPC * i2];
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              x[k][2 * i2] -= w[k] * v.x;
              x[k][2 * i2 + 1] -= w[k] * v.y;
            }
          }
        }
      }
    }
    if (tid == 0) fl += 4.0 * (R - K) * (C - c - 1) + 3.0 * (R - K);
    ++K;
    if (K == R) break;        // the column space is all of R^R: the remaining columns are final as they are
    if (K == kmax && c + 1 < C) {
      // rank limit reached: one pass decides whether anything right of c still sticks out of the span
      int big = 0;
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int i2 = 0; i2 < RPT / 2; ++i2) {
          const int r0 = 2 * TPC * i2 + 2 * part;
          if (r0 >= K) n0 += x[k][2 * i2] * x[k][2 * i2];
          if (r0 + 1 >= K) n1 += x[k][2 * i2 + 1] * x[k][2 * i2 + 1];
        }
        double nn = n0 + n1;
#pragma unroll
        for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
        if (act[k] && nn > thr2) big = 1;
      }
      more = __syncthreads_or(big);
      break;
    }
  }
  __syncthreads();
  if (!more) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int col = col0 + k * NCOL;
      const int lim = (col < C) ? s.perm[col] : 0;      // first row that is not an R entry of the column
      const bool acc = (col < C) && s.order[col] != 0;
      const double beta = (col < C) ? s.sig[col] : 0.0;
#pragma unroll
      for (int idx = 0; idx < RPT; ++idx) {
        int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
        if (col < C && r < K) s.X[r * LD + col] = (r < lim) ? x[k][idx] : ((acc && r == lim) ? beta : 0.0);
      }
    }
  }
  if (tid == 0) {
    s.iscr[0] = K; s.iscr[1] = more;
    s.cnt.flops_qr += fl;
    s.cnt.n_qr_reflectors += K;
  }
  __syncthreads();
}

#endif
}  // namespace mdopt
