// EXPERIMENT (not built): software-pipelined QRCP with LAPACK-style norm down-dating.
// Measured on B200 (d=5, chi=64, batch 2072): 770 syndromes/s, 3586 cycles per reflector, against 874 /
// 3120 for the exact-norm version in block_ops.cuh.  Correct (GPU parity tests pass), but the exact-norm
// recomputations run in divergent 8-lane groups and the owner's chain grows by its own update and norm, which
// costs more than the removed norm pass saves.  Kept for the record; see profiles/r01_qr_experiments.md.
#pragma once
#include "../block_ops.cuh"
namespace mdopt {
#ifndef MDOPT_EMU
// -------------------------------------------------------------------------------------------------
// qrcp_reg: same factorisation, software-pipelined.  Column norms are DOWN-DATED (LAPACK dlaqp2: nn -= r^2
// with r the new R-factor entry, exact recomputation when the running value has lost sqrt(eps) of the last
// exact one), so the pivot for reflector j+1 is known as soon as the dot products with v_j are reduced --
// before the rank-1 update itself.  Per reflector:
//   phase 1  dots w = tau v_j^T x, norm down-date, per-warp key                      | barrier A
//   phase 2  pivot owner: own update, exact norm, v_{j+1}  ||  everyone else: x -= w v_j | barrier B
// The owner's serial section (sqrt, reciprocal, scaling) overlaps the other warps' updates, there is no
// separate norm pass and no "row > j" masking outside the owner / recompute paths.
// -------------------------------------------------------------------------------------------------
template <int CHI, int TPC, int CPT>
__device__ __noinline__ void qrcp_reg_dd(Smem<CHI>& s, const double* orig, int R, int C, int kmax, double rel_tol) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int NT = (CHI * 8 < 64) ? 64 : CHI * 8;
  constexpr int NCOL = NT / TPC;            // columns covered by one pass of the thread groups
  constexpr int RPT = N2 / TPC;             // register rows per thread and column
  static_assert(RPT <= 16, "the pivot-slot switch has 16 cases");
  constexpr int NP = RPT / 2;               // register pairs; pair i2 = matrix rows [2*TPC*i2, 2*TPC*(i2+1))
  constexpr int CHP = NP < 4 ? NP : 4;      // pairs per skip chunk
  constexpr int NCHK = NP / CHP;
  constexpr int CROWS = 2 * TPC * CHP;      // matrix rows per chunk
  constexpr int NWARP = NT / 32;
  const int tid = threadIdx.x;
  const int col0 = tid / TPC, part = tid % TPC, lane = tid & 31;
  const int gbase = lane & ~(TPC - 1);
  const unsigned gmask = ((TPC == 32) ? 0xffffffffu : ((1u << TPC) - 1u)) << gbase;
  double* Vall = s.P2;
  unsigned* keys = reinterpret_cast<unsigned*>(s.vn1);  // one key per warp
  auto norm_key = [](double nn, int c) -> unsigned {
    return ((unsigned)__double2hiint(nn) & ~0xFFu) | (unsigned)(0xFF - c);
  };
  auto warp_key_publish = [&](unsigned k) {
    k = __reduce_max_sync(0xffffffffu, k);
    if (lane == 0) keys[tid >> 5] = k;
  };
  double x[CPT][RPT];
  double nn_me[CPT];   // running SQUARED norm of the not-yet-eliminated part of each owned column
  double nn_ref[CPT];  // its value at the last exact computation
  double w[CPT];       // tau * v^T x of the reflector whose update is still pending
  bool open[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int col = col0 + k * NCOL;
    open[k] = col < C;
    w[k] = 0.0;
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
      nn_me[k] = nn; nn_ref[k] = nn;
      const unsigned kk = open[k] ? norm_key(nn, col0 + k * NCOL) : 0u;
      kmax_key = kk > kmax_key ? kk : kmax_key;
    }
    warp_key_publish(kmax_key);
  }
  __syncthreads();
  // x[k] -= w[k] * v for every owned column (closed columns carry w == 0); chunks above the reflector's
  // pivot row (v == 0) or past the last matrix row are skipped
#define MD_QR_AXPY(VP, JROW)                                                          \
  _Pragma("unroll") for (int c = 0; c < NCHK; ++c) {                                  \
    if (CROWS * (c + 1) > (JROW) && CROWS * c < R) {                                  \
      _Pragma("unroll") for (int i2 = c * CHP; i2 < (c + 1) * CHP; ++i2) {            \
        const double2 v = (VP)[TPC * i2];                                             \
        _Pragma("unroll") for (int k = 0; k < CPT; ++k) {                             \
          x[k][2 * i2] -= w[k] * v.x;                                                 \
          x[k][2 * i2 + 1] -= w[k] * v.y;                                             \
        }                                                                             \
      }                                                                               \
    }                                                                                 \
  }
  // a[k] = x[k][slot] for a warp-uniform register slot: one indexed jump instead of RPT selects
#define MD_QR_SLOT(I)                                                                 \
  case I:                                                                             \
    if constexpr ((I) < RPT) {                                                        \
      _Pragma("unroll") for (int k = 0; k < CPT; ++k) a[k] = x[k][(I) < RPT ? (I) : 0]; \
    }                                                                                 \
    break;
#define MD_QR_PICK(SLOT)                                                              \
  switch (SLOT) {                                                                     \
    MD_QR_SLOT(0) MD_QR_SLOT(1) MD_QR_SLOT(2) MD_QR_SLOT(3) MD_QR_SLOT(4) MD_QR_SLOT(5) MD_QR_SLOT(6)   \
    MD_QR_SLOT(7) MD_QR_SLOT(8) MD_QR_SLOT(9) MD_QR_SLOT(10) MD_QR_SLOT(11) MD_QR_SLOT(12)              \
    MD_QR_SLOT(13) MD_QR_SLOT(14) MD_QR_SLOT(15)                                      \
    default: break;                                                                   \
  }
  double thresh = 0.0;
  const int jmax = min(R, C);
  int K = 0, more = 0;
  for (int j = 0;; ++j) {
    // ---- pivot for reflector j: max over the per-warp keys, redundantly per warp ----
    unsigned key = (lane < NWARP) ? keys[lane] : 0u;
    key = __reduce_max_sync(0xffffffffu, key);
    const double best = key ? __hiloint2double((int)(key & ~0xFFu), 0) : -1.0;  // rounded down by < 2.5e-4
    const int bi = 0xFF - (int)(key & 0xFFu);
    if (j == 0) thresh = best * rel_tol * rel_tol;  // keys carry squared norms
    const bool stop = (j >= jmax) || !(best > thresh);
    const bool trunc = !stop && (j >= kmax);
    const bool last = stop || trunc;
    const double2* vp = reinterpret_cast<const double2*>(Vall + (j > 0 ? j - 1 : 0) * N2 + 2 * part);
    const int jslot = ((j / (2 * TPC)) << 1) | (j & 1), jpart = (j >> 1) % TPC;
    const int ko = (!last && bi >= NCOL) ? bi / NCOL : 0;
    const bool owner = !last && (col0 + ko * NCOL == bi);
    if (owner) {
      // pending update of the pivot column, then its exact remaining norm, then v_j and tau_j
      if (j > 0) { MD_QR_AXPY(vp, j - 1) }
      double a[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) a[k] = 0.0;
      MD_QR_PICK(jslot)
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        if (k == ko) {
          double nn = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < NP; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            if (r0 >= j) nn += x[k][2 * i2] * x[k][2 * i2];
            if (r0 + 1 >= j) nn += x[k][2 * i2 + 1] * x[k][2 * i2 + 1];
          }
#pragma unroll
          for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(gmask, nn, off);
          const double aj = __shfl_sync(gmask, a[k], gbase | jpart);
          double beta = -copysign(sqrt(nn), aj);
          double d = aj - beta;
          double tau = 0.0, scal = 0.0;
          if (nn > 0.0) {
            const double inv = 1.0 / (beta * d);
            tau = -d * d * inv;   // (beta - a) / beta
            scal = beta * inv;    // 1 / (a - beta)
          } else {
            beta = aj;            // nothing left to eliminate: H = I
          }
#pragma unroll
          for (int i2 = 0; i2 < NP; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            double2 v;
            v.x = (r0 > j) ? x[k][2 * i2] * scal : ((r0 == j) ? 1.0 : 0.0);
            v.y = (r0 + 1 > j) ? x[k][2 * i2 + 1] * scal : ((r0 + 1 == j) ? 1.0 : 0.0);
            *reinterpret_cast<double2*>(Vall + j * N2 + r0) = v;
            if (r0 == j) x[k][2 * i2] = beta; else if (r0 > j) x[k][2 * i2] = 0.0;
            if (r0 + 1 == j) x[k][2 * i2 + 1] = beta; else if (r0 + 1 > j) x[k][2 * i2 + 1] = 0.0;
          }
          open[k] = false;
          if (part == 0) s.tau[j] = tau;
        }
      }
      // the owner group's columns are up to date now (its other columns went through the same update)
#pragma unroll
      for (int k = 0; k < CPT; ++k) w[k] = 0.0;
    }
    if (j > 0) {
      bool pend = false;
#pragma unroll
      for (int k = 0; k < CPT; ++k) pend |= (w[k] != 0.0);
      if (__any_sync(0xffffffffu, pend)) { MD_QR_AXPY(vp, j - 1) }
    }
    if (last) { more = trunc ? 1 : 0; break; }
    __syncthreads();  // B: v_j, tau_j visible; every column carries H_0 .. H_{j-1}
    bool any_open = false;
#pragma unroll
    for (int k = 0; k < CPT; ++k) any_open |= open[k];
    if (__any_sync(0xffffffffu, any_open)) {
      const double tau = s.tau[j];
      const double2* vj = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
#pragma unroll
      for (int k = 0; k < CPT; ++k) w[k] = 0.0;
#pragma unroll
      for (int c = 0; c < NCHK; ++c) {
        if (CROWS * (c + 1) > j && CROWS * c < R) {
          double w0[CPT], w1[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) { w0[k] = 0.0; w1[k] = 0.0; }
#pragma unroll
          for (int i2 = c * CHP; i2 < (c + 1) * CHP; ++i2) {
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
      double a[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) a[k] = 0.0;
      MD_QR_PICK(jslot)
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const double xj = __shfl_sync(0xffffffffu, a[k], gbase | jpart);
        w[k] = open[k] ? w[k] * tau : 0.0;
        const double r = xj - w[k];            // new R-factor entry (v_j[j] == 1)
        double nn = nn_me[k] - r * r;
        if (open[k] && nn_ref[k] > 0.0 && !(nn > TOL3Z * nn_ref[k])) {
          // the down-dated value has lost its accuracy: apply the update now and recompute exactly
#pragma unroll
          for (int i2 = 0; i2 < NP; ++i2) {
            const double2 v = vj[TPC * i2];
            x[k][2 * i2] -= w[k] * v.x;
            x[k][2 * i2 + 1] -= w[k] * v.y;
          }
          double ex = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < NP; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            if (r0 > j) ex += x[k][2 * i2] * x[k][2 * i2];
            if (r0 + 1 > j) ex += x[k][2 * i2 + 1] * x[k][2 * i2 + 1];
          }
#pragma unroll
          for (int off = 1; off < TPC; off <<= 1) ex += __shfl_xor_sync(gmask, ex, off);
          nn = ex; nn_ref[k] = ex; w[k] = 0.0;
        }
        nn_me[k] = nn > 0.0 ? nn : 0.0;
        // a column at or below the stop threshold can never become a pivot again (norms only shrink):
        // retire it, its pending update (if any) is still applied through w[k]
        if (!(nn_me[k] > thresh)) open[k] = false;
      }
    }
    {
      unsigned kmax_key = 0u;
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const unsigned kk = open[k] ? norm_key(nn_me[k], col0 + k * NCOL) : 0u;
        kmax_key = kk > kmax_key ? kk : kmax_key;
      }
      warp_key_publish(kmax_key);
    }
    __syncthreads();  // A: keys visible
    K = j + 1;
  }
#undef MD_QR_AXPY
#undef MD_QR_SLOT
#undef MD_QR_PICK
  if (!more) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int col = col0 + k * NCOL;
#pragma unroll
      for (int idx = 0; idx < RPT; ++idx) {
        int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
        if (col < C && r < K) s.X[r * LD + col] = x[k][idx];
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

#endif
}  // namespace mdopt
