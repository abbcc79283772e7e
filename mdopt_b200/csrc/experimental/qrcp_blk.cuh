// EXPERIMENT (not built by default): panel-blocked variant of qrcp_reg (GB pivots per panel, flag chain
// between the pivot columns, one two-pass bulk update per panel with a Gram-based forward substitution).
// Correct (all 28 GPU tests pass with it) but SLOWER on B200 at GB = 4: 5.7k cycles per reflector against
// 3.2k for qrcp_reg -- while the pivot columns hand over, only 8 of 512 threads work, and that serial chain
// costs more than the barriers and shuffle rounds it saves.  Kept for round 2 (a tensor-core bulk update
// would change the balance).  To try it: include after block_ops.cuh and call qrcp_blk<CHI, 8, CPT, 4>.
#pragma once
#include "../block_ops.cuh"
namespace mdopt {
// -------------------------------------------------------------------------------------------------
// Blocked variant of qrcp_reg.  Pivots are chosen GB at a time (the GB largest remaining column norms).
//   panel phase  only the GB pivot columns work: column g applies the panel's earlier reflectors to
//                itself, re-checks its norm against the rank threshold, builds its reflector and the
//                Gram entries v_g . v_a with the earlier ones; columns hand over through shared-memory flags.
//   bulk phase   every other open column applies the whole panel in TWO passes over its registers:
//                W_a = v_a . x for all a at once, forward substitution z_a = tau_a (W_a - sum_{b<a} G_ab z_b),
//                x -= sum_a z_a v_a, and the new squared norm in the same pass.
// Per reflector this costs 2/GB CTA barriers and 2/GB shuffle reductions instead of 2 and 2.
// A pre-selected pivot that turns out rank deficient after the panel's earlier reflectors is dropped.
// Output convention identical to qrcp_reg.
// -------------------------------------------------------------------------------------------------
template <int CHI, int TPC, int CPT, int GB>
__device__ __noinline__ void qrcp_blk(Smem<CHI>& s, const double* orig, int R, int C, int kmax, double rel_tol) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int NT = (CHI * 8 < 64) ? 64 : CHI * 8;
  constexpr int NCOL = NT / TPC;
  constexpr int RPT = N2 / TPC;             // register rows per thread and column
  constexpr int KPL = (N2 + 31) / 32;       // column keys per lane in the selection
  enum { F_READY = 0, F_SKIP = 1, F_MORE = 2 };
  // register slot idx = 2*i2 + e holds matrix row 2*TPC*i2 + 2*part + e
  const int tid = threadIdx.x;
  const int col0 = tid / TPC, part = tid % TPC, lane = tid & 31;
  const unsigned gmask = ((TPC == 32) ? 0xffffffffu : ((1u << TPC) - 1u)) << (lane & ~(TPC - 1));
  double* Vall = s.P2;
  double* Gm = s.rot;                       // Gm[a * GB + b] = v_a . v_b (b < a) of the current panel
  unsigned long long* colkey = reinterpret_cast<unsigned long long*>(s.vn1);  // one key per column
  volatile int* flags = reinterpret_cast<volatile int*>(s.flag);             // [2][GB], by panel parity
  auto norm_key = [](double nn, int c) -> unsigned long long {
    return ((unsigned long long)__double_as_longlong(nn) & ~0xFFull) | (unsigned long long)(0xFF - c);
  };
  double x[CPT][RPT];
  double nn_me[CPT];
  bool open[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int col = col0 + k * NCOL;
    open[k] = col < C;
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) {
      int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
      x[k][idx] = (open[k] && r < R) ? orig[(size_t)r * C + col] : 0.0;
    }
    double nn = 0.0;
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) nn += x[k][idx] * x[k][idx];
#pragma unroll
    for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(0xffffffffu, nn, off);
    nn_me[k] = nn;
    if (part == 0 && col < N2) colkey[col] = open[k] ? norm_key(nn, col) : 0ull;
  }
  if (CPT * NCOL < N2) {  // columns no thread group owns (narrow variant): never pivots
    for (int c = CPT * NCOL + tid; c < N2; c += NT) colkey[c] = 0ull;
  }
  if (tid < 2 * GB) flags[tid] = 0;
  __syncthreads();
  double thresh2 = -1.0;
  int j0 = 0, more = 0, grp = 0;
  bool finished = false;
  while (!finished) {
    // ---- the GB largest column keys, selected redundantly (and identically) by every warp ----
    unsigned long long mykeys[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) mykeys[q] = (lane + 32 * q < N2) ? colkey[lane + 32 * q] : 0ull;
    int mypiv = -1;   // lane g keeps the g-th pivot candidate of the panel
#pragma unroll
    for (int g = 0; g < GB; ++g) {
      unsigned long long best = 0ull;
#pragma unroll
      for (int q = 0; q < KPL; ++q) best = mykeys[q] > best ? mykeys[q] : best;
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
        best = o > best ? o : best;
      }
      const double bnn = best ? __longlong_as_double((long long)(best & ~0xFFull)) : -1.0;
      if (thresh2 < 0.0) thresh2 = bnn * rel_tol * rel_tol;
      if (lane == g) mypiv = (bnn > thresh2) ? 0xFF - (int)(best & 0xFFull) : -1;
#pragma unroll
      for (int q = 0; q < KPL; ++q)
        if (mykeys[q] == best) mykeys[q] = 0ull;
    }
    volatile int* fl = flags + (grp & 1) * GB;
    if (tid < GB) flags[((grp + 1) & 1) * GB + tid] = 0;   // arm the other parity for the next panel
    const int ncand = __popc(__ballot_sync(0xffffffffu, lane < GB && mypiv >= 0));
    if (ncand == 0) finished = true;
    // ================= panel phase: only the owners of the candidates work =================
    bool in_panel[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) in_panel[k] = false;
#pragma unroll 1
    for (int g = 0; g < ncand; ++g) {
      const int c = __shfl_sync(0xffffffffu, mypiv, g);
      const int ko = c / NCOL;
      if (col0 + ko * NCOL != c) continue;
      // I own candidate g.  Wait for candidate g-1 (its flag carries the number of accepted reflectors).
      int acc = 0;
      if (g > 0) {
        int f;
        while ((f = fl[g - 1]) == 0) __nanosleep(20);
        __threadfence_block();
        acc = (f - 1) >> 2;
        if (((f - 1) & 3) == F_MORE) {   // propagate the stop
          __syncwarp(gmask);
          if (part == 0) fl[g] = f;
          continue;
        }
      }
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        if (k != ko) continue;
        in_panel[k] = true;
        const int j = j0 + acc;
        // apply the panel's earlier reflectors j0 .. j-1 to my column
        for (int a = 0; a < acc; ++a) {
          const double2* va = reinterpret_cast<const double2*>(Vall + (j0 + a) * N2 + 2 * part);
          double w = 0.0;
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2) {
            const double2 v = va[TPC * i2];
            w += v.x * x[k][2 * i2] + v.y * x[k][2 * i2 + 1];
          }
#pragma unroll
          for (int off = 1; off < TPC; off <<= 1) w += __shfl_xor_sync(gmask, w, off);
          w *= s.tau[j0 + a];
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2) {
            const double2 v = va[TPC * i2];
            x[k][2 * i2] -= w * v.x;
            x[k][2 * i2 + 1] -= w * v.y;
          }
        }
        // remaining squared norm (rows >= j) and the pivot-row entry
        const int jslot = ((j / (2 * TPC)) << 1) | (j & 1), jpart = (j >> 1) % TPC;
        double nn = 0.0, a_piv = 0.0;
#pragma unroll
        for (int idx = 0; idx < RPT; ++idx) {
          const int r = 2 * TPC * (idx >> 1) + 2 * part + (idx & 1);
          if (r >= j) nn += x[k][idx] * x[k][idx];
          if (idx == jslot) a_piv = x[k][idx];
        }
#pragma unroll
        for (int off = 1; off < TPC; off <<= 1) nn += __shfl_xor_sync(gmask, nn, off);
        a_piv = __shfl_sync(gmask, a_piv, (lane & ~(TPC - 1)) | jpart);
        int st = F_READY;
        if (!(nn > thresh2) || j >= min(R, C)) st = F_SKIP;   // rank deficient after all: never a pivot
        else if (j >= kmax) st = F_MORE;                      // a significant direction beyond kmax
        open[k] = false;
        if (st == F_READY) {
          const double beta = -copysign(sqrt(nn), a_piv);
          const double d = a_piv - beta;
          const double inv = 1.0 / (beta * d);
          const double tau = -d * d * inv;   // (beta - a) / beta
          const double scal = beta * inv;    // 1 / (a - beta)
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2) {
            const int r0 = 2 * TPC * i2 + 2 * part;
            double2 v;
            v.x = (r0 > j) ? x[k][2 * i2] * scal : ((r0 == j) ? 1.0 : 0.0);
            v.y = (r0 + 1 > j) ? x[k][2 * i2 + 1] * scal : ((r0 + 1 == j) ? 1.0 : 0.0);
            *reinterpret_cast<double2*>(Vall + j * N2 + r0) = v;
            if (r0 == j) x[k][2 * i2] = beta; else if (r0 > j) x[k][2 * i2] = 0.0;
            if (r0 + 1 == j) x[k][2 * i2 + 1] = beta; else if (r0 + 1 > j) x[k][2 * i2 + 1] = 0.0;
            // keep v in place of x for the Gram products below?  no: recompute from shared memory
          }
          if (part == 0) s.tau[j] = tau;
          __syncwarp(gmask);
          // Gram entries with the panel's earlier reflectors: G[acc][a] = v_j . v_{j0+a}
          const double2* vme = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
          for (int a = 0; a < acc; ++a) {
            const double2* va = reinterpret_cast<const double2*>(Vall + (j0 + a) * N2 + 2 * part);
            double gsum = 0.0;
#pragma unroll
            for (int i2 = 0; i2 < RPT / 2; ++i2) {
              const double2 u = vme[TPC * i2], v = va[TPC * i2];
              gsum += u.x * v.x + u.y * v.y;
            }
#pragma unroll
            for (int off = 1; off < TPC; off <<= 1) gsum += __shfl_xor_sync(gmask, gsum, off);
            if (part == 0) Gm[acc * GB + a] = gsum;
          }
          acc += 1;
        }
        __threadfence_block();
        __syncwarp(gmask);
        if (part == 0) fl[g] = 1 + (acc << 2) + st;
      }
    }
    __syncthreads();
    // ================= bulk phase: everybody else applies the whole panel =================
    int nb = 0;
    if (ncand > 0) {
      const int f = fl[ncand - 1];
      nb = (f - 1) >> 2;
      if (((f - 1) & 3) == F_MORE) { more = 1; finished = true; }
    }
    if (nb > 0 && !more) {
      bool any_open = false;
#pragma unroll
      for (int k = 0; k < CPT; ++k) any_open |= (open[k] && !in_panel[k]);
      if (__any_sync(0xffffffffu, any_open)) {
        double W[GB][CPT];
#pragma unroll
        for (int a = 0; a < GB; ++a)
#pragma unroll
          for (int k = 0; k < CPT; ++k) W[a][k] = 0.0;
#pragma unroll
        for (int i2 = 0; i2 < RPT / 2; ++i2) {
#pragma unroll
          for (int a = 0; a < GB; ++a) {
            if (a < nb) {
              const double2 v = *reinterpret_cast<const double2*>(Vall + (j0 + a) * N2 + 2 * TPC * i2 + 2 * part);
#pragma unroll
              for (int k = 0; k < CPT; ++k) W[a][k] += v.x * x[k][2 * i2] + v.y * x[k][2 * i2 + 1];
            }
          }
        }
#pragma unroll
        for (int off = 1; off < TPC; off <<= 1) {
#pragma unroll
          for (int a = 0; a < GB; ++a)
#pragma unroll
            for (int k = 0; k < CPT; ++k) W[a][k] += __shfl_xor_sync(0xffffffffu, W[a][k], off);
        }
        // forward substitution: z_a = tau_a (W_a - sum_{b<a} G[a][b] z_b)
#pragma unroll
        for (int a = 0; a < GB; ++a) {
          if (a < nb) {
            const double tau = s.tau[j0 + a];
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              double acc2 = W[a][k];
#pragma unroll
              for (int b = 0; b < a; ++b) acc2 -= Gm[a * GB + b] * W[b][k];
              W[a][k] = (open[k] && !in_panel[k]) ? tau * acc2 : 0.0;
            }
          }
        }
        double nn[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) nn[k] = 0.0;
        const int jlast = j0 + nb - 1;
#pragma unroll
        for (int i2 = 0; i2 < RPT / 2; ++i2) {
          double xa[CPT], xb[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k) { xa[k] = x[k][2 * i2]; xb[k] = x[k][2 * i2 + 1]; }
#pragma unroll
          for (int a = 0; a < GB; ++a) {
            if (a < nb) {
              const double2 v = *reinterpret_cast<const double2*>(Vall + (j0 + a) * N2 + 2 * TPC * i2 + 2 * part);
#pragma unroll
              for (int k = 0; k < CPT; ++k) { xa[k] -= W[a][k] * v.x; xb[k] -= W[a][k] * v.y; }
            }
          }
          const int r0 = 2 * TPC * i2 + 2 * part;
#pragma unroll
          for (int k = 0; k < CPT; ++k) {
            x[k][2 * i2] = xa[k]; x[k][2 * i2 + 1] = xb[k];
            if (r0 > jlast) nn[k] += xa[k] * xa[k];
            if (r0 + 1 > jlast) nn[k] += xb[k] * xb[k];
          }
        }
#pragma unroll
        for (int off = 1; off < TPC; off <<= 1) {
#pragma unroll
          for (int k = 0; k < CPT; ++k) nn[k] += __shfl_xor_sync(0xffffffffu, nn[k], off);
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k)
          if (open[k] && !in_panel[k]) nn_me[k] = nn[k];
      }
    }
    j0 += nb;
    if (j0 >= min(R, C)) finished = true;
    // ---- publish the refreshed keys for the next selection ----
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int col = col0 + k * NCOL;
      if (part == 0 && col < N2) colkey[col] = open[k] ? norm_key(nn_me[k], col) : 0ull;
    }
    __syncthreads();
    ++grp;
  }
  const int K = j0;
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

}  // namespace mdopt
