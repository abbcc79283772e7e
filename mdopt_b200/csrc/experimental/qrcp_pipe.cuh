// EXPERIMENT (not built by default): flag-synchronised, batch-pivoted variant of qrcp_reg.
// Correct (all GPU parity tests pass with it) but measured SLOWER than the two-barrier version on B200
// (5.1k vs 4.7k cycles per reflector at d=5/chi=64): the bound is the issue/FP64 throughput of the
// rank-1 update itself (~2.8k cycles per apply call with 4 warps per scheduler), not the CTA barriers.
// Kept for round 2, when a blocked (DMMA) trailing update makes synchronisation the bottleneck again.
// To try it: include this header after block_ops.cuh and call qrcp_pipe<CHI> in Engine::factor.
#pragma once
#include "../block_ops.cuh"
namespace mdopt {
// -------------------------------------------------------------------------------------------------
// Pipelined variant of qrcp_reg: pivots are chosen GB at a time (the GB largest remaining column norms at
// the start of the group), and inside a group there is NO CTA barrier.  The owner of each pivot column
// publishes its reflector (raw column + tau + scal) through shared memory and a volatile flag; every warp
// applies published reflectors to its own columns at its own pace.  The critical path per reflector is the
// owner warp's "apply previous reflector to my columns, build the next one", not two CTA-wide barriers.
// A pre-selected pivot whose column has meanwhile dropped below the rank threshold is skipped (it does not
// consume a reflector), so the result is a rank-revealing Householder QR with the same stop rule as above;
// only the pivot ORDER differs (norms are refreshed once per group).
//   out: s.X[k*LD + c] coefficients, s.P2 = scaled reflectors Vall[K][N2], s.tau; s.iscr[0] = K, s.iscr[1] = more
// -------------------------------------------------------------------------------------------------
template <int CHI>
__device__ __noinline__ void qrcp_pipe(Smem<CHI>& s, const double* orig, int R, int C, int kmax, double rel_tol) {
  constexpr int LD = Smem<CHI>::LD;
  constexpr int N2 = Smem<CHI>::N2;
  constexpr int RPT = N2 / 4;               // register rows per thread
  constexpr int CH = RPT < 8 ? RPT : 8;     // register rows per chunk
  constexpr int NCHK = RPT / CH;            // chunk c covers matrix rows [4*CH*c, 4*CH*(c+1))
  constexpr int GB = 8;                     // pivots per group
  constexpr int KPL = (N2 + 31) / 32;       // column keys per lane
  enum { F_PENDING = 0, F_READY = 1, F_SKIP = 2, F_MORE = 3 };
  // register slot idx = 2*i2 + e holds matrix row 8*i2 + 2*part + e (pairs of rows -> 16-byte LDS)
  const int tid = threadIdx.x;
  const int col = tid >> 2, part = tid & 3, lane = tid & 31;
  const bool col_ok = col < C;
  double* Vall = s.P2;                      // raw pivot columns while factoring, scaled reflectors on exit
  double* scal_arr = s.vn2;
  unsigned long long* colkey = reinterpret_cast<unsigned long long*>(s.vn1);  // one key per column
  volatile int* flags = reinterpret_cast<volatile int*>(s.flag);             // [2][GB] group-parity double buffer
  auto norm_key = [](double nn, int c) -> unsigned long long {
    return ((unsigned long long)__double_as_longlong(nn) & ~0xFFull) | (unsigned long long)(0xFF - c);
  };
  double x[RPT];
#pragma unroll
  for (int idx = 0; idx < RPT; ++idx) {
    int r = 8 * (idx >> 1) + 2 * part + (idx & 1);
    x[idx] = (col_ok && r < R) ? orig[(size_t)r * C + col] : 0.0;
  }
  double nn_me;   // squared norm of the not-yet-eliminated part of this thread group's column
  double a_me;    // its entry in the next pivot row (valid in all 4 lanes of the group)
  {
    double nn = 0.0;
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) nn += x[idx] * x[idx];
    nn += __shfl_xor_sync(0xffffffffu, nn, 1);
    nn += __shfl_xor_sync(0xffffffffu, nn, 2);
    nn_me = nn;
    a_me = __shfl_sync(0xffffffffu, x[0], lane & ~3);
    if (part == 0) colkey[col] = col_ok ? norm_key(nn_me, col) : 0ull;
    if (tid < 2 * GB) flags[tid] = F_PENDING;
  }
  __syncthreads();
  double thresh2 = -1.0;
  const int jmax = min(R, C);
  int j = 0, more = 0, grp = 0;
  bool open = col_ok, finished = false;
  while (!finished) {
    // ---- the GB largest column keys, selected redundantly (and identically) by every warp ----
    unsigned long long mykeys[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) mykeys[q] = (lane + 32 * q < N2) ? colkey[lane + 32 * q] : 0ull;
    int mypiv = -1;   // lane g keeps the g-th pivot of the group
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
      if (thresh2 < 0.0) thresh2 = bnn * rel_tol * rel_tol;   // first key of the first group = largest norm^2
      if (lane == g) mypiv = (bnn > thresh2) ? 0xFF - (int)(best & 0xFFull) : -1;
#pragma unroll
      for (int q = 0; q < KPL; ++q)
        if (mykeys[q] == best) mykeys[q] = 0ull;   // keys are unique (they embed the column index)
    }
    volatile int* fl = flags + (grp & 1) * GB;
    if (tid < GB) flags[((grp + 1) & 1) * GB + tid] = F_PENDING;  // arm the other parity for the next group
#pragma unroll 1
    for (int g = 0; g < GB && !finished; ++g) {
      const int c = __shfl_sync(0xffffffffu, mypiv, g);
      if (c < 0 || j >= jmax) { finished = true; break; }
      const bool owner = (col == c);
      if (owner) {
        // ---- build reflector j from this column (already updated through reflector j-1) ----
        int st = F_READY;
        if (!(nn_me > thresh2)) st = F_SKIP;          // became rank deficient since the selection
        else if (j >= kmax) st = F_MORE;              // a significant direction beyond kmax
        if (st == F_READY) {
          const double a = a_me;
          const double beta = -copysign(sqrt(nn_me), a);
          const double d = a - beta;
          const double inv = 1.0 / (beta * d);
#pragma unroll
          for (int i2 = 0; i2 < RPT / 2; ++i2)
            *reinterpret_cast<double2*>(Vall + j * N2 + 8 * i2 + 2 * part) = make_double2(x[2 * i2], x[2 * i2 + 1]);
          if (part == 0) {
            s.tau[j] = -d * d * inv;     // (beta - a) / beta
            scal_arr[j] = beta * inv;    // 1 / (a - beta)
          }
          // own column: R[0:j+1, j] stays in registers, zero below (what makes closed columns invariant)
          const int jslot = ((j >> 3) << 1) | (j & 1), jpart = (j >> 1) & 3;
#pragma unroll
          for (int idx = 0; idx < RPT; ++idx) {
            const int r = 8 * (idx >> 1) + 2 * part + (idx & 1);
            if (idx == jslot && part == jpart) x[idx] = beta; else if (r > j) x[idx] = 0.0;
          }
          open = false;
        } else if (st == F_SKIP) {
          open = false;                               // numerically dependent column: never a pivot
        }
        __threadfence_block();
        __syncwarp(0xFu << (lane & ~3));
        if (part == 0) fl[g] = st;
      }
      // ---- wait for the reflector (or the skip / stop notice) ----
      int st;
      while ((st = fl[g]) == F_PENDING) __nanosleep(20);
      __threadfence_block();
      if (st == F_MORE) { more = 1; finished = true; break; }
      if (st == F_SKIP) continue;
      if (__any_sync(0xffffffffu, open)) {
        const double tau = s.tau[j], scal = scal_arr[j];
        const int jslot = ((j >> 3) << 1) | (j & 1), jpart = (j >> 1) & 3;
        const int nslot = (((j + 1) >> 3) << 1) | ((j + 1) & 1), npart = ((j + 1) >> 1) & 3;
        const double2* rc = reinterpret_cast<const double2*>(Vall + j * N2 + 2 * part);
        double w = 0.0, xj = 0.0;
#pragma unroll
        for (int cc = 0; cc < NCHK; ++cc) {
          if (4 * CH * (cc + 1) > j && 4 * CH * cc < R) {
            double w0 = 0.0, w1 = 0.0;
            if (4 * CH * cc > j) {
#pragma unroll
              for (int i2 = cc * CH / 2; i2 < (cc + 1) * CH / 2; ++i2) {
                const double2 v = rc[4 * i2];
                w0 += v.x * x[2 * i2];
                w1 += v.y * x[2 * i2 + 1];
              }
            } else {
#pragma unroll
              for (int i2 = cc * CH / 2; i2 < (cc + 1) * CH / 2; ++i2) {
                const int r0 = 8 * i2 + 2 * part;
                const double2 v = rc[4 * i2];
                if (r0 > j) w0 += v.x * x[2 * i2];
                if (r0 + 1 > j) w1 += v.y * x[2 * i2 + 1];
                if (2 * i2 == jslot) xj = x[2 * i2];
                if (2 * i2 + 1 == jslot) xj = x[2 * i2 + 1];
              }
            }
            w += w0 + w1;
          }
        }
        w += __shfl_xor_sync(0xffffffffu, w, 1);
        w += __shfl_xor_sync(0xffffffffu, w, 2);
        xj = __shfl_sync(0xffffffffu, xj, (lane & ~3) | jpart);
        w = owner ? 0.0 : tau * (xj + scal * w);   // tau * v^T x;  closed columns give exactly 0
        const double ws = w * scal;
        double nn = 0.0, an = 0.0;
#pragma unroll
        for (int cc = 0; cc < NCHK; ++cc) {
          if (4 * CH * (cc + 1) > j && 4 * CH * cc < R) {
            double s0 = 0.0, s1 = 0.0;
            if (4 * CH * cc > j) {
#pragma unroll
              for (int i2 = cc * CH / 2; i2 < (cc + 1) * CH / 2; ++i2) {
                const double2 v = rc[4 * i2];
                const double xa = x[2 * i2] - ws * v.x, xb = x[2 * i2 + 1] - ws * v.y;
                x[2 * i2] = xa; x[2 * i2 + 1] = xb;
                s0 += xa * xa; s1 += xb * xb;
              }
            } else {
#pragma unroll
              for (int i2 = cc * CH / 2; i2 < (cc + 1) * CH / 2; ++i2) {
                const int r0 = 8 * i2 + 2 * part;
                const double2 v = rc[4 * i2];
                double xa = x[2 * i2], xb = x[2 * i2 + 1];
                if (r0 > j) { xa -= ws * v.x; s0 += xa * xa; } else if (r0 == j) xa -= w;
                if (r0 + 1 > j) { xb -= ws * v.y; s1 += xb * xb; } else if (r0 + 1 == j) xb -= w;
                x[2 * i2] = xa; x[2 * i2 + 1] = xb;
              }
            }
            nn += s0 + s1;
          }
          if (4 * CH * cc <= j + 1 && j + 1 < 4 * CH * (cc + 1)) {  // chunk holding the next pivot row
#pragma unroll
            for (int idx = cc * CH; idx < (cc + 1) * CH; ++idx)
              if (idx == nslot) an = x[idx];
          }
        }
        nn += __shfl_xor_sync(0xffffffffu, nn, 1);
        nn += __shfl_xor_sync(0xffffffffu, nn, 2);
        if (open) nn_me = nn;
        a_me = __shfl_sync(0xffffffffu, an, (lane & ~3) | npart);
      }
      ++j;
    }
    // ---- group barrier: publish the refreshed column keys for the next selection ----
    if (part == 0) colkey[col] = open ? norm_key(nn_me, col) : 0ull;
    __syncthreads();
    ++grp;
  }
  const int K = j;
  if (!more) {
#pragma unroll
    for (int idx = 0; idx < RPT; ++idx) {
      int r = 8 * (idx >> 1) + 2 * part + (idx & 1);
      if (col_ok && r < K) s.X[r * LD + col] = x[idx];
    }
    // raw pivot columns -> scaled reflectors v_j = (0.., 1, scal_j * raw[j+1:]) for form_q_reg
    for (int t = tid; t < K * N2; t += blockDim.x) {
      const int jj = t / N2, r = t - jj * N2;
      const double raw = Vall[t];
      Vall[t] = (r > jj) ? raw * scal_arr[jj] : ((r == jj) ? 1.0 : 0.0);
    }
  }
  if (tid == 0) {
    s.iscr[0] = K; s.iscr[1] = more;
    double f = 2.0 * R * C;
    for (int jj = 0; jj < K; ++jj) f += 4.0 * (R - jj) * (C - jj - 1) + 3.0 * (R - jj);
    s.cnt.flops_qr += f;
    s.cnt.n_qr_reflectors += K;
  }
  __syncthreads();
}

}  // namespace mdopt
