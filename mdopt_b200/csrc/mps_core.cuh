// Per-CTA MPS engine: one CTA decodes one syndrome, start to finish.
//
// Replaces, for the batched decode path, the reference's
//   mdopt/contractor/contractor.py:177-378   mps_mpo_contract (zip-up sweep)
//   mdopt/mps/canonical.py:345-420           move_orth_centre
//   mdopt/optimiser/utils.py:284-345         the per-string loop of apply_constraints
//   mdopt/mps/canonical.py:906-991 + 241-272 marginal + dense
//   examples/decoding/decoding.py:1362-1379  tolerant arg-max
//
// Formulation.  The reference builds theta = M . A_{c+1} . W_{c+1} and splits it.
// Whenever the MPO site tensor W, read as a matrix (vL,pU) -> (vR,pD), has orthonormal rows (XOR_BULK,
// SWAP, IDENTITY, XOR_LEFT/COPY_LEFT never occur in that position), theta = Mmat . T with T an isometry,
// so the split of theta is the split of the carried matrix Mmat (2K x w*chi) followed by M' . T.  The
// engine therefore factors Mmat (128 x 128 at chi=64 instead of 128 x 256) and applies the next site
// afterwards.  For non-isometric W (XOR_RIGHT at the end of every string) theta is formed explicitly.
// Orth-centre moves are the same step with an identity MPO, run in a mirrored frame when moving left.
//
// Factorisation modes:
//   MODE_SVD   every split is an SVD with the reference's truncation rule k = min(chi_max, #(s > cut))
//              (utils.py:119); moves use cut = 1e-12 (canonical.py:399-405).  Before any rotation the split checks
//              whether its columns are already mutually orthogonal (probe check, then Gram check): then the SVD is
//              X = (X / sigma) sigma with sigma the column norms, the coefficient matrix is diagonal and the next
//              carried matrix a scaled gather of the next site tensor (layout 3, st.in_x).  With the bit-flip bias
//              this holds for every split of a decode; otherwise one-sided Jacobi sweeps run.
//   MODE_SVD_GRAM  the same with the deterministic Gram check only.
//   MODE_FAST  rank-revealing Householder QR with column pivoting; only when more than chi_max
//              columns carry weight above rank_tol * ||.|| does the step fall back to the Jacobi SVD.
//              Non-truncating steps only re-gauge the bond, which leaves the state unchanged.
#pragma once
#include "block_ops.cuh"

namespace mdopt {


// Per-CTA engine.  All persistent state lives in shared memory (s.st) so that it costs no registers in
// the register-resident QR loops; the Engine object itself is just the reference to the CTA's Smem.
template <int CHI>
struct Engine {
  static constexpr int LD = Smem<CHI>::LD;
  static constexpr int N2 = Smem<CHI>::N2;
  static constexpr int SITE_STRIDE = CHI * 2 * CHI;
#ifndef MDOPT_EMU
  // register-resident QR / Jacobi routines (block_ops.cuh) exist for the shared-memory capacities only
  static constexpr bool REG_PATH = (CHI <= SMEM_CHI_MAX);
#else
  static constexpr bool REG_PATH = false;
#endif

  Smem<CHI>& s;
  EngineState& st;

  MD_DEV Engine(Smem<CHI>& sm) : s(sm), st(sm.st) {}

  // ---- frame helpers: logical site j of the current frame ----
  MD_DEV int phys(int j) const { return st.mirrored ? (st.P.n_sites - 1 - j) : j; }
  MD_DEV int dim_left(int j) const { int p = phys(j); return st.mirrored ? st.bd[p + 1] : st.bd[p]; }
  MD_DEV int dim_right(int j) const { int p = phys(j); return st.mirrored ? st.bd[p] : st.bd[p + 1]; }
  MD_DEV void set_dim_right(int j, int v) { int p = phys(j); if (st.mirrored) st.bd[p] = v; else st.bd[p + 1] = v; }

  MD_DEV void load_w(int tid) {
    MD_PAR_FOR(i, 16) { s.W[i] = (tid < 0) ? ((i == 0 || i == 3) ? 1.0 : 0.0) : st.wtab[tid * 16 + i]; }
    MD_SYNC();
  }
  MD_DEV int w_vl(int tid) const { return tid < 0 ? 1 : st.wdim[tid * 4 + 0]; }
  MD_DEV int w_vr(int tid) const { return tid < 0 ? 1 : st.wdim[tid * 4 + 1]; }
  MD_DEV int w_iso(int tid) const { return tid < 0 ? 1 : st.wdim[tid * 4 + 2]; }
  // W(q, w', pU, pD) with actual dims (vl, vr, 2, 2)
  MD_DEV double w_at(int vr, int q, int wo, int pu, int pd) const { return s.W[((q * vr + wo) * 2 + pu) * 2 + pd]; }

  // ---- carried <- site j with the first MPO tensor applied (contractor.py:276-291, first half) ----
  // out[((k*2 + p')*st.Wq + q)*st.Mr + m] = sum_p A[k,p,m] W(0,q,p,p')
  MD_DEV void carried_from_site(int j, int tid) {
    load_w(tid);
    const int A = dim_left(j), B = dim_right(j), vr = w_vr(tid);
    const double* src = st.sites + (size_t)phys(j) * SITE_STRIDE;
    const int mir = st.mirrored;
    MD_PAR_FOR(t, A * 2 * vr * B) {
      int m = t % B; int r = t / B;
      int q = r % vr; r /= vr;
      int pd = r & 1; int k = r >> 1;
      double a0, a1;
      if (!mir) { a0 = src[(k * 2 + 0) * B + m]; a1 = src[(k * 2 + 1) * B + m]; }
      else { a0 = src[(m * 2 + 0) * A + k]; a1 = src[(m * 2 + 1) * A + k]; }
      st.scr0[t] = a0 * w_at(vr, 0, q, 0, pd) + a1 * w_at(vr, 0, q, 1, pd);
    }
    MD_LEADER { s.cnt.bytes_hbm += 8.0 * (A * 2 * B + A * 2 * vr * B); s.cnt.flops_apply += 3.0 * A * 2 * vr * B; }
    MD_SYNC();
    MD_LEADER { st.K = A; st.Wq = vr; st.Mr = B; st.in_x = 0; }
    MD_SYNC();
  }

  // ---- stage site j (frame layout (a,p,b) flat) into dst (shared) ----
  MD_DEV void stage_site(int j, double* dst, int A, int B) {
    const double* src = st.sites + (size_t)phys(j) * SITE_STRIDE;
    if (!st.mirrored) { MD_PAR_FOR(g, A * 2 * B) dst[g] = src[g]; }
    else {
      // phys tensor has dims (B, 2, A): a transpose.  A warp moves a 4 (a) x 8 (b) patch: the global reads are
      // eight fully used 32-byte sectors, the shared-memory writes collide at most 4-way (a plain linear walk
      // over the source would be 16-way when 2B is a multiple of 16)
      const int na = (A + 3) / 4, nbt = (B + 7) / 8;
      MD_PAR_FOR(g, na * nbt * 2 * 32) {
        const int l = g & 31; int t = g >> 5;
        const int p = t & 1; t >>= 1;
        const int ta = t % na, tb = t / na;
        const int a = 4 * ta + (l & 3), b = 8 * tb + (l >> 2);
        if (a < A && b < B) dst[(a * 2 + p) * B + b] = src[((size_t)b * 2 + p) * A + a];
      }
    }
    MD_LEADER { s.cnt.bytes_hbm += 8.0 * A * 2 * B; }
    MD_SYNC();
  }

  // ---- the same copy issued asynchronously (cp.async) so that it overlaps the factorisation; pair with
  //      stage_wait() before the destination is read or reused ----
  MD_DEV void stage_site_async(int j, double* dst, int A, int B) {
    const double* src = st.sites + (size_t)phys(j) * SITE_STRIDE;
    if (!st.mirrored) { MD_PAR_FOR(g, A * 2 * B) md_async_copy8(dst + g, src + g); }
    else {
      const int na = (A + 3) / 4, nbt = (B + 7) / 8;
      MD_PAR_FOR(g, na * nbt * 2 * 32) {
        const int l = g & 31; int t = g >> 5;
        const int p = t & 1; t >>= 1;
        const int ta = t % na, tb = t / na;
        const int a = 4 * ta + (l & 3), b = 8 * tb + (l >> 2);
        if (a < A && b < B) md_async_copy8(dst + (a * 2 + p) * B + b, src + ((size_t)b * 2 + p) * A + a);
      }
    }
    MD_LEADER { s.cnt.bytes_hbm += 8.0 * A * 2 * B; }
  }
  MD_DEV void stage_wait() { md_async_wait(); MD_SYNC(); }

  // ---- out[((row*2+p')*vr + w')*Bn + m'] = sum_{q,p,m} Min[row, q*st.Mr+m] W(q,w',p,p') As[(m*2+p)*Bn + m'] ----
  MD_DEV void apply_site(const double* Min, int ld, int rows, int vl, int vr, int mr, int Bn,
                         const double* As, double* out) {
    constexpr int TR = 8, TM = 2;
    const int nrt = (rows + TR - 1) / TR, nmt = (Bn + TM - 1) / TM;
    const int ncomb = 2 * vr;
    MD_PAR_FOR(t, nrt * ncomb * nmt) {
      int mt = t % nmt; int r = t / nmt;
      int comb = r % ncomb; int rt = r / ncomb;
      int wo = comb % vr, pd = comb / vr;
      int row0 = rt * TR, m0 = mt * TM;
      double acc[TR][TM];
#pragma unroll
      for (int a = 0; a < TR; ++a)
#pragma unroll
        for (int b = 0; b < TM; ++b) acc[a][b] = 0.0;
      for (int q = 0; q < vl; ++q) {
        for (int pu = 0; pu < 2; ++pu) {
          const double w = w_at(vr, q, wo, pu, pd);
          if (w == 0.0) continue;
          // unscaled partial product of this (q, pU) term; the MPO entry multiplies it once at the end
          double part[TR][TM];
#pragma unroll
          for (int a = 0; a < TR; ++a)
#pragma unroll
            for (int b = 0; b < TM; ++b) part[a][b] = 0.0;
          const double* mrow = Min + row0 * ld + q * mr;
          const double* arow = As + pu * Bn + m0;
          if (row0 + TR <= rows && m0 + TM <= Bn && (Bn & 1) == 0) {
            // interior tile, even row length: the two site-tensor entries come as one 16-byte load (adjacent
            // lanes then read adjacent 16-byte words instead of two stride-2 8-byte streams)
            static_assert(TM == 2, "the vector load below assumes two columns per thread");
            for (int m = 0; m < mr; ++m) {
              const double2 av = *reinterpret_cast<const double2*>(arow + (size_t)m * 2 * Bn);
#pragma unroll
              for (int a = 0; a < TR; ++a) {
                const double x = mrow[a * ld + m];
                part[a][0] += x * av.x;
                part[a][1] += x * av.y;
              }
            }
          } else if (row0 + TR <= rows && m0 + TM <= Bn) {  // interior tile: no bounds checks in the inner loop
            for (int m = 0; m < mr; ++m) {
              double av[TM];
#pragma unroll
              for (int b = 0; b < TM; ++b) av[b] = arow[(size_t)m * 2 * Bn + b];
#pragma unroll
              for (int a = 0; a < TR; ++a) {
                const double x = mrow[a * ld + m];
#pragma unroll
                for (int b = 0; b < TM; ++b) part[a][b] += x * av[b];
              }
            }
          } else {
            for (int m = 0; m < mr; ++m) {
              double av[TM];
#pragma unroll
              for (int b = 0; b < TM; ++b) av[b] = (m0 + b < Bn) ? arow[(size_t)m * 2 * Bn + b] : 0.0;
#pragma unroll
              for (int a = 0; a < TR; ++a) {
                const double x = (row0 + a < rows) ? mrow[a * ld + m] : 0.0;
#pragma unroll
                for (int b = 0; b < TM; ++b) part[a][b] += x * av[b];
              }
            }
          }
#pragma unroll
          for (int a = 0; a < TR; ++a)
#pragma unroll
            for (int b = 0; b < TM; ++b) acc[a][b] += w * part[a][b];
        }
      }
#pragma unroll
      for (int a = 0; a < TR; ++a)
#pragma unroll
        for (int b = 0; b < TM; ++b)
          if (row0 + a < rows && m0 + b < Bn)
            out[(((size_t)(row0 + a) * 2 + pd) * vr + wo) * Bn + m0 + b] = acc[a][b];
    }
    MD_LEADER {
      int nnz = 0;
      for (int i = 0; i < vl * vr * 4; ++i) nnz += (s.W[i] != 0.0);
      s.cnt.flops_apply += 2.0 * rows * Bn * mr * nnz;
      s.cnt.bytes_hbm += 8.0 * rows * 2 * vr * Bn;
    }
    MD_SYNC();
  }

  // ---- apply_site for a DIAGONAL coefficient matrix (orthogonal-columns fast path of the SVD mode): row k of
  //      Min has its only entry in column s.order[k] = q*mr + m, so the contraction is a gather:
  //      out[((k*2+p')*vr + w')*Bn + b] = Min[k, c_k] sum_p W(q,w',p,p') As[(m*2+p)*Bn + b] ----
  //      The coefficient of row k is scale * s.sig[s.order[k]]; the result goes straight into s.X (ld LD), where
  //      the next split looks for it (st.in_x), so the carried matrix never leaves the SM on this path.
  // pairs = 0: M' = diag(sigma) in the sorted order (layout 3).  pairs = 1 (layout 4): the rows of rotated columns
  // carry a second entry at the partner column, M'[k, c] = sigma_c J[c, c], M'[k, partner] = sigma_c J[partner, c].
  MD_DEV void apply_site_diag(double scale, int rows, int vl, int vr, int mr, int Bn, const double* As, int pairs) {
    const int Cn = vr * Bn;
    // one 32-lane group per output row segment (k, p', w'): the index arithmetic is done once per segment and
    // the lanes run along b (coalesced reads of the staged site, conflict-free writes of X)
    MD_PAR_FOR(t, rows * 2 * vr * 32) {
      const int l = t & 31; int r = t >> 5;
      // vl, vr <= 2: no integer divisions
      const int wo = (vr == 2) ? (r & 1) : 0; r = (vr == 2) ? (r >> 1) : r;
      const int pd = r & 1, k = r >> 1;
      const int c = s.order[k];
      const int c2 = pairs ? s.perm[c] : -1;
      const double g = scale * s.sig[c];
      const double ga = (c2 >= 0) ? g * s.vn1[c] : g;
      const int q = (c >= mr) ? 1 : 0, m = c - q * mr;
      const double f0 = ga * w_at(vr, q, wo, 0, pd), f1 = ga * w_at(vr, q, wo, 1, pd);
      const double* a0 = As + (m * 2 + 0) * Bn;
      const double* a1 = As + (m * 2 + 1) * Bn;
      double* dst = s.X + (k * 2 + pd) * LD + wo * Bn;
      if (c2 < 0) {
        for (int b = l; b < Bn; b += 32) dst[b] = f0 * a0[b] + f1 * a1[b];
      } else {
        const double gb = g * s.wv[c];
        const int q2 = (c2 >= mr) ? 1 : 0, m2 = c2 - q2 * mr;
        const double h0 = gb * w_at(vr, q2, wo, 0, pd), h1 = gb * w_at(vr, q2, wo, 1, pd);
        const double* b0 = As + (m2 * 2 + 0) * Bn;
        const double* b1 = As + (m2 * 2 + 1) * Bn;
        for (int b = l; b < Bn; b += 32) dst[b] = (f0 * a0[b] + f1 * a1[b]) + (h0 * b0[b] + h1 * b1[b]);
      }
    }
    MD_LEADER { s.cnt.flops_apply += 4.0 * rows * 2 * Cn; st.in_x = 1; }
    MD_SYNC();
  }

  // entry (k, c) of the coefficient matrix of layouts 3 / 4 (see apply_site_diag)
  MD_DEV double sparse_coef(int lay, int k, int c) const {
    const int a = s.order[k];
    const int b = (lay == 4) ? s.perm[a] : -1;
    if (c == a) return (b >= 0) ? s.sig[a] * s.vn1[a] : s.sig[a];
    if (b >= 0 && c == b) return s.sig[a] * s.wv[a];
    return 0.0;
  }

  MD_DEV void load_x(const double* src, int R, int C) {
    MD_PAR_FOR(t, R * C) { int i = t / C, c = t - i * C; s.X[i * LD + c] = src[t]; }
    MD_LEADER { s.cnt.bytes_hbm += 8.0 * R * C; }
    MD_SYNC();
  }

  // ---- factor X (R x C; a copy lives in `orig`, flat ld C) -> Knew, U (in X), coefficients (in P2) ----
  // Returns Knew.  s.iscr[7] = 1 when U columns are addressed through s.order / scaled by 1/s.sig.
  // Layout of the result (s.iscr[7]):
  //   0  U = s.X[:, 0:st.K] (ld LD),               coefficients = s.P2 (ld LD)     [portable QR path]
  //   1  U = s.X[:, order[k]] / sig[order[k]],   coefficients = s.P2 (ld LD)     [Jacobi SVD path]
  //   2  U = s.P2 flat (ld st.K),                   coefficients = s.X rows (ld LD) [register QR path]
  MD_DEV int factor(const double* orig, int R, int C, int chi_lim, double cut, int is_move) {
    int use_svd = (st.P.mode != MODE_FAST);
    int Knew = 0;
    const int kmax = chi_lim < CHI ? chi_lim : CHI;
    if (!use_svd) {
      long long t0 = MD_CLOCK();
#ifndef MDOPT_EMU
      if constexpr (REG_PATH) {
        // 8 lanes per column; a thread group owns one column for C <= CHI (every centre move) and two
        // (c, c + CHI) for wider matrices
        if (C <= CHI) qrcp_reg<CHI, 8, 1>(s, orig, R, C, kmax, st.P.rank_tol);
        else qrcp_reg<CHI, 8, 2>(s, orig, R, C, kmax, st.P.rank_tol);
      } else
#endif
      {
        load_x(orig, R, C);
        qrcp<CHI>(s, R, C, kmax, st.P.rank_tol);
      }
      MD_LEADER { s.cnt.cyc_qr += (double)(MD_CLOCK() - t0); }
      Knew = s.iscr[0];
      if (s.iscr[1]) {  // more than chi_lim significant directions: genuine truncation needed
        use_svd = 1;   // (the truncation rule counts the step as truncating)
      }
    }
    if (!use_svd) {
      const int lay = REG_PATH ? 2 : 0;
      if (Knew == 0) {  // zero matrix: keep a single zero direction (the state has zero norm)
        if (lay == 2) {
          MD_PAR_FOR(i, R) s.P2[i] = (i == 0) ? 1.0 : 0.0;
          MD_PAR_FOR(c, C) s.X[c] = 0.0;
        } else {
          MD_PAR_FOR(i, R) s.X[i * LD] = (i == 0) ? 1.0 : 0.0;
          MD_PAR_FOR(c, C) s.P2[c] = 0.0;
        }
        MD_LEADER { s.iscr[7] = lay; }
        MD_SYNC();
        return 1;
      }
      long long t1 = MD_CLOCK();
#ifndef MDOPT_EMU
      if constexpr (REG_PATH) {
        form_q_reg<CHI>(s, R, Knew);
      } else
#endif
      {
        extract_coeff<CHI>(s, Knew, C);
        form_q<CHI>(s, R, Knew);
      }
      MD_LEADER { s.iscr[7] = lay; s.cnt.cyc_formq += (double)(MD_CLOCK() - t1); }
      MD_SYNC();
      return Knew;
    }
    // the orthogonal fast path leaves the carried matrix in s.X (st.in_x); then orig (scr0) is stale and is only
    // refreshed from X if the rotations have to run after all
    const int x_loaded = (orig == st.scr0) && st.in_x;
    if (!x_loaded) load_x(orig, R, C);
    // ---- Jacobi SVD path ----
    long long tj = MD_CLOCK();
#ifndef MDOPT_EMU
    // SVD mode tries the orthogonality checks first (see jacobi_columns_dev); the truncating fallback of
    // MODE_FAST comes from a QR gauge and goes straight to the sweeps
    if constexpr (REG_PATH)
      jacobi_columns_dev<CHI>(s, R, C, 60, st.P.mode == MODE_SVD ? 1 : (st.P.mode == MODE_SVD_GRAM ? 2 : 0),
                              x_loaded ? st.scr0 : nullptr);
    else
#endif
    {
      // portable path (host emulation, large-chi tier): the same probe, then the sweeps if it fails
      int ortho = 0;
      // N2 x N2 doubles the split does not read: the flat scratch matrix that is not `orig`
      double* pair_table = (orig == st.scr1) ? st.scr0 : st.scr1;
#ifdef MD_JPROF   // profiling build (scripts/build_variant.sh NAME -DMD_JPROF): probe / pair search / sweeps split of this
                  // phase, reported through the QR counters that the SVD modes leave at zero
      long long tq0 = MD_CLOCK(), tq1 = tq0, tq2 = tq0;
#endif
      if (st.P.mode != MODE_FAST && C >= 2) {
        ortho = ortho_probe<CHI>(s, R, C);
#ifdef MD_JPROF
        tq1 = MD_CLOCK();
#endif
        // failed on a few columns that pair up: one round of disjoint rotations settles the split (layout 4)
#ifdef MDOPT_EMU
        if (!ortho && !emu_pair_fix_off() && pair_fix<CHI>(s, R, C, pair_table)) ortho = 2;
#else
        if (!ortho && pair_fix<CHI>(s, R, C, pair_table)) ortho = 2;
#endif
      }
#ifdef MD_JPROF
      tq2 = MD_CLOCK();
      MD_LEADER { s.cnt.cyc_qr += (double)(tq1 - tq0); s.cnt.cyc_formq += (double)(tq2 - tq1); if (!ortho) s.cnt.n_qr_reflectors += 1.0; }
#endif
      MD_LEADER { s.iscr[9] = ortho; }
      if (!ortho) {
        if (x_loaded) {  // the rotations are about to change X: refresh the flat copy the coefficient product reads
          MD_PAR_FOR(t, R * C) { int r = t / C, c = t - r * C; st.scr0[t] = s.X[r * LD + c]; }
          MD_SYNC();
        }
        // a few rounds of rotations restricted to the offending pairs, each followed by the probe; the cyclic sweeps
        // (C - 1 rounds over every pair, at least twice) only if that does not settle
        int settled = 0;
        int rounds = (st.P.mode != MODE_FAST && C >= 2) ? 12 : 0;   // the probe ran: s.flag holds its verdict per column
#ifdef MDOPT_EMU
        if (emu_pair_fix_off() == 1) rounds = 0;   // 2: only the single-round shortcut is off, so every rotating split comes here
#endif
        for (int it = 0; it < rounds && !settled; ++it) {
          if (!pair_fix<CHI>(s, R, C, pair_table, 1)) break;
          settled = ortho_probe<CHI>(s, R, C);
        }
        if (!settled) jacobi_columns<CHI>(s, R, C);
#ifdef MD_JPROF
        MD_LEADER { s.cnt.cyc_pad += (double)(MD_CLOCK() - tq2); }
#endif
      }
      MD_SYNC();
    }
    MD_LEADER { s.cnt.cyc_jacobi += (double)(MD_CLOCK() - tj); }
    long long tc = MD_CLOCK();
    if (s.iscr[10]) stage_wait();   // the prefetched site tensor has landed in s.P2 (see step())
    if (s.iscr[5]) st.status = ST_NOT_CONVERGED;
#ifndef MDOPT_EMU
    if constexpr (REG_PATH) truncation_rule_dev<CHI>(s, C, chi_lim, cut);
    else
#endif
      truncation_rule<CHI>(s, C, chi_lim, cut);
    Knew = s.iscr[0];
    if (st.trace != nullptr) {   // uniform: the trace buffer is per launch
      if (st.n_traced < st.P.trace_max) {
        double* tr = st.trace + (size_t)st.n_traced * st.P.trace_stride;
        MD_LEADER { tr[0] = R; tr[1] = C; tr[2] = Knew; tr[3] = is_move; }
        MD_PAR_FOR(c, N2) { if (4 + c < st.P.trace_stride) tr[4 + c] = (c < C) ? s.sig[s.order[c]] : 0.0; }
      }
      MD_SYNC();
      MD_LEADER { st.n_traced += 1; }
      MD_SYNC();
    }
    if (Knew > CHI) { st.status = ST_CAPACITY; MD_SYNC(); return 0; }  // would overflow the site stride
    if (Knew == 0) {
      MD_SYNC();
      MD_PAR_FOR(i, R) s.X[i * LD] = (i == 0) ? 1.0 : 0.0;
      MD_PAR_FOR(c, C) s.P2[c] = 0.0;
      MD_LEADER { s.iscr[7] = 0; }
      MD_SYNC();
      return 1;
    }
    MD_PAR_FOR(k, Knew) s.tau[k] = 1.0 / s.sig[s.order[k]];   // for U = X[:, order] / sigma (store_u)
    // coefficients M'[k, c] = (1/sigma_k) sum_i X[i, order[k]] orig[i, c]   (= sigma_k v_k^T)
    if (s.iscr[9]) {
      // no rotation happened: X == orig with orthogonal columns, so M' = diag(sigma) in the sorted order (the
      // off-diagonal dot products are below the check's threshold).  It is not materialised: step() reads
      // s.sig / s.order directly (layout 3).  One round of disjoint pair rotations (s.iscr[9] == 2): M' =
      // sigma J^T has at most two entries per row, read from s.vn1 / s.wv / s.perm (layout 4).
      const int lay34 = (s.iscr[9] == 2) ? 4 : 3;
      MD_LEADER { s.iscr[7] = lay34; }
    } else {
      constexpr int TK = 8;
      const int nkt = (Knew + TK - 1) / TK;
      MD_PAR_FOR(t, nkt * C) {
        int c = t % C, kt = t / C;
        int k0 = kt * TK;
        double acc[TK];
        int col[TK];
#pragma unroll
        for (int a = 0; a < TK; ++a) { acc[a] = 0.0; col[a] = s.order[min(k0 + a, Knew - 1)]; }
        for (int i = 0; i < R; ++i) {
          double xo = orig[(size_t)i * C + c];
#pragma unroll
          for (int a = 0; a < TK; ++a) acc[a] += s.X[i * LD + col[a]] * xo;
        }
#pragma unroll
        for (int a = 0; a < TK; ++a)
          if (k0 + a < Knew) s.P2[(k0 + a) * LD + c] = acc[a] / s.sig[col[a]];
      }
      MD_LEADER { s.cnt.flops_apply += 2.0 * R * C * Knew; s.cnt.bytes_hbm += 8.0 * R * C * nkt; s.iscr[7] = 1; }
    }
    MD_SYNC();
    MD_LEADER { s.cnt.cyc_coef += (double)(MD_CLOCK() - tc); }
    return Knew;
  }

  // ---- site j <- U (R = st.K*2 rows, Knew columns), frame layout (k, r, k') ----
  MD_DEV double u_at(int lay, int row, int kn, int Knew) const {
    if (lay == 2) return s.P2[row * Knew + (kn ^ q_swizzle(row, Knew))];
    if (lay == 1 || lay >= 3) return s.X[row * LD + s.order[kn]] * s.tau[kn];   // s.tau[kn] = 1 / sigma_kn
    return s.X[row * LD + kn];
  }
  MD_DEV void store_u(int j, int Kl, int Knew) {
    double* dst = st.sites + (size_t)phys(j) * SITE_STRIDE;
    const int lay = s.iscr[7];
    // 32-lane groups per output row, lanes along the fastest output index (no per-element divisions)
    if (!st.mirrored) {
      MD_PAR_FOR(t, Kl * 2 * 32) {
        const int l = t & 31, row = t >> 5;
        for (int kn = l; kn < Knew; kn += 32) dst[(size_t)row * Knew + kn] = u_at(lay, row, kn, Knew);
      }
    } else {  // phys tensor dims (Knew, 2, Kl)
      MD_PAR_FOR(t, Knew * 2 * 32) {
        const int l = t & 31; const int r = t >> 5; const int p = r & 1, kn = r >> 1;
        for (int k = l; k < Kl; k += 32) dst[(size_t)r * Kl + k] = u_at(lay, k * 2 + p, kn, Knew);
      }
    }
    MD_LEADER { s.cnt.bytes_hbm += 8.0 * Kl * 2 * Knew; }
    MD_SYNC();
  }

  // ---- one two-site step: absorb site jn (frame index) with MPO tensor tid ----
  MD_DEV void step(int jn, int tid, int chi_lim, double cut, int is_move, int renorm_sv) {
    const int vl = w_vl(tid), vr = w_vr(tid), iso = w_iso(tid);
    const int Bn = dim_right(jn);
    const int R = st.K * 2;
    int C = st.Wq * st.Mr;
    const double* orig = st.scr0;
    load_w(tid);
    if (vl != st.Wq || dim_left(jn) != st.Mr) { st.status = ST_CAPACITY; return; }
    if (!iso) {
      // theta = Mmat . T formed explicitly (contractor.py:328-342), R x (2*vr*Bn)
      if (2 * vr * Bn > N2) { st.status = ST_CAPACITY; return; }
      if (!st.in_x) load_x(st.scr0, R, C);
      stage_site(jn, s.P2, st.Mr, Bn);
      apply_site(s.X, LD, R, vl, vr, st.Mr, Bn, s.P2, st.scr1);
      C = 2 * vr * Bn;
      orig = st.scr1;
    }
    // SVD modes on the shared-memory capacities: the next site tensor is fetched into s.P2 while the split runs
    // (the orthogonal fast path does not touch P2); factor() waits for it before anything else writes P2
    const int pre = REG_PATH && iso && st.P.mode != MODE_FAST;
    if (pre) stage_site_async(jn, s.P2, st.Mr, Bn);
    MD_LEADER { s.iscr[10] = pre; }
    int Knew = factor(orig, R, C, chi_lim, cut, is_move);
    if (st.status == ST_CAPACITY) { if (pre) stage_wait(); return; }
    const int lay = s.iscr[7];
    double* coef = (lay == 2) ? s.X : s.P2;                 // coefficients M' (Knew x C, ld LD); layout 3: diagonal
    double* stage = (lay == 2) ? (s.X + CHI * LD) : ((lay >= 3) ? s.P2 : s.X);  // staging area for the next site
    // layouts 3 / 4 (orthogonal or pair-rotated fast path): M' is not materialised; its rows have norm sigma_k, so
    // ||s|| comes from the truncation rule
    const double dscale = (lay >= 3 && renorm_sv && s.dscr[3] > 0.0) ? 1.0 / s.dscr[3] : 1.0;
    if (renorm_sv && lay < 3) {  // mps_mpo_contract(renormalise=True): s /= ||s|| (utils.py:126-129); ||s|| = ||M'||_F
      MD_PAR_FOR(ch, 256) {
        double acc = 0.0;
        for (int t = ch; t < Knew * C; t += 256) { int k = t / C, c = t - k * C; double v = coef[k * LD + c]; acc += v * v; }
        s.part[ch] = acc;
      }
      MD_SYNC();
      MD_LEADER {
        double acc = 0.0;
        for (int ch = 0; ch < 256; ++ch) acc += s.part[ch];
        s.dscr[4] = (acc > 0.0) ? 1.0 / sqrt(acc) : 1.0;
      }
      MD_SYNC();
      const double inv = s.dscr[4];
      MD_PAR_FOR(t, Knew * C) { int k = t / C, c = t - k * C; coef[k * LD + c] *= inv; }
      MD_SYNC();
    }
    long long ts = MD_CLOCK();
    store_u(jn - 1, st.K, Knew);
    if (iso) {
      // the site tensor the carried matrix absorbs next: staged (transposed in the mirrored frame) -- except on the
      // capacities whose work matrices live in global memory anyway, where the unmirrored site is read in place
      const double* site = stage;
      if constexpr (!Smem<CHI>::MATS_IN_SMEM) {
        if (!st.mirrored) site = st.sites + (size_t)phys(jn) * SITE_STRIDE;
      }
      if (site == stage && !(pre && lay >= 3)) stage_site(jn, stage, st.Mr, Bn);   // layouts 3 / 4: already in s.P2 (prefetched)
      MD_LEADER { s.cnt.cyc_io += (double)(MD_CLOCK() - ts); }
      long long ta = MD_CLOCK();
      if (lay >= 3) apply_site_diag(dscale, Knew, vl, vr, st.Mr, Bn, site, lay == 4);   // result stays in s.X (st.in_x)
      else {
        apply_site(coef, LD, Knew, vl, vr, st.Mr, Bn, site, st.scr0);
        MD_LEADER { st.in_x = 0; }
      }
      MD_LEADER { s.cnt.cyc_apply += (double)(MD_CLOCK() - ta); }
    } else {
      if (lay >= 3) {
        MD_PAR_FOR(t, Knew * C) { int k = t / C, c = t - k * C; st.scr0[t] = dscale * sparse_coef(lay, k, c); }
      } else {
        MD_PAR_FOR(t, Knew * C) { int k = t / C, c = t - k * C; st.scr0[t] = coef[k * LD + c]; }
      }
      MD_LEADER { s.cnt.bytes_hbm += 8.0 * Knew * C; st.in_x = 0; }
      MD_SYNC();
    }
    MD_LEADER { st.K = Knew; st.Wq = vr; st.Mr = Bn; s.cnt.n_steps += 1.0; set_dim_right(jn - 1, Knew); }
    MD_SYNC();
  }

  // ---- carried (st.K,2,1,st.Mr) -> site j, optionally divided by its Frobenius norm (optimiser/utils.py:340-345) ----
  MD_DEV void carried_to_site(int j, int renorm) {
    const int n = st.K * 2 * st.Mr;
    double scale = 1.0;
    // the carried matrix is flat in scr0, or (orthogonal fast path) in s.X with leading dimension LD
    const int inx = st.in_x, Bc = st.Mr;
    auto cv = [&](int t) -> double { return inx ? s.X[(t / Bc) * LD + (t % Bc)] : st.scr0[t]; };
    if (renorm) {
      MD_PAR_FOR(ch, 256) {
        double acc = 0.0;
        for (int t = ch; t < n; t += 256) { double v = cv(t); acc += v * v; }
        s.part[ch] = acc;
      }
      MD_SYNC();
      MD_LEADER {
        double acc = 0.0;
        for (int ch = 0; ch < 256; ++ch) acc += s.part[ch];
        double nrm = sqrt(acc);
        s.dscr[4] = (nrm > 0.0) ? 1.0 / nrm : 1.0;
      }
      MD_SYNC();
      scale = s.dscr[4];
    }
    double* dst = st.sites + (size_t)phys(j) * SITE_STRIDE;
    const int A = st.K, B = st.Mr;
    if (!st.mirrored) { MD_PAR_FOR(g, n) dst[g] = cv(g) * scale; }
    else {
      MD_PAR_FOR(g, n) {  // phys dims (B, 2, A)
        int a = g % A; int r = g / A; int p = r & 1; int b = r >> 1;
        dst[g] = cv((a * 2 + p) * B + b) * scale;
      }
    }
    MD_SYNC();
    MD_LEADER { s.cnt.bytes_hbm += 16.0 * n; st.in_x = 0; }
    MD_SYNC();
  }

  // ---- sweep of an MPO over frame st.sites start..start+len-1 (tids == nullptr: identity MPO = move) ----
  MD_PHASE void sweep(int start, int len, const int* tids, int chi_lim, double cut, int renorm_after,
                    int is_move, int renorm_sv) {
    carried_from_site(start, tids ? tids[0] : -1);
    int next_tid = (tids && len > 1) ? tids[1] : -1;   // the program lives in global memory: fetch one step ahead
    for (int t = 1; t < len; ++t) {
      const int cur = next_tid;
      if (tids && t + 1 < len) next_tid = tids[t + 1];
      step(start + t, cur, chi_lim, cut, is_move, renorm_sv);
      if (st.status == ST_CAPACITY) return;
    }
    if (st.Wq != 1) { st.status = ST_CAPACITY; return; }
    carried_to_site(start + len - 1, renorm_after);
  }

  // ---- move the orthogonality centre (canonical.py:345-420) ----
  MD_DEV void set_frame(int mir) { MD_SYNC(); MD_LEADER { st.mirrored = mir; } MD_SYNC(); }
  MD_DEV void set_centre(int c) { MD_SYNC(); MD_LEADER { st.centre = c; } MD_SYNC(); }

  MD_DEV void move_centre(int target) {
    const int from = st.centre;
    if (target == from) return;
    const int chi_lim = CHI;
    if (target > from) {
      set_frame(0);
      sweep(from, target - from + 1, nullptr, chi_lim, st.P.cut_move, 0, 1, 0);
    } else {
      set_frame(1);
      const int N = st.P.n_sites;
      sweep(N - 1 - from, from - target + 1, nullptr, chi_lim, st.P.cut_move, 0, 1, 0);
      set_frame(0);
    }
    set_centre(target);
  }

  // ---- two-site gate on (site, site+1) with the centre at `site`, then split (depolarising bias:
  //      decoding.py:254-274 -- theta = "ijk,klm,jlno->inom", split with chi=1e4 / cut=1e-12, spectrum and
  //      centre renormalised).  The gate sits in the tensor table as 16 doubles G[j][l][n][o]. ----
  MD_PHASE void gate2(int site, int gid, int renorm) {
    const int L = st.bd[site], M = st.bd[site + 1], Rr = st.bd[site + 2];
    const int R = 2 * L, C = 2 * Rr;
    if (R > N2 || C > N2 || M > CHI) { st.status = ST_CAPACITY; return; }
    load_w(gid);
    const double* A = st.sites + (size_t)site * SITE_STRIDE;
    const double* B = st.sites + (size_t)(site + 1) * SITE_STRIDE;
    double* theta = st.scr1;
    MD_PAR_FOR(t, R * C) {
      int m = t % Rr; int r = t / Rr;
      int o = r & 1; r >>= 1;
      int n = r & 1; int i = r >> 1;
      double acc = 0.0;
      for (int j = 0; j < 2; ++j)
        for (int l = 0; l < 2; ++l) {
          double g = s.W[((j * 2 + l) * 2 + n) * 2 + o];
          if (g == 0.0) continue;
          double inner = 0.0;
          for (int k = 0; k < M; ++k) inner += A[(i * 2 + j) * M + k] * B[((size_t)k * 2 + l) * Rr + m];
          acc += g * inner;
        }
      theta[t] = acc;
    }
    MD_LEADER { s.cnt.flops_apply += 8.0 * R * C * M; s.cnt.bytes_hbm += 8.0 * (L * 2 * M + M * 2 * Rr + R * C); }
    MD_SYNC();
    const int Knew = factor(theta, R, C, CHI, st.P.cut_move, 1);
    if (st.status == ST_CAPACITY) return;
    int lay = s.iscr[7];
    if (lay >= 3) {
      // orthogonal / pair-rotated fast path of the SVD modes: the coefficient matrix is not materialised by
      // factor(); this op reads it as a matrix, so write it out (layout 1: coefficients in s.P2)
      MD_PAR_FOR(t, Knew * C) { int k = t / C, c = t - k * C; s.P2[k * LD + c] = sparse_coef(lay, k, c); }
      MD_LEADER { s.iscr[7] = 1; }
      MD_SYNC();
      lay = 1;
    }
    const double* coef = (lay == 2) ? s.X : s.P2;
    double scale = 1.0;
    if (renorm) {
      MD_PAR_FOR(ch, 256) {
        double acc = 0.0;
        for (int t = ch; t < Knew * C; t += 256) { int k = t / C, c = t - k * C; double v = coef[k * LD + c]; acc += v * v; }
        s.part[ch] = acc;
      }
      MD_SYNC();
      MD_LEADER {
        double acc = 0.0;
        for (int ch = 0; ch < 256; ++ch) acc += s.part[ch];
        s.dscr[4] = 1.0 / sqrt(acc);  // unguarded like the reference (0 -> inf -> NaN)
      }
      MD_SYNC();
      scale = s.dscr[4];
    }
    store_u(site, L, Knew);
    double* dst = st.sites + (size_t)(site + 1) * SITE_STRIDE;
    MD_PAR_FOR(t, Knew * C) { int k = t / C, c = t - k * C; dst[t] = coef[k * LD + c] * scale; }
    MD_SYNC();
    MD_LEADER { st.bd[site + 1] = Knew; st.centre = site + 1; s.cnt.n_steps += 1.0; }
    MD_SYNC();
  }

  // ---- product state + bit-flip bias (mps/utils.py:439-512, decoding.py:140-190) ----
  MD_PHASE void init_product(const uint8_t* sym) {
    const int N = st.P.n_sites;
    const double d = (st.P.bias_p >= 0.0) ? sqrt(1.0 - st.P.bias_p) : 1.0;
    const double o = (st.P.bias_p >= 0.0) ? sqrt(st.P.bias_p) : 0.0;
    const int kL = st.P.n_logical;
    MD_PAR_FOR(j, N) {
      double a0, a1;
      int c = sym[j];
      if (c == 0) { a0 = 1.0; a1 = 0.0; }
      else if (c == 1) { a0 = 0.0; a1 = 1.0; }
      else { a0 = 1.0 / sqrt(2.0); a1 = a0; }
      if (j >= kL && st.P.bias_p >= 0.0) {
        double b0 = a0 * d + a1 * o, b1 = a0 * o + a1 * d;
        a0 = b0; a1 = b1;
      }
      st.sites[(size_t)j * SITE_STRIDE + 0] = a0;
      st.sites[(size_t)j * SITE_STRIDE + 1] = a1;
    }
    MD_PAR_FOR(j, N + 1) st.bd[j] = 1;
    MD_SYNC();
    MD_LEADER { st.centre = 0; st.mirrored = 0; st.status = ST_OK; st.n_traced = 0; }
    MD_SYNC();
  }

  // ---- marginal over st.sites kL..N-1, reverse, dense, |.|, L2 renorm, tolerant arg-max ----
  //   canonical.py:956-989, :150-161, :241-272; decoding.py:1362-1379
  MD_PHASE void marginal_readout(double* out, int* success) {
    // kL = number of leading sites handled by the dense stage; those with their bit in kept_mask set are kept
    // (n_logical of them), the others are traced out like the tail (erasures of qubits < n_logical only)
    const int N = st.P.n_sites, kL = st.P.n_head, ren = st.P.renormalise;
    const unsigned kept = st.P.kept_mask;
    double* T = s.part;        // current last tensor (L x 2), L <= CHI
    double* w = s.wv;          // traced vector
    int L = st.bd[N - 1];
    MD_PAR_FOR(t, L * 2) T[t] = st.sites[(size_t)(N - 1) * SITE_STRIDE + t];  // (L, 2, 1)
    MD_SYNC();
    for (int site = N - 1; site >= kL; --site) {
      MD_PAR_FOR(l, L) w[l] = (T[l * 2] + T[l * 2 + 1]) * (1.0 / sqrt(2.0));
      MD_SYNC();
      const int L2 = st.bd[site - 1];
      const double* A = st.sites + (size_t)(site - 1) * SITE_STRIDE;  // (L2, 2, L)
      MD_PAR_FOR(t, L2 * 2) {
        double acc = 0.0;
        for (int l = 0; l < L; ++l) acc += A[(size_t)t * L + l] * w[l];
        T[t] = acc;
      }
      MD_LEADER { s.cnt.bytes_hbm += 8.0 * L2 * 2 * L; s.cnt.flops_apply += 2.0 * L2 * 2 * L; }
      MD_SYNC();
      if (ren) {
        MD_LEADER {
          double acc = 0.0;
          for (int t = 0; t < L2 * 2; ++t) acc += T[t] * T[t];
          s.dscr[5] = sqrt(acc);
        }
        MD_SYNC();
        const double nrm = s.dscr[5];
        MD_PAR_FOR(t, L2 * 2) T[t] = T[t] / nrm;  // unguarded like the reference: 0/0 -> NaN
        MD_SYNC();
      }
      L = L2;
    }
    // the merged last logical tensor goes back into the MPS (what CanonicalMPS.marginal returns)
    MD_PAR_FOR(t, L * 2) st.sites[(size_t)(kL - 1) * SITE_STRIDE + t] = T[t];
    MD_LEADER { st.bd[kL] = 1; }
    MD_SYNC();
    // Dense stage over the kL leading sites, meeting in the middle so that up to 12 kept sites fit the shared
    // work pages (the reference's dense read-out limit, decoding.py:1358-1379): the left part accumulates
    // D[(kept bits of sites < jm), b], the right part E[b, (kept bits of sites >= jm)], and the result is
    // out[iL + nL * iR] = sum_b D[iL, b] E[b, iR] (later sites are more significant, as after reverse()).
    int nkept_total = 0;
    for (int j = 0; j < kL; ++j) nkept_total += (kept >> j) & 1u;
    int jm = 0;  // first site of the right part: the left part takes the first floor(nkept / 2) kept sites
    {
      int seen = 0;
      const int want = (nkept_total <= 6) ? nkept_total : nkept_total / 2;
      while (jm < kL && seen < want) { seen += (kept >> jm) & 1u; ++jm; }
      if (nkept_total <= 6) jm = kL;  // small read-outs stay one-sided (right part empty)
    }
    double* D0 = s.X;
    double* D1 = s.X + (N2 * LD) / 2;
    double* E0 = s.P2;
    double* E1 = s.P2 + (CHI * LD) / 2;
    const int page = (CHI * LD) / 2;  // doubles per page (the smaller of the two pairs)
    int nb = 1, nidx = 1;
    MD_LEADER { D0[0] = 1.0; }
    MD_SYNC();
    for (int j = 0; j < jm; ++j) {
      const int Bl = nb;
      const int Br = (j == kL - 1) ? 1 : st.bd[j + 1];
      const double* A = st.sites + (size_t)j * SITE_STRIDE;
      const int last = (j == kL - 1);
      const int keep = (kept >> j) & 1u;
      if ((size_t)nidx * (keep ? 2 : 1) * Br > (size_t)page) { st.status = ST_CAPACITY; MD_SYNC(); return; }
      if (keep) {
        MD_PAR_FOR(t, nidx * 2 * Br) {
          int b2 = t % Br; int r = t / Br; int idx = r % nidx; int p = r / nidx;
          double acc = 0.0;
          for (int b = 0; b < Bl; ++b) {
            double a = last ? T[b * 2 + p] : A[((size_t)b * 2 + p) * Br + b2];
            acc += D0[idx * Bl + b] * a;
          }
          D1[(p * nidx + idx) * Br + b2] = acc;
        }
      } else {  // traced head site: contract its physical leg with (1, 1)/sqrt(2)
        MD_PAR_FOR(t, nidx * Br) {
          int b2 = t % Br; int idx = t / Br;
          double acc = 0.0;
          for (int b = 0; b < Bl; ++b) {
            double a = last ? (T[b * 2] + T[b * 2 + 1]) : (A[((size_t)b * 2) * Br + b2] + A[((size_t)b * 2 + 1) * Br + b2]);
            acc += D0[idx * Bl + b] * a;
          }
          D1[idx * Br + b2] = acc * (1.0 / sqrt(2.0));
        }
      }
      MD_SYNC();
      double* tmp = D0; D0 = D1; D1 = tmp;
      if (keep) nidx *= 2;
      nb = Br;
    }
    // right part, from the last head site down to jm: E[b, iR], the newly absorbed (earlier) site is the least
    // significant bit of iR
    int nR = 1, bR = 1;  // E0 is (bR x nR); empty right part: the 1 x 1 identity
    MD_LEADER { E0[0] = 1.0; }
    MD_SYNC();
    for (int j = kL - 1; j >= jm; --j) {
      const int Bl = st.bd[j];                 // left bond of site j
      const int Br = bR;                       // = right bond of site j (1 for the last head site)
      const double* A = st.sites + (size_t)j * SITE_STRIDE;
      const int last = (j == kL - 1);
      const int keep = (kept >> j) & 1u;
      if ((size_t)Bl * nR * (keep ? 2 : 1) > (size_t)page) { st.status = ST_CAPACITY; MD_SYNC(); return; }
      if (keep) {
        MD_PAR_FOR(t, Bl * 2 * nR) {
          int iR = t % nR; int r = t / nR; int p = r & 1; int bl = r >> 1;
          double acc = 0.0;
          for (int b = 0; b < Br; ++b) {
            double a = last ? T[bl * 2 + p] : A[((size_t)bl * 2 + p) * Br + b];
            acc += a * E0[b * nR + iR];
          }
          E1[bl * (2 * nR) + (2 * iR + p)] = acc;
        }
      } else {
        MD_PAR_FOR(t, Bl * nR) {
          int iR = t % nR; int bl = t / nR;
          double acc = 0.0;
          for (int b = 0; b < Br; ++b) {
            double a = last ? (T[bl * 2] + T[bl * 2 + 1]) : (A[((size_t)bl * 2) * Br + b] + A[((size_t)bl * 2 + 1) * Br + b]);
            acc += a * E0[b * nR + iR];
          }
          E1[bl * nR + iR] = acc * (1.0 / sqrt(2.0));
        }
      }
      MD_SYNC();
      double* tmp = E0; E0 = E1; E1 = tmp;
      if (keep) nR *= 2;
      bR = Bl;
    }
    // combine: |sum_b D[iL, b] E[b, iR]|, its 2-norm and its maximum (block reductions through s.part)
    const int nL = nidx, ncls = nL * nR;
    MD_PAR_FOR(i, ncls) {
      const int iL = i % nL, iR = i / nL;
      double acc = 0.0;
      for (int b = 0; b < nb; ++b) acc += D0[iL * nb + b] * E0[b * nR + iR];
      out[i] = fabs(acc);
    }
    MD_SYNC();
    MD_PAR_FOR(ch, 64) {
      double nrm = 0.0, top = 0.0; int bad = 0;
      for (int i = ch; i < ncls; i += 64) { double v = out[i]; nrm += v * v; bad |= (v != v); top = fmax(top, v); }
      s.part[ch] = nrm; s.part[64 + ch] = top; s.part[128 + ch] = bad ? 1.0 : 0.0;
    }
    MD_SYNC();
    MD_LEADER {
      double nrm = 0.0, top = 0.0, bad = 0.0;
      for (int ch = 0; ch < 64; ++ch) { nrm += s.part[ch]; top = fmax(top, s.part[64 + ch]); bad += s.part[128 + ch]; }
      nrm = sqrt(nrm);
      s.dscr[4] = (ren && nrm > 1e-17) ? 1.0 / nrm : 1.0;
      s.dscr[5] = top;
      s.iscr[8] = (bad != 0.0 || nrm != nrm) ? 1 : 0;
    }
    MD_SYNC();
    const double inv = s.dscr[4];
    MD_PAR_FOR(i, ncls) out[i] = out[i] * inv;
    MD_SYNC();
    MD_LEADER {
      if (s.iscr[8]) { *success = 0; if (st.status == ST_OK) st.status = ST_ZERO_NORM; }
      else {
        const double top = s.dscr[5] * inv;
        const double eps = fmax(st.P.readout_rel * top, st.P.readout_abs);
        *success = (out[0] >= top - eps) ? 1 : 0;
      }
    }
    MD_SYNC();
  }

  // ---- run the error-independent program on one syndrome ----
  MD_DEV void run(const int* prog, const uint8_t* sym, double* out, int* success) {
    init_product(sym);
    run_ops(prog, out, success);
  }

  // ---- run a program on the MPS already present in st.sites / st.bd / st.centre ----
  MD_DEV void run_ops(const int* prog, double* out, int* success) {
    const long long t_all = MD_CLOCK();
    int pc = 0;
    while (pc < st.P.n_ops) {
      const int op = prog[pc];
      if (op == OP_END) break;
      if (op == OP_SWEEP) {
        const int start = prog[pc + 1], len = prog[pc + 2], ren = prog[pc + 3], rsv = prog[pc + 4];
        move_centre(start);
        if (st.status != ST_CAPACITY) {
          sweep(start, len, prog + pc + 5, st.P.chi_max, st.P.cut_sweep, ren, 0, rsv);
          set_centre(start + len - 1);
        }
        pc += 5 + len;
      } else if (op == OP_MOVE) {
        move_centre(prog[pc + 1]);
        pc += 2;
      } else if (op == OP_GATE2) {
        move_centre(prog[pc + 1]);
        if (st.status != ST_CAPACITY) gate2(prog[pc + 1], prog[pc + 2], prog[pc + 3]);
        pc += 4;
      } else if (op == OP_COMPRESS) {
        // compress_bond (canonical.py:684-784, "svd"): centre to `bond`, split theta(bond, bond+1) with the
        // user's chi_max / cut (traced), then hand the spectrum back to site `bond` (U.S | V, :766-767)
        const int bond = prog[pc + 1], rsv = prog[pc + 2];
        move_centre(bond);
        if (st.status != ST_CAPACITY) {
          set_frame(0);
          sweep(bond, 2, nullptr, st.P.chi_max, st.P.cut_sweep, 0, 0, rsv);
          set_centre(bond + 1);
        }
        if (st.status != ST_CAPACITY) {
          const double cut_back = st.P.cut_sweep < st.P.cut_move ? st.P.cut_sweep : st.P.cut_move;
          set_frame(1);
          sweep(st.P.n_sites - 2 - bond, 2, nullptr, CHI, cut_back, 0, 1, 0);
          set_frame(0);
          set_centre(bond);
        }
        pc += 3;
      } else if (op == OP_MARGINAL) {
        if (out != nullptr) marginal_readout(out, success);
        pc += 1;
      } else {
        st.status = ST_CAPACITY;
        break;
      }
      if (st.status == ST_CAPACITY) break;
    }
    MD_LEADER { s.cnt.cyc_total += (double)(MD_CLOCK() - t_all); }
  }
};

}  // namespace mdopt
