/*
 * mdopt_b200 -- C ABI of the B200 (sm_100a) kernels for mdopt's batched MPS-MPO decode path.
 *
 * The reference (quicophy/mdopt v1.1.1) is pure Python and has no FFI of its own; its only dispatch
 * point is the array shim mdopt/backend/array.py:11-115.  These entry points are what a binding for
 * the decode path replaces (see INTEGRATION.md for the ctypes stub a maintainer would add):
 *
 *   mdopt_decode_batch_*      examples/decoding/decoding.py:1164-1404  decode_css (per syndrome), fanned
 *                             out by examples/decoding/quantum_surface.py:242-246 (Pool.starmap);
 *                             internally mdopt/optimiser/utils.py:216-359 apply_constraints,
 *                             mdopt/contractor/contractor.py:177-378 mps_mpo_contract,
 *                             mdopt/mps/canonical.py:345-420 move_orth_centre, :906-991 marginal.
 *   mdopt_svd_batch_device    mdopt/utils/utils.py:29-138 svd / :290-383 split_two_site_tensor
 *   mdopt_site_apply_batch_device   mdopt/contractor/contractor.py:328-342 (theta build, one site)
 *   mdopt_readout_batch_device      examples/decoding/decoding.py:1362-1379 (abs, L2 renorm, arg-max)
 *
 * Conventions: plain pointers and sizes, no allocation inside, the CUDA stream is passed as void*
 * (cudaStream_t).  "_device" functions take device pointers; "_host" functions take host pointers and
 * issue the host<->device copies themselves on the given stream using caller-provided device scratch.
 * All functions return 0 on success, a negative MDOPT_E_* code on a usage error, or a positive
 * cudaError_t.  Per-item numerical status is reported in the status arrays (MDOPT_ST_*).
 */
#ifndef MDOPT_B200_H
#define MDOPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDOPT_ST_OK 0            /* decoded */
#define MDOPT_ST_NOT_CONVERGED 1 /* a Jacobi SVD hit its sweep limit (result still returned) */
#define MDOPT_ST_CAPACITY 2      /* a tensor exceeded the compiled capacity; result invalid */
#define MDOPT_ST_ZERO_NORM 3     /* the state lost all weight; result is NaN like the reference's */
#define MDOPT_ST_NOT_MONOMIAL 4  /* monomial engine only: a tensor left the sparse structure it verifies at every step
                                    (or the program has a two-site gate); result invalid, re-run on the dense engine */

#define MDOPT_E_BADARG (-1)
#define MDOPT_E_CAPACITY (-2)
#define MDOPT_E_NOGPU (-3)

#define MDOPT_MODE_FAST 0 /* rank-revealing QR steps, Jacobi SVD only where truncation is active */
#define MDOPT_MODE_SVD 1  /* every split is an SVD with the reference's truncation rule; a split whose columns are
                             already orthogonal (probe check, then Gram check) skips the Jacobi rotations */
#define MDOPT_MODE_SVD_GRAM 2 /* same, with the deterministic Gram check only */

/* Program opcodes (int32 stream); see mdopt_b200/schedule.py for the builder. */
#define MDOPT_OP_END 0
#define MDOPT_OP_SWEEP 1    /* [1, start, len, renorm_centre_after, renorm_spectrum, tensor_id x len] */
#define MDOPT_OP_MOVE 2     /* [2, target_site] */
#define MDOPT_OP_MARGINAL 3 /* [3] marginalise sites n_logical..N-1 and read out */
#define MDOPT_OP_GATE2 4    /* [4, site, gate_id, renormalise] two-site gate G[j][l][n][o] (16 doubles in wtab) on
                               (site, site+1), then split with chi=capacity / cut_move (depolarising bias) */

#define MDOPT_OP_MARGINAL_DMRG 6 /* [6, sweeps] monomial engine only: marginalise sites n_logical..N-1, then the
                                   Dephasing-DMRG read-out from |+...+> (decoding.py:1382-1400,
                                   mdopt/optimiser/dephasing_dmrg.py:147-374); dist[b] = the n_logical bits of the main
                                   component in the order of the reversed logical chain, success[b] = overlap with
                                   |0...0> (1 or 0).  For more than 12 logical sites, where the output stride is
                                   n_logical doubles instead of 2^n_logical */
#define MDOPT_OP_COMPRESS 5 /* [5, bond, renorm_spectrum] compress_bond (mdopt/mps/canonical.py:684-784, "svd"): centre to
                               `bond`, split theta(bond, bond+1) with chi_max / cut_sweep, leave U.S at `bond` */

typedef struct mdopt_decode_desc {
  int32_t n_sites;      /* MPS length N = 2n + k */
  int32_t n_logical;    /* k = kx + kz logical sites, 1 <= k <= 12 (the reference's dense read-out limit,
                           decoding.py:1358; needs chi_cap >= 2^(k-1) for k <= 6, 2^ceil(k/2) above) */
  int32_t chi_max;      /* decode_css(chi_max=...) */
  int32_t mode;         /* MDOPT_MODE_* */
  int32_t renormalise;  /* decode_css(renormalise=...) */
  int32_t trace_stride; /* doubles per traced split (0: no trace); layout [rows, cols, kept, is_move, s...] */
  int32_t trace_max;    /* traced splits per syndrome */
  int32_t n_ops;        /* ints in the program */
  double cut_sweep;     /* decode_css(cut=...) */
  double cut_move;      /* 1e-12, mdopt/utils/utils.py:293 default used by move_orth_centre */
  double rank_tol;      /* MDOPT_MODE_FAST: relative rank threshold (default 1e-14) */
  double bias_p;        /* bit-flip bias probability, < 0: none */
  int32_t n_head;       /* read-out: sites 0..n_head-1 form the dense stage (0: use n_logical) */
  uint32_t kept_mask;   /* read-out: bit j set = site j (< n_head) is kept, clear = traced out; exactly
                           n_logical bits set, bit n_head-1 set.  0: the first n_logical sites.  Only erasures of
                           qubits with index < n_logical change it (decoding.py:1341-1347 passes qubit indices to
                           marginal() as site indices) */
  double readout_rel;   /* success = v[0] >= max(v) - max(readout_rel * max(v), readout_abs) */
  double readout_abs;   /* decode_css: 1e-9 / 1e-12 (decoding.py:1367-1376); decode_custom: 1e-12 / 0 (:1600-1604) */
} mdopt_decode_desc;

int mdopt_b200_abi_version(void);

/* Number of SMs of the current device (<= 0: no usable GPU). */
int mdopt_device_sm_count(void);

/* Smallest compiled capacity (8, 16, 32 or 64) that holds chi; 0 if none. */
int mdopt_chi_capacity(int chi);

/* Resident CTAs the decode kernel uses on the current device for a capacity. */
int mdopt_decode_num_ctas(int chi_cap);

/* Bytes of device workspace the decode kernel needs for n_ctas CTAs. */
size_t mdopt_decode_workspace_bytes(int chi_cap, int n_sites, int n_ctas);

/*
 * Decode `batch` syndromes that share one program.
 *   program   int32[n_ops]            opcode stream (device)
 *   wtab      double[n_tensors*16]    MPO site tensors (vL,vR,pU,pD) row-major, zero padded (device)
 *   wdim      int32[n_tensors*4]      vL, vR, isometric flag, 0 (device)
 *   symbols   uint8[batch*n_sites]    per-site product-state symbol 0='0', 1='1', 2='+' (device)
 *   dist      double[batch*2^k]       out: |marginal| per logical class (device)
 *   success   int32[batch]            out: tolerant arg-max flag (device)
 *   status    int32[batch]            out: MDOPT_ST_* (device)
 *   trace     double[batch*trace_max*trace_stride] or NULL (device)
 *   counters  double[16] or NULL      out: summed work counters {flops_qr, flops_jacobi, flops_apply,
 *                                     bytes, steps, truncating steps, jacobi sweeps, qr reflectors, then
 *                                     leader-thread cycle counts: qr, form_q, jacobi, apply, io, coef,
 *                                     total, 0} (device)
 *   workspace mdopt_decode_workspace_bytes() bytes (device)
 */
int mdopt_decode_batch_device(const mdopt_decode_desc* desc, int chi_cap, int n_ctas, const int32_t* program,
                              const double* wtab, const int32_t* wdim, int n_tensors, const uint8_t* symbols,
                              int batch, double* dist, int32_t* success, int32_t* status, double* trace,
                              double* counters, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Same, with the per-syndrome inputs and outputs in HOST memory (pinned for overlap).  d_symbols /
 * d_dist / d_success / d_status are caller-provided device staging buffers of the same sizes.  The call
 * enqueues H2D, the kernel and D2H on `stream` and synchronises the stream before returning.
 */
int mdopt_decode_batch_host(const mdopt_decode_desc* desc, int chi_cap, int n_ctas, const int32_t* d_program,
                            const double* d_wtab, const int32_t* d_wdim, int n_tensors, const uint8_t* h_symbols,
                            int batch, double* h_dist, int32_t* h_success, int32_t* h_status, uint8_t* d_symbols,
                            double* d_dist, int32_t* d_success, int32_t* d_status, double* d_counters,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * The monomial ("trellis") engine for the same decode (mdopt_b200/csrc/mono_core.cuh): one warp per syndrome, every
 * site tensor held as 2*chi (value, target) pairs.  It covers programs made of MDOPT_OP_SWEEP / _MOVE / _MARGINAL on
 * product-state inputs (the bit-flip-bias decode, decoding.py:1267-1274, 1285-1339) and verifies the structure it
 * relies on at every step; a syndrome that leaves it gets MDOPT_ST_NOT_MONOMIAL and must be re-run through
 * mdopt_decode_batch_*.  Arguments as for mdopt_decode_batch_device / _host, with warps in place of CTAs;
 * desc->mode is ignored; counters = {0, 0, 0, bytes, steps, truncating steps, 0, fallbacks, 0...}.
 * Capacities: 8, 16, ..., 512.  n_logical may be up to 64 here: above 12 the program must end with
 * MDOPT_OP_MARGINAL_DMRG and dist holds n_logical doubles per syndrome.
 */
int mdopt_mono_chi_capacity(int chi);
int mdopt_mono_num_warps(int chi_cap);
size_t mdopt_mono_workspace_bytes(int chi_cap, int n_sites, int n_warps);
int mdopt_mono_decode_batch_device(const mdopt_decode_desc* desc, int chi_cap, int n_warps, const int32_t* program,
                                   const double* wtab, const int32_t* wdim, int n_tensors, const uint8_t* symbols,
                                   int batch, double* dist, int32_t* success, int32_t* status, double* trace,
                                   double* counters, void* workspace, size_t workspace_bytes, void* stream);
int mdopt_mono_decode_batch_host(const mdopt_decode_desc* desc, int chi_cap, int n_warps, const int32_t* d_program,
                                 const double* d_wtab, const int32_t* d_wdim, int n_tensors, const uint8_t* h_symbols,
                                 int batch, double* h_dist, int32_t* h_success, int32_t* h_status, uint8_t* d_symbols,
                                 double* d_dist, int32_t* d_success, int32_t* d_status, double* d_counters,
                                 void* workspace, size_t workspace_bytes, void* stream);

/*
 * Run a program on caller-owned MPS buffers, in place (single-state API: mps_mpo_contract
 * contractor.py:177-378, apply_constraints optimiser/utils.py:216-359, move_orth_centre
 * canonical.py:345-420, marginal canonical.py:906-991).
 *   sites_io  double[batch][n_sites][mdopt_site_stride(chi_cap)]  packed (vL,p,vR) row-major per site
 *   bonds_io  int32[batch][n_sites+1]   bond dimensions, bonds_io[0] = bonds_io[n_sites] = 1
 *   centre_io int32[batch]              orthogonality centre in / out
 *   dist/success may be NULL when the program has no MDOPT_OP_MARGINAL (a chi_max in the descriptor that
 *   exceeds chi_cap is allowed here: it is clamped by the capacity check on every tensor).
 */
size_t mdopt_site_stride(int chi_cap);
size_t mdopt_mps_program_workspace_bytes(int chi_cap, int batch);
int mdopt_mps_program_device(const mdopt_decode_desc* desc, int chi_cap, const int32_t* program, const double* wtab,
                             const int32_t* wdim, int n_tensors, int batch, double* sites_io, int32_t* bonds_io,
                             int32_t* centre_io, double* dist, int32_t* success, int32_t* status, double* trace,
                             void* workspace, size_t workspace_bytes, void* stream);

/*
 * Batched truncated SVD, one CTA per matrix (mdopt/utils/utils.py:29-138).
 *   a        double[batch*m*n] row-major, m,n <= 2*chi_cap
 *   u        double[batch*m*kmax], s double[batch*kmax], vh double[batch*kmax*n], kmax = min(m,n)
 *   kept     int32[batch]      min(chi_max, #(s > cut))
 *   trunc    double[batch]     sum of discarded s^2
 *   Columns >= kept[b] of u / rows >= kept[b] of vh / entries >= kept[b] of s are zero.
 *   renormalise != 0 divides the kept s by their 2-norm.
 */
int mdopt_svd_batch_device(int chi_cap, int batch, int m, int n, const double* a, int chi_max, double cut,
                           int renormalise, double* u, double* s, double* vh, int32_t* kept, double* trunc,
                           int32_t* status, void* stream);

/*
 * Batched one-site contraction theta[b] = M[b] . (A[b] x W) (contractor.py:328-342):
 *   m_in  double[batch][rows][vl*mr], a_site double[batch][mr][2][bn], w double[vl*vr*4]
 *   theta double[batch][rows][2][vr][bn]
 */
int mdopt_site_apply_batch_device(int chi_cap, int batch, int rows, int vl, int vr, int mr, int bn,
                                  const double* m_in, const double* a_site, const double* w, double* theta,
                                  void* stream);

/*
 * Batched read-out (decoding.py:1362-1379): dist[b] = |v[b]| (/ ||v[b]||_2 if renormalise and > 1e-17),
 * success[b] = dist[b][0] >= max(dist[b]) - max(1e-9*max, 1e-12).
 */
int mdopt_readout_batch_device(int batch, int n_classes, const double* v, int renormalise, double* dist,
                               int32_t* success, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDOPT_B200_H */
