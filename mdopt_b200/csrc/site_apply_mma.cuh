// K1 on the Blackwell data path: batched theta-build theta[b] = M[b] . (A[b] x W) (contractor.py:328-342) with
//   * operands staged by the TMA engine: one `cp.async.bulk.shared::cluster.global` per operand row into padded
//     shared-memory rows, completion counted on an mbarrier (SASS: UBLKCP + SYNCS), and
//   * the contraction on the FP64 tensor pipe: `mma.sync.aligned.m8n8k4.f64` tiles (SASS: DMMA), 16 x 32 outputs per
//     warp task, accumulators in registers.
// Shapes it takes: rows % 16 == 0, bn % 32 == 0, mr % 4 == 0 (every full-chi step of a decode); other shapes run the
// scalar kernel in kernels.cuh.  Row paddings (+4 doubles for M, +2 for A) make the 64-bit fragment loads of a
// half-warp hit 16 distinct bank pairs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mdopt {
namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned) global -> shared, completing on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int CHI>
struct SiteApplyMmaSmem {
  static constexpr int LDM = 2 * CHI + 4;   // padded row of M (cin <= 2*CHI doubles)
  static constexpr int LDA = CHI + 2;       // padded row of A (bn <= CHI doubles)
  alignas(16) double Ms[2 * CHI * LDM];
  alignas(16) double As[2 * CHI * LDA];
  double W[16];
  alignas(8) uint64_t bar;
};

__host__ __device__ inline bool site_apply_mma_ok(int rows, int vl, int vr, int mr, int bn) {
  return rows % 16 == 0 && bn % 32 == 0 && mr % 4 == 0 && vl >= 1 && vl <= 2 && vr >= 1 && vr <= 2;
}

template <int CHI>
__global__ void __launch_bounds__(512, 1)
site_apply_mma_kernel(int batch, int rows, int vl, int vr, int mr, int bn, const double* __restrict__ m_in,
                      const double* __restrict__ a_site, const double* __restrict__ w, double* __restrict__ theta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = SiteApplyMmaSmem<CHI>;
  SM& s = *reinterpret_cast<SM*>(smem_raw);
  constexpr int LDM = SM::LDM, LDA = SM::LDA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int cin = vl * mr;
  if (tid < 16) s.W[tid] = (tid < vl * vr * 4) ? w[tid] : 0.0;
  if (tid == 0) mbar_init(&s.bar, 1);
  __syncthreads();
  uint32_t parity = 0;
  const uint32_t row_m_bytes = (uint32_t)cin * 8u, row_a_bytes = (uint32_t)bn * 8u;
  const int n_a_rows = mr * 2;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double* mb = m_in + (size_t)b * rows * cin;
    const double* ab = a_site + (size_t)b * n_a_rows * bn;
    // ---- stage both operands: one bulk copy per row, all counted on one mbarrier phase ----
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the last item are done
    if (tid == 0) mbar_expect_tx(&s.bar, (uint32_t)rows * row_m_bytes + (uint32_t)n_a_rows * row_a_bytes);
    for (int r = tid; r < rows + n_a_rows; r += blockDim.x) {
      if (r < rows) bulk_g2s(s.Ms + (size_t)r * LDM, mb + (size_t)r * cin, row_m_bytes, &s.bar);
      else bulk_g2s(s.As + (size_t)(r - rows) * LDA, ab + (size_t)(r - rows) * bn, row_a_bytes, &s.bar);
    }
    mbar_wait(&s.bar, parity);
    parity ^= 1u;
    // ---- warp tasks: (16 rows) x (32 columns) of one (p', w') block ----
    const int nrt = rows / 16, nct = bn / 32, ncomb = 2 * vr;
    double* out = theta + (size_t)b * rows * 2 * vr * bn;
    for (int task = warp; task < nrt * nct * ncomb; task += nwarp) {
      const int ct = task % nct;
      const int rt = (task / nct) % nrt;
      const int comb = task / (nct * nrt);
      const int wo = comb % vr, pd = comb / vr;
      const int row0 = rt * 16, col0 = ct * 32;
      double acc[2][4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
      for (int q = 0; q < vl; ++q)
        for (int pu = 0; pu < 2; ++pu) {
          const double wv = s.W[((q * vr + wo) * 2 + pu) * 2 + pd];
          if (wv == 0.0) continue;
          const double* mrow0 = s.Ms + (size_t)(row0 + g) * LDM + q * mr + t;
          const double* mrow1 = mrow0 + (size_t)8 * LDM;
          const double* arow = s.As + (size_t)(t * 2 + pu) * LDA + col0 + g;
#pragma unroll 4
          for (int m = 0; m < mr; m += 4) {
            const double a0 = mrow0[m] * wv, a1 = mrow1[m] * wv;
            const double* ar = arow + (size_t)m * 2 * LDA;
            const double b0 = ar[0], b1 = ar[8], b2 = ar[16], b3 = ar[24];
            dmma_8x8x4(acc[0][0], a0, b0); dmma_8x8x4(acc[0][1], a0, b1);
            dmma_8x8x4(acc[0][2], a0, b2); dmma_8x8x4(acc[0][3], a0, b3);
            dmma_8x8x4(acc[1][0], a1, b0); dmma_8x8x4(acc[1][1], a1, b1);
            dmma_8x8x4(acc[1][2], a1, b2); dmma_8x8x4(acc[1][3], a1, b3);
          }
        }
      // C fragment: rows g (+8), columns 2t, 2t+1 of each 8-column tile -> one 16-byte store per tile
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = row0 + 8 * i + g;
        double* orow = out + (((size_t)row * 2 + pd) * vr + wo) * bn + col0 + 2 * t;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(orow + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
    __syncthreads();   // every warp is done reading the operands before the next item's copies land on them
  }
}

}  // namespace
}  // namespace mdopt
