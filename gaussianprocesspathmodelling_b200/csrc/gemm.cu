// FP64 tensor-core tile GEMM for sm_100a: 128x128 CTA tile, operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) through a 6-stage mbarrier ring, one producer warp and
// eight consumer warps issuing mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) on 64x32 warp tiles.
//
// FP64 has no tcgen05 kind, so the accumulators live in registers (128 per thread), not TMEM.
#include "gemm.cuh"

namespace gpm {

constexpr int STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int GEMM_THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int GEMM_SMEM = STAGES * 2 * SLAB_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8 + NB * 4 * 8;

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // swizzle needs 1024 B
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_full = base + STAGES * 2 * SLAB_BYTES;
  const uint32_t bar_empty = bar_full + STAGES * 8;
  double* red = reinterpret_cast<double*>(gen + STAGES * 2 * SLAB_BYTES + 2 * STAGES * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // tile coordinates
  int ti, tj;
  if (p.tri) {
    const int b = blockIdx.x;
    int i = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= b) i++;
    while (i * (i + 1) / 2 > b) i--;
    ti = i;
    tj = b - i * (i + 1) / 2;
  } else {
    ti = blockIdx.x % p.tiles_m;
    tj = blockIdx.x / p.tiles_m;
  }
  const long long bz = blockIdx.y;
  const int a_row = p.a_row0 + ti * NB + (int)(bz * p.batch_a_rows);
  const int b_row = p.b_row0 + tj * p.b_tile_rows + (int)(bz * p.batch_b_rows);
  const int nslab = p.klen / SLAB_K;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, CONSUMER_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===== TMA producer: one elected lane =====
    if (lane == 0) {
      prefetch_tmap(&mapA);
      prefetch_tmap(&mapB);
      for (int s = 0; s < nslab; s++) {
        const int st = s % STAGES;
        if (s >= STAGES) mbar_wait(bar_empty + st * 8, ((s / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(bar_full + st * 8, 2 * SLAB_BYTES);
        const uint32_t dst = base + st * 2 * SLAB_BYTES;
        tma_load_2d(dst, &mapA, p.a_col0 + s * SLAB_K, a_row, bar_full + st * 8);
        tma_load_2d(dst + SLAB_BYTES, &mapB, p.b_col0 + s * SLAB_K, b_row, bar_full + st * 8);
      }
    }
    return;
  }

  // ===== consumers: 2 (m) x 4 (n) warps, warp tile 64 x 32 =====
  const int wm = warp >> 2, wn = warp & 3;
  const int g = lane >> 2, q = lane & 3;
  uint32_t off[4];
#pragma unroll
  for (int t = 0; t < 4; t++) off[t] = frag_off(g, q, t);
  const uint32_t a_warp = wm * 64 * 128, b_warp = SLAB_BYTES + wn * 32 * 128;

  double acc[8][4][2];
#pragma unroll
  for (int mt = 0; mt < 8; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  for (int s = 0; s < nslab; s++) {
    const int st = s % STAGES;
    mbar_wait(bar_full + st * 8, (s / STAGES) & 1);
    const uint32_t sa = base + st * 2 * SLAB_BYTES + a_warp;
    const uint32_t sb = base + st * 2 * SLAB_BYTES + b_warp;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      double a[8], b[4];
#pragma unroll
      for (int mt = 0; mt < 8; mt++) a[mt] = lds_f64(sa + mt * 1024 + off[t]);
#pragma unroll
      for (int nt = 0; nt < 4; nt++) b[nt] = lds_f64(sb + nt * 1024 + off[t]);
#pragma unroll
      for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + st * 8);
  }

  // ===== epilogue: registers -> global (16-byte stores, 64 B per row segment per quad) =====
  const long long crow_base = p.c_row0 + (long long)ti * NB + wm * 64 + bz * p.batch_c_rows;
  const long long ccol_base = p.c_col0 + (long long)tj * NB + wn * 32 + 2 * q;
  const long long rows_end = p.c_rows_end + bz * p.batch_c_rows;
#pragma unroll
  for (int mt = 0; mt < 8; mt++) {
    const long long row = crow_base + mt * 8 + g;
    double sq = 0.0;
    if (row < rows_end) {
      double* crow = p.C + row * p.ldc;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const long long col = ccol_base + nt * 8;
        double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
        if (col + 1 < p.c_cols_end) {
          double2* ptr = reinterpret_cast<double2*>(crow + col);
          if (p.epi == EPI_SUB) {
            const double2 c = *ptr;
            v0 = c.x - v0;
            v1 = c.y - v1;
          }
          *ptr = make_double2(v0, v1);
          sq += v0 * v0 + v1 * v1;
        } else if (col < p.c_cols_end) {
          if (p.epi == EPI_SUB) v0 = crow[col] - v0;
          crow[col] = v0;
          sq += v0 * v0;
        }
      }
    }
    if (p.rowsq) {
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      if (q == 0) red[(wm * 64 + mt * 8 + g) * 4 + wn] = sq;
    }
  }
  if (p.rowsq) {
    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
    if (tid < NB) {
      const long long row = p.c_row0 + (long long)ti * NB + tid + bz * p.batch_c_rows;
      if (row < rows_end) {
        const double s4 = (red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]);
        p.rowsq[row + bz * p.batch_rowsq] += s4;
      }
    }
  }
}

int launch_gemm(gpm_handle_impl* h, const CUtensorMap& mapA, const CUtensorMap& mapB,
                const GemmArgs& args, int batch, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    GPM_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    attr_set = true;
  }
  if (args.klen <= 0 || args.klen % SLAB_K != 0) {
    set_error("gemm: contraction length %d is not a positive multiple of %d", args.klen, SLAB_K);
    return 998;
  }
  const int gx = gemm_grid_x(args);
  if (gx <= 0 || batch <= 0) return 0;
  dim3 grid(gx, batch);
  gemm_nt_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(mapA, mapB, args);
  GPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace gpm
