// FP64 tensor-core tile GEMM (NT form) used by the blocked Cholesky, the blocked TRSM of the
// posterior variance and the batched fits:   C_tile (op)= A_rows * B_rows^T
// where both operands are row-major with the contraction index contiguous.
#pragma once
#include "common.cuh"

namespace gpm {

enum { EPI_STORE = 0, EPI_SUB = 1, EPI_NEG = 2 };   // C = acc;  C = C - acc;  C = -acc

struct GemmArgs {
  double* C;              // output matrix
  long long ldc;
  double* rowsq;          // optional: rowsq[c_row] += sum over the tile's columns of out^2
  int tiles_m, tiles_n;   // tile grid; with tri != 0 only tiles ti >= tj of a tiles_m x tiles_m grid
  int tri;
  int a_row0, a_col0;     // element origin of tile (0,*) in A and of the contraction range
  int b_row0, b_col0;     // element origin of tile (*,0) in B
  int b_tile_rows;        // B rows advance per tj (NB, or 0 when every tile uses the same B block)
  int klen;               // contraction length (multiple of SLAB_K)
  long long c_row0, c_col0;          // origin of tile (0,0) in C
  long long c_rows_end, c_cols_end;  // exclusive store bounds in C
  int epi;                // EPI_STORE: C = acc;  EPI_SUB: C = C - acc
  // batch (blockIdx.y): rows added per batch index to A, B and C
  long long batch_a_rows, batch_b_rows, batch_c_rows;
  long long batch_rowsq;
  // Variance-sweep mode (sweep_nblk > 0): a "tile" is a 128-row block of query points and the CTA runs
  // the whole blocked forward substitution on it, block column by block column:
  //   D(0), U(1), D(1), U(2), D(2), ...   with  U(k): W[:,k] -= W[:,0:k] L[k,0:k]^T   (B from mapB)
  //                                             D(k): W[:,k]  = W[:,k] inv(L_kk)^T     (B from mapB2, rowsq)
  // Each op reads what the previous one wrote (through global memory, generic -> async proxy fence).
  int sweep_nblk;
  int sweep_tri;          // sweep mode on a block-upper-triangular right-hand side (row block t is zero left of
                          // block column t): row block t starts at op D(t) and contracts from column t*NB
  int tri_b;              // tile mode: the B block is lower triangular (an inverted diagonal block, B[n][c] = 0 for
                          // c > n): a consumer warp skips the slabs that are all zero for its 32 columns
  // tile mode: triangular operands let a tile skip the structurally-zero part of the contraction
  int kstart_mode;        // 1: start at ti*NB (A zero left of its diagonal block)
  int kend_mode;          // 1: stop at (ti+1)*NB (A zero right of its diagonal block); 2: stop at (tj+1)*NB (same for B)
  long long batch_cols;   // columns added per batch index to the A, B and C column origins
  int tiles_per_cta;      // filled by launch_gemm: consecutive tiles one CTA works through
  int max_tiles_per_cta;  // 0 = default (16); the look-ahead Cholesky caps it so that SMs free up regularly
  // tile mode, EPI_STORE: forward substitution fused into the panel solves of a factorisation.  After the tile
  // L_ik = K_ik inv(L_kk)^T is formed (still in registers) the running right-hand side is updated,
  //   r_i -= L_ik z_k,   z_k = inv(L_kk) r_k  (written by potf2 before this launch),
  // so the forward solve needs no pass of its own over L.  rhs_r / rhs_z: (rows x rhs_R) row-major.
  double* rhs_r;
  const double* rhs_z;
  int rhs_R;
  long long rhs_z_row0;   // first row of z_k
  long long rhs_r_row0;   // first row of the residual block of tile ti = 0
  long long rhs_rows_end; // rows per matrix (exclusive bound for the residual rows)
  long long batch_rhs_rows;
  // raw operand pointers (optional): with them a launch of at most 37 tiles on one matrix takes the latency
  // kernel of gemm_small.cu (four 64 x 64 quarters per tile, bitwise the same arithmetic)
  const double* small_A;
  const double* small_B;
  long long small_lda, small_ldb;
  long long small_a_rows_end, small_b_rows_end;   // rows of the matrices behind A and B (reads are clamped to them)
#ifdef GPM_GEMM_TIMING
  int dbg_launch;         // instrumentation builds: slot of this launch in the phase-stamp table
#endif
  int diag_lower;         // tile mode: tiles with ti == tj are symmetric (SYRK) and only their lower triangle is
                          // needed: the 8x8 sub-tiles strictly above the diagonal are not computed
};

// number of CTAs along x for the given args
inline int gemm_grid_x(const GemmArgs& a) {
  return a.tri ? a.tiles_m * (a.tiles_m + 1) / 2 : a.tiles_m * a.tiles_n;
}

// mapC describes the matrix args.C points into (used to prefetch C tiles by TMA when epi == EPI_SUB)
bool gemm_small_eligible(const gpm_handle_impl* h, const GemmArgs& a, int batch);
int launch_gemm_small(gpm_handle_impl* h, const GemmArgs& a, cudaStream_t stream);
// large batches: 64 x 128 half-tiles, two CTAs per SM (gemm_half.cu)
bool gemm_half_eligible(const gpm_handle_impl* h, const GemmArgs& a, int batch);
int launch_gemm_half(gpm_handle_impl* h, const GemmArgs& a, int batch, cudaStream_t stream);

int launch_gemm(gpm_handle_impl* h, const CUtensorMap& mapA, const CUtensorMap& mapB,
                const CUtensorMap& mapC, const GemmArgs& args, int batch, cudaStream_t stream,
                const CUtensorMap* mapB2 = nullptr);

}  // namespace gpm
