// fp32 SIMT GEMM family with fused operand prologues and output epilogues.
//
// One kernel template covers the three operand layouts a Linear layer needs (forward, backward-data,
// weight-gradient) so that BatchNorm, ReLU, the residual add, the ReLU bit mask and the split
// reduction of weight gradients never make their own pass over HBM (the reference runs each as a
// separate ATen kernel).  fp32 FFMA is used on purpose: the parity budget is 1e-5
// relative (BASELINE.json north_star), which TF32/bf16 tensor-core products alone do not meet.
//
// Tile: 256 threads = 8 warps; a CTA owns (8*RM) x 128 outputs; warp w owns RM rows, lane l owns
// four columns; the reduction is walked in slices of 32 staged in shared memory (register
// double-buffered global loads, one __syncthreads per slice).  Per 4 reduction steps a thread issues
// RM + 4 LDS.128 for 16*RM FFMA, which is FFMA-bound for RM = 8.
#include "rc_gemm_tile.cuh"

namespace rc {

template <int RM, int AL, int BL, int kRK>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const GemmP p) {
  extern __shared__ __align__(16) float smem[];
  gemm_tile<RM, AL, BL, kRK>(p, blockIdx, smem);       // (waits for the kernel before it inside: after the early B fetch)
}

// `stages`: 2 when a CTA walks more than one reduction slice (double buffering), else 1.  The single-slice GEMMs of
// the reference-shape step then take 72 KB instead of 143 KB: three CTAs fit an SM, so the CTAs of the next kernel in
// the chain can already be resident (programmatic dependent launch) while this one is still running.
template <int RM, int AL, int BL, int kRK>
static size_t gemm_smem_bytes(int stages) {
  constexpr int kPadK = kRK + 4;
  constexpr int BM = 8 * RM;
  constexpr int SA = (AL == RC_A_ROW) ? kPadK : (BM + 4);
  constexpr int A_ROWS = (AL == RC_A_ROW) ? BM : kRK;
  constexpr int SB = (BL == RC_B_COL) ? kPadK : kBN;
  constexpr int B_ROWS = (BL == RC_B_COL) ? kBN : kRK;
  size_t tiles = (size_t)(2 * A_ROWS * SA + stages * B_ROWS * SB) * sizeof(float);    // A keeps both stages: B follows them
  size_t epi = (size_t)17 * kBN * sizeof(float);
  size_t cs = (size_t)kRK * BM * sizeof(float);
  size_t part = (RM == 1 && kRK == 128 && AL == RC_A_ROW) ? (size_t)8 * 8 * kBN * sizeof(float) : 0;   // warp-split partials
  size_t need = tiles > epi ? tiles : epi;
  need = need > cs ? need : cs;
  return need > part ? need : part;
}

template <int RM, int AL, int BL, int kRK>
static int gemm_launch(const GemmP& p, dim3 grid, cudaStream_t s) {
  static bool attr_set = false;
  const int slices = p.tiles1 + p.tiles2;
  const int per_cta = p.g.splits > 1 ? ceil_div(slices, p.g.splits) : slices;
  const size_t smem = gemm_smem_bytes<RM, AL, BL, kRK>(per_cta > 1 ? 2 : 1);
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_kernel<RM, AL, BL, kRK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)gemm_smem_bytes<RM, AL, BL, kRK>(2));
    attr_set = true;
  }
  launch_pdl(gemm_kernel<RM, AL, BL, kRK>, grid, dim3(kGemmThreads), smem, s, p);
  return check_launch("gemm_kernel");
}

template <int AL, int BL>
static int gemm_dispatch_rm(int rm, int rk, const GemmP& p, dim3 grid, cudaStream_t s) {
  if (AL == RC_A_ROW && rk == 128) {
    if (rm == 1) return gemm_launch<1, AL, BL, 128>(p, grid, s);
    return gemm_launch<2, AL, BL, 128>(p, grid, s);
  }
  switch (rm) {
    case 1: return gemm_launch<1, AL, BL, 32>(p, grid, s);
    case 2: return gemm_launch<2, AL, BL, 32>(p, grid, s);
    case 4: return gemm_launch<4, AL, BL, 32>(p, grid, s);
    default: return gemm_launch<8, AL, BL, 32>(p, grid, s);
  }
}

// slice length: the long slice only where it removes round trips (row-major A, tiny row tiles, K > 32)
static int choose_rk(const rc_gemm* g, int rm) {
  return (g->a_layout == RC_A_ROW && rm <= 2 && (g->k > 32 || g->k2 > 32) && g->splits <= 1) ? 128 : 32;
}

static int choose_rm(const rc_gemm* g) {
  if (g->rows_per_warp == 1 || g->rows_per_warp == 2 || g->rows_per_warp == 4 || g->rows_per_warp == 8) return g->rows_per_warp;
  const long long col_tiles = ceil_div(g->n > 0 ? g->n : 1, kBN);
  const long long sp = g->splits > 1 ? g->splits : 1;
  const int cands[4] = {8, 4, 2, 1};
  for (int c = 0; c < 4; ++c) {
    const long long ctas = (long long)ceil_div(g->m > 0 ? g->m : 1, 8 * cands[c]) * col_tiles * sp;
    // weight-gradient GEMMs run beside the data-gradient chain and cost it through the issue slots they take: 124 CTAs of
    // 32 output rows did better than 248 of 16 at the reference shape (250.5 vs 254.7 us per step; 8 rows: 284, 64 rows: 272)
    const long long enough = g->a_layout == RC_A_RED ? 96 : kNumSMs;
    if (ctas >= enough) return cands[c];
  }
  return 1;
}


}  // namespace rc

using namespace rc;

extern "C" int rc_gemm_row_tile(const rc_gemm* g) {
  if (!g) return -1;
  if (g->tc_ws && gemm_tc_kind(g) == 1) return 64;     // statistics tiles of the tensor-core kernel: half a 128-row tile
  return 8 * choose_rm(g);
}

/* debug: device buffer of 3 x 16 x 2 int64; CTA 0 of the tensor-core rows kernel records, for its first 16 tiles, the
 * clock at which its producer / MMA / epilogue role began and ended the tile (NULL switches it off) */
extern "C" void rc_debug_tc_trace(void* device_buf) { gemm_tc_set_trace(static_cast<long long*>(device_buf)); }

extern "C" size_t rc_gemm_tc_workspace(const rc_gemm* g) { return g ? gemm_tc_workspace(g) : 0; }

extern "C" int rc_gemm_tc_wgrad_splits(int m, int n, int k) {
  rc_gemm g = {};
  g.m = m; g.n = n; g.k = k; g.a_layout = RC_A_RED; g.b_layout = RC_B_RED;
  return gemm_tc_kind(&g) == 2 ? gemm_tc_wgrad_splits(&g) : 0;
}

extern "C" int rc_gemm_run(const rc_gemm* g, void* stream) {
  if (!g || !g->a.ptr || !g->b.ptr || !g->d) return fail(RC_ERR_ARG, "rc_gemm_run: null operand");
  if (g->m < 0 || g->n < 0 || g->k < 0 || g->k2 < 0) return fail(RC_ERR_ARG, "rc_gemm_run: negative size");
  if ((g->a_layout != RC_A_ROW && g->a_layout != RC_A_RED) || (g->b_layout != RC_B_COL && g->b_layout != RC_B_RED))
    return fail(RC_ERR_ARG, "rc_gemm_run: bad layout");
  if (g->k2 > 0 && (!g->a2 || !g->b2)) return fail(RC_ERR_ARG, "rc_gemm_run: second segment without operands");
  if (g->bits_out && g->b_layout != RC_B_COL) return fail(RC_ERR_ARG, "rc_gemm_run: bits_out needs the forward layout");
  if ((g->epi == RC_EPI_BN_STATS || g->epi == RC_EPI_BN_RELU_BWD) && !g->stats) return fail(RC_ERR_ARG, "rc_gemm_run: stats missing");
  if ((g->epi == RC_EPI_MASK_POS || g->epi == RC_EPI_BN_RELU_BWD) && !g->e_aux) return fail(RC_ERR_ARG, "rc_gemm_run: e_aux missing");
  if ((g->epi == RC_EPI_RELU_RES || g->epi == RC_EPI_ADD_RES) && !g->res) return fail(RC_ERR_ARG, "rc_gemm_run: res missing");
  if (g->epi == RC_EPI_BN_RELU_BWD && (!g->e_p0 || !g->e_p1 || !g->e_p2 || !g->e_p3)) return fail(RC_ERR_ARG, "rc_gemm_run: BN vectors missing");
  if (g->splits > 1 && g->epi != RC_EPI_NONE) return fail(RC_ERR_ARG, "rc_gemm_run: split reduction only with RC_EPI_NONE");
  if (g->splits > 1 && g->bias) return fail(RC_ERR_ARG, "rc_gemm_run: split reduction cannot add a bias");
  if (g->colsum_a && g->a_layout != RC_A_RED) return fail(RC_ERR_ARG, "rc_gemm_run: colsum_a needs A stored [r][i]");
  if (g->a.op == RC_OP_GINE_AGGR) {
    if (g->a_layout != RC_A_ROW || !g->a.idx0 || !g->a.idx1 || !g->a.aux || !g->a.p0 || !g->a.p1 || !g->a.p2)
      return fail(RC_ERR_ARG, "rc_gemm_run: RC_OP_GINE_AGGR needs A stored [i][r] and rowptr / col / attr / w_edge / b_edge / eps");
    if (g->k % 4 || g->a.ld % 4 || !aligned16(g->a.ptr) || !aligned16(g->a.p0) || !aligned16(g->a.p1) ||
        (g->a_out && (g->ld_a_out % 4 || !aligned16(g->a_out))))
      return fail(RC_ERR_ARG, "rc_gemm_run: RC_OP_GINE_AGGR needs 4 | k, 4 | ld and 16-byte aligned x, w_edge, b_edge, a_out");
  }
  if (g->b.op == RC_OP_GINE_AGGR) return fail(RC_ERR_ARG, "rc_gemm_run: RC_OP_GINE_AGGR is an A-operand prologue");
  if (g->a_layout == RC_A_RED && (g->epi != RC_EPI_NONE || g->bits_out))
    return fail(RC_ERR_ARG, "rc_gemm_run: weight-gradient GEMMs (A stored [r][i]) take RC_EPI_NONE and no bits_out");
  if (g->b.op != RC_OP_NONE && g->a_layout != RC_A_RED)
    return fail(RC_ERR_ARG, "rc_gemm_run: a B-operand prologue exists for weight-gradient GEMMs only (A stored [r][i])");
  if (g->m == 0 || g->n == 0) return RC_OK;
  {
    const int tc = gemm_tc_kind(g);
    if ((tc == 1 && g->tc_ws) || tc == 2) return gemm_tc_run(g, static_cast<cudaStream_t>(stream));
  }
  GemmP p;
  p.g = *g;
  auto vec_ok = [](const float* ptr, int ld) { return ptr != nullptr && (ld % 4 == 0) && aligned16(ptr); };
  p.a_vec = vec_ok(g->a.ptr, g->a.ld);
  if (g->a.op == RC_OP_AFFINE2) p.a_vec = p.a_vec && vec_ok(g->a.aux, g->a.ld_aux);
  p.b_vec = vec_ok(g->b.ptr, g->b.ld);
  if (g->b.op == RC_OP_AFFINE2) p.b_vec = p.b_vec && vec_ok(g->b.aux, g->b.ld_aux);
  p.a2_vec = g->k2 > 0 ? vec_ok(g->a2, g->lda2) : 0;
  p.b2_vec = g->k2 > 0 ? vec_ok(g->b2, g->ldb2) : 0;
  p.d_vec = vec_ok(g->d, g->ldd) && (g->splits <= 1 || g->split_stride % 4 == 0);
  const int rm = choose_rm(g);
  const int rk = choose_rk(g, rm);
  p.tiles1 = ceil_div(g->k, rk);
  p.tiles2 = g->k2 > 0 ? ceil_div(g->k2, rk) : 0;
  static const bool early_ok = !(getenv("RC_GEMM_B_EARLY") && getenv("RC_GEMM_B_EARLY")[0] == '0');   // A/B switch
  p.b_early = (early_ok && g->b_static && g->b.op == RC_OP_NONE) ? 1 : 0;
  dim3 grid(ceil_div(g->m, 8 * rm), ceil_div(g->n, kBN), g->splits > 1 ? g->splits : 1);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g->a_layout == RC_A_ROW && g->b_layout == RC_B_COL) return gemm_dispatch_rm<RC_A_ROW, RC_B_COL>(rm, rk, p, grid, s);
  if (g->a_layout == RC_A_ROW && g->b_layout == RC_B_RED) return gemm_dispatch_rm<RC_A_ROW, RC_B_RED>(rm, rk, p, grid, s);
  if (g->a_layout == RC_A_RED && g->b_layout == RC_B_RED) return gemm_dispatch_rm<RC_A_RED, RC_B_RED>(rm, rk, p, grid, s);
  return fail(RC_ERR_ARG, "rc_gemm_run: layout combination (A red-major, B col) is not instantiated");
}
