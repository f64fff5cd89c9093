// Step program: the whole training step as ONE persistent cooperative kernel.
//
// At the reference shape (8 graphs x 122 stations) every kernel of the step is a few microseconds of work; as
// separate launches (even inside a CUDA graph) the step is bound by ~50 dependent launch / drain / cold-start
// gaps.  Here the host records the same entry-point calls into a program (rc_prog_begin .. rc_prog_end) and the
// device runs it in one launch: one CTA pair per SM stays resident, walks the ops of a phase as virtual CTAs
// (the same tile functions the standalone kernels call) and meets the others at a grid-wide barrier before the
// next phase.  Ops recorded on lane 1 (weight / bias gradients, their reductions) share phases with the
// critical-path ops they are independent of, so they fill SMs the small critical-path ops leave idle.
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "rc_crps_tile.cuh"
#include "rc_deepsets_tile.cuh"
#include "rc_gemm_tile.cuh"
#include "rc_gine_tile.cuh"
#include "rc_misc_tile.cuh"
#include "rc_prog.h"

namespace cg = cooperative_groups;

namespace rc {

// ------------------------------------------------------------------------------------------------ recorder
struct Recorder {
  bool active = false;
  int lane = 0;
  int main_phase = -1, side_phase = -1;
  size_t smem_max = 0;
  std::vector<Op> ops;
};
static thread_local Recorder g_rec;

bool recording() { return g_rec.active; }

int record_op(int type, int variant, dim3 grid, size_t smem_bytes, const void* params, size_t bytes) {
  if (!g_rec.active) return fail(RC_ERR_ARG, "record_op outside rc_prog_begin / rc_prog_end");
  if (bytes > (size_t)kOpParamBytes) return fail(RC_ERR_ARG, "record_op: parameter block of %zu bytes", bytes);
  Op op;
  memset(&op, 0, sizeof(op));
  op.type = type; op.variant = variant;
  op.gx = (int)grid.x; op.gy = (int)grid.y; op.gz = (int)grid.z;
  op.smem_bytes = (int)smem_bytes;
  memcpy(op.params, params, bytes);
  if (g_rec.lane == 0) {
    op.phase = ++g_rec.main_phase;
  } else {
    g_rec.side_phase = std::max(g_rec.main_phase + 1, g_rec.side_phase + 1);
    op.phase = g_rec.side_phase;
  }
  g_rec.smem_max = std::max(g_rec.smem_max, smem_bytes);
  g_rec.ops.push_back(op);
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------ interpreter
// Every op body is a separate (non-inlined) device function: each gets its own register allocation under the
// kernel's 128-register cap instead of one allocation for the union of all of them.
#define RC_NOINLINE __device__ __noinline__
// (one CTA per SM when a long-slice GEMM is part of the program: its double-buffered tiles take 139 KB)

template <int RM, int AL, int BL, int RK>
RC_NOINLINE void call_gemm(const GemmP* __restrict__ p, const uint3 bid, float* smem) { gemm_tile<RM, AL, BL, RK>(*p, bid, smem); }

template <int AL, int BL>
__device__ __forceinline__ void run_gemm(const Op& op, const uint3 bid, float* smem) {
  const GemmP* p = reinterpret_cast<const GemmP*>(op.params);
  if (AL == RC_A_ROW && (op.variant & 64)) {      // long reduction slice: one memory round trip for K <= 128
    if ((op.variant & 15) == 1) call_gemm<1, AL, BL, 128>(p, bid, smem);
    else call_gemm<2, AL, BL, 128>(p, bid, smem);
    return;
  }
  switch (op.variant & 15) {
    case 1: call_gemm<1, AL, BL, 32>(p, bid, smem); break;
    case 2: call_gemm<2, AL, BL, 32>(p, bid, smem); break;
    default: call_gemm<4, AL, BL, 32>(p, bid, smem); break;
  }
}

RC_NOINLINE void call_gine_fwd(const GineFwdP* __restrict__ p, uint3 bid, uint3 gdim) { gine_fwd_tile<1>(*p, bid, gdim); }
RC_NOINLINE void call_gine_bwd(const GineBwdP* __restrict__ p, uint3 bid, uint3 gdim, float* smem) { gine_bwd_tile<1>(*p, bid, gdim, smem); }
RC_NOINLINE void call_gine_fin(const GineFinP* __restrict__ p, uint3 bid, uint3 gdim) { gine_fin_tile(*p, bid, gdim); }
RC_NOINLINE void call_bn_stats(const BnStatsFinP* __restrict__ p, uint3 bid, uint3 gdim) { bn_stats_fin_tile(*p, bid, gdim); }
RC_NOINLINE void call_bn_eval(const BnEvalP* __restrict__ p, uint3 bid, uint3 gdim) { bn_eval_tile(*p, bid, gdim); }
RC_NOINLINE void call_bn_bwd(const BnBwdFinP* __restrict__ p, uint3 bid, uint3 gdim) { bn_bwd_fin_tile(*p, bid, gdim); }
RC_NOINLINE void call_reduce(const ReduceP* __restrict__ p, uint3 bid, uint3 gdim) { reduce_tile(*p, bid, gdim); }
RC_NOINLINE void call_ds_fwd(const DsFwdP* __restrict__ p, uint3 bid, uint3 gdim, float* smem) { ds_fwd_tile(*p, bid, gdim, smem); }
template <int KQ>
RC_NOINLINE void call_ds_bwd(const DsBwdP* __restrict__ p, uint3 bid, uint3 gdim, float* smem) { ds_bwd_tile<KQ>(*p, bid, gdim, smem); }
RC_NOINLINE void call_crps_count(const CrpsCountP* __restrict__ p, uint3 bid, uint3 gdim) { crps_count_tile(*p, bid, gdim); }
template <int W>
RC_NOINLINE void call_crps_main(const CrpsMainP* __restrict__ p, uint3 bid, uint3 gdim) { crps_main_tile<W>(*p, bid, gdim); }
RC_NOINLINE void call_crps_final(const CrpsFinalP* __restrict__ p, uint3 bid, uint3 gdim) { crps_final_tile(*p, bid, gdim); }
RC_NOINLINE void call_adamw_tick(const AdamTickP* __restrict__ p, uint3 bid, uint3 gdim) { adamw_tick_tile(*p, bid, gdim); }
RC_NOINLINE void call_adamw(const AdamP* __restrict__ p, uint3 bid, uint3 gdim) { adamw_tile(*p, bid, gdim); }

__device__ __forceinline__ void run_tile(const Op& op, const uint3 bid, float* smem) {
  const uint3 gdim = make_uint3(op.gx, op.gy, op.gz);
  switch (op.type) {
    case OP_GEMM: {
      const int al = (op.variant >> 4) & 1, bl = (op.variant >> 5) & 1;
      if (al == RC_A_ROW && bl == RC_B_COL) run_gemm<RC_A_ROW, RC_B_COL>(op, bid, smem);
      else if (al == RC_A_ROW) run_gemm<RC_A_ROW, RC_B_RED>(op, bid, smem);
      else run_gemm<RC_A_RED, RC_B_RED>(op, bid, smem);
      break;
    }
    case OP_GINE_FWD: call_gine_fwd(reinterpret_cast<const GineFwdP*>(op.params), bid, gdim); break;
    case OP_GINE_BWD: call_gine_bwd(reinterpret_cast<const GineBwdP*>(op.params), bid, gdim, smem); break;
    case OP_GINE_FIN: call_gine_fin(reinterpret_cast<const GineFinP*>(op.params), bid, gdim); break;
    case OP_BN_STATS_FIN: call_bn_stats(reinterpret_cast<const BnStatsFinP*>(op.params), bid, gdim); break;
    case OP_BN_EVAL_PREP: call_bn_eval(reinterpret_cast<const BnEvalP*>(op.params), bid, gdim); break;
    case OP_BN_BWD_FIN: call_bn_bwd(reinterpret_cast<const BnBwdFinP*>(op.params), bid, gdim); break;
    case OP_REDUCE: call_reduce(reinterpret_cast<const ReduceP*>(op.params), bid, gdim); break;
    case OP_DS_FWD: call_ds_fwd(reinterpret_cast<const DsFwdP*>(op.params), bid, gdim, smem); break;
    case OP_DS_BWD:
      if (op.variant == 5) call_ds_bwd<5>(reinterpret_cast<const DsBwdP*>(op.params), bid, gdim, smem);
      else call_ds_bwd<1>(reinterpret_cast<const DsBwdP*>(op.params), bid, gdim, smem);
      break;
    case OP_CRPS_COUNT: call_crps_count(reinterpret_cast<const CrpsCountP*>(op.params), bid, gdim); break;
    case OP_CRPS_MAIN: {
      const CrpsMainP* p = reinterpret_cast<const CrpsMainP*>(op.params);
      switch (op.variant) {
        case 0: call_crps_main<2>(p, bid, gdim); break;
        case 1: call_crps_main<3>(p, bid, gdim); break;
        case 2: call_crps_main<4>(p, bid, gdim); break;
        default: call_crps_main<5>(p, bid, gdim); break;
      }
      break;
    }
    case OP_CRPS_FINAL: call_crps_final(reinterpret_cast<const CrpsFinalP*>(op.params), bid, gdim); break;
    case OP_ADAMW_TICK: call_adamw_tick(reinterpret_cast<const AdamTickP*>(op.params), bid, gdim); break;
    case OP_ADAMW: call_adamw(reinterpret_cast<const AdamP*>(op.params), bid, gdim); break;
    default: break;   // OP_NOP: barrier only
  }
}

// ops are sorted by phase; phase_start[ph] .. phase_start[ph+1] are the ops of phase ph
__global__ void __launch_bounds__(256, 1)
prog_kernel(const Op* __restrict__ ops, const int* __restrict__ phase_start, int n_phases) {
  extern __shared__ __align__(16) float smem[];
  cg::grid_group grid = cg::this_grid();
  for (int ph = 0; ph < n_phases; ++ph) {
    const int lo = phase_start[ph], hi = phase_start[ph + 1];
    const Op& last = ops[hi - 1];
    const int total = last.tile_begin + last.gx * last.gy * last.gz;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int i = lo;
      while (i + 1 < hi && ops[i + 1].tile_begin <= t) ++i;
      const Op& op = ops[i];
      const int local = t - op.tile_begin;
      const uint3 bid = make_uint3(local % op.gx, (local / op.gx) % op.gy, local / (op.gx * op.gy));
      run_tile(op, bid, smem);
      __syncthreads();                 // the next tile reuses this CTA's shared memory
    }
    grid.sync();
  }
}

static bool op_supported(const Op& op) {
  switch (op.type) {
    case OP_GEMM: { const int rm = op.variant & 15; return rm == 1 || rm == 2 || rm == 4; }
    case OP_GINE_FWD: case OP_GINE_BWD: return op.variant == 1;
    case OP_DS_BWD: return op.variant == 5 || op.variant == 1;
    default: return true;
  }
}

}  // namespace rc

using namespace rc;

extern "C" int rc_prog_begin(void) {
  if (g_rec.active) return fail(RC_ERR_ARG, "rc_prog_begin: already recording");
  g_rec = Recorder();
  g_rec.active = true;
  return RC_OK;
}

extern "C" int rc_prog_lane(int lane) {
  if (!g_rec.active) return fail(RC_ERR_ARG, "rc_prog_lane: not recording");
  if (lane != 0 && lane != 1) return fail(RC_ERR_ARG, "rc_prog_lane: lane must be 0 or 1");
  g_rec.lane = lane;
  return RC_OK;
}

extern "C" int rc_prog_nop(void) {     /* an empty phase: measures the cost of one grid-wide barrier */
  int dummy = 0;
  return record_op(OP_NOP, 0, dim3(1), 0, &dummy, sizeof(dummy));
}

extern "C" int rc_prog_join(void) {
  if (!g_rec.active) return fail(RC_ERR_ARG, "rc_prog_join: not recording");
  g_rec.main_phase = std::max(g_rec.main_phase, g_rec.side_phase);
  g_rec.lane = 0;
  return RC_OK;
}

extern "C" int rc_prog_abort(void) {
  g_rec = Recorder();
  return RC_OK;
}

extern "C" size_t rc_prog_bytes(void) {
  const int n_phases = std::max(g_rec.main_phase, g_rec.side_phase) + 1;
  return g_rec.ops.size() * sizeof(Op) + (size_t)(n_phases + 1) * sizeof(int) + 16;
}

// Stops recording and uploads the program.  info = {n_ops, n_phases, smem_bytes, ops_bytes (offset of the phase table)}.
extern "C" int rc_prog_end(void* device_buf, size_t bytes, int* info) {
  if (!g_rec.active) return fail(RC_ERR_ARG, "rc_prog_end: not recording");
  g_rec.active = false;
  if (!device_buf || !info) return fail(RC_ERR_ARG, "rc_prog_end: null pointer");
  if (bytes < rc_prog_bytes()) return fail(RC_ERR_WORKSPACE, "rc_prog_end: buffer %zu < %zu", bytes, rc_prog_bytes());
  std::vector<Op>& ops = g_rec.ops;
  for (const Op& op : ops)
    if (!op_supported(op)) return fail(RC_ERR_ARG, "rc_prog_end: op type %d variant %d is not part of the step program", op.type, op.variant);
  std::stable_sort(ops.begin(), ops.end(), [](const Op& a, const Op& b) { return a.phase < b.phase; });
  const int n_phases = ops.empty() ? 0 : ops.back().phase + 1;
  std::vector<int> start(n_phases + 1, 0);
  size_t i = 0;
  for (int ph = 0; ph < n_phases; ++ph) {
    start[ph] = (int)i;
    int tiles = 0;
    while (i < ops.size() && ops[i].phase == ph) {
      ops[i].tile_begin = tiles;
      tiles += ops[i].gx * ops[i].gy * ops[i].gz;
      ++i;
    }
    if (start[ph] == (int)i) return fail(RC_ERR_ARG, "rc_prog_end: empty phase %d", ph);
  }
  start[n_phases] = (int)ops.size();
  const size_t ops_bytes = (ops.size() * sizeof(Op) + 15) & ~(size_t)15;
  cudaError_t e = cudaMemcpy(device_buf, ops.data(), ops.size() * sizeof(Op), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(static_cast<char*>(device_buf) + ops_bytes, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_prog_end: %s", cudaGetErrorString(e));
  info[0] = (int)ops.size(); info[1] = n_phases; info[2] = (int)g_rec.smem_max; info[3] = (int)ops_bytes;
  g_rec = Recorder();
  return RC_OK;
}

extern "C" int rc_prog_run(const void* device_buf, const int* info, void* stream) {
  if (!device_buf || !info) return fail(RC_ERR_ARG, "rc_prog_run: null pointer");
  if (g_rec.active) return fail(RC_ERR_ARG, "rc_prog_run: cannot run while recording");
  const int n_phases = info[1], smem = info[2];
  if (n_phases == 0) return RC_OK;
  static int max_smem_set = -1;
  static int blocks_per_sm = 0, num_sms = 0;
  if (smem > max_smem_set) {
    cudaError_t e = cudaFuncSetAttribute(prog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_prog_run: %s", cudaGetErrorString(e));
    max_smem_set = smem;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, prog_kernel, 256, smem);
    if (blocks_per_sm < 1) return fail(RC_ERR_CUDA, "rc_prog_run: the program kernel does not fit on an SM");
    if (blocks_per_sm > 2) blocks_per_sm = 2;
  }
  const Op* ops = static_cast<const Op*>(device_buf);
  const int* phase_start = reinterpret_cast<const int*>(static_cast<const char*>(device_buf) + info[3]);
  int np = n_phases;
  void* args[] = {(void*)&ops, (void*)&phase_start, (void*)&np};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)prog_kernel, dim3(num_sms * blocks_per_sm), dim3(256), args, (size_t)smem,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_prog_run: %s", cudaGetErrorString(e));
  return check_launch("prog_kernel");
}
