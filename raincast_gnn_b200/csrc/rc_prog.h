// Step-program recorder (host side).  While a thread is recording (rc_prog_begin .. rc_prog_end) the library's
// entry points do not launch: they append (op type, variant, virtual grid, parameter block) to a program that
// rc_prog_run later executes inside ONE persistent cooperative kernel with grid-wide barriers between phases.
#pragma once
#include <stddef.h>

#include "rc_common.cuh"

namespace rc {

enum OpType {
  OP_GEMM = 1, OP_GINE_FWD, OP_GINE_BWD, OP_GINE_FIN, OP_BN_STATS_FIN, OP_BN_EVAL_PREP, OP_BN_BWD_FIN, OP_REDUCE,
  OP_DS_FWD, OP_DS_BWD, OP_CRPS_COUNT, OP_CRPS_MAIN, OP_CRPS_FINAL, OP_ADAMW_TICK, OP_ADAMW, OP_NOP
};

constexpr int kOpParamBytes = 832;

struct alignas(16) Op {
  int type, variant, phase, gx, gy, gz;
  int tile_begin;          // first virtual CTA of this op inside its phase
  int smem_bytes;
  alignas(16) unsigned char params[kOpParamBytes];
};

bool recording();
// Appends an op; RC_ERR_ARG if the op cannot be part of a program (unsupported variant, parameter block too big).
int record_op(int type, int variant, dim3 grid, size_t smem_bytes, const void* params, size_t bytes);

}  // namespace rc
