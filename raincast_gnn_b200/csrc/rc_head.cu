// Head Linear + links + CRPS + their backward as ONE kernel for small batches (the reference shape: 976 nodes).
//
//   raw = h W^T + b                       models/gnn.py:123,139 (`aggr`, H -> C)
//   pred = links(raw), loss = mean CRPS   models/model_utils.py:89-113, models/loss.py:203-272 (and :12-68, :346-369)
//   d raw, d h = d raw W, partial d W = d raw^T h, d b = column sums of d raw
//
// In the step these were three dependent launches (head GEMM, one-CTA CRPS kernel, backward-data GEMM) plus a
// weight-gradient GEMM on the side stream; every one of them is a few microseconds of launch and dependency latency around
// a few hundred kFLOP.  Here a warp owns one node: lane l holds columns 4 l .. 4 l + 3 of every 128-column chunk of the
// node's hidden row and of the C weight rows, the C dot products are warp sums, every lane evaluates the node's CRPS and
// gradient (same values, one SIMT pass), and d h is written back in the layout it was read in.  The count of valid targets
// - the denominator of the mean, needed before the first gradient is written - is an input of the step: every CTA counts
// all of y itself, before it waits for the kernel in front of it (programmatic dependent launch), like the weight rows.
// The CTA's weight-gradient partial goes through shared memory ([8 warps][C][H]) and is reduced with the other partial
// sums of the step (rc_reduce_segments); the loss partials are summed in a fixed order by the CTA that finishes last.
#include "rc_common.cuh"
#include "rc_crps_node.cuh"

namespace rc {

constexpr int kHeadRows = 8;          // nodes per CTA (one per warp)
constexpr int kHeadMaxNodes = 16384;  // every CTA counts the valid targets of the whole batch

struct HeadCrpsP {
  const float* h;
  const float* w;
  const float* b;
  const float* y;
  float* d_h;
  float* partials;
  double* loss_partial;
  double* loss_out;
  int* n_valid;
  int m, hidden, kind;
  float u_fixed, xi, t;
};

static __device__ unsigned int g_head_arrivals = 0;

template <int WIDTH, int CH>
__global__ void __launch_bounds__(256) head_crps_kernel(const HeadCrpsP p) {
  constexpr int H = 128 * CH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ float red[kHeadRows][WIDTH][H];
  __shared__ float red_b[kHeadRows][WIDTH];
  __shared__ double sh_loss[kHeadRows];
  __shared__ int sh_cnt[kHeadRows];
  __shared__ int s_cnt, s_last;

  // ---- before the wait: parameters and the step's targets
  float4 wv[WIDTH][CH];
#pragma unroll
  for (int c = 0; c < WIDTH; ++c)
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) wv[c][ch] = ldg4(p.w + (size_t)c * H + 128 * ch + 4 * lane);
  const float bias = lane < WIDTH ? __ldg(p.b + lane) : 0.f;
  int cnt = 0;
  for (int i = tid; i < p.m; i += 256) cnt += !isnan(__ldg(p.y + i));
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) sh_cnt[warp] = cnt;
  __syncthreads();
  if (tid == 0) {
    int s = 0;
    for (int w8 = 0; w8 < kHeadRows; ++w8) s += sh_cnt[w8];
    s_cnt = s;
  }
  __syncthreads();
  const int n_valid = s_cnt;
  const float inv_n = n_valid > 0 ? 1.0f / (float)n_valid : 0.0f;
  const int row = blockIdx.x * kHeadRows + warp;
  const float yi = row < p.m ? __ldg(p.y + row) : nanf("");

  pdl_entry();

  float g[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float loss = 0.0f;
  float4 hv[CH];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch) hv[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < p.m) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) hv[ch] = ld4(p.h + (size_t)row * H + 128 * ch + 4 * lane);
    float raw[WIDTH];
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) {
      float s = 0.f;
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        s += hv[ch].x * wv[c][ch].x + hv[ch].y * wv[c][ch].y + hv[ch].z * wv[c][ch].z + hv[ch].w * wv[c][ch].w;
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // (xor: every lane ends with the same sum)
      raw[c] = s + __shfl_sync(0xffffffffu, bias, c);
    }
    if (!isnan(yi)) loss = crps_node_k<WIDTH - 2>(raw, yi, 1, p.u_fixed, p.xi, p.t, g);
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) g[c] *= inv_n;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < WIDTH; ++c) {
        d.x = fmaf(g[c], wv[c][ch].x, d.x); d.y = fmaf(g[c], wv[c][ch].y, d.y);
        d.z = fmaf(g[c], wv[c][ch].z, d.z); d.w = fmaf(g[c], wv[c][ch].w, d.w);
      }
      st4(p.d_h + (size_t)row * H + 128 * ch + 4 * lane, d);
    }
  }
  // ---- this CTA's weight / bias gradient partial and loss partial (rows beyond m contribute zeros)
#pragma unroll
  for (int c = 0; c < WIDTH; ++c)
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
      st4(&red[warp][c][128 * ch + 4 * lane], make_float4(g[c] * hv[ch].x, g[c] * hv[ch].y, g[c] * hv[ch].z, g[c] * hv[ch].w));
  if (lane < WIDTH) {
    float gl = 0.f;
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) gl = lane == c ? g[c] : gl;
    red_b[warp][lane] = gl;
  }
  if (lane == 0) sh_loss[warp] = (double)loss;
  __syncthreads();
  float* part = p.partials + (size_t)blockIdx.x * (WIDTH * H + WIDTH);
  for (int e = tid; e < WIDTH * H + WIDTH; e += 256) {
    float s = 0.f;
    if (e < WIDTH * H) {
#pragma unroll
      for (int w8 = 0; w8 < kHeadRows; ++w8) s += red[w8][e / H][e % H];
    } else {
#pragma unroll
      for (int w8 = 0; w8 < kHeadRows; ++w8) s += red_b[w8][e - WIDTH * H];
    }
    part[e] = s;
  }
  if (tid == 0) {
    double s = 0.0;
    for (int w8 = 0; w8 < kHeadRows; ++w8) s += sh_loss[w8];
    p.loss_partial[blockIdx.x] = s;
    __threadfence();
    s_last = atomicAdd(&g_head_arrivals, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- the last CTA: mean over the valid nodes (fixed order: thread t sums partials t, t + 256, ..; thread 0 sums the threads)
  __threadfence();
  __shared__ double sh_tot[256];
  double s = 0.0;
  for (int i = tid; i < (int)gridDim.x; i += 256) s += __ldcg(p.loss_partial + i);
  sh_tot[tid] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 256; ++i) tot += sh_tot[i];
    p.loss_out[0] = n_valid > 0 ? tot / (double)n_valid : nan("");      // mean of an empty selection is NaN in torch too
    p.n_valid[0] = n_valid;
    g_head_arrivals = 0;
  }
}

template <int WIDTH>
static void head_launch(const HeadCrpsP& p, int blocks, cudaStream_t s) {
  if (p.hidden == 128) launch_pdl(head_crps_kernel<WIDTH, 1>, dim3(blocks), dim3(256), 0, s, p);
  else                 launch_pdl(head_crps_kernel<WIDTH, 2>, dim3(blocks), dim3(256), 0, s, p);
}

}  // namespace rc

using namespace rc;

extern "C" int rc_head_crps_blocks(int num_nodes, int hidden) {
  if (num_nodes < 1 || num_nodes > kHeadMaxNodes || (hidden != 128 && hidden != 256)) return 0;   // 0: the fused kernel does not apply
  return ceil_div(num_nodes, kHeadRows);
}

extern "C" int rc_head_crps_fwd_bwd(const float* h, const float* w, const float* b, const float* y, float* d_h, float* partials,
                                    double* loss_partials, double* loss_out, int32_t* n_valid, int num_nodes, int hidden, int kind,
                                    float u_fixed, float xi, float t, void* stream) {
  if (!h || !w || !b || !y || !d_h || !partials || !loss_partials || !loss_out || !n_valid)
    return fail(RC_ERR_ARG, "rc_head_crps_fwd_bwd: null pointer");
  if (kind < 0 || kind > 3) return fail(RC_ERR_ARG, "rc_head_crps_fwd_bwd: kind %d", kind);
  if (kind >= 2 && (xi == 1.0f || xi == 2.0f || xi == 0.0f)) return fail(RC_ERR_ARG, "rc_head_crps_fwd_bwd: xi must not be 0, 1 or 2");
  const int blocks = rc_head_crps_blocks(num_nodes, hidden);
  if (blocks == 0) return fail(RC_ERR_ARG, "rc_head_crps_fwd_bwd: needs 1 <= nodes <= %d and hidden 128 or 256 (got %d, %d)", kHeadMaxNodes, num_nodes, hidden);
  if (!aligned16(h) || !aligned16(w) || !aligned16(d_h)) return fail(RC_ERR_ARG, "rc_head_crps_fwd_bwd: h, w and d_h must be 16-byte aligned");
  const HeadCrpsP p{h, w, b, y, d_h, partials, loss_partials, loss_out, n_valid, num_nodes, hidden, kind, u_fixed, xi, t};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (kind) {
    case 0: head_launch<2>(p, blocks, s); break;
    case 1: head_launch<3>(p, blocks, s); break;
    case 2: head_launch<4>(p, blocks, s); break;
    default: head_launch<5>(p, blocks, s); break;
  }
  return check_launch("head_crps_kernel");
}
