// Station-graph construction and the dst-sorted CSR / transpose / reverse-edge layout.
// Host builders (used by the dataset/loader, which batches on the CPU like the reference's PyG
// DataLoader, train.py:155-156) and a device builder with identical, bit-exact output.
#include <math.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "rc_common.cuh"

using namespace rc;

// ------------------------------------------------------------------------------------------------ host: radius graph
// utils/data.py:261-284: selection D[i,j] <= max_dist (diagonal excluded), row-major order.
namespace {

template <class Visit>
void dense_visit(const float* dist, int n, float max_dist, Visit&& visit) {
  for (int i = 0; i < n; ++i) {
    const float* row = dist + (size_t)i * n;
    for (int j = 0; j < n; ++j)
      if (j != i && row[j] <= max_dist) visit(i, j, row[j]);
  }
}

struct CellGrid {
  double x0, y0, cell;
  int nx, ny;
  std::vector<int> start, items;   // nodes of a cell in ascending id order
  int cell_of(double v, double lo, int cnt) const {
    int c = (int)floor((v - lo) / cell);
    return c < 0 ? 0 : (c >= cnt ? cnt - 1 : c);
  }
  CellGrid(const double* xy, int n, double radius) {
    double x1 = -std::numeric_limits<double>::infinity(), y1 = x1;
    x0 = y0 = std::numeric_limits<double>::infinity();
    for (int i = 0; i < n; ++i) {
      x0 = std::min(x0, xy[2 * i]); x1 = std::max(x1, xy[2 * i]);
      y0 = std::min(y0, xy[2 * i + 1]); y1 = std::max(y1, xy[2 * i + 1]);
    }
    cell = radius * 1.001 + 1e-12;
    // cap the grid at ~4 cells per node so that tiny radii do not explode the table
    double span = std::max(x1 - x0, y1 - y0);
    double min_cell = span / (2.0 * sqrt((double)std::max(n, 1)) + 1.0);
    if (cell < min_cell) cell = min_cell;
    nx = std::max(1, (int)floor((x1 - x0) / cell) + 1);
    ny = std::max(1, (int)floor((y1 - y0) / cell) + 1);
    start.assign((size_t)nx * ny + 1, 0);
    for (int i = 0; i < n; ++i) start[(size_t)cell_of(xy[2 * i + 1], y0, ny) * nx + cell_of(xy[2 * i], x0, nx) + 1]++;
    for (size_t c = 0; c < (size_t)nx * ny; ++c) start[c + 1] += start[c];
    items.resize(n);
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) items[cur[(size_t)cell_of(xy[2 * i + 1], y0, ny) * nx + cell_of(xy[2 * i], x0, nx)]++] = i;
  }
};

template <class Visit>
void coords_visit(const double* xy, int n, double max_dist, Visit&& visit) {
  if (n == 0) return;
  CellGrid grid(xy, n, max_dist);
  const float md = (float)max_dist;   // numpy compares the float32 matrix against a weak python float
  const int reach = std::max(1, (int)ceil(max_dist * 1.001 / grid.cell));
  std::vector<std::pair<int, float>> cand;
  for (int i = 0; i < n; ++i) {
    cand.clear();
    const int cx = grid.cell_of(xy[2 * i], grid.x0, grid.nx), cy = grid.cell_of(xy[2 * i + 1], grid.y0, grid.ny);
    for (int yy = std::max(0, cy - reach); yy <= std::min(grid.ny - 1, cy + reach); ++yy)
      for (int xx = std::max(0, cx - reach); xx <= std::min(grid.nx - 1, cx + reach); ++xx) {
        const size_t c = (size_t)yy * grid.nx + xx;
        for (int k = grid.start[c]; k < grid.start[c + 1]; ++k) {
          const int j = grid.items[k];
          if (j == i) continue;
          const double dx = xy[2 * i] - xy[2 * j], dy = xy[2 * i + 1] - xy[2 * j + 1];
          const float d = (float)sqrt(dx * dx + dy * dy);
          if (d <= md) cand.emplace_back(j, d);
        }
      }
    std::sort(cand.begin(), cand.end());
    for (auto& c : cand) visit(i, c.first, c.second);
  }
}

template <class Walker>
int radius_fill(Walker&& walk, int n, int64_t n_edges, int64_t* edge_index, float* edge_attr, const char* who) {
  float top = -std::numeric_limits<float>::infinity();
  int64_t cnt = 0;
  walk([&](int, int, float d) { top = std::max(top, d); ++cnt; });
  if (cnt + n != n_edges) return fail(RC_ERR_ARG, "%s: n_edges %lld but the graph has %lld", who, (long long)n_edges, (long long)(cnt + n));
  if (cnt == 0) top = 1.0f;                               // utils/data.py:269
  int64_t* src = edge_index;
  int64_t* dst = edge_index + n_edges;
  int64_t e = 0;
  walk([&](int i, int j, float d) {
    src[e] = i; dst[e] = j;
    edge_attr[e] = 1.0f / (d / top);                      // (d / max)^-1 in float32, utils/data.py:272
    ++e;
  });
  for (int i = 0; i < n; ++i, ++e) { src[e] = i; dst[e] = i; edge_attr[e] = 1.0f; }   // self loops, :278-282
  return RC_OK;
}

}  // namespace

extern "C" int rc_radius_graph_count_host(const float* dist, int n, float max_dist, int64_t* n_edges) {
  if (!dist || !n_edges || n < 0) return fail(RC_ERR_ARG, "rc_radius_graph_count_host: bad argument");
  int64_t cnt = 0;
  dense_visit(dist, n, max_dist, [&](int, int, float) { ++cnt; });
  *n_edges = cnt + n;
  return RC_OK;
}

extern "C" int rc_radius_graph_fill_host(const float* dist, int n, float max_dist, int64_t n_edges,
                                         int64_t* edge_index, float* edge_attr) {
  if (!dist || !edge_index || !edge_attr || n < 0) return fail(RC_ERR_ARG, "rc_radius_graph_fill_host: bad argument");
  return radius_fill([&](auto&& v) { dense_visit(dist, n, max_dist, v); }, n, n_edges, edge_index, edge_attr,
                     "rc_radius_graph_fill_host");
}

extern "C" int rc_radius_graph_coords_count_host(const double* xy, int n, double max_dist, int64_t* n_edges) {
  if (!xy || !n_edges || n < 0 || !(max_dist >= 0)) return fail(RC_ERR_ARG, "rc_radius_graph_coords_count_host: bad argument");
  int64_t cnt = 0;
  coords_visit(xy, n, max_dist, [&](int, int, float) { ++cnt; });
  *n_edges = cnt + n;
  return RC_OK;
}

extern "C" int rc_radius_graph_coords_fill_host(const double* xy, int n, double max_dist, int64_t n_edges,
                                                int64_t* edge_index, float* edge_attr) {
  if (!xy || !edge_index || !edge_attr || n < 0) return fail(RC_ERR_ARG, "rc_radius_graph_coords_fill_host: bad argument");
  return radius_fill([&](auto&& v) { coords_visit(xy, n, max_dist, v); }, n, n_edges, edge_index, edge_attr,
                     "rc_radius_graph_coords_fill_host");
}

// ------------------------------------------------------------------------------------------------ host: CSR layout
static bool csr_ptrs_ok(const rc_csr* o, int64_t n_edges) {
  if (!o || !o->rowptr || !o->t_rowptr) return false;
  if (n_edges == 0) return true;   // an empty graph has no per-edge arrays
  return o->col && o->attr && o->perm && o->t_dst && o->t_attr && o->t_perm && o->t_slot && o->rev;
}

extern "C" int rc_csr_build_host(const int64_t* edge_index, const float* edge_attr, int64_t n_edges, int num_nodes,
                                 const rc_csr* out) {
  if (((!edge_index || !edge_attr) && n_edges > 0) || !csr_ptrs_ok(out, n_edges) || n_edges < 0 || num_nodes < 0)
    return fail(RC_ERR_ARG, "rc_csr_build_host: bad argument");
  if (n_edges > 0x7fffffffLL) return fail(RC_ERR_ARG, "rc_csr_build_host: more than 2^31-1 edges");
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + n_edges;
  const int m = num_nodes;
  const int e_n = (int)n_edges;
  for (int e = 0; e < e_n; ++e)
    if (src[e] < 0 || src[e] >= m || dst[e] < 0 || dst[e] >= m)
      return fail(RC_ERR_GRAPH, "rc_csr_build_host: edge %d = (%lld,%lld) outside [0,%d)", e, (long long)src[e], (long long)dst[e], m);
  for (int i = 0; i <= m; ++i) out->rowptr[i] = out->t_rowptr[i] = 0;
  for (int e = 0; e < e_n; ++e) { out->rowptr[dst[e] + 1]++; out->t_rowptr[src[e] + 1]++; }
  for (int i = 0; i < m; ++i) { out->rowptr[i + 1] += out->rowptr[i]; out->t_rowptr[i + 1] += out->t_rowptr[i]; }
  std::vector<int32_t> cur(out->rowptr, out->rowptr + m), tcur(out->t_rowptr, out->t_rowptr + m), inv(e_n);
  for (int e = 0; e < e_n; ++e) {                       // stable counting sort: edge order kept inside a row
    const int p = cur[dst[e]]++;
    out->perm[p] = e; out->col[p] = (int32_t)src[e]; out->attr[p] = edge_attr[e]; inv[e] = p;
    const int q = tcur[src[e]]++;
    out->t_perm[q] = e; out->t_dst[q] = (int32_t)dst[e]; out->t_attr[q] = edge_attr[e];
  }
  for (int q = 0; q < e_n; ++q) out->t_slot[q] = inv[out->t_perm[q]];
  for (int i = 0; i < m; ++i)
    for (int s = out->rowptr[i]; s < out->rowptr[i + 1]; ++s) {
      const int c = out->col[s];
      int r = -1;
      for (int k = out->rowptr[c]; k < out->rowptr[c + 1]; ++k)
        if (out->col[k] == i) { r = k; break; }
      out->rev[s] = r;
    }
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------ device: CSR layout
namespace rc {

__global__ void csr_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int e_n, int m,
                                 int* cnt_dst, int* cnt_src, int* err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_n) return;
  const long long s = src[e], d = dst[e];
  if (s < 0 || s >= m || d < 0 || d >= m) { atomicExch(err, 1); return; }
  atomicAdd(&cnt_dst[d], 1);      // integer atomics: the result does not depend on arrival order
  atomicAdd(&cnt_src[s], 1);
}

// exclusive scan of cnt[0..m) into ptr[0..m]; blockIdx.x selects one of the two arrays.
__global__ void __launch_bounds__(1024) csr_scan_kernel(const int* cnt_a, int* ptr_a, const int* cnt_b, int* ptr_b, int m) {
  const int* cnt = blockIdx.x == 0 ? cnt_a : cnt_b;
  int* ptr = blockIdx.x == 0 ? ptr_a : ptr_b;
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < m; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < m ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;   // exclusive prefix of warp totals
    }
    __syncthreads();
    const int excl = carry + warp_tot[wid] + incl - v;
    if (i < m) ptr[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) ptr[m] = carry;
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int e_n,
                                const int* __restrict__ rowptr, const int* __restrict__ t_rowptr, int* cur_dst,
                                int* cur_src, int* perm, int* t_perm, const int* err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_n || *err) return;   // invalid node ids were found: leave the outputs untouched
  const int d = (int)dst[e], s = (int)src[e];
  perm[rowptr[d] + atomicAdd(&cur_dst[d], 1)] = e;       // arrival order is arbitrary; rows are sorted next
  t_perm[t_rowptr[s] + atomicAdd(&cur_src[s], 1)] = e;
}

// one thread per row (dst rows, then src rows): sort the row's edge ids ascending == stable sort
__global__ void csr_sort_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ t_rowptr, int m, int* perm,
                                     int* t_perm, const int* err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 2 * m || *err) return;
  const int* ptr = r < m ? rowptr : t_rowptr;
  int* arr = r < m ? perm : t_perm;
  const int row = r < m ? r : r - m;
  const int lo = ptr[row], hi = ptr[row + 1];
  for (int i = lo + 1; i < hi; ++i) {
    const int v = arr[i];
    int j = i - 1;
    while (j >= lo && arr[j] > v) { arr[j + 1] = arr[j]; --j; }
    arr[j + 1] = v;
  }
}

__global__ void csr_emit_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                const float* __restrict__ edge_attr, int e_n, const int* __restrict__ perm,
                                const int* __restrict__ t_perm, int* col, float* attr, int* t_dst, float* t_attr, int* inv, const int* err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= e_n || *err) return;
  const int e = perm[p];
  col[p] = (int)src[e];
  attr[p] = edge_attr[e];
  inv[e] = p;
  const int f = t_perm[p];
  t_dst[p] = (int)dst[f];
  t_attr[p] = edge_attr[f];
}

__global__ void csr_link_kernel(const int64_t* __restrict__ dst, int e_n, const int* __restrict__ rowptr,
                                const int* __restrict__ col, const int* __restrict__ perm, const int* __restrict__ t_perm,
                                const int* __restrict__ inv, int* t_slot, int* rev, const int* err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= e_n || *err) return;
  t_slot[p] = inv[t_perm[p]];
  const int i = (int)dst[perm[p]];   // row of slot p
  const int c = col[p];
  int r = -1;
  for (int k = rowptr[c]; k < rowptr[c + 1]; ++k)
    if (col[k] == i) { r = k; break; }
  rev[p] = r;
}

}  // namespace rc

extern "C" size_t rc_csr_build_workspace(int64_t n_edges, int num_nodes) {
  return ((size_t)2 * (num_nodes > 0 ? num_nodes : 0) + (size_t)(n_edges > 0 ? n_edges : 0) + 4) * sizeof(int32_t);
}

extern "C" int rc_csr_build(const int64_t* edge_index, const float* edge_attr, int64_t n_edges, int num_nodes,
                            const rc_csr* out, void* workspace, size_t workspace_bytes, int32_t* err_flag, void* stream) {
  if (((!edge_index || !edge_attr) && n_edges > 0) || !csr_ptrs_ok(out, n_edges) || !workspace || !err_flag || n_edges < 0 || num_nodes < 0)
    return fail(RC_ERR_ARG, "rc_csr_build: bad argument");
  if (n_edges > 0x7fffffffLL) return fail(RC_ERR_ARG, "rc_csr_build: more than 2^31-1 edges");
  if (workspace_bytes < rc_csr_build_workspace(n_edges, num_nodes)) return fail(RC_ERR_WORKSPACE, "rc_csr_build: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int m = num_nodes, e_n = (int)n_edges;
  int* cur_dst = static_cast<int*>(workspace);
  int* cur_src = cur_dst + m;
  int* inv = cur_src + m;
  const int64_t* src = edge_index;
  const int64_t* dst = edge_index + n_edges;
  const int eb = ceil_div(e_n > 0 ? e_n : 1, 256);
  cudaMemsetAsync(cur_dst, 0, (size_t)2 * m * sizeof(int), s);
  cudaMemsetAsync(err_flag, 0, sizeof(int), s);
  csr_count_kernel<<<eb, 256, 0, s>>>(src, dst, e_n, m, cur_dst, cur_src, err_flag);
  if (int e = check_launch("csr_count_kernel")) return e;
  csr_scan_kernel<<<2, 1024, 0, s>>>(cur_dst, out->rowptr, cur_src, out->t_rowptr, m);
  if (int e = check_launch("csr_scan_kernel")) return e;
  cudaMemsetAsync(cur_dst, 0, (size_t)2 * m * sizeof(int), s);
  csr_fill_kernel<<<eb, 256, 0, s>>>(src, dst, e_n, out->rowptr, out->t_rowptr, cur_dst, cur_src, out->perm, out->t_perm, err_flag);
  if (int e = check_launch("csr_fill_kernel")) return e;
  csr_sort_rows_kernel<<<ceil_div(2 * m > 0 ? 2 * m : 1, 256), 256, 0, s>>>(out->rowptr, out->t_rowptr, m, out->perm, out->t_perm, err_flag);
  if (int e = check_launch("csr_sort_rows_kernel")) return e;
  csr_emit_kernel<<<eb, 256, 0, s>>>(src, dst, edge_attr, e_n, out->perm, out->t_perm, out->col, out->attr, out->t_dst, out->t_attr, inv, err_flag);
  if (int e = check_launch("csr_emit_kernel")) return e;
  csr_link_kernel<<<eb, 256, 0, s>>>(dst, e_n, out->rowptr, out->col, out->perm, out->t_perm, inv, out->t_slot, out->rev, err_flag);
  return check_launch("csr_link_kernel");
}
