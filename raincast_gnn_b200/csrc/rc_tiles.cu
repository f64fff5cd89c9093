// Station tiles: clusters of rows of a gather matrix (the dst-sorted CSR for the GINE forward, the src-sorted
// transpose for its backward) whose gathered rows fit in one CTA's shared memory.
//
// PyG's GINEConv gathers x[edge_index[0]] edge by edge (call site models/gnn.py:27-29); a station graph is a
// radius graph (utils/data.py:261-284), so rows that are neighbours in the graph share most of their sources.
// The tiler grows clusters by breadth-first search over the graph itself (no coordinates needed): a tile is a
// list of rows plus the union of their sources, the union capped at `max_src` staged rows.  Batched reference
// graphs (block-diagonal, 122 stations each) come out as one tile per graph with an empty halo; the 100k-node
// config-4 graph comes out with ~2.2 staged rows per owned row at max_src = 224.
//
// Layout written (tile-major; everything a CTA needs for one tile is two contiguous ranges, so a tile is staged with
// one bulk copy of its block plus one bulk copy per gathered row):
//   tile_stage_ptr[T+1]  range into stage_id                       stage_id[...]  rows to stage: the tile's own rows
//                                                                                  (in tile order), then its halo
//   tile_blk_ptr[T+1]    range into blocks, in 16-byte units       blocks[...]    per tile, 16-byte records:
//        header  {rows owned, rows staged, edges, 0}
//        one record per owned row  {node id, byte offset of its first edge inside the block, degree, 0}
//        edge records, two per 16 bytes: {byte offset of the staged source row = index * row_bytes, attr bits};
//        every row starts on a 16-byte boundary (odd degrees are followed by one unused 8-byte pad)
// Edges of a row keep the CSR slot order (= the reference's edge order), so the kernels sum in the same order as the
// untiled ones.
#include <string.h>

#include <algorithm>
#include <vector>

#include "rc_common.cuh"

extern "C" int rc_gine_tiles_build_host(const int32_t* rowptr, const int32_t* col, const float* attr, int num_nodes,
                                        int64_t n_edges, int max_src, int max_block_bytes, int row_bytes,
                                        int32_t* tile_stage_ptr, int32_t* tile_blk_ptr, int32_t* stage_id, int32_t* blocks,
                                        int32_t* n_tiles_out, int64_t* n_staged_out, int64_t* n_block_units_out,
                                        int32_t* max_staged_out, int32_t* max_block_bytes_out) {
  using rc::fail;
  if (!rowptr || !col || !attr || !tile_stage_ptr || !tile_blk_ptr || !stage_id || !blocks || !n_tiles_out || !n_staged_out ||
      !n_block_units_out || !max_staged_out || !max_block_bytes_out || num_nodes < 0 || n_edges < 0 || max_src < 1 ||
      max_block_bytes < 48 || row_bytes < 16 || row_bytes % 16)
    return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: bad argument");
  const int n = num_nodes;
  if (rowptr[0] != 0 || (n > 0 && rowptr[n] != n_edges)) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: rowptr does not span n_edges");
  if ((int64_t)max_src * row_bytes > INT32_MAX) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: max_src * row_bytes overflows");
  std::vector<int32_t> tile_of(n, -1), stamp(n, -1), qstamp(n, -1), loc(n, 0);
  std::vector<int32_t> queue, rows, srcs, seeds;
  queue.reserve(1024); rows.reserve(1024); srcs.reserve(1024);
  size_t seed_head = 0;
  int next_unassigned = 0;
  int tid = 0;
  int64_t n_staged = 0, units = 0;     // units: 16-byte records written to `blocks`
  int max_staged = 0, max_blk = 0;
  tile_stage_ptr[0] = 0;
  tile_blk_ptr[0] = 0;
  // bytes a row adds to its tile's block: its record + its edges padded to a whole number of 16-byte records
  auto row_block_bytes = [&](int v) { return 16 + 16 * ((rowptr[v + 1] - rowptr[v] + 1) / 2); };
  while (true) {
    // ---- seed: a frontier row an earlier tile could not take (keeps tiles packed against each other), else the
    // lowest unassigned id
    int seed = -1;
    while (seed_head < seeds.size()) {
      const int c = seeds[seed_head++];
      if (tile_of[c] < 0) { seed = c; break; }
    }
    if (seed < 0) {
      while (next_unassigned < n && tile_of[next_unassigned] >= 0) ++next_unassigned;
      if (next_unassigned >= n) break;
      seed = next_unassigned;
    }
    rows.clear(); srcs.clear(); queue.clear();
    int nsrc = 0, blk_bytes = 16;
    size_t head = 0;
    queue.push_back(seed);
    qstamp[seed] = tid;
    while (true) {
      for (; head < queue.size() && nsrc < max_src; ++head) {
        const int v = queue[head];
        if (tile_of[v] >= 0) continue;
        const int b = rowptr[v], e = rowptr[v + 1];
        int extra = stamp[v] != tid ? 1 : 0;
        for (int s = b; s < e; ++s) {
          const int u = col[s];
          if (u < 0 || u >= n) return fail(RC_ERR_GRAPH, "rc_gine_tiles_build_host: column id %d outside [0, %d)", u, n);
          if (stamp[u] != tid && u != v) {
            // duplicates inside one row (multi-edges) must count once: mark provisionally with -2 - tid
            if (stamp[u] != -2 - tid) { stamp[u] = -2 - tid; ++extra; }
          }
        }
        const bool fits = nsrc + extra <= max_src && blk_bytes + row_block_bytes(v) <= max_block_bytes;
        for (int s = b; s < e; ++s) {           // settle the provisional marks
          const int u = col[s];
          if (stamp[u] == -2 - tid) {
            if (fits) { stamp[u] = tid; srcs.push_back(u); } else stamp[u] = -1;
          }
        }
        if (!fits) {
          if (rows.empty())
            return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: row %d (%d edges, %d distinct rows) exceeds max_src=%d / max_block_bytes=%d",
                        v, e - b, extra, max_src, max_block_bytes);
          seeds.push_back(v);
          continue;
        }
        if (stamp[v] != tid) { stamp[v] = tid; srcs.push_back(v); }
        nsrc += extra;
        blk_bytes += row_block_bytes(v);
        tile_of[v] = tid;
        rows.push_back(v);
        for (int s = b; s < e; ++s) {
          const int u = col[s];
          if (tile_of[u] < 0 && qstamp[u] != tid) { qstamp[u] = tid; queue.push_back(u); }
        }
      }
      // rows still queued when the tile filled up seed later tiles
      for (; head < queue.size(); ++head)
        if (tile_of[queue[head]] < 0) seeds.push_back(queue[head]);
      // a component that ended early: small ones share a tile with the next component
      if (nsrc * 4 >= max_src) break;
      while (next_unassigned < n && tile_of[next_unassigned] >= 0) ++next_unassigned;
      int cand = next_unassigned;                 // skip rows this tile already tried and could not take
      while (cand < n && (tile_of[cand] >= 0 || qstamp[cand] == tid)) ++cand;
      if (cand >= n) break;
      queue.push_back(cand);
      qstamp[cand] = tid;
    }
    // ---- emit the tile: own rows first (longest row first, so that warps claiming rows in order finish together),
    // halo after (insertion order)
    const int nrows = (int)rows.size();
    std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) { return rowptr[a + 1] - rowptr[a] > rowptr[b + 1] - rowptr[b]; });
    for (int r = 0; r < nrows; ++r) { loc[rows[r]] = r; stage_id[n_staged + r] = rows[r]; }
    int nh = 0;
    for (int u : srcs)
      if (tile_of[u] != tid) { loc[u] = nrows + nh; stage_id[n_staged + nrows + nh] = u; ++nh; }
    int32_t* blk = blocks + 4 * units;
    int tile_edges = 0;
    int eoff = 16 + 16 * nrows;                   // byte offset of the next row's first edge inside the block
    for (int r = 0; r < nrows; ++r) {
      const int v = rows[r];
      const int deg = rowptr[v + 1] - rowptr[v];
      int32_t* rec = blk + 4 + 4 * r;
      rec[0] = v; rec[1] = eoff; rec[2] = deg; rec[3] = 0;
      int32_t* ed = blk + eoff / 4;
      for (int k = 0; k < deg; ++k) {
        const int s = rowptr[v] + k;
        ed[2 * k] = loc[col[s]] * row_bytes;
        memcpy(&ed[2 * k + 1], &attr[s], sizeof(float));
      }
      if (deg & 1) { ed[2 * deg] = 0; ed[2 * deg + 1] = 0; }
      eoff += 16 * ((deg + 1) / 2);
      tile_edges += deg;
    }
    blk[0] = nrows; blk[1] = nrows + nh; blk[2] = tile_edges; blk[3] = 0;
    n_staged += nrows + nh;
    units += eoff / 16;
    if (nrows + nh > max_staged) max_staged = nrows + nh;
    if (eoff > max_blk) max_blk = eoff;
    ++tid;
    tile_stage_ptr[tid] = (int32_t)n_staged;
    tile_blk_ptr[tid] = (int32_t)units;
    if (n_staged > INT32_MAX || units > INT32_MAX) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: graph too large for 32-bit tile offsets");
  }
  *n_tiles_out = tid;
  *n_staged_out = n_staged;
  *n_block_units_out = units;
  *max_staged_out = max_staged;
  *max_block_bytes_out = max_blk;
  return RC_OK;
}
