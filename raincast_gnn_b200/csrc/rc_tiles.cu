// Station tiles: clusters of rows of a gather matrix (the dst-sorted CSR for the GINE forward, the src-sorted
// transpose for its backward) whose gathered rows fit in one CTA's shared memory, with the rows of a cluster
// grouped three by three so that a warp reads each gathered row from shared memory once per group.
//
// PyG's GINEConv gathers x[edge_index[0]] edge by edge (call site models/gnn.py:27-29); a station graph is a
// radius graph (utils/data.py:261-284), so rows that are neighbours in the graph share most of their sources.
//   - tiles: grown by breadth-first search over the graph itself (no coordinates needed): a list of rows plus the
//     union of their sources, the union capped at `max_src` staged rows.  Batched reference graphs (block-diagonal,
//     122 stations each) come out as one tile per graph with an empty halo.
//   - row groups: inside a tile, a row is grouped with the two ungrouped neighbours that share most sources with it.
//     The group's edges are stored per DISTINCT source ("entry": staged-row offset + one attribute per row of the
//     group), sorted by which rows of the group use that source (7 classes).  The kernels run one straight-line
//     loop per class: one 16-byte record load and one row load per entry, no per-edge test, and 2.2 edges per row
//     load on the config-4 graph.
//
// Layout written (tile-major; everything a CTA needs for one tile is two contiguous ranges):
//   tile_stage_ptr[T+1]  range into stage_id                      stage_id[...]  rows to stage: the tile's own rows
//                                                                                 (group by group), then its halo
//   tile_blk_ptr[T+1]    range into blocks, in 16-byte units      blocks[...]    per tile, 16-byte records:
//        header  {rows owned, rows staged, groups, edges}
//        three records per group
//            {node id of row 0, 1, 2 (-1: no such row), byte offset of the group's first entry inside the block}
//            {n1 | n2 << 16, n3 | n4 << 16, n5 | n6 << 16, n7}     entries per class; class = bit mask of the rows
//            {degree of row 0, 1, 2 as float, byte offset of row 0 in the staged rows (row k: + k * row_bytes)}
//        entries {staged-row index * row_bytes, attr for row 0, 1, 2 (float32 bits; 0 where the row has no such
//        edge)}, class 1 first, class 7 last; groups with most edges first (warps claim groups in order).
// A (source, row) pair that occurs more than once (multi-edge) gets one entry per occurrence.
#include <string.h>

#include <algorithm>
#include <vector>

#include "rc_common.cuh"

namespace {

struct Entry {
  int32_t loc;       // staged-row index
  int32_t src;       // node id (before loc is known)
  int32_t mask;
  float a[3];
};

struct Group {
  int32_t row[3];
  int32_t nrow;
  int32_t first_entry, n_entry;   // into the tile's entry list
  int32_t pairs;
  int32_t cnt[8];
};

}  // namespace

extern "C" int rc_gine_tiles_build_host(const int32_t* rowptr, const int32_t* col, const float* attr, int num_nodes,
                                        int64_t n_edges, int max_src, int max_block_bytes, int row_bytes,
                                        int32_t* tile_stage_ptr, int32_t* tile_blk_ptr, int32_t* stage_id, int32_t* blocks,
                                        int32_t* n_tiles_out, int64_t* n_staged_out, int64_t* n_block_units_out,
                                        int32_t* max_staged_out, int32_t* max_block_bytes_out, int64_t* n_entries_out) {
  using rc::fail;
  if (!rowptr || !col || !attr || !tile_stage_ptr || !tile_blk_ptr || !stage_id || !blocks || !n_tiles_out || !n_staged_out ||
      !n_block_units_out || !max_staged_out || !max_block_bytes_out || !n_entries_out || num_nodes < 0 || n_edges < 0 ||
      max_src < 1 || max_block_bytes < 80 || row_bytes < 512 || row_bytes % 512)
    return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: bad argument");
  const int n = num_nodes;
  if (rowptr[0] != 0 || (n > 0 && rowptr[n] != n_edges)) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: rowptr does not span n_edges");
  if ((int64_t)max_src * row_bytes > INT32_MAX) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: max_src * row_bytes overflows");
  for (int64_t s = 0; s < n_edges; ++s)
    if (col[s] < 0 || col[s] >= n) return fail(RC_ERR_GRAPH, "rc_gine_tiles_build_host: column id %d outside [0, %d)", col[s], n);
  std::vector<int32_t> tile_of(n, -1), stamp(n, -1), qstamp(n, -1), loc(n, 0), gstamp(n, -1), smark(n, -1), ent_of(n, -1),
      estamp(n, -1);
  std::vector<int32_t> queue, rows, srcs, seeds, singles, order;
  std::vector<Entry> entries, gent;
  std::vector<Group> groups;
  queue.reserve(1024); rows.reserve(1024); srcs.reserve(1024);
  size_t seed_head = 0;
  int next_unassigned = 0;
  int tid = 0, gen = 0, mark = 0;
  int64_t n_staged = 0, units = 0, n_entries = 0;     // units: 16-byte records written to `blocks`
  int max_staged = 0, max_blk = 0;
  tile_stage_ptr[0] = 0;
  tile_blk_ptr[0] = 0;
  // estimate of the bytes a row adds to its tile's block while the tile grows (its share of a group + one entry per
  // two edges); the exact size is known after grouping, and a tile that comes out too large gives rows back
  auto row_block_bytes = [&](int v) { return 16 + 8 * (rowptr[v + 1] - rowptr[v]); };

  // groups + entries of the rows in `rows` (tile `tid`); returns the block size in bytes
  auto group_tile = [&]() -> int {
    groups.clear(); entries.clear(); singles.clear();
    auto close_group = [&](const int32_t* r, int nr) {
      Group g{};
      g.nrow = nr;
      for (int k = 0; k < 3; ++k) g.row[k] = k < nr ? r[k] : -1;
      g.first_entry = (int32_t)entries.size();
      ++mark;
      gent.clear();
      for (int k = 0; k < nr; ++k) {
        const int v = r[k];
        for (int s = rowptr[v]; s < rowptr[v + 1]; ++s) {
          const int u = col[s];
          int e = -1;
          if (estamp[u] == mark) {
            // first entry of this source, then (multi-edges only) later ones
            for (int j = ent_of[u]; j < (int)gent.size(); ++j)
              if (gent[j].src == u && !(gent[j].mask & (1 << k))) { e = j; break; }
          } else {
            estamp[u] = mark;
            ent_of[u] = (int32_t)gent.size();
          }
          if (e < 0) {
            e = (int)gent.size();
            gent.push_back(Entry{0, u, 0, {0.f, 0.f, 0.f}});
          }
          gent[e].mask |= 1 << k;
          gent[e].a[k] = attr[s];
          ++g.pairs;
        }
      }
      for (const Entry& e : gent) ++g.cnt[e.mask];
      g.n_entry = (int32_t)gent.size();
      entries.insert(entries.end(), gent.begin(), gent.end());
      groups.push_back(g);
    };
    for (int v : rows) {
      if (gstamp[v] == tid) continue;
      ++mark;
      for (int s = rowptr[v]; s < rowptr[v + 1]; ++s) smark[col[s]] = mark;
      // the two ungrouped rows of this tile, among v's sources, that share most sources with v
      int best[2] = {-1, -1}, score[2] = {-1, -1};
      for (int s = rowptr[v]; s < rowptr[v + 1]; ++s) {
        const int u = col[s];
        if (u == v || tile_of[u] != tid || gstamp[u] == tid || u == best[0] || u == best[1]) continue;
        int sc = 0;
        for (int q = rowptr[u]; q < rowptr[u + 1]; ++q) sc += smark[col[q]] == mark;
        if (sc > score[0]) { best[1] = best[0]; score[1] = score[0]; best[0] = u; score[0] = sc; }
        else if (sc > score[1]) { best[1] = u; score[1] = sc; }
      }
      if (best[0] < 0) { singles.push_back(v); gstamp[v] = tid; continue; }
      int32_t r[3] = {v, best[0], best[1]};
      const int nr = best[1] < 0 ? 2 : 3;
      for (int k = 0; k < nr; ++k) gstamp[r[k]] = tid;
      close_group(r, nr);
    }
    for (size_t i = 0; i < singles.size(); i += 3)
      close_group(singles.data() + i, (int)std::min<size_t>(3, singles.size() - i));
    return 16 + 48 * (int)groups.size() + 16 * (int)entries.size();
  };

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
    ++gen;
    int nsrc = 0, blk_bytes = 16;
    size_t head = 0;
    queue.push_back(seed);
    qstamp[seed] = tid;
    while (true) {
      for (; head < queue.size() && nsrc < max_src; ++head) {
        const int v = queue[head];
        if (tile_of[v] >= 0) continue;
        const int b = rowptr[v], e = rowptr[v + 1];
        int extra = stamp[v] != gen ? 1 : 0;
        for (int s = b; s < e; ++s) {
          const int u = col[s];
          if (stamp[u] != gen && u != v) {
            // duplicates inside one row (multi-edges) must count once: mark provisionally with -2 - gen
            if (stamp[u] != -2 - gen) { stamp[u] = -2 - gen; ++extra; }
          }
        }
        const bool fits = nsrc + extra <= max_src && blk_bytes + row_block_bytes(v) <= max_block_bytes;
        for (int s = b; s < e; ++s) {           // settle the provisional marks
          const int u = col[s];
          if (stamp[u] == -2 - gen) {
            if (fits) { stamp[u] = gen; srcs.push_back(u); } else stamp[u] = -1;
          }
        }
        if (!fits) {
          if (rows.empty() && nsrc + extra > max_src)
            return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: row %d gathers %d distinct rows, more than max_src=%d", v, extra, max_src);
          if (rows.empty()) {            // the estimate says no, the exact size decides below
            if (stamp[v] != gen) { stamp[v] = gen; srcs.push_back(v); }
            for (int s = b; s < e; ++s)
              if (stamp[col[s]] != gen) { stamp[col[s]] = gen; srcs.push_back(col[s]); }
            tile_of[v] = tid;
            rows.push_back(v);
            nsrc = max_src;              // close the tile
            break;
          }
          seeds.push_back(v);
          continue;
        }
        if (stamp[v] != gen) { stamp[v] = gen; srcs.push_back(v); }
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
    // ---- group the rows; a block that comes out larger than the estimate gives its last rows back
    int exact = group_tile();
    while (exact > max_block_bytes) {
      if (rows.size() == 1)
        return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: row %d (%d edges) needs a block of %d bytes, more than max_block_bytes=%d",
                    rows[0], rowptr[rows[0] + 1] - rowptr[rows[0]], exact, max_block_bytes);
      const size_t keep = std::max<size_t>(1, std::min(rows.size() - 1, rows.size() * (size_t)max_block_bytes / exact));
      for (size_t i = keep; i < rows.size(); ++i) { tile_of[rows[i]] = -1; seeds.push_back(rows[i]); }
      rows.resize(keep);
      for (int v : rows) gstamp[v] = -1;
      // the staged set shrinks with the rows
      ++gen;
      srcs.clear();
      for (int v : rows) {
        if (stamp[v] != gen) { stamp[v] = gen; srcs.push_back(v); }
        for (int s = rowptr[v]; s < rowptr[v + 1]; ++s)
          if (stamp[col[s]] != gen) { stamp[col[s]] = gen; srcs.push_back(col[s]); }
      }
      exact = group_tile();
    }
    // ---- emit the tile: own rows first, group by group (groups with most edges first, so that warps claiming
    // groups in order finish together), halo after (insertion order)
    const int nrows = (int)rows.size(), ngroups = (int)groups.size();
    order.resize(ngroups);
    for (int i = 0; i < ngroups; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return groups[a].pairs > groups[b].pairs; });
    int r = 0;
    for (int gi : order)
      for (int k = 0; k < groups[gi].nrow; ++k) { loc[groups[gi].row[k]] = r; stage_id[n_staged + r] = groups[gi].row[k]; ++r; }
    int nh = 0;
    for (int u : srcs)
      if (tile_of[u] != tid) { loc[u] = nrows + nh; stage_id[n_staged + nrows + nh] = u; ++nh; }
    int32_t* blk = blocks + 4 * units;
    int tile_edges = 0;
    int eoff = 16 + 48 * ngroups;                 // byte offset of the next group's first entry inside the block
    for (int i = 0; i < ngroups; ++i) {
      Group& g = groups[order[i]];
      int32_t* rec = blk + 4 + 12 * i;
      rec[0] = g.row[0]; rec[1] = g.row[1]; rec[2] = g.row[2]; rec[3] = eoff;
      rec[4] = g.cnt[1] | (g.cnt[2] << 16); rec[5] = g.cnt[3] | (g.cnt[4] << 16); rec[6] = g.cnt[5] | (g.cnt[6] << 16); rec[7] = g.cnt[7];
      for (int k = 0; k < 3; ++k) {
        const float d = g.row[k] >= 0 ? (float)(rowptr[g.row[k] + 1] - rowptr[g.row[k]]) : 0.f;
        memcpy(&rec[8 + k], &d, sizeof(float));
      }
      rec[11] = loc[g.row[0]] * row_bytes;
      Entry* eb = entries.data() + g.first_entry;
      for (int j = 0; j < g.n_entry; ++j) eb[j].loc = loc[eb[j].src];
      std::stable_sort(eb, eb + g.n_entry, [](const Entry& a, const Entry& b) { return a.mask != b.mask ? a.mask < b.mask : a.loc < b.loc; });
      int32_t* ed = blk + eoff / 4;
      for (int j = 0; j < g.n_entry; ++j) {
        ed[4 * j] = eb[j].loc * row_bytes;
        memcpy(&ed[4 * j + 1], eb[j].a, 3 * sizeof(float));
      }
      for (int c = 1; c < 8; ++c)
        if (g.cnt[c] > 0xffff) return fail(RC_ERR_ARG, "rc_gine_tiles_build_host: more than 65535 entries of one class in a group");
      eoff += 16 * g.n_entry;
      tile_edges += g.pairs;
    }
    blk[0] = nrows; blk[1] = nrows + nh; blk[2] = ngroups; blk[3] = tile_edges;
    n_staged += nrows + nh;
    units += eoff / 16;
    n_entries += (int64_t)entries.size();
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
  *n_entries_out = n_entries;
  return RC_OK;
}

// Walks the tiles the way the kernels do and checks them against the CSR they were built from: every row owned by
// exactly one group, every (source, attr bits) of a row present in the group's entries exactly as often as in the
// CSR row, offsets inside the staged range, class counts and degrees consistent.
extern "C" int rc_gine_tiles_verify_host(const int32_t* rowptr, const int32_t* col, const float* attr, int num_nodes,
                                         int64_t n_edges, int n_tiles, int max_staged, int max_block_bytes, int row_bytes,
                                         const int32_t* tile_stage_ptr, const int32_t* tile_blk_ptr, const int32_t* stage_id,
                                         const int32_t* blocks) {
  using rc::fail;
  if (!rowptr || !tile_stage_ptr || !tile_blk_ptr || !stage_id || !blocks || num_nodes < 0 || n_tiles < 0 || row_bytes <= 0)
    return fail(RC_ERR_ARG, "rc_gine_tiles_verify_host: bad argument");
  std::vector<char> seen(num_nodes, 0);
  std::vector<std::pair<int32_t, uint32_t>> want, got;
  int64_t edges = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int32_t* stage = stage_id + tile_stage_ptr[t];
    const int nst = tile_stage_ptr[t + 1] - tile_stage_ptr[t];
    const int32_t* blk = blocks + 4 * (int64_t)tile_blk_ptr[t];
    const int blk_bytes = 16 * (tile_blk_ptr[t + 1] - tile_blk_ptr[t]);
    if (nst > max_staged || blk_bytes > max_block_bytes) return fail(RC_ERR_GRAPH, "tile %d: %d staged rows / %d block bytes exceed the stated maxima", t, nst, blk_bytes);
    if (blk[1] != nst) return fail(RC_ERR_GRAPH, "tile %d: header says %d staged rows, range says %d", t, blk[1], nst);
    const int nrows = blk[0], ngroups = blk[2];
    int owned = 0, tile_edges = 0, prev_pairs = INT32_MAX;
    for (int i = 0; i < ngroups; ++i) {
      const int32_t* rec = blk + 4 + 12 * i;
      const int cnt[8] = {0, rec[4] & 0xffff, (int)((uint32_t)rec[4] >> 16), rec[5] & 0xffff, (int)((uint32_t)rec[5] >> 16),
                          rec[6] & 0xffff, (int)((uint32_t)rec[6] >> 16), rec[7]};
      int off = rec[3], pairs = 0;
      for (int k = 0; k < 3; ++k) {
        const int v = rec[k];
        if (v < 0) continue;
        if (v >= num_nodes || seen[v]) return fail(RC_ERR_GRAPH, "tile %d group %d: row %d out of range or owned twice", t, i, v);
        seen[v] = 1;
        const int self = rec[11] / row_bytes;
        int kk = 0;                                   // rows of a group are staged consecutively
        for (int j = 0; j < k; ++j) kk += rec[j] >= 0;
        if (rec[11] % row_bytes || self + kk >= nrows || stage[self + kk] != v)
          return fail(RC_ERR_GRAPH, "tile %d group %d: row %d is not staged where the record says", t, i, v);
        float d;
        memcpy(&d, &rec[8 + k], sizeof(float));
        if (d != (float)(rowptr[v + 1] - rowptr[v])) return fail(RC_ERR_GRAPH, "tile %d group %d: degree of row %d", t, i, v);
        ++owned;
      }
      for (int k = 0; k < 3; ++k) {
        const int v = rec[k];
        got.clear(); want.clear();
        int o = off;
        for (int c = 1; c < 8; ++c)
          for (int j = 0; j < cnt[c]; ++j, o += 16) {
            if (o + 16 > blk_bytes) return fail(RC_ERR_GRAPH, "tile %d group %d: entries run past the block", t, i);
            const int32_t* e = blk + o / 4;
            if (e[0] % row_bytes || e[0] < 0 || e[0] / row_bytes >= nst) return fail(RC_ERR_GRAPH, "tile %d group %d: staged offset %d", t, i, e[0]);
            if (c & (1 << k)) {
              if (v < 0) return fail(RC_ERR_GRAPH, "tile %d group %d: class %d uses an absent row", t, i, c);
              got.emplace_back(stage[e[0] / row_bytes], (uint32_t)e[1 + k]);
            }
          }
        if (v < 0) continue;
        for (int s = rowptr[v]; s < rowptr[v + 1]; ++s) {
          uint32_t bits;
          memcpy(&bits, &attr[s], 4);
          want.emplace_back(col[s], bits);
        }
        std::sort(got.begin(), got.end());
        std::sort(want.begin(), want.end());
        if (got != want) return fail(RC_ERR_GRAPH, "tile %d group %d: edges of row %d differ from the CSR row", t, i, v);
        pairs += (int)got.size();
      }
      if (pairs > prev_pairs) return fail(RC_ERR_GRAPH, "tile %d: groups are not ordered by edge count", t);
      prev_pairs = pairs;
      tile_edges += pairs;
    }
    if (owned != nrows || tile_edges != blk[3]) return fail(RC_ERR_GRAPH, "tile %d: %d rows / %d edges in groups, header says %d / %d", t, owned, tile_edges, nrows, blk[3]);
    for (int r = 0; r < nst; ++r)
      if (stage[r] < 0 || stage[r] >= num_nodes) return fail(RC_ERR_GRAPH, "tile %d: staged id %d", t, stage[r]);
    edges += tile_edges;
  }
  for (int v = 0; v < num_nodes; ++v)
    if (!seen[v]) return fail(RC_ERR_GRAPH, "row %d is in no tile", v);
  if (edges != n_edges) return fail(RC_ERR_GRAPH, "tiles hold %lld edges, the CSR %lld", (long long)edges, (long long)n_edges);
  return RC_OK;
}
