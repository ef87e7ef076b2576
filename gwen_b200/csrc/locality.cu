// K0r: locality tiles for graphs whose node NUMBERING carries no locality (a mesh with permuted ids, an
// unstructured grid in file order).  The tiled aggregation kernel (aggregate.cu) stages the distinct source
// rows of a destination tile in shared memory once; that only pays when the destinations of a tile share
// their sources, i.e. when a tile is a compact patch of the graph.  This pass finds such patches from the CSR
// alone -- no coordinates -- and hands the tile plan builder a processing order and tile boundaries:
//
//   1. seeds = a maximal independent set of the `radius`-th power of the graph (Luby rounds: a node becomes a
//      seed when its priority -- a bijective hash of its id, so ties cannot happen and the result does not
//      depend on scheduling -- is the largest among the undecided nodes within `radius` hops; everything
//      within `radius` hops of a new seed is then covered).  Seeds are > radius hops apart and every node is
//      within radius hops of one, so on a surface mesh the cells below have bounded, similar sizes.
//   2. cells = Voronoi regions of the seeds: level-synchronous pull BFS from all seeds at once, a node
//      takes the smallest cell index among the neighbours assigned one level earlier.
//   3. order = nodes sorted by (cell, depth, id) (one radix sort); cells larger than `cap_rows` are cut
//      into equal chunks along that order (inner ball first), small neighbours-in-the-order are packed into
//      one tile up to `merge_rows` rows.
//
//   4. balance: the staged kernel deals tile t to CTA t mod G, so the tiles are renumbered by size -- sorted
//      by descending row count and dealt in snake order over `deal` = G CTAs (round r left-to-right when r is
//      even, right-to-left when odd) -- which takes the heaviest CTA from ~1.2x to ~1.02x the mean load.
//
// Everything is integer work over the CSR: thread per node, 4-byte loads of the neighbour list, a handful of
// sweeps per Luby round; one-time per graph like K0 itself.  Deterministic: every kernel computes a pure
// function of the previous sweep's arrays (double-buffered), the growth step only reads labels that were
// final before the launch.  The sort and the scans are CUB library calls.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.cuh"

namespace gwen {
namespace {

constexpr int kLt = 256;
enum : uint8_t { kUndecided = 0, kSeed = 1, kCovered = 2 };

// murmur3 finaliser: a bijection of uint32, 0 -> 0, so prio(v) = mix(v + 1) is unique and non-zero
__device__ __forceinline__ uint32_t loc_mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85ebca6bu;
  x ^= x >> 13;
  x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

__global__ void k_loc_prio(const uint8_t* __restrict__ state, int64_t n, uint32_t* __restrict__ m) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  m[v] = state[v] == kUndecided ? loc_mix(uint32_t(v) + 1u) : 0u;
}

// one hop of max-propagation along the CSR's edges (source -> destination)
__global__ void k_loc_sweep(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int64_t n,
                            const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  uint32_t best = in[v];
  const int32_t e1 = rowptr[v + 1];
  for (int32_t e = rowptr[v]; e < e1; ++e) {
    const int32_t u = src[e];
    if (u < n) best = max(best, in[u]);
  }
  out[v] = best;
}

// undecided nodes that hold the largest priority of their radius-ball become seeds; flag = 1 at new seeds
__global__ void k_loc_pick(uint8_t* __restrict__ state, const uint32_t* __restrict__ mk, int64_t n,
                           uint32_t* __restrict__ flag) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  uint32_t f = 0;
  if (state[v] == kUndecided && mk[v] == loc_mix(uint32_t(v) + 1u)) {
    state[v] = kSeed;
    f = 1;
  }
  flag[v] = f;
}

__global__ void k_loc_cover(uint8_t* __restrict__ state, const uint32_t* __restrict__ fk, int64_t n) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  if (state[v] == kUndecided && fk[v]) state[v] = kCovered;
}

// after the last round: whatever is still undecided becomes a seed of its own (tiny cells, packed later)
__global__ void k_loc_seed_flags(const uint8_t* __restrict__ state, int64_t n, int32_t* __restrict__ is_seed) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v > n) return;
  is_seed[v] = (v < n && state[v] != kCovered) ? 1 : 0;
}

__global__ void k_loc_cells_init(const uint8_t* __restrict__ state, const int32_t* __restrict__ seed_idx,
                                 int64_t n, int32_t* __restrict__ cell, int32_t* __restrict__ depth) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  const bool s = state[v] != kCovered;
  cell[v] = s ? seed_idx[v] : -1;
  depth[v] = s ? 0 : -1;
}

// level d of the multi-source BFS: an unassigned node joins the smallest cell among the neighbours that were
// assigned at level d - 1.  Labels of level d - 1 were final before this launch; a neighbour assigned by
// this launch shows depth -1 or d, never d - 1, so concurrent writes cannot change what a thread decides.
__global__ void k_loc_grow(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int64_t n,
                           int32_t d, volatile int32_t* cell, volatile int32_t* depth) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  if (depth[v] >= 0) return;
  int32_t best = INT32_MAX;
  const int32_t e1 = rowptr[v + 1];
  for (int32_t e = rowptr[v]; e < e1; ++e) {
    const int32_t u = src[e];
    if (u < n && depth[u] == d - 1) best = min(best, cell[u]);
  }
  if (best != INT32_MAX) {
    cell[v] = best;
    depth[v] = d;
  }
}

// nodes no seed reaches (possible only when the edge list is not symmetric) share one extra cell;
// sort keys (cell, depth, id) and the cell histogram
__global__ void k_loc_keys(int32_t* __restrict__ cell, int32_t* __restrict__ depth, int64_t n,
                           const int32_t* __restrict__ n_seeds, uint64_t* __restrict__ keys,
                           int32_t* __restrict__ vals, int32_t* __restrict__ cell_size,
                           int32_t* __restrict__ status) {
  const int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (v >= n) return;
  int32_t c = cell[v], d = depth[v];
  if (c < 0) {
    c = *n_seeds;
    d = 0;
    cell[v] = c;
    depth[v] = d;
    atomicAdd(&status[3], 1);
  }
  keys[v] = (uint64_t(uint32_t(c)) << 8) | uint64_t(min(d, 255));
  vals[v] = static_cast<int32_t>(v);
  atomicAdd(&cell_size[c], 1);
}

// one thread walks the cells in index order and cuts the sorted node list into tiles: a cell larger than
// merge_rows stands alone (cut into equal chunks of at most cap_rows), smaller ones are packed together while
// the tile stays within merge_rows.  The next cell's size is loaded one iteration ahead.
__global__ void k_loc_pack(const int32_t* __restrict__ cell_size, const int32_t* __restrict__ n_seeds,
                           int32_t merge_rows, int32_t cap_rows, int32_t* __restrict__ tile_ptr,
                           int32_t* __restrict__ status) {
  if (blockIdx.x || threadIdx.x) return;
  const int32_t cells = *n_seeds + (status[3] > 0 ? 1 : 0);
  int32_t nt = 0, pos = 0, cur = 0, biggest = 0;
  tile_ptr[0] = 0;
  int32_t next = cells > 0 ? cell_size[0] : 0;
  for (int32_t c = 0; c < cells; ++c) {
    const int32_t s = next;
    if (c + 1 < cells) next = cell_size[c + 1];
    biggest = max(biggest, s);
    if (s == 0) continue;
    if (s > merge_rows) {
      if (cur > 0) {
        tile_ptr[++nt] = pos;
        cur = 0;
      }
      const int32_t chunks = (s + cap_rows - 1) / cap_rows;
      for (int32_t j = 1; j <= chunks; ++j)
        tile_ptr[++nt] = pos + static_cast<int32_t>(int64_t(s) * j / chunks);
      pos += s;
    } else {
      if (cur + s > merge_rows) {
        tile_ptr[++nt] = pos;
        cur = 0;
      }
      cur += s;
      pos += s;
    }
  }
  if (cur > 0) tile_ptr[++nt] = pos;
  status[0] = cells;
  status[1] = nt;
  status[2] = biggest;
}

// ---- step 4: tiles renumbered by size, dealt in snake order ----------------------------------------------
__global__ void k_loc_tile_keys(const int32_t* __restrict__ tile_ptr, const int32_t* __restrict__ status,
                                int64_t cap, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= cap) return;
  const int32_t nt = status[1];
  keys[t] = t < nt ? uint32_t(INT32_MAX - (tile_ptr[t + 1] - tile_ptr[t])) : 0xFFFFFFFFu;
  vals[t] = static_cast<int32_t>(t);
}

__global__ void k_loc_deal(const uint32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                           const int32_t* __restrict__ status, int32_t deal, int64_t cap,
                           int32_t* __restrict__ new_size, int32_t* __restrict__ new_of_old) {
  const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (k > cap) return;
  const int32_t nt = status[1];
  if (k >= nt) {
    new_size[k] = 0;
    return;
  }
  const int32_t r = static_cast<int32_t>(k / deal), p = static_cast<int32_t>(k % deal);
  const int32_t cnt = min(deal, nt - r * deal);
  const int32_t t2 = r * deal + ((r & 1) ? cnt - 1 - p : p);
  new_size[t2] = INT32_MAX - static_cast<int32_t>(keys[k]);
  new_of_old[vals[k]] = t2;
}

__global__ void k_loc_reorder(const int32_t* __restrict__ old_ptr, const int32_t* __restrict__ new_ptr,
                              const int32_t* __restrict__ new_of_old, const int32_t* __restrict__ status,
                              const int32_t* __restrict__ order_in, int64_t n, int32_t* __restrict__ order_out) {
  const int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  int lo = 0, hi = status[1];  // last tile with old_ptr[t] <= q
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (old_ptr[mid] <= q) lo = mid; else hi = mid;
  }
  order_out[new_ptr[new_of_old[lo]] + (q - old_ptr[lo])] = order_in[q];
}

struct LocWs {
  size_t off_state, off_a, off_b, off_flag, off_seed, off_size, off_keys_in, off_keys_out, off_vals_in,
      off_otmp, off_optr, off_cub, cub_bytes, bytes;
  int bits;
};

cudaError_t loc_ws(int64_t n, LocWs* p) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nn = static_cast<size_t>(n + 2);
  p->off_state = take(nn);
  p->off_a = take(nn * 4);
  p->off_b = take(nn * 4);
  p->off_flag = take(nn * 4);      // is_seed
  p->off_seed = take(nn * 4);      // exclusive scan of is_seed; [n] = number of seeds
  p->off_size = take(nn * 4);
  p->off_keys_in = take(nn * 8);
  p->off_keys_out = take(nn * 8);
  p->off_vals_in = take(nn * 4);
  p->off_otmp = take(nn * 4);      // order before the tiles are renumbered
  p->off_optr = take(nn * 4);      // tile_ptr before the tiles are renumbered
  int nb = 1;
  while ((int64_t(1) << nb) < n + 2) ++nb;
  p->bits = 8 + nb;
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t err = cub::DeviceRadixSort::SortPairs(
      nullptr, sort_bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
      static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(n), 0, p->bits);
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(n + 1));
  if (err != cudaSuccess) return err;
  size_t sort2_bytes = 0, scan2_bytes = 0;
  err = cub::DeviceRadixSort::SortPairs(nullptr, sort2_bytes, static_cast<const uint32_t*>(nullptr),
                                        static_cast<uint32_t*>(nullptr), static_cast<const int32_t*>(nullptr),
                                        static_cast<int32_t*>(nullptr), static_cast<int>(n + 1), 0, 32);
  if (err != cudaSuccess) return err;
  err = cub::DeviceScan::ExclusiveSum(nullptr, scan2_bytes, static_cast<int32_t*>(nullptr),
                                      static_cast<int32_t*>(nullptr), static_cast<int>(n + 2));
  if (err != cudaSuccess) return err;
  p->cub_bytes = std::max(std::max(sort_bytes, scan_bytes), std::max(sort2_bytes, scan2_bytes));
  p->off_cub = take(p->cub_bytes);
  p->bytes = off;
  return cudaSuccess;
}

}  // namespace
}  // namespace gwen

using namespace gwen;

extern "C" int gwen_locality_workspace_bytes(int64_t n, size_t* out) {
  GWEN_CHECK_ARG(out && n > 0 && n < INT32_MAX - 2, "bad arguments");
  LocWs p;
  GWEN_CUDA(loc_ws(n, &p));
  *out = p.bytes;
  return GWEN_OK;
}

extern "C" int gwen_locality_tiles(const int32_t* rowptr, const int32_t* src, int64_t n, int32_t radius,
                                   int32_t rounds, int32_t merge_rows, int32_t cap_rows, int32_t deal,
                                   int32_t* order,
                                   int32_t* tile_ptr, int32_t* cell, int32_t* depth, int32_t* status,
                                   void* ws, size_t ws_bytes, void* stream) {
  GWEN_CHECK_ARG(rowptr && src && order && tile_ptr && cell && depth && status && ws, "null pointer");
  GWEN_CHECK_ARG(n > 0 && n < INT32_MAX - 2, "bad node count");
  GWEN_CHECK_ARG(radius >= 1 && radius <= 250 && rounds >= 1 && rounds <= 64, "radius in [1, 250], rounds in [1, 64]");
  GWEN_CHECK_ARG(merge_rows >= 1 && cap_rows >= merge_rows, "need 1 <= merge_rows <= cap_rows");
  GWEN_CHECK_ARG(deal >= 0, "deal must be >= 0");
  LocWs p;
  GWEN_CUDA(loc_ws(n, &p));
  if (ws_bytes < p.bytes) return set_err(GWEN_E_WORKSPACE, "workspace %zu < required %zu", ws_bytes, p.bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  uint8_t* state = reinterpret_cast<uint8_t*>(base + p.off_state);
  uint32_t* a = reinterpret_cast<uint32_t*>(base + p.off_a);
  uint32_t* b = reinterpret_cast<uint32_t*>(base + p.off_b);
  int32_t* is_seed = reinterpret_cast<int32_t*>(base + p.off_flag);
  int32_t* seed_idx = reinterpret_cast<int32_t*>(base + p.off_seed);
  int32_t* cell_size = reinterpret_cast<int32_t*>(base + p.off_size);
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(base + p.off_keys_in);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(base + p.off_keys_out);
  int32_t* vals_in = reinterpret_cast<int32_t*>(base + p.off_vals_in);
  void* cub_ws = base + p.off_cub;
  const unsigned grid = static_cast<unsigned>(ceil_div(n + 1, kLt));
  GWEN_CUDA(cudaMemsetAsync(state, 0, static_cast<size_t>(n + 2), st));
  GWEN_CUDA(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
  GWEN_CUDA(cudaMemsetAsync(cell_size, 0, static_cast<size_t>(n + 2) * 4, st));
  // 1. Luby rounds on the radius-th power of the graph
  for (int r = 0; r < rounds; ++r) {
    k_loc_prio<<<grid, kLt, 0, st>>>(state, n, a);
    for (int h = 0; h < radius; ++h) {
      k_loc_sweep<<<grid, kLt, 0, st>>>(rowptr, src, n, a, b);
      std::swap(a, b);
    }
    k_loc_pick<<<grid, kLt, 0, st>>>(state, a, n, b);
    std::swap(a, b);
    for (int h = 0; h < radius; ++h) {
      k_loc_sweep<<<grid, kLt, 0, st>>>(rowptr, src, n, a, b);
      std::swap(a, b);
    }
    k_loc_cover<<<grid, kLt, 0, st>>>(state, a, n);
    GWEN_LAUNCH_CHECK("k_loc_* (seed rounds)");
  }
  // 2. cells
  k_loc_seed_flags<<<grid, kLt, 0, st>>>(state, n, is_seed);
  size_t cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, is_seed, seed_idx, static_cast<int>(n + 1), st));
  k_loc_cells_init<<<grid, kLt, 0, st>>>(state, seed_idx, n, cell, depth);
  for (int d = 1; d <= radius; ++d) k_loc_grow<<<grid, kLt, 0, st>>>(rowptr, src, n, d, cell, depth);
  GWEN_LAUNCH_CHECK("k_loc_grow");
  // 3. order and tiles
  k_loc_keys<<<grid, kLt, 0, st>>>(cell, depth, n, seed_idx + n, keys_in, vals_in, cell_size, status);
  GWEN_LAUNCH_CHECK("k_loc_keys");
  cub_bytes = p.cub_bytes;
  int32_t* order_tmp = reinterpret_cast<int32_t*>(base + p.off_otmp);
  int32_t* old_ptr = reinterpret_cast<int32_t*>(base + p.off_optr);
  GWEN_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, vals_in,
                                            deal ? order_tmp : order, static_cast<int>(n), 0, p.bits, st));
  k_loc_pack<<<1, 32, 0, st>>>(cell_size, seed_idx + n, merge_rows, cap_rows, deal ? old_ptr : tile_ptr, status);
  GWEN_LAUNCH_CHECK("k_loc_pack");
  if (!deal) return GWEN_OK;
  // 4. tiles renumbered by size (the Luby / cell arrays are free again: reuse them)
  uint32_t* tk_in = a;
  uint32_t* tk_out = b;
  int32_t* tv_in = is_seed;
  int32_t* tv_out = vals_in;
  int32_t* new_size = cell_size;
  int32_t* new_of_old = seed_idx;
  const int64_t cap = n + 1;  // more than the tile count can be (<= number of cells + n / cap_rows + 1 <= n + 1)
  k_loc_tile_keys<<<grid, kLt, 0, st>>>(old_ptr, status, cap, tk_in, tv_in);
  GWEN_LAUNCH_CHECK("k_loc_tile_keys");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, tk_in, tk_out, tv_in, tv_out,
                                            static_cast<int>(cap), 0, 32, st));
  k_loc_deal<<<static_cast<unsigned>(ceil_div(cap + 1, kLt)), kLt, 0, st>>>(tk_out, tv_out, status, deal, cap,
                                                                              new_size, new_of_old);
  GWEN_LAUNCH_CHECK("k_loc_deal");
  cub_bytes = p.cub_bytes;
  GWEN_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, new_size, tile_ptr, static_cast<int>(n + 2), st));
  k_loc_reorder<<<grid, kLt, 0, st>>>(old_ptr, tile_ptr, new_of_old, status, order_tmp, n, order);
  GWEN_LAUNCH_CHECK("k_loc_reorder");
  return GWEN_OK;
}
