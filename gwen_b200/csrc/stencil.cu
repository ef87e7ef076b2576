// K1s: aggregation fast path for mesh graphs (the COSMO-shaped grids of BASELINE configs 2-5).
//
// When the graph is exactly the 8-neighbour H x W mesh with GCN normalisation (K0 detects this),
//   out[r, c] = dis[r, c] * sum_{dr, dc in {-1,0,1}} dis[r+dr, c+dc] * x[r+dr, c+dc]     (valid nodes)
// is a separable 3 x 3 box filter of y = dis * x.  The CSR kernels read 9 staged source rows per
// destination through the 128 B/clk shared-memory pipe (the measured bound of k_agg_tiled, LSU data
// pipe 72 %); here a sub-warp slides along a destination row holding the last three COLUMN sums
//   s[t] = (dis*x)[r-1, t] + (dis*x)[r, t] + (dis*x)[r+1, t]
// in registers, so each destination costs 3 shared-memory row reads, and the kernel goes back to
// being HBM-bound.
//
// Data movement: work item = (tile of TH x TW destinations, batch b, feature slab).  One lane
// issues ONE cp.async.bulk.tensor.4d per item: the box {slab, TW+2, TH+2, 1} of x viewed as
// [B][Hs][W][F], plus one 2-D box of the (padded) dis array.  Coordinates start one row/column
// outside the tile; TMA zero-fills everything outside the mesh, which IS the truncated stencil of
// border nodes (their missing neighbours contribute 0).  Warp-specialised, NS-stage full/empty
// mbarrier ring as in k_agg_tiled.
//
// Summation order differs from the CSR order (column sums first), so results are not bitwise
// equal to gwen_aggregate_fwd: they agree to fp32 rounding (tests: <= 1e-6 normalised) and are
// deterministic and independent of tiling / partitioning (each destination's order is fixed by
// geometry alone).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"
#include "stencil_common.cuh"

namespace gwen {
using namespace st;
namespace {

struct StencilArgs {
  void* out;
  const float* bias;
  int64_t batch, feat, ldo, o_bstride;
  int hd, w, row_off;        // destination rows, mesh width, source row of destination row 0
  int tw, tiles_x, num_tiles, slabs, slab_elems, relu, num_stages;
  uint32_t x_bytes, dis_pitch, stage_bytes;  // bytes of the x box, dis box row pitch, one stage
  const float* disb;       // zero-bordered dis, [hs + 2][disb_pitch]
  int64_t disb_pitch;
  // peer-halo mode (gwen_grid_stencil_peer_fwd): the kernel fetches its two halo rows itself
  int peer;                 // 0 = plain launch
  unsigned char* x_base;    // local x (halo row 0 at offset 0 of each batch slice)
  const unsigned char* up_src;    // neighbour rows to copy (device pointers into PEER memory)
  const unsigned char* down_src;
  int64_t x_bstride_bytes, up_bstride_bytes, down_bstride_bytes, row_bytes, bottom_off_bytes;
  uint32_t* flag_up_remote;   // where this rank announces "my x is ready" to the neighbours
  uint32_t* flag_down_remote;
  uint32_t* ctl;              // local control words, see gwen_halo_peers
  uint64_t spin_timeout_ns;   // bound on every wait of the protocol (a sticky error word is set instead of hanging)
  uint32_t wait_ns;           // poll interval of the consumer warps on the full barrier (0 = spin)
};

// control words of the peer-halo protocol (ctl[] in local device memory, zero-initialised)
enum { CTL_FROM_UP = 0, CTL_FROM_DOWN = 1, CTL_EPOCH = 2, CTL_HALO_DONE = 3, CTL_CTAS_DONE = 4, CTL_ERROR = 5 };
enum { PEER_ERR_NEIGHBOUR_FLAG = 1, PEER_ERR_HALO_DONE = 2 };

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {  // uncached at every level we control
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// LPR lanes cover one slab (LPR * 16 bytes); 32/LPR sub-warps per warp, each walks one unit
// (destination row r of the tile, columns [seg*SEG, seg*SEG + SEG)).
template <typename T, int LPR>
__global__ void __launch_bounds__(576, 1)
    k_grid_stencil(const __grid_constant__ CUtensorMap xmap, StencilArgs a) {
  constexpr int VN = V16<T>::N;
  constexpr int RPW = 32 / LPR;
  constexpr int MAXS = 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAXS], empty_bar[MAXS];
  const int ns = a.num_stages;
  const uint32_t stage0 = (smem_u32(smem_raw) + 127u) & ~127u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ncw = nwarps - (a.peer ? 2 : 1);  // consumers | producer warp | (peer mode) halo warp
  const uint32_t epoch1 = a.peer ? *reinterpret_cast<volatile uint32_t*>(a.ctl + CTL_EPOCH) + 1u : 0u;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&xmap);
    for (int i = 0; i < ns; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), ncw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int per_tile = static_cast<int>(a.batch) * a.slabs;
  // Work ITEMS (batch x tile x slab), not tiles, are dealt round-robin: 949 tiles over 148 CTAs would
  // leave some CTAs 7 tiles and others 6 (9 % tail); 3796 items leave 26 vs 25.  Item order: batch
  // slowest, then tile, slab fastest -- CTAs that run together fetch all slabs of ~37 neighbouring tiles
  // of ONE member: whole rows at a time from DRAM, and the halos of vertical neighbours (25 tiles later)
  // are still in L2.
  const int64_t total_items = int64_t(a.num_tiles) * per_tile;
  const int64_t n_items = total_items > int64_t(blockIdx.x)
                              ? (total_items - 1 - int64_t(blockIdx.x)) / int64_t(gridDim.x) + 1 : 0;
  constexpr uint32_t kSlabBytes = LPR * 16u;                     // = a.slab_elems * sizeof(T)
  const uint32_t srow_bytes = uint32_t(a.tw + 2) * kSlabBytes;  // one staged mesh row

  // Peer mode walks the interior tile rows (of all members) first and the two tile rows that read a halo
  // row last (tile row 0, then the bottom one), so the halo fetch runs under the interior work.
  const int tiles_y = a.num_tiles / a.tiles_x;
  const int n_interior = a.peer ? (tiles_y > 2 ? (tiles_y - 2) * a.tiles_x : 0) : a.num_tiles;
  const int64_t n_int_items = int64_t(a.batch) * n_interior * a.slabs;
  auto decode = [&](int64_t gi, int& b, int& t, int& slab) {
    if (gi < n_int_items) {
      const int64_t per_b = int64_t(n_interior) * a.slabs;
      b = int(gi / per_b);
      const int r = int(gi % per_b);
      slab = r % a.slabs;
      t = r / a.slabs + (a.peer ? a.tiles_x : 0);
    } else {                                       // boundary tiles (peer mode only)
      const int nb = a.num_tiles - n_interior;
      const int64_t g2 = gi - n_int_items, per_b = int64_t(nb) * a.slabs;
      b = int(g2 / per_b);
      const int r = int(g2 % per_b);
      slab = r % a.slabs;
      const int q = r / a.slabs;
      t = q < a.tiles_x ? q : (tiles_y - 1) * a.tiles_x + (q - a.tiles_x);
    }
  };

  if (warp == ncw + 1) {
    // ===== halo warp (peer mode): announce, then pull this CTA's share of both halo rows =====
    if (blockIdx.x == 0 && lane == 0) {
      __threadfence_system();
      if (a.flag_up_remote) st_release_sys(a.flag_up_remote, epoch1);
      if (a.flag_down_remote) st_release_sys(a.flag_down_remote, epoch1);
    }
    const int64_t units_row = a.row_bytes / 16;
    const int64_t units = units_row * a.batch;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const unsigned char* src = side ? a.down_src : a.up_src;
      if (!src) continue;
      if (lane == 0) {
        // bounded: a neighbour that launched a different sequence of peer calls (or died) must give a
        // diagnosable failure, not a hung GPU -- on timeout the sticky error word is set and we go on
        const uint64_t t0 = globaltimer_ns();
        while (int32_t(ld_acquire_sys(a.ctl + (side ? CTL_FROM_DOWN : CTL_FROM_UP)) - epoch1) < 0) {
          __nanosleep(64);
          if (globaltimer_ns() - t0 > a.spin_timeout_ns) {
            atomicOr(a.ctl + CTL_ERROR, uint32_t(PEER_ERR_NEIGHBOUR_FLAG));
            break;
          }
        }
      }
      __syncwarp();
      const int64_t sb = side ? a.down_bstride_bytes : a.up_bstride_bytes;
      unsigned char* dst = a.x_base + (side ? a.bottom_off_bytes : 0);
      constexpr int U = 4;
      for (int64_t u0 = (int64_t(blockIdx.x) * 32 + lane); u0 < units; u0 += int64_t(gridDim.x) * 32 * U) {
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int64_t u = u0 + int64_t(k) * gridDim.x * 32;
          if (u < units) v[k] = ld_peer_v4(src + (u / units_row) * sb + (u % units_row) * 16);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int64_t u = u0 + int64_t(k) * gridDim.x * 32;
          if (u < units)
            *reinterpret_cast<uint4*>(dst + (u / units_row) * a.x_bstride_bytes + (u % units_row) * 16) = v[k];
        }
      }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) red_release_gpu_add(a.ctl + CTL_HALO_DONE, 1u);
  } else if (warp == ncw) {
    if (lane == 0) {  // ===== producer =====
      int st = 0;
      uint32_t round = 0;
      bool halo_ready = !a.peer;
      for (int64_t it = 0; it < n_items; ++it) {
        const int64_t gi = int64_t(blockIdx.x) + it * int64_t(gridDim.x);   // global item index
        int b, t, slab;
        decode(gi, b, t, slab);
        if (!halo_ready && gi >= n_int_items) {  // first tile that reads a halo row
          const uint64_t t0 = globaltimer_ns();
          while (ld_acquire_gpu(a.ctl + CTL_HALO_DONE) < gridDim.x) {
            __nanosleep(32);
            if (globaltimer_ns() - t0 > a.spin_timeout_ns) {   // a CTA of this grid never ran: see CTL_ERROR
              atomicOr(a.ctl + CTL_ERROR, uint32_t(PEER_ERR_HALO_DONE));
              break;
            }
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");  // generic stores -> TMA reads
          halo_ready = true;
        }
        const int r0 = (t / a.tiles_x) * TH, c0 = (t % a.tiles_x) * a.tw;
        if (round > 0) mbar_wait(smem_u32(&empty_bar[st]), (round - 1) & 1u);
        const uint32_t bar = smem_u32(&full_bar[st]);
        const uint32_t dst = stage0 + uint32_t(st) * a.stage_bytes;
        mbar_expect_tx(bar, a.x_bytes + uint32_t(TH + 2) * a.dis_pitch);
        tma_load_4d(dst, &xmap, slab * a.slab_elems, c0 - 1, r0 + a.row_off - 1, b, bar);
        // dis box: TH+2 rows of the zero-bordered dis array (element [r+1][c+1] = dis[r][c]),
        // so mesh node (r0+row_off-1+i, c0-1+t) sits at disb[(r0+row_off+i) * pitch + c0 + t].
        const float* dsrc = a.disb + int64_t(r0 + a.row_off) * a.disb_pitch + c0;
        for (int i = 0; i < TH + 2; ++i)
          bulk_g2s(dst + a.x_bytes + uint32_t(i) * a.dis_pitch, dsrc + int64_t(i) * a.disb_pitch,
                   a.dis_pitch, bar);
        if (++st == ns) { st = 0; ++round; }
      }
    }
  } else {
  // ===== consumers =====
  const int sub = lane / LPR, l = lane % LPR;
  const int nseg = a.tw / SEG;
  const int units = TH * nseg;
  int st = 0;
  uint32_t round = 0;
  for (int64_t it = 0; it < n_items; ++it) {
    const int64_t gi = int64_t(blockIdx.x) + it * int64_t(gridDim.x);
    int b, t, slab;
    decode(gi, b, t, slab);
    const int r0 = (t / a.tiles_x) * TH, c0 = (t % a.tiles_x) * a.tw;
    const int rows_valid = min(TH, a.hd - r0), cols_valid = min(a.tw, a.w - c0);
    const int64_t col = int64_t(slab) * a.slab_elems + int64_t(l) * VN;
    const bool on = col < a.feat;
    const uint32_t xs = stage0 + uint32_t(st) * a.stage_bytes;
    const uint32_t ds = xs + a.x_bytes;
    constexpr int VP = VN / 2;  // packed fp32 pairs per 16 bytes
    uint64_t bv[VP];
#pragma unroll
    for (int k = 0; k < VP; ++k)
      bv[k] = (a.bias && on) ? pk2(__ldg(a.bias + col + 2 * k), __ldg(a.bias + col + 2 * k + 1)) : 0ull;
    T* ob = static_cast<T*>(a.out) + b * a.o_bstride + col;
    const bool relu = a.relu != 0;
    mbar_wait_backoff(smem_u32(&full_bar[st]), round & 1u, a.wait_ns);
    if (on) {
      for (int u = warp * RPW + sub; u < units; u += ncw * RPW) {
        const int r = u / nseg, cb = (u - r * nseg) * SEG;
        const int nvalid = cols_valid - cb;  // destinations of this unit inside the mesh
        if (r >= rows_valid || nvalid <= 0) continue;
        // staged rows r, r+1, r+2 hold mesh rows (r0+r-1 .. r0+r+1); staged column t = mesh c0-1+t
        const uint32_t x0 = xs + uint32_t(r) * srow_bytes + uint32_t(cb) * kSlabBytes + uint32_t(l) * 16u;
        const uint32_t x1 = x0 + srow_bytes, x2 = x1 + srow_bytes;
        const uint32_t q0 = ds + uint32_t(r) * a.dis_pitch + uint32_t(cb) * 4u;
        const uint32_t q1 = q0 + a.dis_pitch, q2 = q1 + a.dis_pitch;
        T* op = ob + (int64_t(r0 + r) * a.w + (c0 + cb)) * a.ldo;  // destination column cb
        uint64_t s0[VP], s1[VP], s2[VP];  // column sums of staged columns t-2, t-1, t
        float dmid_prev = 0.0f;           // dis of the destination under the window centre
#pragma unroll
        for (int tt = 0; tt < SEG + 2; ++tt) {
          const uint4 v0 = lds128(x0 + uint32_t(tt) * kSlabBytes);
          const uint4 v1 = lds128(x1 + uint32_t(tt) * kSlabBytes);
          const uint4 v2 = lds128(x2 + uint32_t(tt) * kSlabBytes);
          const float d0 = lds32(q0 + uint32_t(tt) * 4u);
          const float d1 = lds32(q1 + uint32_t(tt) * 4u);
          const float d2 = lds32(q2 + uint32_t(tt) * 4u);
          uint64_t f0[VP], f1[VP], f2[VP];
          V16<T>::unpack2(v0, f0);
          V16<T>::unpack2(v1, f1);
          V16<T>::unpack2(v2, f2);
          const uint64_t e0 = pk2(d0, d0), e1 = pk2(d1, d1), e2 = pk2(d2, d2);
#pragma unroll
          for (int k = 0; k < VP; ++k) {
            s0[k] = s1[k];
            s1[k] = s2[k];
            s2[k] = fma2(e2, f2[k], fma2(e1, f1[k], mul2(e0, f0[k])));
          }
          if (tt >= 2) {  // destination column cb + tt - 2 (window = staged columns tt-2 .. tt)
            if (tt - 2 < nvalid) {
              const uint64_t dm = pk2(dmid_prev, dmid_prev);
              uint64_t o[VP];
#pragma unroll
              for (int k = 0; k < VP; ++k) o[k] = fma2(dm, add2(add2(s0[k], s1[k]), s2[k]), bv[k]);
              stg128(op, V16<T>::pack2(o, relu));
            }
            op += a.ldo;
          }
          dmid_prev = d1;  // dis of mesh node (r0+r, c0-1+tt): the centre of the NEXT window
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty_bar[st]));
    if (++st == ns) { st = 0; ++round; }
  }
  }  // roles
  if (a.peer) {
    // last CTA out re-arms the protocol for the next launch: every CTA has read the epoch and
    // passed (or never needed) the halo wait by the time all of them have finished.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(a.ctl + CTL_CTAS_DONE, 1u) == gridDim.x - 1) {
        a.ctl[CTL_HALO_DONE] = 0;
        a.ctl[CTL_CTAS_DONE] = 0;
        __threadfence();
        *reinterpret_cast<volatile uint32_t*>(a.ctl + CTL_EPOCH) = epoch1;
      }
    }
  }
}

inline int env_int2(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (!v) return dflt;
  int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

template <typename T, int LPR>
int launch_stencil(const CUtensorMap& xmap, const StencilArgs& a, size_t smem, cudaStream_t st) {
  auto kern = k_grid_stencil<T, LPR>;
  GWEN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  // experiment knobs (developer): consumer warps per CTA and CTAs per SM
  static const int cw_env = env_int2("GWEN_STENCIL_CWARPS", 16, 2, 16);
  static const int cps_env = env_int2("GWEN_STENCIL_CTAS_PER_SM", 1, 1, 4);
  const int cps = a.peer ? 1 : cps_env;
  const int64_t items = int64_t(a.num_tiles) * a.batch * a.slabs;
  const int grid = static_cast<int>(std::min<int64_t>(items, std::max(1, (sm_count() - sm_reserve()) * cps)));
  // consumer sub-warps = TH * tw / SEG units when possible: 16 consumer warps + 1 producer warp
  if (!a.peer) {
    kern<<<grid, 32 * (cw_env + 1), smem, st>>>(xmap, a);
    GWEN_LAUNCH_CHECK("k_grid_stencil");
    return GWEN_OK;
  }
  // Peer mode: the CTAs wait for one another (every CTA publishes its halo share before the boundary
  // tiles start), so the whole grid must be co-resident.  Clamp it to what the occupancy calculator
  // says fits and launch COOPERATIVELY: the runtime then either schedules all CTAs together or
  // fails the launch -- it never runs a part of the grid that would spin on the rest.
  int per_sm = 0;
  GWEN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 576, smem));
  if (per_sm < 1) return set_err(GWEN_E_NOSUPPORT, "peer stencil does not fit on an SM");
  const int grid_peer = std::min(grid, per_sm * sm_count());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid_peer);
  cfg.blockDim = dim3(576);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  // developer measurement only (GWEN_PEER_NO_COOP=1): the same grid without the cooperative attribute, to see what
  // the co-residency guarantee costs per launch; never the default -- without it a partly scheduled grid would spin
  static const bool no_coop = getenv("GWEN_PEER_NO_COOP") != nullptr;
  cfg.numAttrs = no_coop ? 0 : 1;
  GWEN_CUDA(cudaLaunchKernelEx(&cfg, kern, xmap, a));
  GWEN_LAUNCH_CHECK("k_grid_stencil (peer)");
  return GWEN_OK;
}

}  // namespace
}  // namespace gwen

using namespace gwen;

static int stencil_fwd(const void* x, void* out, const float* dis_padded, int64_t dis_pitch,
                       int64_t dis_rows, int64_t batch, int64_t hs, int64_t hd, int64_t w, int64_t row_off,
                       int64_t feat, int64_t ldx, int64_t x_bstride, int64_t ldo, int64_t o_bstride,
                       int dtype, const float* bias, int epilogue, int32_t slab_elems,
                       int32_t tile_w, const gwen_halo_peers* peers, void* stream) {
  GWEN_CHECK_ARG(batch >= 0 && hs >= 0 && hd >= 0 && w >= 0 && feat >= 0, "negative size");
  if (batch == 0 || hd == 0 || w == 0 || feat == 0) return GWEN_OK;
  GWEN_CHECK_ARG(x && out && dis_padded, "null pointer");
  GWEN_CHECK_ARG(dtype == GWEN_F32 || dtype == GWEN_BF16, "unknown dtype %d", dtype);
  GWEN_CHECK_ARG(row_off >= 0 && row_off + hd <= hs + 1, "row_off outside the source rows");
  // every tile bulk-copies TH + 2 dis rows starting at bordered row r0 + row_off, r0 <= round_up(hd, TH) - TH
  GWEN_CHECK_ARG(dis_rows >= (hd + TH - 1) / TH * TH + row_off + 2,
                 "bordered dis has %lld rows, the launch reads %lld", (long long)dis_rows,
                 (long long)((hd + TH - 1) / TH * TH + row_off + 2));
  GWEN_CHECK_ARG(dis_pitch % 4 == 0 && aligned16(dis_padded),
                 "bordered dis needs a pitch that is a multiple of 4 floats and a 16-byte base");
  GWEN_CHECK_ARG(hs * w < INT32_MAX && batch < 65536, "mesh too large");
  const int esz = dtype == GWEN_F32 ? 4 : 2;
  const int vn = 16 / esz;
  if (feat % vn || ldx % vn || ldo % vn || x_bstride % vn || o_bstride % vn || !aligned16(x) ||
      !aligned16(out))
    return set_err(GWEN_E_ALIGN, "grid stencil needs 16-byte aligned rows (feat %% %d == 0)", vn);
  // defaults from the cfg 2 / cfg 3 sweeps (tools/sweep_agg.py): fp32 16-wide tiles with 64-float
  // slabs (4 stages of 46 KB), bf16 32-wide tiles with 256-element slabs
  int tw = tile_w > 0 ? tile_w : (dtype == GWEN_F32 ? 16 : 32);
  GWEN_CHECK_ARG(tw % SEG == 0 && tw >= SEG && tw <= 128, "tile_w must be a multiple of %d", SEG);
  int lpr = slab_elems > 0 ? static_cast<int>(slab_elems / vn) : (dtype == GWEN_F32 ? 16 : 32);
  if (lpr != 8 && lpr != 16 && lpr != 32)
    return set_err(GWEN_E_BADARG, "slab_elems must be %d, %d or %d", 8 * vn, 16 * vn, 32 * vn);
  while (lpr > 8 && lpr * vn / 2 >= feat) lpr /= 2;
  const uint32_t dis_box_w = static_cast<uint32_t>((tw + 2 + 3) / 4 * 4);
  auto stage_for = [&](int lpr_) {
    size_t xb = size_t(TH + 2) * (tw + 2) * lpr_ * 16;
    return align_up(xb + size_t(TH + 2) * dis_box_w * 4, 128);
  };
  const size_t smem_cap = 226 * 1024;
  while (lpr > 8 && 2 * stage_for(lpr) + 256 > smem_cap) lpr /= 2;
  if (2 * stage_for(lpr) + 256 > smem_cap)
    return set_err(GWEN_E_NOSUPPORT, "stencil tile does not fit in shared memory");
  static const int stage_hint = env_int2("GWEN_STENCIL_STAGES", 4, 2, 8);
  int ns = 2;
  while (ns < stage_hint && (ns + 1) * stage_for(lpr) + 256 <= smem_cap) ++ns;
  const int slab = lpr * vn;
  // x viewed as [B][Hs][W][F]; box {slab, tw+2, TH+2, 1}, no swizzle, zero fill outside the mesh
  auto enc = tensor_map_encoder();
  if (!enc) return set_err(GWEN_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap xmap;
  {
    cuuint64_t dims[4] = {cuuint64_t(feat), cuuint64_t(w), cuuint64_t(hs), cuuint64_t(batch)};
    cuuint64_t strides[3] = {cuuint64_t(ldx) * esz, cuuint64_t(w) * ldx * esz,
                             cuuint64_t(batch > 1 ? x_bstride : hs * w * ldx) * esz};
    cuuint32_t box[4] = {cuuint32_t(slab), cuuint32_t(tw + 2), cuuint32_t(TH + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&xmap, dtype == GWEN_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                     4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GWEN_E_CUDA, "x tensor map encode failed (%d)", int(r));
  }
  const int tiles_x = static_cast<int>(ceil_div(w, tw)), tiles_y = static_cast<int>(ceil_div(hd, TH));
  StencilArgs a{};
  a = StencilArgs{out, bias, batch, feat, ldo, o_bstride, static_cast<int>(hd), static_cast<int>(w),
                static_cast<int>(row_off), tw, tiles_x, tiles_x * tiles_y,
                static_cast<int>(ceil_div(feat, slab)), slab, (epilogue & GWEN_EPI_RELU) ? 1 : 0, ns,
                static_cast<uint32_t>(size_t(TH + 2) * (tw + 2) * lpr * 16), dis_box_w * 4,
                static_cast<uint32_t>(stage_for(lpr)), dis_padded, dis_pitch};
  if (peers) {
    // x is this rank's band [B][hd + 2][w][feat]: local row 0 / hd + 1 are the halo rows
    GWEN_CHECK_ARG(row_off == 1 && hs == hd + 2 && ldx == feat, "peer mode needs x = [B, hd + 2, w, feat], row_off 1");
    GWEN_CHECK_ARG(peers->ctl && aligned16(peers->up_row) && aligned16(peers->down_row), "bad halo peers");
    a.peer = 1;
    a.x_base = static_cast<unsigned char*>(const_cast<void*>(x));
    a.up_src = static_cast<const unsigned char*>(peers->up_row);
    a.down_src = static_cast<const unsigned char*>(peers->down_row);
    a.row_bytes = w * feat * esz;
    a.x_bstride_bytes = (batch > 1 ? x_bstride : hs * w * ldx) * esz;
    a.up_bstride_bytes = peers->up_bstride * esz;
    a.down_bstride_bytes = peers->down_bstride * esz;
    a.bottom_off_bytes = (hd + 1) * a.row_bytes;
    a.flag_up_remote = peers->up_flag;
    a.flag_down_remote = peers->down_flag;
    a.ctl = peers->ctl;
    static const int timeout_ms = env_int2("GWEN_PEER_TIMEOUT_MS", 10000, 1, 3600000);
    a.spin_timeout_ns = uint64_t(timeout_ms) * 1000000ull;
  }
  a.wait_ns = wait_backoff_ns();
  if (dis_pitch < int64_t(tiles_x - 1) * tw + dis_box_w)
    return set_err(GWEN_E_BADARG, "bordered dis pitch %lld < %lld needed for tile width %d",
                   (long long)dis_pitch, (long long)(int64_t(tiles_x - 1) * tw + dis_box_w), tw);
  const size_t smem = ns * stage_for(lpr) + 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == GWEN_F32) {
    if (lpr == 32) return launch_stencil<float, 32>(xmap, a, smem, st);
    if (lpr == 16) return launch_stencil<float, 16>(xmap, a, smem, st);
    return launch_stencil<float, 8>(xmap, a, smem, st);
  }
  if (lpr == 32) return launch_stencil<__nv_bfloat16, 32>(xmap, a, smem, st);
  if (lpr == 16) return launch_stencil<__nv_bfloat16, 16>(xmap, a, smem, st);
  return launch_stencil<__nv_bfloat16, 8>(xmap, a, smem, st);
}

extern "C" int gwen_grid_stencil_fwd(const void* x, void* out, const float* dis_padded,
                                     int64_t dis_pitch, int64_t dis_rows, int64_t batch, int64_t hs, int64_t hd,
                                     int64_t w, int64_t row_off, int64_t feat, int64_t ldx,
                                     int64_t x_bstride, int64_t ldo, int64_t o_bstride, int dtype,
                                     const float* bias, int epilogue, int32_t slab_elems,
                                     int32_t tile_w, void* stream) {
  return stencil_fwd(x, out, dis_padded, dis_pitch, dis_rows, batch, hs, hd, w, row_off, feat, ldx, x_bstride,
                     ldo, o_bstride, dtype, bias, epilogue, slab_elems, tile_w, nullptr, stream);
}

extern "C" int gwen_grid_stencil_peer_fwd(void* x, void* out, const float* dis_padded,
                                          int64_t dis_pitch, int64_t dis_rows, int64_t batch, int64_t hd, int64_t w,
                                          int64_t feat, int64_t x_bstride, int64_t ldo,
                                          int64_t o_bstride, int dtype, const float* bias,
                                          int epilogue, int32_t slab_elems, int32_t tile_w,
                                          const gwen_halo_peers* peers, void* stream) {
  GWEN_CHECK_ARG(peers != nullptr, "null halo peers");
  return stencil_fwd(x, out, dis_padded, dis_pitch, dis_rows, batch, hd + 2, hd, w, 1, feat, feat, x_bstride,
                     ldo, o_bstride, dtype, bias, epilogue, slab_elems, tile_w, peers, stream);
}
