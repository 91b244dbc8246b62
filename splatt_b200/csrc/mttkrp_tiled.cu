// 3-mode root MTTKRP with the LEAF factor staged tile by tile in shared memory.
//
// The generic root kernel sits on the measured ceiling of its access pattern: two
// factor rows per nonzero gathered through L2 -> SM.  When one SM's
// nonzeros touch every leaf row several times (nnz per SM >> rows of the leaf factor),
// half of those gathers can be served from shared memory instead:
//
//   * the stream is built "CTA-tiled" (stream_build.cu): the records of CTA r's range
//     are regrouped by leaf tile, seg_off[r * ntiles + t] marks the segments;
//   * one persistent CTA per SM walks its range tile by tile; a producer warp streams
//     the leaf-factor tiles (TMA bulk copies, double buffered, mbarrier full/empty);
//   * consumer warps stage their records through a private TMA ring as before, read
//     the leaf row from the tile (LDS.128) and gather only the parent row from L2;
//   * slice / sub-range ends reduce into the output with red.global.add.f64.
//
// Same results as the generic kernel (linearity); chosen by spb200_launch_mttkrp when the
// stream carries the tiling and the launch parameters fit.
#include "mttkrp_kernels.cuh"

namespace spb200 {

constexpr int kTW  = 24;          // consumer warps per CTA (+ 1 producer warp)
constexpr int kTRS = 96;          // records per warp per staging round
constexpr int kTB  = 4;           // records whose gathers are issued together

struct TiledArgs {
  const SpRec *    rec;
  const uint32_t * rootid;
  const uint32_t * seg_off;       // this grid's ranges: [gridDim.x * ntiles + 1]
  const double *   leaf;
  const double *   parent;
  double *         out;
  uint32_t         ntiles, tile_rows, leaf_rows;
  uint32_t         tile_bytes;    // bytes reserved per tile buffer (tile_rows * pitch)
  int              ldm, ncols, col0;
};

__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

template <int L>
__global__ void __launch_bounds__((kTW + 1) * 32, 1) mttkrp_tiled_root3(const TiledArgs a) {
  constexpr int G  = 32 / L;
  constexpr int NG = kTW * G;     // lane groups per CTA

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char * tiles = smem;                                   // 2 x tile_bytes
  SpRec *    ring = reinterpret_cast<SpRec *>(smem + 2 * a.tile_bytes);       // [kTW][2][kTRS]
  uint64_t * bars = reinterpret_cast<uint64_t *>(ring + kTW * 2 * kTRS);
  uint64_t * tile_full  = bars;            // [2]
  uint64_t * tile_empty = bars + 2;        // [2]
  uint64_t * rec_full   = bars + 4;        // [kTW][2]

  const int      warp  = threadIdx.x >> 5;
  const int      lane  = threadIdx.x & 31;
  const uint32_t pitch = static_cast<uint32_t>(a.ldm) * 8u;
  const uint32_t NT    = a.ntiles;
  const uint32_t * so  = a.seg_off + static_cast<size_t>(blockIdx.x) * NT;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&tile_full[b], 1); mbar_init(&tile_empty[b], kTW); }
    for (int w = 0; w < kTW * 2; ++w) mbar_init(&rec_full[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  auto tile_rows_of = [&](uint32_t t) {
    const uint32_t first = t * a.tile_rows;
    return min(a.tile_rows, a.leaf_rows - first);
  };

  // ------------------------------------------------------------------ producer warp
  if (warp == kTW) {
    if (lane == 0) {
      for (uint32_t t = 0; t < NT; ++t) {
        const uint32_t b = t & 1u;
        if (t >= 2) mbar_wait(&tile_empty[b], ((t - 2) >> 1) & 1u);   // all warps left tile t-2
        const uint32_t bytes = tile_rows_of(t) * pitch;
        mbar_arrive_expect_tx(&tile_full[b], bytes);
        const char * src = reinterpret_cast<const char *>(a.leaf) +
                           static_cast<size_t>(t) * a.tile_rows * pitch;
        // bulk copies of at most 32 KB each
        for (uint32_t off = 0; off < bytes; off += 32768u)
          tma_bulk_g2s(tiles + b * a.tile_bytes + off, src + off, min(32768u, bytes - off),
                       &tile_full[b]);
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  const int  grp    = lane / L;
  const int  gl     = lane % L;
  const bool act    = (2 * gl) < a.ncols;
  const bool leader = (gl == 0);
  const int  colx   = act ? (a.col0 + 2 * gl) : a.col0;
  const char * pbase = reinterpret_cast<const char *>(a.parent + colx);
  char *       obase = reinterpret_cast<char *>(a.out + colx);
  const uint32_t tcol = static_cast<uint32_t>(colx) * 8u;      // byte offset of this lane's columns
  SpRec *    myring = ring + warp * 2 * kTRS;
  uint64_t * mybars = rec_full + warp * 2;

  // part of segment t that belongs to lane-group gi / to this warp
  auto part = [&](uint32_t t, uint32_t g0, uint32_t g1, uint32_t & lo, uint32_t & hi) {
    const uint32_t s0 = so[t], len = so[t + 1] - s0;
    lo = s0 + static_cast<uint32_t>(static_cast<unsigned long long>(g0) * len / NG);
    hi = s0 + static_cast<uint32_t>(static_cast<unsigned long long>(g1) * len / NG);
  };

  // issue side: rounds are enumerated tile-major, at least one (possibly empty) per tile
  uint32_t it = 0, ioff = 0, ij = 0;
  auto issue_next = [&]() {
    if (it >= NT) return;
    uint32_t ws, we;
    part(it, warp * G, (warp + 1) * G, ws, we);
    const uint32_t rs  = ws + ioff;
    const uint32_t cnt = (rs < we) ? min(static_cast<uint32_t>(kTRS), we - rs) : 0u;
    if (lane == 0) {
      uint64_t * bar = &mybars[ij & 1u];
      if (cnt) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(bar, cnt * 16u);
        tma_bulk_g2s(myring + (ij & 1u) * kTRS, a.rec + rs, cnt * 16u, bar);
      } else {
        mbar_arrive(bar);
      }
    }
    ++ij;
    ioff += kTRS;
    if (ws + ioff >= we) { ++it; ioff = 0; }
  };
  issue_next();
  issue_next();

  const double2 zero2 = make_double2(0.0, 0.0);
  double2 acc1 = zero2, acc0 = zero2;      // fiber / slice partial sums
  uint32_t j = 0;                          // rounds consumed

  for (uint32_t t = 0; t < NT; ++t) {
    mbar_wait(&tile_full[t & 1u], (t >> 1) & 1u);
    const unsigned char * tile = tiles + (t & 1u) * a.tile_bytes + tcol;
    const uint32_t kbase = t * a.tile_rows;
    uint32_t ws, we, gs, ge;
    part(t, warp * G, (warp + 1) * G, ws, we);
    part(t, warp * G + grp, warp * G + grp + 1, gs, ge);
    const uint32_t nr = (we > ws) ? (we - ws + kTRS - 1) / kTRS : 1u;
    for (uint32_t i = 0; i < nr; ++i, ++j) {
      mbar_wait(&mybars[j & 1u], (j >> 1) & 1u);
      SpRec *        buf = myring + (j & 1u) * kTRS;
      const uint32_t rs  = ws + i * kTRS;
      const uint32_t re  = min(we, rs + kTRS);
      const uint32_t lo  = max(gs, rs), hi = min(ge, re);
      // the group's last record of this tile closes the slice (sub-range boundary)
      if (leader && hi > lo && hi == ge)
        buf[hi - 1 - rs].aux = (buf[hi - 1 - rs].aux & SPB200_IDX_MASK) | (2u << SPB200_IDX_BITS);
      __syncwarp();
      if (act && hi > lo) {
        uint32_t n = lo;
        for (; n + kTB <= hi; n += kTB) {
          uint4   q[kTB];
          double2 b[kTB], r[kTB];
          uint32_t any = 0;
#pragma unroll
          for (int u = 0; u < kTB; ++u) {
            q[u] = *reinterpret_cast<const uint4 *>(&buf[n + u - rs]);
            any |= q[u].w;
          }
#pragma unroll
          for (int u = 0; u < kTB; ++u)
            if (q[u].w >> SPB200_IDX_BITS) r[u] = ld_row_na(pbase, q[u].w & SPB200_IDX_MASK, pitch);
#pragma unroll
          for (int u = 0; u < kTB; ++u)
            b[u] = *reinterpret_cast<const double2 *>(tile + static_cast<size_t>(q[u].z - kbase) * pitch);
          if ((any >> (SPB200_IDX_BITS + 1)) == 0) {
#pragma unroll
            for (int u = 0; u < kTB; ++u) {
              const double v = __hiloint2double(static_cast<int>(q[u].y), static_cast<int>(q[u].x));
              acc1           = fma2(v, b[u], acc1);
              if (q[u].w >> SPB200_IDX_BITS) { acc0 = fma2(acc1, r[u], acc0); acc1 = zero2; }
            }
          } else {
#pragma unroll
            for (int u = 0; u < kTB; ++u) {
              const double   v = __hiloint2double(static_cast<int>(q[u].y), static_cast<int>(q[u].x));
              const uint32_t c = q[u].w >> SPB200_IDX_BITS;
              acc1             = fma2(v, b[u], acc1);
              if (c) {
                acc0 = fma2(acc1, r[u], acc0);
                acc1 = zero2;
                if (c >= 2) {
                  red_row(obase, __ldg(&a.rootid[n + u]), pitch, acc0);
                  acc0 = zero2;
                }
              }
            }
          }
        }
        for (; n < hi; ++n) {
          const uint4    q = *reinterpret_cast<const uint4 *>(&buf[n - rs]);
          const double   v = __hiloint2double(static_cast<int>(q.y), static_cast<int>(q.x));
          const uint32_t c = q.w >> SPB200_IDX_BITS;
          const double2  b = *reinterpret_cast<const double2 *>(tile + static_cast<size_t>(q.z - kbase) * pitch);
          acc1             = fma2(v, b, acc1);
          if (c) {
            acc0 = fma2(acc1, ld_row_na(pbase, q.w & SPB200_IDX_MASK, pitch), acc0);
            acc1 = zero2;
            if (c >= 2) {
              red_row(obase, __ldg(&a.rootid[n]), pitch, acc0);
              acc0 = zero2;
            }
          }
        }
      }
      __syncwarp();
      issue_next();        // refill the stage just consumed with round j + 2
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&tile_empty[t & 1u]);    // this warp is done with tile t
  }
}

}  // namespace spb200

// smem the kernel needs for a given tile size
static size_t tiled_smem_bytes(uint32_t tile_bytes) {
  return 2 * (size_t)tile_bytes + sizeof(SpRec) * spb200::kTW * 2 * spb200::kTRS +
         sizeof(uint64_t) * (4 + 2 * spb200::kTW) + 128;
}

uint32_t spb200_tiled_rows_for(int ncolumns) {
  const size_t pitch = (size_t)(ncolumns + (ncolumns & 1)) * 8;
  const size_t fixed = tiled_smem_bytes(0);
  const size_t avail = (227 * 1024 - fixed) / 2;
  return (uint32_t)(avail / pitch);
}

bool spb200_tiled_applicable(const FiberStream & s, int kind, int ncolumns, int ldm) {
  if (s.nmodes != 3 || kind != SPB200_KIND_ROOT || !s.seg_off || !s.rootid || s.ntiles == 0) return false;
  const int rpad = ncolumns + (ncolumns & 1);
  if (rpad > 64) return false;                       // single column pass only
  const size_t tile_bytes = (size_t)s.ktile_rows * ldm * 8;
  return tiled_smem_bytes((uint32_t)tile_bytes) <= 227 * 1024;
}

int spb200_launch_tiled_root3(const FiberStream & s, int ncolumns, int ldm, uint64_t leaf_rows,
                                   const double * leaf, const double * parent, double * d_out,
                                   cudaStream_t stream) {
  using namespace spb200;
  TiledArgs a;
  a.rec = s.rec; a.rootid = s.rootid; a.seg_off = s.seg_off;
  a.leaf = leaf; a.parent = parent; a.out = d_out;
  a.ntiles = s.ntiles; a.tile_rows = s.ktile_rows; a.leaf_rows = (uint32_t)leaf_rows;
  a.ldm = ldm; a.col0 = 0; a.ncols = ncolumns + (ncolumns & 1);
  a.tile_bytes = s.ktile_rows * (uint32_t)ldm * 8u;
  const size_t smem = tiled_smem_bytes(a.tile_bytes);
  const int threads = (kTW + 1) * 32;
  const unsigned grid = s.kranges;
#define SPB200_TILED_LAUNCH(LL)                                                                   \
  do {                                                                                            \
    static bool set[64] = {false};          /* function attributes are per device */            \
    int dev_ = 0;                                                                                 \
    SPB200_CUDA_OK(cudaGetDevice(&dev_));                                                         \
    if (dev_ < 0 || dev_ >= 64 || !set[dev_]) {                                                   \
      SPB200_CUDA_OK(cudaFuncSetAttribute(mttkrp_tiled_root3<LL>,                                 \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      if (dev_ >= 0 && dev_ < 64) set[dev_] = true;                                               \
    }                                                                                             \
    mttkrp_tiled_root3<LL><<<grid, threads, smem, stream>>>(a);                                   \
  } while (0)
  if (a.ncols <= 8) SPB200_TILED_LAUNCH(4);
  else if (a.ncols <= 16) SPB200_TILED_LAUNCH(8);
  else if (a.ncols <= 32) SPB200_TILED_LAUNCH(16);
  else SPB200_TILED_LAUNCH(32);
#undef SPB200_TILED_LAUNCH
  spb200_count_launches(1);
  SPB200_CUDA_OK(cudaGetLastError());
  return SPLATT_SUCCESS;
}
