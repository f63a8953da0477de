// bb_xfer: the fused batched put/get kernel (sm_100a).
//
// Replaces the reference's per-shard UCX RMA calls (blackbird_client.cpp:231-237 put,
// :315-327 get) with ONE persistent launch per batch.  Per 16 KiB tile:
//
//   warp 0  producer : descriptor lookup (32 tiles looked up in parallel), zero-pads short
//                      tiles, `cp.async.bulk` global -> shared (TMA, mbarrier complete_tx).
//                      The source may be local HBM (put) or a peer GPU's slab over NVLink (get).
//   warp 1  hash     : one elected thread issues 4x `tcgen05.mma.kind::i8` (M128 N16 K32) that
//                      read the landed tile *in place* as the A operand against the BBH64 weight
//                      matrix; accumulators live in TMEM (16 columns per stage).
//   warp 2  store    : `cp.async.bulk` shared -> global to 1..3 destinations (local, peer-mapped,
//                      i.e. replica fan-out with a single HBM read), or `multimem.st` to an NVLS
//                      multicast address (one store, N replicas).
//   warps 4-7 epilogue: `tcgen05.ld` the 128x16 s32 accumulators, fold them into the
//                      position-dependent 64-bit digest, atomically reduce per object; the last
//                      contributor finalises, compares with the expected digest (get/verify).
//
// Payload bytes never pass through registers on the BBH64/unicast path: TMA in, tensor core
// reads shared memory, TMA out.  Stages recycle through full -> (acc_full) -> empty mbarriers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "common/tchash_def.h"
#include "kernels/ptx.cuh"
#include "kernels/xfer.h"

namespace bb::gpu {
namespace {

using namespace bb::ptx;

constexpr int kStages = 8;
constexpr int kThreads = 256;
constexpr int kStoreLag = 2;  // stores may trail this many tiles before their stage is released
constexpr uint32_t kTmemCols = 128;  // kStages * 16 accumulator columns
static_assert(kStages * tchash::kN == kTmemCols);
static_assert(kTileBytes == tchash::kTileBytes);

__constant__ uint64_t c_col_mul[tchash::kN];

struct StageMeta {
  uint64_t dst[kMaxDst];
  uint32_t desc;
  uint32_t tile_in_obj;
  uint32_t bytes;
  uint32_t ndst_flags;  // ndst | flags << 8
};

struct __align__(1024) Smem {
  uint8_t tile[kStages][kTileBytes];
  uint8_t w[2048];
  uint64_t full[kStages];
  uint64_t acc_full[kStages];
  uint64_t empty[kStages];
  StageMeta meta[kStages];
  uint32_t dirty[kStages];
  uint32_t tmem_base;
};

struct Params {
  const XferDesc* descs;
  const uint32_t* tile_start;
  uint32_t ndesc;
  uint32_t total_tiles;
  unsigned long long* sum_ws;
  uint32_t* done_ws;
  uint64_t* digest_out;
  uint32_t* status_out;
  uint32_t* debug_d;
};

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), o);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), o);
    v += (static_cast<uint64_t>(hi) << 32) | lo;
  }
  return v;
}

// Adds a warp's partial digest of `ntiles_part` tile-quadrants to object `d`; the last
// contributor (4 quadrants x ntiles) finalises, verifies and re-zeroes the workspace.
__device__ __forceinline__ void flush_partial(const Params& p, uint32_t d, uint64_t acc, uint32_t ntiles_part,
                                              uint32_t lane) {
  if (d == 0xFFFFFFFFu) return;
  const uint64_t total = warp_sum64(acc);
  if (lane == 0) {
    atomicAdd(&p.sum_ws[d], static_cast<unsigned long long>(total));
    __threadfence();
    const uint32_t prev = atomicAdd(&p.done_ws[d], ntiles_part);
    const uint32_t ntiles = __ldg(&p.tile_start[d + 1]) - __ldg(&p.tile_start[d]);
    if (prev + ntiles_part == 4u * ntiles) {
      __threadfence();
      const uint64_t sum = atomicExch(&p.sum_ws[d], 0ull);
      p.done_ws[d] = 0;
      const XferDesc* desc = &p.descs[d];
      const uint64_t digest = tchash::finalize(sum, desc->nbytes);
      p.digest_out[d] = digest;
      p.status_out[d] = ((desc->flags & XFER_VERIFY) && digest != desc->expect) ? 1u : 0u;
    }
  }
}

template <int ALGO>
__global__ void __launch_bounds__(kThreads, 1) bb_xfer_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t G = gridDim.x;
  const uint32_t my_tiles = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + G - 1) / G : 0;
  constexpr bool kHash = (ALGO == ALGO_BBH64);

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.empty[i], kHash ? 5 : 1);  // store warp + 4 epilogue warps
      s.dirty[i] = kTileBytes;
    }
    fence_mbar_init();
  }
  if constexpr (kHash) {
    // BBH64 weight matrix W[k][n] as the UMMA B operand (N=16 rows, K-major, no swizzle).
    for (uint32_t o = threadIdx.x; o < 2048; o += kThreads)
      s.w[o] = static_cast<uint8_t>(tchash::weight(tchash::off_to_k(o), tchash::off_to_row(o)));
    fence_proxy_async_smem();
    if (warp == 3) tmem_alloc<kTmemCols>(&s.tmem_base);
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (kHash) tc_fence_after();
  const uint32_t tmem_base = kHash ? s.tmem_base : 0;

  if (warp == 0) {
    // ================================================================ TMA producer
    uint32_t it = 0;
    for (uint32_t base = 0; base < my_tiles; base += 32) {
      const uint32_t idx = base + lane;
      uint64_t src = 0, nbytes = 0, dst0 = 0, dst1 = 0, dst2 = 0;
      uint32_t d = 0, ti = 0, ndst_flags = 0;
      if (idx < my_tiles) {
        const uint32_t t = blockIdx.x + idx * G;
        uint32_t lo = 0, hi = p.ndesc;  // tile_start[lo] <= t < tile_start[hi]
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (__ldg(&p.tile_start[mid]) <= t) lo = mid; else hi = mid;
        }
        d = lo;
        ti = t - __ldg(&p.tile_start[d]);
        const uint4* q = reinterpret_cast<const uint4*>(&p.descs[d]);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
        src = (static_cast<uint64_t>(q0.y) << 32) | q0.x;
        dst0 = (static_cast<uint64_t>(q0.w) << 32) | q0.z;
        dst1 = (static_cast<uint64_t>(q1.y) << 32) | q1.x;
        dst2 = (static_cast<uint64_t>(q1.w) << 32) | q1.z;
        nbytes = (static_cast<uint64_t>(q2.y) << 32) | q2.x;
        ndst_flags = (q2.w & 0xFFu) | (q3.z << 8);
      }
      const uint32_t cnt = min(32u, my_tiles - base);
      for (uint32_t i = 0; i < cnt; ++i, ++it) {
        const uint32_t stage = it % kStages;
        const uint32_t par = (it / kStages) & 1u;
        const uint64_t src_i = shfl64(src, i);
        const uint64_t nbytes_i = shfl64(nbytes, i);
        const uint32_t ti_i = __shfl_sync(0xffffffffu, ti, i);
        const uint64_t off = static_cast<uint64_t>(ti_i) * kTileBytes;
        const uint32_t bytes = static_cast<uint32_t>(min(static_cast<uint64_t>(kTileBytes), nbytes_i - off));
        const uint32_t b16 = bytes & ~15u;
        const uint32_t tail = bytes & 15u;
        mbar_wait(&s.empty[stage], par ^ 1u);
        bool wrote = false;
        if constexpr (kHash) {
          const uint32_t dirty = s.dirty[stage];
          for (uint32_t o = b16 + lane * 16; o < dirty; o += 512) {
            *reinterpret_cast<uint4*>(&s.tile[stage][o]) = make_uint4(0, 0, 0, 0);
            wrote = true;
          }
        }
        if (lane < tail) {
          s.tile[stage][b16 + lane] = *reinterpret_cast<const uint8_t*>(src_i + off + b16 + lane);
          wrote = true;
        }
        if (wrote) fence_proxy_async_smem();
        __syncwarp();
        const uint64_t d0 = shfl64(dst0, i), d1 = shfl64(dst1, i), d2 = shfl64(dst2, i);
        const uint32_t d_i = __shfl_sync(0xffffffffu, d, i);
        const uint32_t nf_i = __shfl_sync(0xffffffffu, ndst_flags, i);
        if (lane == 0) {
          s.dirty[stage] = (bytes + 15u) & ~15u;
          StageMeta& m = s.meta[stage];
          m.dst[0] = d0; m.dst[1] = d1; m.dst[2] = d2;
          m.desc = d_i; m.tile_in_obj = ti_i; m.bytes = bytes; m.ndst_flags = nf_i;
          mbar_arrive_expect_tx(&s.full[stage], b16);
          if (b16) bulk_g2s(s.tile[stage], reinterpret_cast<const void*>(src_i + off), b16, &s.full[stage]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================================================ tensor-core hash issuer
    if constexpr (kHash) {
      constexpr uint32_t idesc = umma_idesc_i8(tchash::kRows, tchash::kN, false, false);
      const uint32_t w_addr = smem_u32(s.w);
      for (uint32_t it = 0; it < my_tiles; ++it) {
        const uint32_t stage = it % kStages;
        const uint32_t par = (it / kStages) & 1u;
        mbar_wait(&s.full[stage], par);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(s.tile[stage]);
          const uint32_t tmem_d = tmem_base + stage * tchash::kN;
#pragma unroll
          for (uint32_t j = 0; j < tchash::kK / 32; ++j) {
            // each MMA consumes K=32 bytes = two 16-byte K chunks (LBO 128 B apart)
            mma_i8_ss(tmem_d, umma_desc_kmajor_noswizzle(a_addr + j * 256, 128, 1024),
                      umma_desc_kmajor_noswizzle(w_addr + j * 256, 128, 1024), idesc, j > 0 ? 1u : 0u);
          }
          tc_commit(&s.acc_full[stage]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================================================================ store warp
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kStages;
      const uint32_t par = (it / kStages) & 1u;
      mbar_wait(&s.full[stage], par);
      const StageMeta m = s.meta[stage];
      const uint64_t off = static_cast<uint64_t>(m.tile_in_obj) * kTileBytes;
      const uint32_t b16 = m.bytes & ~15u;
      const uint32_t tail = m.bytes & 15u;
      const uint32_t ndst = m.ndst_flags & 0xFFu;
      const uint32_t flags = m.ndst_flags >> 8;
      if (flags & XFER_MULTIMEM) {
        uint8_t* mc = reinterpret_cast<uint8_t*>(m.dst[0] + off);
        for (uint32_t o = lane * 16; o < b16; o += 512)
          multimem_st_v4(mc + o, *reinterpret_cast<const uint4*>(&s.tile[stage][o]));
      } else if (lane == 0) {
        for (uint32_t r = 0; r < ndst; ++r)
          if (b16) bulk_s2g(reinterpret_cast<void*>(m.dst[r] + off), s.tile[stage], b16);
      }
      if (lane == 0) bulk_commit();
      if (lane < tail) {
        const uint8_t b = s.tile[stage][b16 + lane];
        for (uint32_t r = 0; r < ndst; ++r) *reinterpret_cast<uint8_t*>(m.dst[r] + off + b16 + lane) = b;
      }
      __syncwarp();
      if (it >= kStoreLag && lane == 0) {
        bulk_wait_read<kStoreLag>();
        mbar_arrive(&s.empty[(it - kStoreLag) % kStages]);
      }
    }
    if (lane == 0) {
      bulk_wait_read<0>();
      const uint32_t first = my_tiles > kStoreLag ? my_tiles - kStoreLag : 0;
      for (uint32_t it = first; it < my_tiles; ++it) mbar_arrive(&s.empty[it % kStages]);
      bulk_wait<0>();  // writes performed before the kernel retires
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue: TMEM -> digest
    if constexpr (kHash) {
      const uint32_t q = warp & 3u;
      const uint32_t row = q * 32 + lane;
      uint32_t cur_d = 0xFFFFFFFFu, cur_tiles = 0;
      uint64_t acc = 0;
      for (uint32_t it = 0; it < my_tiles; ++it) {
        const uint32_t stage = it % kStages;
        const uint32_t par = (it / kStages) & 1u;
        mbar_wait(&s.acc_full[stage], par);
        mbar_wait(&s.full[stage], par);  // already complete; acquires the producer's StageMeta writes
        tc_fence_after();
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + ((q * 32u) << 16) + stage * tchash::kN, r);
        tmem_ld_wait();
        const uint32_t d = s.meta[stage].desc;
        const uint32_t ti = s.meta[stage].tile_in_obj;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[stage]);
        if (d != cur_d) {
          flush_partial(p, cur_d, acc, cur_tiles, lane);
          cur_d = d;
          acc = 0;
          cur_tiles = 0;
        }
        uint64_t rr = 0;
#pragma unroll
        for (int n = 0; n < 16; ++n) rr += static_cast<uint64_t>(r[n]) * c_col_mul[n];
        acc += tchash::row_contrib(rr, static_cast<uint64_t>(ti) * tchash::kRows + row);
        ++cur_tiles;
        if (p.debug_d) {
          const uint64_t t = blockIdx.x + static_cast<uint64_t>(it) * G;
          uint32_t* o = p.debug_d + (t * tchash::kRows + row) * tchash::kN;
#pragma unroll
          for (int n = 0; n < 16; ++n) o[n] = r[n];
        }
      }
      flush_partial(p, cur_d, acc, cur_tiles, lane);
    }
  }

  // ------------------------------------------------------------------ teardown
  if constexpr (kHash) tc_fence_before();
  __syncthreads();
  if constexpr (kHash) {
    if (warp == 3) {
      tc_fence_after();
      tmem_dealloc<kTmemCols>(tmem_base);
    }
  }
}

int g_sm_count[16] = {0};
bool g_const_init[16] = {false};

int sm_count(int dev) {
  if (dev < 0 || dev >= 16) dev = 0;
  if (!g_sm_count[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

template <int ALGO>
cudaError_t launch_t(const XferLaunch& l, const Params& p, int grid, cudaStream_t st) {
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bb_xfer_kernel<ALGO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(Smem)));
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  bb_xfer_kernel<ALGO><<<grid, kThreads, sizeof(Smem), st>>>(p);
  return cudaGetLastError();
}

}  // namespace

int xfer_smem_bytes(int) { return static_cast<int>(sizeof(Smem)); }

int launch_xfer(const XferLaunch& l) {
  if (l.ndesc == 0 || l.total_tiles == 0) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev >= 0 && dev < 16 && !g_const_init[dev]) {
    uint64_t h[tchash::kN];
    for (uint32_t n = 0; n < tchash::kN; ++n) h[n] = tchash::col_mul(n);
    e = cudaMemcpyToSymbol(c_col_mul, h, sizeof h);
    if (e != cudaSuccess) return static_cast<int>(e);
    g_const_init[dev] = true;
  }
  Params p;
  p.descs = l.descs;
  p.tile_start = l.tile_start;
  p.ndesc = l.ndesc;
  p.total_tiles = l.total_tiles;
  p.sum_ws = reinterpret_cast<unsigned long long*>(l.sum_ws);
  p.done_ws = l.done_ws;
  p.digest_out = l.digest_out;
  p.status_out = l.status_out;
  p.debug_d = l.debug_d;
  int grid = l.max_ctas > 0 ? l.max_ctas : sm_count(dev);
  grid = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(grid), l.total_tiles));
  cudaStream_t st = static_cast<cudaStream_t>(l.stream);
  switch (l.algo) {
    case ALGO_NONE: e = launch_t<ALGO_NONE>(l, p, grid, st); break;
    case ALGO_BBH64: e = launch_t<ALGO_BBH64>(l, p, grid, st); break;
    default: return static_cast<int>(cudaErrorNotSupported);
  }
  return static_cast<int>(e);
}

}  // namespace bb::gpu
