// bb_xfer_fp8: fused MXFP8 transfer kernels (sm_100a; SURVEY K12 "fp8 pack fused into put / unpack into get").
//
//   put  (PACK)   : TMA loads a 32 KiB bf16 tile -> 8 pack warps compute the E8M0 block scales and the E4M3
//                   payload *in shared memory* -> the 16 KiB payload tile is hashed in place by tcgen05.mma
//                   (BBH64, same tile hash as bb_xfer) -> TMA stores payload + 512 B of scales to the
//                   (peer) slab.  The bf16 data is read from HBM once and only the 0.52x packed bytes cross NVLink.
//   get  (UNPACK) : TMA loads payload tile + scales from the slab -> tensor-core hash of the payload tile ->
//                   8 unpack warps expand to bf16 in shared memory -> TMA stores the 32 KiB tile to the caller.
//
// The stored object keeps the layout of the unfused path ([E4M3 payload n][E8M0 scales n/32], common/mxfp8.h), and so
// does its digest: BBH64 is a sum of per-tile terms, this kernel accumulates the payload tiles (tile index 0..T-1)
// and returns the *unfinalised* sum per object; the scales region (tiles T..) is hashed as an XFER_RAW_SUM slice
// by bb_xfer and the host adds the two sums and finalises.  Objects must hold a multiple of 16384 elements (whole
// payload tiles); other shapes take the unfused path.
//
// Warp roles (12 warps): 0 producer, 1 MMA issuer, 2 store, 3 accumulator, 4-7 epilogue (TMEM -> row hashes),
// 8-11 + the idle lanes of nothing else: convert warps (pack / unpack).  4-stage ring:
// per stage 32 KiB wide tile + 16 KiB payload tile + 512 B scales.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>

#include "common/tchash_def.h"
#include "kernels/ptx.cuh"
#include "kernels/xfer.h"

namespace bb::gpu {
namespace {

using namespace bb::ptx;

constexpr int kFpStages = 4;
constexpr int kFpThreads = 384;
constexpr int kConvertWarps = 4;            // warps 8..11
constexpr uint32_t kWideBytes = 2 * kTileBytes;  // bf16 side of one payload tile
constexpr uint32_t kScaleBytes = kTileBytes / 32;
constexpr uint32_t kFpTmemCols = 64;        // kFpStages * 16 accumulator columns
constexpr int kFpStoreLag = 1;
static_assert(kFpStages * tchash::kN == kFpTmemCols);

__constant__ uint64_t c_fp_col_mul[tchash::kN];

struct FpMeta {
  uint64_t dst_payload;  // pack: payload destination; unpack: bf16 destination
  uint64_t dst_scales;   // pack: scales destination
  uint32_t desc;
  uint32_t tile_in_obj;
  uint32_t obj_ntiles;
  uint32_t pad;
};

struct FpLookup {
  FpMeta m;
  uint64_t src;         // pack: bf16 source; unpack: payload source
  uint64_t src_scales;  // unpack: scales source
};

struct __align__(1024) SmemFp {
  uint8_t tile[kFpStages][kTileBytes];
  uint8_t wide[kFpStages][kWideBytes];
  uint8_t scales[kFpStages][kScaleBytes];
  uint8_t w[2048];
  uint64_t loaded[kFpStages];     // TMA landed (pack: wide; unpack: tile + scales)
  uint64_t converted[kFpStages];  // convert warps done (pack: tile + scales written; unpack: wide written)
  uint64_t acc_full[kFpStages];
  uint64_t epi_done[kFpStages];
  uint64_t empty[kFpStages];
  FpMeta meta[kFpStages];
  uint64_t part[kFpStages][4];
  FpLookup lk[32];
  uint32_t tmem_base;
};

struct FpParams {
  const XferDesc* descs;
  const uint32_t* tile_start;
  uint32_t ndesc;
  uint32_t total_tiles;
  unsigned long long* sum_ws;  // [ndesc] raw BBH64 sums of the payload tiles (zero on entry)
};

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), o);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), o);
    v += (static_cast<uint64_t>(hi) << 32) | lo;
  }
  return v;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t b16) { return __uint_as_float(b16 << 16); }

// One 16-byte chunk (8 bf16) per lane, 4 lanes per 32-element block: identical arithmetic to mxfp8_pack_kernel
// (mxfp8.cu) and to the CPU reference (common/mxfp8.cpp).
__device__ __forceinline__ void pack_chunk(const uint4 v, uint2* payload_out, uint8_t* scale_out, bool write_scale) {
  float f[8];
  f[0] = bf16_bits_to_float(v.x & 0xFFFFu); f[1] = bf16_bits_to_float(v.x >> 16);
  f[2] = bf16_bits_to_float(v.y & 0xFFFFu); f[3] = bf16_bits_to_float(v.y >> 16);
  f[4] = bf16_bits_to_float(v.z & 0xFFFFu); f[5] = bf16_bits_to_float(v.z >> 16);
  f[6] = bf16_bits_to_float(v.w & 0xFFFFu); f[7] = bf16_bits_to_float(v.w >> 16);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = fabsf(f[i]);
    amax = (a > amax) ? a : amax;  // NaN compares false and is skipped
  }
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
  amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
  int e = 127;
  if (amax > 0.f && amax < __int_as_float(0x7F800000)) {
    const int be = (__float_as_int(amax) >> 23) & 0xFF;
    e = (be == 0 ? -127 : be - 127) - 8 + 127;
    e = max(0, min(254, e));
  }
  const float inv = __int_as_float((254 - e) << 23);
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo |= static_cast<uint32_t>(__nv_cvt_float_to_fp8(f[i] * inv, __NV_SATFINITE, __NV_E4M3)) << (8 * i);
    hi |= static_cast<uint32_t>(__nv_cvt_float_to_fp8(f[4 + i] * inv, __NV_SATFINITE, __NV_E4M3)) << (8 * i);
  }
  *payload_out = make_uint2(lo, hi);
  if (write_scale) *scale_out = static_cast<uint8_t>(e);
}

__device__ __forceinline__ uint4 unpack_chunk(const uint2 p, uint8_t scale) {
  const float s = __int_as_float(static_cast<int>(scale) << 23);  // 2^(e-127); e == 0 flushes to zero
  uint32_t out[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w = i < 2 ? p.x : p.y;
    const uint8_t b0 = (w >> (16 * (i & 1))) & 0xFF, b1 = (w >> (16 * (i & 1) + 8)) & 0xFF;
    const float f0 = __half2float(__half(__nv_cvt_fp8_to_halfraw(b0, __NV_E4M3))) * s;
    const float f1 = __half2float(__half(__nv_cvt_fp8_to_halfraw(b1, __NV_E4M3))) * s;
    const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
    out[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}

template <bool UNPACK>
__global__ void __launch_bounds__(kFpThreads, 1) bb_xfer_fp8_kernel(const __grid_constant__ FpParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemFp& s = *reinterpret_cast<SmemFp*>(smem_raw);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;

  const uint32_t tpc = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t0 = blockIdx.x * tpc;
  const uint32_t my_tiles = t0 < p.total_tiles ? min(tpc, p.total_tiles - t0) : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFpStages; ++i) {
      mbar_init(&s.loaded[i], 1);
      mbar_init(&s.converted[i], kConvertWarps);
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.epi_done[i], 4);
      mbar_init(&s.empty[i], 2);  // store warp + accumulator
    }
    fence_mbar_init();
  }
  for (uint32_t o = threadIdx.x; o < 2048; o += kFpThreads)
    s.w[o] = static_cast<uint8_t>(tchash::weight(tchash::off_to_k(o), tchash::off_to_row(o)));
  fence_proxy_async_smem();
  if (warp == 3) tmem_alloc<kFpTmemCols>(&s.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s.tmem_base;

  if (warp == 0) {
    // ================================================================ TMA producer
    uint32_t it = 0;
    for (uint32_t base = 0; base < my_tiles; base += 32) {
      const uint32_t idx = base + lane;
      if (idx < my_tiles) {
        const uint32_t t = t0 + idx;
        uint32_t lo = 0, hi = p.ndesc;
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (__ldg(&p.tile_start[mid]) <= t) lo = mid; else hi = mid;
        }
        const uint32_t first = __ldg(&p.tile_start[lo]);
        const uint32_t next = __ldg(&p.tile_start[lo + 1]);
        const uint4* q = reinterpret_cast<const uint4*>(&p.descs[lo]);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
        const uint64_t src = (static_cast<uint64_t>(q0.y) << 32) | q0.x;
        const uint64_t d0 = (static_cast<uint64_t>(q0.w) << 32) | q0.z;
        const uint64_t d1 = (static_cast<uint64_t>(q1.y) << 32) | q1.x;
        const uint32_t ti = t - first;
        FpLookup& e = s.lk[lane];
        e.m.desc = lo;
        e.m.tile_in_obj = ti;
        e.m.obj_ntiles = next - first;
        if constexpr (UNPACK) {
          e.src = src + static_cast<uint64_t>(ti) * kTileBytes;          // payload tile in the slab
          e.src_scales = d1 + static_cast<uint64_t>(ti) * kScaleBytes;   // its scales
          e.m.dst_payload = d0 + static_cast<uint64_t>(ti) * kWideBytes;  // bf16 destination
          e.m.dst_scales = 0;
        } else {
          e.src = src + static_cast<uint64_t>(ti) * kWideBytes;           // bf16 source tile
          e.src_scales = 0;
          e.m.dst_payload = d0 + static_cast<uint64_t>(ti) * kTileBytes;
          e.m.dst_scales = d1 + static_cast<uint64_t>(ti) * kScaleBytes;
        }
      }
      __syncwarp();
      const uint32_t cnt = min(32u, my_tiles - base);
      for (uint32_t i = 0; i < cnt; ++i, ++it) {
        const uint32_t stage = it % kFpStages;
        const uint32_t par = (it / kFpStages) & 1u;
        mbar_wait(&s.empty[stage], par ^ 1u);
        const FpLookup& e = s.lk[i];
        if (lane < 2) reinterpret_cast<uint4*>(&s.meta[stage])[lane] = reinterpret_cast<const uint4*>(&e.m)[lane];
        __syncwarp();
        if (lane == 0) {
          if constexpr (UNPACK) {
            mbar_arrive_expect_tx(&s.loaded[stage], kTileBytes + kScaleBytes);
            bulk_g2s(s.tile[stage], reinterpret_cast<const void*>(e.src), kTileBytes, &s.loaded[stage]);
            bulk_g2s(s.scales[stage], reinterpret_cast<const void*>(e.src_scales), kScaleBytes, &s.loaded[stage]);
          } else {
            mbar_arrive_expect_tx(&s.loaded[stage], kWideBytes);
            bulk_g2s(s.wide[stage], reinterpret_cast<const void*>(e.src), kWideBytes, &s.loaded[stage]);
          }
        }
        __syncwarp();
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================================================ tensor-core hash issuer (payload tile)
    constexpr uint32_t idesc = umma_idesc_i8(tchash::kRows, tchash::kN, false, false);
    const uint32_t w_addr = smem_u32(s.w);
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kFpStages;
      const uint32_t par = (it / kFpStages) & 1u;
      if constexpr (UNPACK) mbar_wait(&s.loaded[stage], par);     // payload tile landed by TMA
      else mbar_wait(&s.converted[stage], par);                   // payload tile written by the pack warps
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(s.tile[stage]);
        const uint32_t tmem_d = tmem_base + stage * tchash::kN;
#pragma unroll
        for (uint32_t j = 0; j < tchash::kK / 32; ++j)
          mma_i8_ss(tmem_d, umma_desc_kmajor_noswizzle(a_addr + j * 256, 128, 1024), umma_desc_kmajor_noswizzle(w_addr + j * 256, 128, 1024),
                    idesc, j > 0 ? 1u : 0u);
        tc_commit(&s.acc_full[stage]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ================================================================ store warp
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kFpStages;
      const uint32_t par = (it / kFpStages) & 1u;
      mbar_wait(&s.converted[stage], par);
      if constexpr (!UNPACK) mbar_wait(&s.loaded[stage], par);  // (already complete) acquires the producer's FpMeta
      else mbar_wait(&s.loaded[stage], par);
      const FpMeta& m = s.meta[stage];
      if (lane == 0) {
        if constexpr (UNPACK) {
          bulk_s2g(reinterpret_cast<void*>(m.dst_payload), s.wide[stage], kWideBytes);
        } else {
          bulk_s2g(reinterpret_cast<void*>(m.dst_payload), s.tile[stage], kTileBytes);
          bulk_s2g(reinterpret_cast<void*>(m.dst_scales), s.scales[stage], kScaleBytes);
        }
        bulk_commit();
      }
      __syncwarp();
      if (it >= kFpStoreLag && lane == 0) {
        bulk_wait_read<kFpStoreLag>();
        mbar_arrive(&s.empty[(it - kFpStoreLag) % kFpStages]);
      }
    }
    if (lane == 0) {
      bulk_wait_read<0>();
      const uint32_t first = my_tiles > kFpStoreLag ? my_tiles - kFpStoreLag : 0;
      for (uint32_t it = first; it < my_tiles; ++it) mbar_arrive(&s.empty[it % kFpStages]);
      bulk_wait<0>();
    }
  } else if (warp == 3) {
    // ================================================================ accumulator: per-object raw sums
    uint32_t cur_d = 0xFFFFFFFFu;
    uint64_t acc = 0;
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kFpStages;
      const uint32_t par = (it / kFpStages) & 1u;
      mbar_wait(&s.epi_done[stage], par);
      const uint32_t d = s.meta[stage].desc;
      const uint64_t sum = s.part[stage][0] + s.part[stage][1] + s.part[stage][2] + s.part[stage][3];
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.empty[stage]);
      if (d != cur_d) {
        if (cur_d != 0xFFFFFFFFu && lane == 0) atomicAdd(&p.sum_ws[cur_d], static_cast<unsigned long long>(acc));
        cur_d = d;
        acc = 0;
      }
      acc += sum;
    }
    if (cur_d != 0xFFFFFFFFu && lane == 0) atomicAdd(&p.sum_ws[cur_d], static_cast<unsigned long long>(acc));
  } else if (warp < 8) {
    // ================================================================ epilogue warps 4..7: TMEM -> row hashes
    const uint32_t q = warp & 3u;
    const uint32_t row = q * 32 + lane;
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kFpStages;
      const uint32_t par = (it / kFpStages) & 1u;
      mbar_wait(&s.acc_full[stage], par);
      mbar_wait(&s.loaded[stage], par);  // acquires the producer's FpMeta writes
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem_base + ((q * 32u) << 16) + stage * tchash::kN, r);
      tmem_ld_wait();
      tc_fence_before();
      const uint32_t ti = s.meta[stage].tile_in_obj;
      uint64_t rr = 0;
#pragma unroll
      for (int n = 0; n < 16; ++n) rr += static_cast<uint64_t>(r[n]) * c_fp_col_mul[n];
      const uint64_t tot = warp_sum64(tchash::row_contrib(rr, static_cast<uint64_t>(ti) * tchash::kRows + row));
      if (lane == 0) {
        s.part[stage][q] = tot;
        mbar_arrive(&s.epi_done[stage]);
      }
    }
  } else {
    // ================================================================ convert warps 8..11
    const uint32_t cw = warp - 8;  // 0..3
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t stage = it % kFpStages;
      const uint32_t par = (it / kFpStages) & 1u;
      mbar_wait(&s.loaded[stage], par);
      // 2048 16-byte bf16 chunks per tile; warp cw takes chunks [cw*512, cw*512+512) -> 16 iterations of 32 lanes
      if constexpr (UNPACK) {
        const uint2* pay = reinterpret_cast<const uint2*>(s.tile[stage]);
        uint4* out = reinterpret_cast<uint4*>(s.wide[stage]);
#pragma unroll 4
        for (uint32_t k = 0; k < 16; ++k) {
          const uint32_t c = cw * 512 + k * 32 + lane;
          out[c] = unpack_chunk(pay[c], s.scales[stage][c >> 2]);
        }
      } else {
        const uint4* in = reinterpret_cast<const uint4*>(s.wide[stage]);
        uint2* pay = reinterpret_cast<uint2*>(s.tile[stage]);
#pragma unroll 4
        for (uint32_t k = 0; k < 16; ++k) {
          const uint32_t c = cw * 512 + k * 32 + lane;
          pack_chunk(in[c], &pay[c], &s.scales[stage][c >> 2], (c & 3u) == 0);
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> tensor core / TMA store readers
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.converted[stage]);
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc<kFpTmemCols>(tmem_base);
  }
}

struct FpDeviceState {
  bool consts = false;
  bool attr[2] = {false, false};
  int sm_count = 0;
};
FpDeviceState g_fp_dev[16];
std::mutex g_fp_mu;

}  // namespace

int xfer_fp8_smem_bytes() { return static_cast<int>(sizeof(SmemFp)); }

int launch_xfer_fp8(const XferLaunch& l, bool unpack) {
  if (l.ndesc == 0 || l.total_tiles == 0) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev < 0 || dev >= 16) return static_cast<int>(cudaErrorInvalidDevice);
  FpDeviceState& ds = g_fp_dev[dev];
  {
    std::lock_guard<std::mutex> lk(g_fp_mu);
    if (!ds.consts) {
      uint64_t h[tchash::kN];
      for (uint32_t n = 0; n < tchash::kN; ++n) h[n] = tchash::col_mul(n);
      e = cudaMemcpyToSymbol(c_fp_col_mul, h, sizeof h);
      if (e != cudaSuccess) return static_cast<int>(e);
      int n = 0;
      if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
      ds.sm_count = n;
      ds.consts = true;
    }
    if (!ds.attr[unpack ? 1 : 0]) {
      e = unpack ? cudaFuncSetAttribute(bb_xfer_fp8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmemFp)))
                 : cudaFuncSetAttribute(bb_xfer_fp8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmemFp)));
      if (e != cudaSuccess) return static_cast<int>(e);
      ds.attr[unpack ? 1 : 0] = true;
    }
  }
  FpParams p;
  p.descs = l.descs;
  p.tile_start = l.tile_start;
  p.ndesc = l.ndesc;
  p.total_tiles = l.total_tiles;
  p.sum_ws = reinterpret_cast<unsigned long long*>(l.sum_ws);
  int grid = l.max_ctas > 0 ? l.max_ctas : std::min(ds.sm_count, 128);
  grid = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(grid), l.total_tiles));
  cudaStream_t st = static_cast<cudaStream_t>(l.stream);
  if (unpack) bb_xfer_fp8_kernel<true><<<grid, kFpThreads, sizeof(SmemFp), st>>>(p);
  else bb_xfer_fp8_kernel<false><<<grid, kFpThreads, sizeof(SmemFp), st>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace bb::gpu
