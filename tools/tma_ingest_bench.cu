// tma_ingest_bench.cu - how fast can ONE SM pull GEMM operand tiles out of L2? (B200 microbenchmark, tools only)
//
// The wide rollout's tcgen05 GEMM needs 64 KB of operands (A/B x hi/lo, 128 rows x 128 B each) per 768 MMA cycles
// and runs at ~2.5x that. This program times the same shared-memory ring WITHOUT any MMA, in several load styles:
//   0  4 x cp.async.bulk.tensor.2d (128 rows x 32 floats, SWIZZLE_128B) per stage, one producer lane   (= the GEMM)
//   1  same, two producer lanes in two warps (2 loads each)
//   2  4 x cp.async.bulk 1-D, contiguous 16 KB chunks (pre-tiled global layout)
//   3  cp.async 16 B (LDGSTS) by 8 warps, completion through the mbarrier
//   4  half by TMA (mode 0, 2 loads) + half by LDGSTS (8 warps)
//   5  cluster of 2, A tile multicast (each CTA issues 64 of the 128 A rows to both), B tile private
//   6  mode 0 with 32 KB stages x 6 (same bytes in flight, finer granularity)
// Output: bytes per cycle per SM (mean over CTAs) for a grid of 64 and of 148 CTAs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_ingest_bench tools/tma_ingest_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "W_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni W_DONE;\n\t"
      "bra.uni W_LOOP;\n\t"
      "W_DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

constexpr int kThreads = 64 + 256;  // producer warp, consumer warp, 8 "epilogue" warps (LDGSTS producers in modes 3/4)
constexpr int kTile = 128 * 32 * 4;  // 16 KB: 128 rows x 32 floats

struct Args {
  const float* src[4];  // 4 operand arrays [rows][K]
  int rows, K, n_kb, mode;
  long long* cycles;
};

template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(kThreads, 1)
ingest_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
              const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3,
              const __grid_constant__ CUtensorMap mh, Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024 - (raw & 1023)) & 1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mode = a.mode;
  const bool cluster = mode == 5;
  const uint32_t rank = cluster ? cluster_rank() : 0;
  const int tile = cluster ? (blockIdx.x >> 1) : blockIdx.x;
  const int row0 = (tile * 128) % a.rows;
  const int brow0 = ((tile * 37 + (cluster ? rank : 0)) * 128) % a.rows;  // "weights" rows: another tile
  if (threadIdx.x == 0) {
    const uint32_t full_count = mode == 1 ? 2 : (mode == 3 ? 256 : (mode == 4 ? 257 : 1));
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], full_count);
      mbar_init(&empty[s], cluster ? 2 : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (cluster) cluster_sync();
  const long long t0 = clock64();
  if (warp == 0 || (mode == 1 && warp == 2)) {
    if (lane == 0) {
      for (int kb = 0; kb < a.n_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* st = smem + s * STAGE_BYTES;
        const int k0 = (kb * 32) % a.K;
        if (mode == 0) {
          mbar_expect_tx(&full[s], 4 * kTile);
          tma_2d(st, &m0, &full[s], k0, row0);
          tma_2d(st + kTile, &m1, &full[s], k0, row0);
          tma_2d(st + 2 * kTile, &m2, &full[s], k0, brow0);
          tma_2d(st + 3 * kTile, &m3, &full[s], k0, brow0);
        } else if (mode == 6) {
          mbar_expect_tx(&full[s], 2 * kTile);
          if (kb & 1) {
            tma_2d(st, &m0, &full[s], k0, row0);
            tma_2d(st + kTile, &m1, &full[s], k0, row0);
          } else {
            tma_2d(st, &m2, &full[s], k0, brow0);
            tma_2d(st + kTile, &m3, &full[s], k0, brow0);
          }
        } else if (mode == 1) {
          mbar_expect_tx(&full[s], 2 * kTile);
          if (warp == 0) {
            tma_2d(st, &m0, &full[s], k0, row0);
            tma_2d(st + kTile, &m1, &full[s], k0, row0);
          } else {
            tma_2d(st + 2 * kTile, &m2, &full[s], k0, brow0);
            tma_2d(st + 3 * kTile, &m3, &full[s], k0, brow0);
          }
        } else if (mode == 2) {
          mbar_expect_tx(&full[s], 4 * kTile);
          // pre-tiled layout: tile (row block, k block) is one contiguous 16 KB chunk
          const size_t kblocks = a.K / 32;
          const size_t ta = (static_cast<size_t>(row0 / 128) * kblocks + k0 / 32) * (kTile / 4);
          const size_t tb = (static_cast<size_t>(brow0 / 128) * kblocks + k0 / 32) * (kTile / 4);
          bulk_1d(st, a.src[0] + ta, kTile, &full[s]);
          bulk_1d(st + kTile, a.src[1] + ta, kTile, &full[s]);
          bulk_1d(st + 2 * kTile, a.src[2] + tb, kTile, &full[s]);
          bulk_1d(st + 3 * kTile, a.src[3] + tb, kTile, &full[s]);
        } else if (mode == 4) {
          mbar_expect_tx(&full[s], 2 * kTile);
          tma_2d(st, &m0, &full[s], k0, row0);
          tma_2d(st + kTile, &m1, &full[s], k0, row0);
        } else if (mode == 5) {
          mbar_expect_tx(&full[s], 4 * kTile);
          // A: this CTA issues rows [64 rank, 64 rank + 64) of the shared tile to both CTAs; B: private
          tma_2d_mc(st + rank * (kTile / 2), &mh, &full[s], k0, row0 + 64 * rank, 0x3);
          tma_2d_mc(st + kTile + rank * (kTile / 2), &mh, &full[s], k0, (row0 + 4096) % a.rows + 64 * rank, 0x3);
          tma_2d(st + 2 * kTile, &m2, &full[s], k0, brow0);
          tma_2d(st + 3 * kTile, &m3, &full[s], k0, brow0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < a.n_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (cluster) mbar_arrive_remote(&empty[s], rank ^ 1);
      }
    }
  } else if (mode == 3 || mode == 4) {
    // LDGSTS producers: 256 threads copy 2 (mode 4) or 4 (mode 3) tiles of 16 KB per stage, 16 B per request
    const int t = threadIdx.x - 64;
    const int n_tiles = mode == 3 ? 4 : 2;
    const int first = mode == 3 ? 0 : 2;
    for (int kb = 0; kb < a.n_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      unsigned char* st = smem + s * STAGE_BYTES;
      const int k0 = (kb * 32) % a.K;
      for (int i = 0; i < n_tiles; ++i) {
        const int arr = first + i;
        const int r0 = arr < 2 ? row0 : brow0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int idx = t + j * 256;  // 1024 chunks of 16 B: row = idx / 8, chunk = idx % 8
          const int row = idx >> 3, ch = idx & 7;
          const float* src = a.src[arr] + static_cast<size_t>(r0 + row) * a.K + k0 + ch * 4;
          cp_async16(st + arr * kTile + row * 128 + ((ch ^ (row & 7)) << 4), src);
        }
      }
      cp_async_arrive(&full[s]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) a.cycles[blockIdx.x] = t1 - t0;
  if (cluster) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn fn, float* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {K * sizeof(float)};
  const cuuint32_t box[2] = {32, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

template <int STAGES, int STAGE_BYTES>
static void run(int mode, int grid, const CUtensorMap* maps, const CUtensorMap& mh, Args a) {
  auto k = ingest_kernel<STAGES, STAGE_BYTES>;
  const int smem = STAGES * STAGE_BYTES + 1024 + 256;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  a.mode = mode;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = mode == 5 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k, maps[0], maps[1], maps[2], maps[3], mh, a));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), a.cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0, mx = 0;
  for (long long c : cyc) {
    mean += c;
    if (c > mx) mx = c;
  }
  mean /= grid;
  const double bytes = static_cast<double>(a.n_kb) * STAGE_BYTES;
  printf("mode %d grid %3d stages %d x %3d KB: %.1f us, per-CTA mean %.0f cyc (max %.0f) -> %.1f B/cyc/SM, %.2f TB/s aggregate\n",
         mode, grid, STAGES, STAGE_BYTES / 1024, ms * 1e3, mean, mx, bytes / mean, bytes * grid / (ms * 1e-3) / 1e12);
}

int main() {
  const int rows = 16384, K = 512;
  float* buf[4];
  for (int i = 0; i < 4; ++i) {
    CK(cudaMalloc(&buf[i], static_cast<size_t>(rows) * K * sizeof(float)));
    CK(cudaMemset(buf[i], 0, static_cast<size_t>(rows) * K * sizeof(float)));
  }
  long long* cycles;
  CK(cudaMalloc(&cycles, 1024 * sizeof(long long)));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  CUtensorMap maps[4];
  for (int i = 0; i < 4; ++i) maps[i] = make_map(fn, buf[i], rows, K, 128);
  CUtensorMap mh = make_map(fn, buf[0], rows, K, 64);
  Args a{};
  for (int i = 0; i < 4; ++i) a.src[i] = buf[i];
  a.rows = rows;
  a.K = K;
  a.n_kb = 256;  // 16 MB per CTA
  a.cycles = cycles;
  for (int grid : {64, 148}) {
    for (int mode : {0, 1, 2, 3, 4, 5}) run<3, 4 * kTile>(mode, grid, maps, mh, a);
    Args b = a;
    b.n_kb = 512;
    run<6, 2 * kTile>(6, grid, maps, mh, b);
  }
  return 0;
}
