// Micro-benchmarks behind test hooks (not on the product path): how fast can one SM pull
// DRAM-resident rows through TMA?  Used to separate "TMA / DRAM limit" from "pipeline limit" in
// the feed-bound conv layers.
#include "common.h"
#include "tc_gemm.cuh"

bool tc_make_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                    uint32_t box_rows);

namespace {

// grid = SMs; each CTA streams boxes of `box_rows` x 128 B (tile t covers rows [t*adv, t*adv+box_rows))
// through `stages` shared-memory slots; warp 0 lane 0 produces, warp 1 lane 0 consumes (waits, optionally
// reads one word, releases).  mode 1: plain 16-byte loads by all threads instead of TMA.
__global__ void __launch_bounds__(256, 1)
tma_stream_kernel(const __grid_constant__ CUtensorMap tm, const uint4* __restrict__ base, int box_rows, int adv,
                  int tiles, int stages, int mode, unsigned int* sink, uint4* wbuf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int box_bytes = box_rows * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * box_bytes);
  uint64_t* empty = full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mode == 1) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const uint4* src = base + (size_t)t * adv * 8;
      for (int i = threadIdx.x; i < box_rows * 8; i += 256) {
        const uint4 v = __ldcs(src + i);
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
      }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
    return;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      tc::mbar_wait(&empty[s], ph ^ 1u, nullptr);
      tc::mbar_expect_tx(&full[s], (uint32_t)box_bytes);
      int done = 0;
      while (done < box_rows) {   // boxes of at most 256 rows
        const int r = min(256, box_rows - done);
        (void)r;
        tc::tma_load_2d(smem + s * box_bytes + done * 128, &tm, &full[s], 0, t * adv + done);
        done += 256;
      }
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp >= 4 && mode >= 2) {
    // epilogue-like writers: 128 threads write 128 rows x 128 B per tile (as many bytes as one
    // 128-row tile reads).  mode 2: lane = row, 8 x 16-byte stores per row (the conv epilogue's
    // pattern: 32 requests of 16 B per instruction); mode 3: 8 lanes per row (coalesced 128 B).
    const int wt = threadIdx.x - 128;
    const uint4 v = make_uint4(wt, 1, 2, 3);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      uint4* dst = wbuf + (size_t)t * 128 * 8;
      if (mode == 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[wt * 8 + c] = v;
      } else if (mode == 3) {
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[c * 128 + wt] = v;
      } else {
        // mode 4 / 5: 2 / 4 adjacent lanes write one row's 32 / 64 contiguous bytes per instruction
        const int lanes = mode == 4 ? 2 : 4;
        const int sub = wt % lanes, rowg = wt / lanes;             // 128 / lanes row groups
        for (int half = 0; half < 8 / lanes; ++half)               // column groups of `lanes` pieces
#pragma unroll
          for (int pass = 0; pass < lanes; ++pass) {               // rows rowg + pass * (128 / lanes)
            const int row = rowg + pass * (128 / lanes);
            dst[row * 8 + half * lanes + sub] = v;
          }
      }
    }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0; unsigned int x = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      tc::mbar_wait(&full[s], ph, nullptr);
      x ^= *reinterpret_cast<volatile unsigned int*>(smem + s * box_bytes);
      tc::mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (x == 0x12345678u) sink[0] = 1;
  }
}


// 2-D tiled variant: the tensor is [rows][pitch8 u64]; tile (ty, tx) loads `nx` side-by-side boxes of
// bh x bw8 (overfetch = halo) and optionally writes its adv_h x adv_w8 region of a second tensor of the
// same geometry.  wmode 1: thread = 64-byte pixel, four 16-byte stores (the SCRFD epilogue's pattern);
// wmode 2: consecutive threads write consecutive 16-byte pieces of a row.
__global__ void __launch_bounds__(256, 1)
tma_tile_kernel(const __grid_constant__ CUtensorMap tm, int pitch8, int bw8, int bh, int nx, int adv_w8, int adv_h,
                int tiles_x, int tiles, int stages, int wmode, unsigned int* sink, uint4* wbuf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int box_bytes = bh * bw8 * 8;
  const int stage_bytes = (box_bytes * nx + 1023) / 1024 * 1024;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int ty = t / tiles_x, tx = t - ty * tiles_x;
      tc::mbar_wait(&empty[s], ph ^ 1u, nullptr);
      tc::mbar_expect_tx(&full[s], (uint32_t)(box_bytes * nx));
      for (int k = 0; k < nx; ++k)
        tc::tma_load_2d(smem + s * stage_bytes + k * box_bytes, &tm, &full[s], tx * adv_w8 + k * (adv_w8 / nx) - 4,
                        ty * adv_h - 1);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0; unsigned int x = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      tc::mbar_wait(&full[s], ph, nullptr);
      x ^= *reinterpret_cast<volatile unsigned int*>(smem + s * stage_bytes);
      tc::mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (x == 0x12345678u) sink[0] = 1;
  } else if (warp >= 4 && wmode) {
    const int wt = threadIdx.x - 128;
    const uint4 v = make_uint4(wt, 1, 2, 3);
    const size_t pitch16 = (size_t)pitch8 / 2;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int ty = t / tiles_x, tx = t - ty * tiles_x;
      uint4* dst = wbuf + (size_t)ty * adv_h * pitch16 + (size_t)tx * (adv_w8 / 2);
      if (wmode == 1) {
        const int ppr = adv_w8 / 8;   // 64-byte pixels per tile row
        for (int px = wt; px < adv_h * ppr; px += 128) {
          const int row = px / ppr, col = px - row * ppr;
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[row * pitch16 + col * 4 + c] = v;
        }
      } else {
        const int upr = adv_w8 / 2;   // 16-byte units per tile row
        for (int u = wt; u < adv_h * upr; u += 128) {
          const int row = u / upr, col = u - row * upr;
          dst[row * pitch16 + col] = v;
        }
      }
    }
  }
}

// Write-only patterns over a 1 GiB buffer, 512 contiguous bytes per thread (= the IResNet stem: 4 pixels
// x 64 bf16 channels).  mode 0: each lane writes its own 512 B as 32 x 16 B in (pass, pixel, half) order
// (32 lines touched per instruction); mode 1: 2 lanes share 32 B; mode 2: 8 lanes share a 128-byte line;
// mode 3: the warp writes its 16 KiB span fully coalesced.
__global__ void __launch_bounds__(128) write_pattern_kernel(uint4* __restrict__ buf, int mode) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint4 v = make_uint4((unsigned)gid, 1, 2, 3);
  uint4* wbase = buf + (gid - lane) * 32;      // the warp's 16 KiB span (1024 x 16 B)
  if (mode == 0) {
    uint4* o = buf + gid * 32;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass)
#pragma unroll
      for (int j = 0; j < 4; ++j) { o[j * 8 + pass * 2] = v; o[j * 8 + pass * 2 + 1] = v; }
  } else if (mode == 1) {
    // lane pair (2k, 2k+1): instruction (pass, j, which) writes the 32 B of pixel j of lane 2k+which
#pragma unroll
    for (int pass = 0; pass < 4; ++pass)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int which = 0; which < 2; ++which)
          wbase[((lane & ~1) + which) * 32 + j * 8 + pass * 2 + (lane & 1)] = v;
  } else if (mode == 2) {
    // 8-lane group g: instruction (r, j) writes the whole 128-byte pixel j of lane 8g + r
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) wbase[((lane & ~7) + r) * 32 + j * 8 + (lane & 7)] = v;
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) wbase[k * 32 + lane] = v;
  }
}

}  // namespace

extern "C" FR_API int fr_debug_tma_stream(fr_ctx* ctx, long long total_rows, int box_rows, int adv, int stages,
                                          int mode, int iters, float* ms_out, double* gbs_out) {
  if (!ctx || box_rows <= 0 || box_rows > 256 && box_rows % 256 != 0 || stages < 1 || stages > 16)
    return FR_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  void* buf = nullptr;
  void* wbuf = nullptr;
  unsigned int* sink = nullptr;
  if (cudaMalloc(&buf, (size_t)total_rows * 128) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess ||
      cudaMalloc(&wbuf, (size_t)total_rows * 128) != cudaSuccess)
    return fr_fail(ctx, FR_ERR_CUDA, "microbench allocation failed");
  cudaMemset(buf, 1, (size_t)total_rows * 128);
  CUtensorMap tm;
  if (!tc_make_map_2d(&tm, buf, (uint64_t)total_rows, 64, 64, (uint32_t)std::min(box_rows, 256)))
    return fr_fail(ctx, FR_ERR_CUDA, "tensor map failed");
  const int tiles = (int)((total_rows - box_rows) / adv);
  const int smem = stages * box_rows * 128 + 1024 + 256;
  cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  tma_stream_kernel<<<sms, 256, smem, ctx->stream>>>(tm, (const uint4*)buf, box_rows, adv, tiles, stages, mode, sink, (uint4*)wbuf);
  cudaEventRecord(e0, ctx->stream);
  for (int i = 0; i < iters; ++i)
    tma_stream_kernel<<<sms, 256, smem, ctx->stream>>>(tm, (const uint4*)buf, box_rows, adv, tiles, stages, mode, sink, (uint4*)wbuf);
  cudaEventRecord(e1, ctx->stream);
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink); cudaFree(wbuf);
  if (err != cudaSuccess) return fr_fail(ctx, FR_ERR_CUDA, cudaGetErrorString(err));
  if (ms_out) *ms_out = ms / iters;
  if (gbs_out) *gbs_out = (double)tiles * box_rows * 128 / (ms / iters * 1e-3) / 1e9;
  return FR_OK;
}

bool tc_make_map_2d_u64(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                        uint32_t box_rows);

extern "C" FR_API int fr_debug_tma_tiles(fr_ctx* ctx, int rows, int pitch8, int bw8, int bh, int nx, int adv_w8,
                                         int adv_h, int stages, int wmode, int iters, float* ms_out,
                                         double* gbs_out) {
  if (!ctx || bw8 > 256 || bh > 256 || stages < 1 || stages > 16 || pitch8 % adv_w8 || rows % adv_h || adv_w8 % nx)
    return FR_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  void* buf = nullptr;
  void* wbuf = nullptr;
  unsigned int* sink = nullptr;
  const size_t bytes = (size_t)rows * pitch8 * 8;
  if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess ||
      cudaMalloc(&wbuf, bytes) != cudaSuccess)
    return fr_fail(ctx, FR_ERR_CUDA, "microbench allocation failed");
  cudaMemset(buf, 1, bytes);
  CUtensorMap tm;
  if (!tc_make_map_2d_u64(&tm, buf, (uint64_t)pitch8, (uint64_t)rows, (uint32_t)bw8, (uint32_t)bh))
    return fr_fail(ctx, FR_ERR_CUDA, "tensor map failed");
  const int tiles_x = pitch8 / adv_w8, tiles = tiles_x * (rows / adv_h);
  const int stage_bytes = (bh * bw8 * 8 * nx + 1023) / 1024 * 1024;
  const int smem = stages * stage_bytes + 1024 + 256;
  if (smem > 227 * 1024) return FR_ERR_INVALID_ARG;
  cudaFuncSetAttribute(tma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < iters + 1; ++i) {
    if (i == 1) cudaEventRecord(e0, ctx->stream);
    tma_tile_kernel<<<sms, 256, smem, ctx->stream>>>(tm, pitch8, bw8, bh, nx, adv_w8, adv_h, tiles_x, tiles, stages,
                                                      wmode, sink, (uint4*)wbuf);
  }
  cudaEventRecord(e1, ctx->stream);
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink); cudaFree(wbuf);
  if (err != cudaSuccess) return fr_fail(ctx, FR_ERR_CUDA, cudaGetErrorString(err));
  if (ms_out) *ms_out = ms / iters;
  if (gbs_out) *gbs_out = (double)bytes / (ms / iters * 1e-3) / 1e9;   // unique bytes read (= written) per second
  return FR_OK;
}

extern "C" FR_API int fr_debug_write_pattern(fr_ctx* ctx, int mode, int iters, float* ms_out, double* gbs_out) {
  if (!ctx) return FR_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  const size_t bytes = (size_t)1 << 30;
  void* buf = nullptr;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) return fr_fail(ctx, FR_ERR_CUDA, "microbench allocation failed");
  const unsigned blocks = (unsigned)(bytes / 512 / 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < iters + 1; ++i) {
    if (i == 1) cudaEventRecord(e0, ctx->stream);
    write_pattern_kernel<<<blocks, 128, 0, ctx->stream>>>((uint4*)buf, mode);
  }
  cudaEventRecord(e1, ctx->stream);
  cudaError_t err = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  if (err != cudaSuccess) return fr_fail(ctx, FR_ERR_CUDA, cudaGetErrorString(err));
  if (ms_out) *ms_out = ms / iters;
  if (gbs_out) *gbs_out = (double)bytes / (ms / iters * 1e-3) / 1e9;
  return FR_OK;
}
