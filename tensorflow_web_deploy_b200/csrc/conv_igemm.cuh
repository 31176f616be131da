// Implicit-GEMM convolution + BatchNorm(scale/shift) + ReLU for sm_100a.
//
//   D[M = n*oh*ow, N = cout]  =  shift[N]  +  sum over (tap r,s ; channel block)  A_tap[M, 64] * (scale * W)[N, 64]^T
//
// Restates ops.conv2d (/root/reference/inception/slim/ops.py:229-245): tf.nn.conv2d (NHWC x HWIO,
// no bias) -> batch_norm inference branch (ops.py:122-128, folded to y = x*scale + shift) -> relu.
//
// * A tiles (128 output pixels x 64 input channels of ONE filter tap) are gathered by the TMA unit in
//   im2col mode straight from the NHWC activation tensor: no im2col buffer exists in HBM, SAME padding
//   and ragged channel counts are the unit's out-of-bounds zero fill.  1x1/stride-1 layers use the
//   plain 2-D tiled mode on the [pixels, channels] view.
// * W tiles ([n_tile x 64] of the packed K-major weight matrix, BatchNorm scale already multiplied in on the
//   host) come in by 2-D tiled TMA.
// * Both land in shared memory with the 128-byte swizzle and feed tcgen05.mma (kind::f16, M=128,
//   N=n_tile<=256, K=16) issued by one elected lane; the fp32 accumulator lives in TMEM.
// * The BatchNorm shift enters through the tensor core as well: every tile starts with ONE K=16 MMA of a
//   constant [128 x 16] tile whose first two K columns are 1 against a [n_tile x 16] tile holding
//   (hi, lo) 16-bit halves of shift — so the accumulator already is conv*scale + shift and the epilogue is
//   ReLU + convert only.  MEASURED (profiles/r01_epilogue.md): with per-column scale/shift/floor vectors read
//   from shared memory by every lane, each 128 x 64 chunk cost 768 cycles of shared-memory bandwidth (24
//   LDS.128 per thread), more than its TMEM read (512), and the shared-memory pipe is the same one the MMA
//   operands come through: every short-K layer was epilogue-bound.
// * Epilogue warps pull the accumulator with tcgen05.ld, ReLU, convert to bf16/fp16 and store into the channel
//   slice of the (possibly concatenated) NHWC output — tf.concat (inception_model.py:125...) never runs as
//   a kernel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace icp {

constexpr int kBM = 128;                    // output pixels per CTA tile (UMMA M)
constexpr int kBK = 64;                     // K elements per pipeline stage (one 128-B swizzle row)
constexpr int kABytes = kBM * kBK * 2;      // 16 KiB
constexpr int kConvThreads = 192;           // one-tile kernel: warps 0 TMA, 1 MMA, 2..5 epilogue
constexpr int kPersistThreads = 416;        // warps 0,11 A producers; 6,12 W producers; 1 MMA issuer; 2..5, 7..10 epilogue
constexpr int kEpiThreads = 256;            // two epilogue warps per TMEM lane quarter (one per half of a 64-column chunk)
constexpr uint32_t kOutStageBytes = 16384;  // epilogue staging
constexpr uint32_t kOnesBytes = 4096;       // [128 x 16] 16-bit "ones" A tile of the bias MMA (32-byte rows, SWIZZLE_32B)
constexpr int kMaxStages = 12;

// Debug timeline (tools/timeline.py): clock64 stamps of CTA 0 into a [rows][512] buffer.  Compiled in only with
// -DICP_TIMELINE_BUILD (tools/build_timeline.sh): the branches alone cost the MMA-issue loop ~1 % of a step.
#ifdef ICP_TIMELINE_BUILD
#define ICP_STAMP(buf, row, idx, pred)                                                        \
  do {                                                                                        \
    if ((buf) && blockIdx.x == 0 && (pred) && (idx) < 512) (buf)[(row) * 512 + (idx)] = clock64(); \
  } while (0)
#else
#define ICP_STAMP(buf, row, idx, pred) \
  do {                                 \
  } while (0)
#endif

struct ConvArgs {
  int m_total;     // valid output pixels (n * oh * ow)
  int oh, ow;      // output spatial size
  int stride;
  int pad_h, pad_w;
  int kw;          // filter width (S)
  int taps;        // R * S
  int cblocks;     // ceil(cin / ck), ck = row_bytes / 2 channels per K block
  int tail_ksteps; // K=16 MMAs issued for the LAST channel block of a tap: ceil((cin - (cblocks-1)*ck) / 16); the
                   // remaining slices of that block are zero padding on both operands
  int tail_rb;     // operand row bytes of that last block: 32 / 64 when <= 16 / 32 channels remain (its tiles then
                   // come through tm_a_tail / tm_b_tail with SWIZZLE_32B/64B boxes), else row_bytes
  int row_bytes;   // bytes per operand row in smem: 32 / 64 / 128 (SWIZZLE_32B/64B/128B)
  int n_tile;      // output channels per CTA, multiple of 16, <= 256
  int cout;        // output channels of the layer: the last N tile may be narrower than n_tile (704 = 256 + 256 + 192)
  int tmem_cols;   // power of two >= max(32, n_tile)
  int stages;
  int im2col;      // 1: A via im2col TMA, 0: A via 2-D tiled TMA
  int fmt;         // 0 fp16, 1 bf16
  void* out;       // NHWC output buffer (16-bit elements)
  int ldo;         // channels in the output buffer (concat width)
  int coff;        // first channel this layer owns in the output buffer
  // Horizontally fused sibling 1x1 convs (one GEMM over the concatenated output channels):
  // columns >= n_split go to a second buffer (out2/ldo2/coff2).
  void* out2;
  int ldo2;
  int coff2;
  int n_split;         // multiple of 16; >= cout when there is no second destination
  // ReLU applies to output columns [relu_lo, relu_hi) (multiples of 16); the rest stay linear
  // (pool-after-projection: BatchNorm shift + ReLU of those channels run after the pool).
  int relu_lo, relu_hi;
  int epi_staged;      // one-tile kernel: 1 = epilogue through shared staging + coalesced stores, 0 = row-per-thread stores
  int epi_teams;       // persistent kernel: 2 = two independent 4-warp epilogue teams (own staging each), 1 = one 8-warp team
  int pdl_late;        // 1: griddepcontrol.launch_dependents is issued when this CTA's LAST tile's MMAs are in flight (not at kernel start)
  int dual;            // persistent kernel (epi_teams == 2 only): 1 = a work item is TWO consecutive M tiles that share every
                       //   weight tile (stage = A0 | A1 | W); the two TMEM halves are simultaneous accumulators, not a double buffer
  long long* dbg;      // optional timeline buffer (CTA 0 only): [role][event] clock64 stamps, see tools/timeline.py
};

__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  if (fmt) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 16 fp32 accumulator values -> 8 packed 16-bit pairs, optional ReLU (packed max against +0 after the conversion:
// half the instructions of an fp32 max per element, same result).
__device__ __forceinline__ void convert16(const uint32_t* v, uint32_t* q, bool relu, int fmt) {
  if (fmt) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      if (relu) h = __hmax2(h, z);
      q[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  } else {
    const __half2 z = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      if (relu) h = __hmax2(h, z);
      q[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
}

// The constant A tile of the bias MMA: row r = (1, 1, 0, ..., 0) as 16 x 16-bit, 32-byte rows in the SWIZZLE_32B
// layout (16-byte chunk index XOR address bit 7, i.e. XOR bit 2 of the row).  Written by 128 threads (r = 0..127)
// with generic stores; the caller fences to the async proxy before the tensor core reads it.
__device__ __forceinline__ void write_ones_row(uint8_t* ones, int r, int fmt) {
  const uint32_t one = fmt ? 0x3F80u : 0x3C00u;
  const int sw = (r >> 2) & 1;
  *reinterpret_cast<uint4*>(ones + r * 32 + ((0 ^ sw) << 4)) = make_uint4(one | (one << 16), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(ones + r * 32 + ((1 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
}

// ---------------------------------------------------------------------------------------------------
// One tile per CTA (used for narrow layers, n_tile < 128, where two or three short-lived CTAs per SM overlap
// each other's prologue/epilogue better than one persistent CTA).
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue.  K blocks: block -1 is the bias
// block (only the [n_tile x 16] shift tile travels, into the stage's W slot), blocks 0..num_kb-1 the filter.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kConvThreads)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                  const __grid_constant__ CUtensorMap tm_a_tail, const __grid_constant__ CUtensorMap tm_b_tail,
                  const __grid_constant__ CUtensorMap tm_bias, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  const int ck = a.row_bytes / 2;
  const uint32_t a_bytes = static_cast<uint32_t>(kBM) * rb;
  const uint32_t b_bytes = static_cast<uint32_t>(a.n_tile) * rb;
  const uint32_t stage_tx = a_bytes + b_bytes;                       // bytes TMA delivers per stage
  const uint32_t stage_bytes = (stage_tx + 1023u) & ~1023u;          // stage pitch keeps every tile on a swizzle period
  const uint32_t ring_bytes = a.stages * stage_bytes < kOutStageBytes ? kOutStageBytes : a.stages * stage_bytes;
  const uint32_t ones_base = base + ring_bytes;
  const uint32_t bar_base = ones_base + kOnesBytes;
  const uint32_t tmem_full_bar = bar_base + 16u * a.stages;
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (a.stages + s); };

  const int m0 = blockIdx.x * kBM;
  const int n_base = blockIdx.y * a.n_tile;
  const int num_kb = a.taps * a.cblocks;
  const bool narrow_tail = a.tail_rb != a.row_bytes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
    ptx::prefetch_tmap(&tm_bias);
    ptx::prefetch_tmap(&tm_a_tail);
    ptx::prefetch_tmap(&tm_b_tail);
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    write_ones_row(gbase + (ones_base - base), static_cast<int>(threadIdx.x) - 64, a.fmt);
    ptx::fence_proxy_async_smem();
  }
  ptx::pdl_launch();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  ptx::pdl_wait();            // the predecessor's output (this layer's input) is complete from here on

  // Role loops run warp-converged; one elected lane issues the TMA / MMA / commit instructions (under a plain
  // `if (lane == 0)` the compiler wraps each of them in an ELECT + branch loop).
  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int bw = 0, bh = 0, bn = 0;
    if (a.im2col) {
      const int q = m0 % a.ow;
      const int t = m0 / a.ow;
      bw = q * a.stride - a.pad_w;
      bh = (t % a.oh) * a.stride - a.pad_h;
      bn = t / a.oh;
    }
    int s = 0;
    uint32_t ph = 0;                       // stage index / phase kept incrementally: no division in the loop
    uint32_t sa = base;
    const uint32_t tail_tx = static_cast<uint32_t>(kBM + a.n_tile) * a.tail_rb;
    const int cb_last = a.cblocks - 1;
    auto advance = [&]() {
      sa += stage_bytes;
      if (++s == a.stages) { s = 0; ph ^= 1u; sa = base; }
    };
    // bias block: only the shift tile travels (the first pass over the ring never waits)
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(full_bar(0), static_cast<uint32_t>(a.n_tile) * 32u);
      ptx::tma_load_2d(sa + a_bytes, &tm_bias, full_bar(0), 0, n_base);
    }
    __syncwarp();
    advance();
    int cb = 0, tr = 0, ts = 0, kcol = 0;
    for (int kb = 0; kb < num_kb; ++kb, kcol += ck) {
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      if (ptx::elect_one()) {
        const bool tail = narrow_tail && cb == cb_last;
        const CUtensorMap* ma = tail ? &tm_a_tail : &tm_a;
        ptx::mbar_expect_tx(full_bar(s), tail ? tail_tx : stage_tx);
        if (a.im2col) {
          ptx::tma_load_im2col_4d(sa, ma, full_bar(s), cb * ck, bw, bh, bn, static_cast<uint16_t>(ts),
                                  static_cast<uint16_t>(tr));
        } else {
          ptx::tma_load_2d(sa, ma, full_bar(s), cb * ck, m0);
        }
        ptx::tma_load_2d(sa + a_bytes, tail ? &tm_b_tail : &tm_b, full_bar(s), kcol, n_base);
      }
      __syncwarp();
      if (++cb == a.cblocks) {
        cb = 0;
        if (++ts == a.kw) { ts = 0; ++tr; }
      }
      advance();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.n_tile);
    const int ksteps = a.row_bytes / 32;
    const uint64_t hi_main = ptx::make_smem_desc_rows(0, rb);              // descriptor minus the address field
    const uint64_t hi_tail = ptx::make_smem_desc_rows(0, a.tail_rb);
    const uint64_t hi_32 = ptx::make_smem_desc_rows(0, 32);
    const uint32_t a_units = a_bytes >> 4;
    const uint32_t stage_units = stage_bytes >> 4;
    const uint32_t base_units = (base & 0x3FFFFu) >> 4;
    const int cb_last = a.cblocks - 1;
    const bool uniform = (a.tail_ksteps == ksteps) && !narrow_tail;        // every block is a full one
    int s = 0;
    uint32_t ph = 0;
    uint32_t su = base_units;                                              // stage start, 16-byte units
    auto advance = [&]() {
      su += stage_units;
      if (++s == a.stages) { s = 0; ph ^= 1u; su = base_units; }
    };
    // bias block: accumulator := shift (hi + lo halves) broadcast down the 128 rows
    ptx::mbar_wait(full_bar(0), 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      ptx::umma_f16(tmem_base, hi_32 | ((ones_base & 0x3FFFFu) >> 4), hi_32 | (su + a_units), idesc, 0u);
      ptx::umma_commit(empty_bar(0));
    }
    __syncwarp();
    advance();
    int cb = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        // +32 B per K=16 slice inside the swizzled row: +2 in the (addr>>4) field
        if (uniform || cb != cb_last) {
          const uint64_t da = hi_main | su, db = hi_main | (su + a_units);
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, da + 2u * k, db + 2u * k, idesc, 1u);
          } else {
            for (int k = 0; k < ksteps; ++k) ptx::umma_f16(tmem_base, da + 2u * k, db + 2u * k, idesc, 1u);
          }
        } else {
          const uint64_t da = hi_tail | su, db = hi_tail | (su + a_units);
          for (int k = 0; k < a.tail_ksteps; ++k) ptx::umma_f16(tmem_base, da + 2u * k, db + 2u * k, idesc, 1u);
        }
        ptx::umma_commit(empty_bar(s));      // smem slot free once these MMAs retire
        if (kb == num_kb - 1) ptx::umma_commit(tmem_full_bar);       // accumulator complete
      }
      __syncwarp();
      if (++cb == a.cblocks) cb = 0;
      advance();
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int et = threadIdx.x - 64;         // 0..127
    const int quarter = warp & 3;            // TMEM lanes this warp may touch: [32q, 32q+32)
    const int row = quarter * 32 + lane;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    if (!a.epi_staged) {
      // Row-per-thread stores straight from registers: no barriers, lowest latency for this one-tile-per-CTA kernel
      // (MEASURED on conv3, 73x73 1x1 64->80: 0.031 ms against 0.050 ms for the staged variant below).
      const int m = m0 + row;
      uint16_t* orow = reinterpret_cast<uint16_t*>(a.out) + static_cast<size_t>(m) * a.ldo + a.coff + n_base;
      uint16_t* orow2 = reinterpret_cast<uint16_t*>(a.out2) + static_cast<size_t>(m) * a.ldo2 + a.coff2 + n_base - a.n_split;
      const bool valid = m < a.m_total;
      for (int c0 = 0; c0 < a.n_tile; c0 += 64) {
        const int cp = (a.n_tile - c0 >= 64) ? 4 : (a.n_tile - c0) / 16;
        uint32_t v[4][16];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
          if (hh < cp) ptx::tmem_ld_32x32b_x16(taddr + c0 + hh * 16, v[hh]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          if (hh < cp) {
            const int col = n_base + c0 + hh * 16;
            uint32_t q[8];
            convert16(v[hh], q, col >= a.relu_lo && col < a.relu_hi, a.fmt);
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>((col < a.n_split ? orow : orow2) + c0 + hh * 16);
              dst[0] = make_uint4(q[0], q[1], q[2], q[3]);
              dst[1] = make_uint4(q[4], q[5], q[6], q[7]);
            }
          }
        }
      }
    } else {
      // TMEM -> registers -> ReLU -> 16-bit -> shared staging (a 128 x 64-column chunk; it aliases the pipeline
      // ring, idle once tmem_full has fired because this CTA owns a single tile) -> global with 8 consecutive lanes
      // per pixel, so each warp store covers whole 128-byte lines instead of 32 different ones.
      uint8_t* stage_g = gbase;                // start of the (now idle) ring, at least kOutStageBytes long
      const int cl = et & 7;                   // chunk lane of the copy-out: 8 lanes cover one pixel's 128 bytes
      for (int c0 = 0; c0 < a.n_tile; c0 += 64) {
        const int cp = (a.n_tile - c0 >= 64) ? 4 : (a.n_tile - c0) / 16;
        asm volatile("bar.sync 1, 128;" ::: "memory");                 // staging free
        uint32_t v[4][16];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh)
          if (hh < cp) ptx::tmem_ld_32x32b_x16(taddr + c0 + hh * 16, v[hh]);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          if (hh < cp) {
            const int col = n_base + c0 + hh * 16;
            uint32_t q[8];
            convert16(v[hh], q, col >= a.relu_lo && col < a.relu_hi, a.fmt);
            const int ch = hh * 2;
            *reinterpret_cast<uint4*>(stage_g + row * 128 + ((ch ^ (row & 7)) << 4)) = make_uint4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<uint4*>(stage_g + row * 128 + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(q[4], q[5], q[6], q[7]);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (cl < 2 * cp) {
          const int col = n_base + c0 + cl * 8;
          uint16_t* dst_base = (col < a.n_split)
                                   ? reinterpret_cast<uint16_t*>(a.out) + a.coff + col
                                   : reinterpret_cast<uint16_t*>(a.out2) + a.coff2 + col - a.n_split;
          const size_t ld = (col < a.n_split) ? a.ldo : a.ldo2;
          uint4 val[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (et >> 3) + 16 * i;
            val[i] = *reinterpret_cast<const uint4*>(stage_g + r * 128 + ((cl ^ (r & 7)) << 4));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = m0 + (et >> 3) + 16 * i;
            if (m < a.m_total) *reinterpret_cast<uint4*>(dst_base + static_cast<size_t>(m) * ld) = val[i];
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------
// Persistent, cross-tile pipelined version (the one the tower uses for n_tile >= 128).
//
// One CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (t -> m_tile = t / n_tiles,
// n_tile index = t % n_tiles, so CTAs running side by side share A and W tiles in L2).  The roles run their
// own tile loops and meet only through mbarriers, so
//   * the TMA producers run ahead across tile boundaries through the smem ring (loads for tile i+1 are in
//     flight while tile i is still multiplying) — this is what hides HBM/L2 latency for the short-K layers,
//   * the accumulator is double-buffered in TMEM (2 x tmem_cols columns): the epilogue of tile i
//     (tcgen05.ld -> ReLU/convert -> TMA stores) overlaps the MMAs of tile i+1,
//   * barrier init, TMEM allocation and descriptor prefetch are paid once per SM, not once per tile.
// Warp roles: 0, 11 = A producers; 6, 12 = W producers; 1 = TMEM owner + MMA issuer; 2..5 and 7..10 = epilogue.
// Every tile's K blocks: block -1 is the bias block (only the [n_tile x 16] shift tile travels, into the stage's
// W slot; the A producer just arrives), blocks 0..num_kb-1 the filter taps x channel blocks.
// ---------------------------------------------------------------------------------------------------
// kCS > 1: launched as clusters of kCS CTAs.  The CTAs own kCS consecutive M tiles of a group (same N tile, same K blocks);
// each loads 1/kCS of every weight tile and multicasts it into all the CTAs' shared memory, so the weight bytes crossing
// L2->SM — the larger half of the traffic of the tap-heavy layers, which are bound by exactly that (profiles/
// r01_epilogue.md) — shrink.  A stage is free only when ALL the CTAs' MMAs have released it (tcgen05.commit multicast onto
// every CTA's empty barrier), because the peers write into it.  kCS = 2 (default where used) and 8 (opt-in, measured
// slower: engine.cu make_conv) are instantiated.
template <int kCS>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_igemm_persistent_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                             const __grid_constant__ CUtensorMap tm_a_tail, const __grid_constant__ CUtensorMap tm_b_tail,
                             const __grid_constant__ CUtensorMap tm_b_half, const __grid_constant__ CUtensorMap tm_b_tail_half,
                             const __grid_constant__ CUtensorMap tm_bias,
                             const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_o2,
                             const __grid_constant__ CUtensorMap tm_o64, const __grid_constant__ CUtensorMap tm_o64_2,
                             const ConvArgs a, const int m_tiles, const int n_tiles) {
  constexpr bool kCluster = kCS > 1;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kCS) - 1u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  const int ck = a.row_bytes / 2;
  const uint32_t a_bytes = static_cast<uint32_t>(kBM) * rb;
  const uint32_t b_bytes = static_cast<uint32_t>(a.n_tile) * rb;
  const bool dual = !kCluster && a.dual != 0;
  const uint32_t mts = dual ? 2u : 1u;                         // M tiles per work item
  const uint32_t a_all = a_bytes * mts;                        // A bytes of a stage (W tile follows)
  const uint32_t stage_bytes = (a_all + b_bytes + 1023u) & ~1023u; // stage pitch keeps every tile on a swizzle period
  const uint32_t out_stage = base + a.stages * stage_bytes;    // 16 KiB epilogue staging
  const uint32_t ones_base = out_stage + static_cast<uint32_t>(a.epi_teams) * kOutStageBytes;  // 4 KiB constant A tile of the bias MMA
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (a.stages + s); };
  const uint32_t acc_bar = bar_base + 16u * a.stages;          // tmem_full[2], tmem_empty[2]
  auto tmem_full = [&](int i) { return acc_bar + 8u * i; };
  auto tmem_empty = [&](int i) { return acc_bar + 16u + 8u * i; };
  const uint32_t tmem_slot = acc_bar + 32u;

  // work item t -> (M tile or, in a cluster, pair of M tiles; N tile).  tile_of(t) = this CTA's M tile.
  uint32_t crank = 0;
  if constexpr (kCluster) crank = ptx::cluster_ctarank();
  const int grp = kCluster ? kCS : (dual ? 2 : 1);              // M tiles per work item (cluster: one per CTA)
  const int total_tiles = ((m_tiles + grp - 1) / grp) * n_tiles;
  const int num_kb = a.taps * a.cblocks;
  const bool narrow_tail = a.tail_rb != a.row_bytes;
  const int first = static_cast<int>(blockIdx.x) / kCS;
  const int step = static_cast<int>(gridDim.x) / kCS;
  auto tile_m0 = [&](int t) { return ((t / n_tiles) * grp + static_cast<int>(crank)) * kBM; };

  if (!a.pdl_late) ptx::pdl_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
    ptx::prefetch_tmap(&tm_bias);
    ptx::prefetch_tmap(&tm_a_tail);
    ptx::prefetch_tmap(&tm_b_tail);
    ptx::prefetch_tmap(&tm_b_half);
    ptx::prefetch_tmap(&tm_b_tail_half);
    ptx::prefetch_tmap(&tm_o);
    ptx::prefetch_tmap(&tm_o2);
    ptx::prefetch_tmap(&tm_o64);
    ptx::prefetch_tmap(&tm_o64_2);
    for (int s = 0; s < a.stages; ++s) {
      ptx::mbar_init(full_bar(s), 2);          // arrivals per stage: the A producer and the W producer
      ptx::mbar_init(empty_bar(s), kCS);   // cluster: every CTA's MMA issuer releases a stage
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), kEpiThreads);   // every epilogue thread arrives
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2 && warp <= 5) {
    write_ones_row(gbase + (ones_base - base), static_cast<int>(threadIdx.x) - 64, a.fmt);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCluster) ptx::cluster_sync();   // the peer's barriers exist before anything is signalled at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  // prologue above overlapped the predecessor's tail; its output is complete from here on.  (Letting the weight producers
  // skip this wait — weights do not depend on the predecessor — was MEASURED neutral: 2.274 vs 2.272 ms/step.)
  ptx::pdl_wait();

  if (warp == 0 || warp == 11) {
    // ---------------- TMA producers #1: A tiles (warp-converged; one elected lane issues) ----------------
    // MEASURED (timeline): one warp gets an im2col TMA out every ~476 cycles (a 2-D tiled one every ~353) however
    // deep the ring is, and that cadence — not bytes, not the MMAs (128 cycles per K block at n_tile = 64) — set
    // the ~518-cycle K-block period of every layer.  Two warps per operand take alternate K blocks.
    const int mine = warp == 0 ? 0 : 1;          // parity of the ring passes this warp serves
    int gpar = 0;                                // parity of the current block in the global block sequence
    uint32_t it = 0;
    int s = 0;
    uint32_t ph = 0;                       // ring position kept incrementally: no division in the K loop
    uint32_t sa = base;
    const uint32_t a_tail_bytes = static_cast<uint32_t>(kBM) * a.tail_rb;
    const int cb_last = a.cblocks - 1;
    auto advance = [&]() {
      sa += stage_bytes;
      if (++s == a.stages) { s = 0; ph ^= 1u; sa = base; }
    };
    for (int t = first; t < total_tiles; t += step) {
      const int m0 = tile_m0(t);
      int bw = 0, bh = 0, bn = 0, bw1 = 0, bh1 = 0, bn1 = 0;
      if (a.im2col) {
        const int q = m0 % a.ow;
        const int r = m0 / a.ow;
        bw = q * a.stride - a.pad_w;
        bh = (r % a.oh) * a.stride - a.pad_h;
        bn = r / a.oh;
        if (dual) {                               // second M tile of the item: base pixel m0 + 128
          const int q1 = (m0 + kBM) % a.ow;
          const int r1 = (m0 + kBM) / a.ow;
          bw1 = q1 * a.stride - a.pad_w;
          bh1 = (r1 % a.oh) * a.stride - a.pad_h;
          bn1 = r1 / a.oh;
        }
      }
      // bias block: the ones tile is resident, nothing to load
      if (gpar == mine) {
        ptx::mbar_wait(empty_bar(s), ph ^ 1u);
        if (ptx::elect_one()) ptx::mbar_arrive(full_bar(s));
        __syncwarp();
      }
      gpar ^= 1;
      advance();
      int cb = 0, tr = 0, ts = 0;
      for (int kb = 0; kb < num_kb; ++kb, ++it, gpar ^= 1) {
        if (gpar == mine) {
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          ICP_STAMP(a.dbg, 0, it, lane == 0);
          if (ptx::elect_one()) {
            const bool tail = narrow_tail && cb == cb_last;
            const CUtensorMap* ma = tail ? &tm_a_tail : &tm_a;
            ptx::mbar_expect_tx(full_bar(s), (tail ? a_tail_bytes : a_bytes) * mts);
            if (a.im2col) {
              ptx::tma_load_im2col_4d(sa, ma, full_bar(s), cb * ck, bw, bh, bn, static_cast<uint16_t>(ts),
                                      static_cast<uint16_t>(tr));
              if (dual)
                ptx::tma_load_im2col_4d(sa + a_bytes, ma, full_bar(s), cb * ck, bw1, bh1, bn1, static_cast<uint16_t>(ts),
                                        static_cast<uint16_t>(tr));
            } else {
              ptx::tma_load_2d(sa, ma, full_bar(s), cb * ck, m0);
              if (dual) ptx::tma_load_2d(sa + a_bytes, ma, full_bar(s), cb * ck, m0 + kBM);
            }
          }
          __syncwarp();
        }
        if (++cb == a.cblocks) {
          cb = 0;
          if (++ts == a.kw) { ts = 0; ++tr; }
        }
        advance();
      }
    }
  } else if (warp == 6 || warp == 12) {
    // ---------------- TMA producers #2: weight tiles (alternate K blocks, like the A producers) ----------------
    const int mine = warp == 6 ? 0 : 1;
    int gpar = 0;
    uint32_t it = 0;
    int s = 0;
    uint32_t ph = 0;
    uint32_t sb = base + a_all;
    const uint32_t b_tail_bytes = static_cast<uint32_t>(a.n_tile) * a.tail_rb;
    const int cb_last = a.cblocks - 1;
    auto advance = [&]() {
      sb += stage_bytes;
      if (++s == a.stages) { s = 0; ph ^= 1u; sb = base + a_all; }
    };
    for (int t = first; t < total_tiles; t += step) {
      const int n_base = (t % n_tiles) * a.n_tile;
      // bias block: the [n_tile x 16] shift tile
      if (gpar == mine) {
        ptx::mbar_wait(empty_bar(s), ph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full_bar(s), static_cast<uint32_t>(a.n_tile) * 32u);
          ptx::tma_load_2d(sb, &tm_bias, full_bar(s), 0, n_base);
        }
        __syncwarp();
      }
      gpar ^= 1;
      advance();
      int cb = 0, kcol = 0;
      for (int kb = 0; kb < num_kb; ++kb, ++it, kcol += ck, gpar ^= 1) {
        if (gpar == mine) {
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          ICP_STAMP(a.dbg, 1, it, lane == 0);
          if (ptx::elect_one()) {
            const bool tail = narrow_tail && cb == cb_last;
            ptx::mbar_expect_tx(full_bar(s), tail ? b_tail_bytes : b_bytes);
            if constexpr (kCluster) {
              // this CTA's share of the weight rows, delivered to every CTA (each barrier receives all shares' bytes)
              const uint32_t part_rows = static_cast<uint32_t>(a.n_tile) / kCS;
              ptx::tma_load_2d_mcast(sb + crank * part_rows * (tail ? static_cast<uint32_t>(a.tail_rb) : rb),
                                     tail ? &tm_b_tail_half : &tm_b_half, full_bar(s), kcol,
                                     n_base + static_cast<int>(crank * part_rows), kMask);
            } else {
              ptx::tma_load_2d(sb, tail ? &tm_b_tail : &tm_b, full_bar(s), kcol, n_base);
            }
          }
          __syncwarp();
        }
        if (++cb == a.cblocks) cb = 0;
        advance();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: warp-converged loop, one elected lane issues ----------------
    // The loop body is on the critical path of every K block (MEASURED: a dozen extra integer instructions here cost
    // 1-2 % of a whole step), so descriptors are built from hoisted halves and the stage address is kept
    // incrementally.
    const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.n_tile);
    const int ksteps = a.row_bytes / 32;
    const uint64_t hi_main = ptx::make_smem_desc_rows(0, rb);              // descriptor minus the address field
    const uint64_t hi_tail = ptx::make_smem_desc_rows(0, a.tail_rb);
    const uint64_t hi_32 = ptx::make_smem_desc_rows(0, 32);
    const uint64_t d_ones = hi_32 | ((ones_base & 0x3FFFFu) >> 4);
    const uint32_t a_units = a_all >> 4;                                    // W tile offset inside a stage
    const uint32_t a1_units = a_bytes >> 4;                                 // dual: second A tile offset
    const uint32_t stage_units = stage_bytes >> 4;
    const uint32_t base_units = (base & 0x3FFFFu) >> 4;
    const int cb_last = a.cblocks - 1;
    const bool uniform = (a.tail_ksteps == ksteps) && !narrow_tail;        // every block is a full one
    uint32_t it = 0, lt = 0;
    int s = 0;
    uint32_t ph = 0;
    uint32_t su = base_units;                                              // stage start, 16-byte units
    auto commit_empty = [&](int s_) {
      if constexpr (kCluster) ptx::umma_commit_mcast(empty_bar(s_), kMask);
      else ptx::umma_commit(empty_bar(s_));
    };
    auto advance = [&]() {
      su += stage_units;
      if (++s == a.stages) { s = 0; ph ^= 1u; su = base_units; }
    };
    for (int t = first; t < total_tiles; t += step, ++lt) {
      // ICP_PDL_POINT=2: trigger BEFORE the last tile's MMAs.  MEASURED 2.278 vs 2.271 ms/step at batch 128, 0.472 vs 0.468 at 8:
      // the successor's CTAs arrive too early again; the default point is after the last MMAs (below).
      if (a.pdl_late == 2 && t + step >= total_tiles) ptx::pdl_launch();
      // dual: both TMEM halves belong to this item (barrier pair 0, one phase per item); else they alternate per tile
      const uint32_t as = dual ? 0u : (lt & 1u);
      ptx::mbar_wait(tmem_empty(as), ((dual ? lt : (lt >> 1)) & 1u) ^ 1u);   // epilogue has drained this accumulator
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + as * a.tmem_cols;
      const uint32_t tacc1 = tmem_base + a.tmem_cols;          // dual: accumulator of the item's second M tile
      // bias block: accumulator := shift (hi + lo halves) broadcast down the 128 rows
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        ptx::umma_f16(tacc, d_ones, hi_32 | (su + a_units), idesc, 0u);
        if (dual) ptx::umma_f16(tacc1, d_ones, hi_32 | (su + a_units), idesc, 0u);
        commit_empty(s);
      }
      __syncwarp();
      advance();
      int cb = 0;
      auto issue_block = [&](uint32_t su_, int cb_) {       // the K=16 MMAs of one K block (elected lane only)
        if (uniform || cb_ != cb_last) {
          const uint64_t da = hi_main | su_, db = hi_main | (su_ + a_units);
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16(tacc, da + 2u * k, db + 2u * k, idesc, 1u);
            if (dual) {
              const uint64_t da1 = hi_main | (su_ + a1_units);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16(tacc1, da1 + 2u * k, db + 2u * k, idesc, 1u);
            }
          } else {
            for (int k = 0; k < ksteps; ++k) ptx::umma_f16(tacc, da + 2u * k, db + 2u * k, idesc, 1u);
            if (dual) {
              const uint64_t da1 = hi_main | (su_ + a1_units);
              for (int k = 0; k < ksteps; ++k) ptx::umma_f16(tacc1, da1 + 2u * k, db + 2u * k, idesc, 1u);
            }
          }
        } else {
          const uint64_t da = hi_tail | su_, db = hi_tail | (su_ + a_units);
          for (int k = 0; k < a.tail_ksteps; ++k) ptx::umma_f16(tacc, da + 2u * k, db + 2u * k, idesc, 1u);
          if (dual) {
            // narrow tail tiles: the second A tile sits at the FULL tile pitch (a_bytes), rows of tail_rb bytes
            const uint64_t da1 = hi_tail | (su_ + a1_units);
            for (int k = 0; k < a.tail_ksteps; ++k) ptx::umma_f16(tacc1, da1 + 2u * k, db + 2u * k, idesc, 1u);
          }
        }
      };
      int kb = 0;
      // two K blocks per trip of the issue loop: one elect / branch round for 8 MMAs and 2 commits (MEASURED: +10 % on
      // the n_tile = 64 layers, whose K-block period is set by this loop rather than by the MMAs; neutral for wider tiles)
      for (; kb + 1 < num_kb; kb += 2, it += 2) {
        const int s0 = s;
        const uint32_t su0 = su, ph0 = ph;
        const int cb0 = cb;
        if (++cb == a.cblocks) cb = 0;
        advance();
        ptx::mbar_wait(full_bar(s0), ph0);
        ptx::mbar_wait(full_bar(s), ph);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_block(su0, cb0);
          commit_empty(s0);
          issue_block(su, cb);
          commit_empty(s);
        }
        __syncwarp();
        if (++cb == a.cblocks) cb = 0;
        advance();
      }
      for (; kb < num_kb; ++kb, ++it) {
        ptx::mbar_wait(full_bar(s), ph);
        ptx::tc_fence_after();
        ICP_STAMP(a.dbg, 2, it, lane == 0);
        if (ptx::elect_one()) {
          issue_block(su, cb);
          commit_empty(s);
        }
        __syncwarp();
        ICP_STAMP(a.dbg, 3, it, lane == 0);
        if (++cb == a.cblocks) cb = 0;
        advance();
      }
      if (ptx::elect_one()) ptx::umma_commit(tmem_full(as));
      __syncwarp();
    }
    // late trigger: this CTA has issued all its MMAs; what is left is the last tile's epilogue.  The successor grid may now
    // be scheduled (once every CTA got here or exited): its launch latency and prologue overlap these final epilogues
    // without its CTAs having occupied SMs during the main loop (the cost of triggering at kernel start, engine.cu).
    if (a.pdl_late == 1) ptx::pdl_launch();
  } else {
    if (a.epi_teams == 2) {
      // ---------------- epilogue (warps 2..5 and 7..10) ----------------
      // TWO independent teams of four warps (one warp per TMEM lane quarter each): the 64-column chunks of the tile
      // stream are dealt to the teams alternately (running chunk counter, so single-chunk tiles alternate too), each team
      // with its own 16 KiB staging buffer, its own named barrier and its own TMA-store thread.  MEASURED before (one
      // 8-warp team, r01_epilogue.md): ~1350 cycles per 128 x 64 chunk as ONE dependent chain (TMEM wait -> convert ->
      // staging free -> STS -> proxy fence -> barrier -> TMA store issue by one thread -> its read of the staging), so
      // every short-K layer (35x35 1x1s, conv3: <= 2500 cycles of loads + MMAs per tile against 2700-5400 of epilogue)
      // was epilogue-bound; two chains in flight halve the period without adding work.
      const int team = warp >= 7 ? 1 : 0;
      const int tt = team ? static_cast<int>(threadIdx.x) - 224 : static_cast<int>(threadIdx.x) - 64;   // 0..127 in the team
      const int et = tt + 128 * team;
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t my_stage = out_stage + static_cast<uint32_t>(team) * kOutStageBytes;
      const uint32_t bar_id = 1u + static_cast<uint32_t>(team);
      uint32_t lt = 0, gchunk = 0;
      for (int t = first; t < total_tiles; t += step, ++lt) {
        const int m0 = tile_m0(t);
        const int n_base = (t % n_tiles) * a.n_tile;
        // columns of this tile that exist: the weight / shift boxes of a narrower last tile are zero-filled past Cout
        // (the MMAs run at full width), the epilogue simply stops at `cols`
        const int cols = a.cout - n_base < a.n_tile ? a.cout - n_base : a.n_tile;
        const int nchunks = (cols + 63) >> 6;
        // dual: the item's two M tiles sit in the two TMEM halves (the second may lie past the last tile: skipped)
        const int subs = dual ? ((m0 / kBM + 1 < m_tiles) ? 2 : 1) : 1;
        const int tchunks = nchunks * subs;
        const uint32_t as = dual ? 0u : (lt & 1u);
        ICP_STAMP(a.dbg, 6, lt, et == 0);
        ptx::mbar_wait(tmem_full(as), (dual ? lt : (lt >> 1)) & 1u);
        ptx::tc_fence_after();
        ICP_STAMP(a.dbg, 4, lt, et == 0);
        const uint32_t taddr0 = tmem_base + as * a.tmem_cols + (static_cast<uint32_t>(quarter * 32) << 16);
        // my chunks of this item: cc with (gchunk + cc) odd/even = team; the last of them releases the accumulator(s)
        int last_mine = -1;
        for (int cc = tchunks - 1; cc >= 0; --cc)
          if (((gchunk + static_cast<uint32_t>(cc)) & 1u) == static_cast<uint32_t>(team)) { last_mine = cc; break; }
        if (last_mine < 0) {                         // nothing of this item is mine: hand the accumulator back at once
          ptx::tc_fence_before();
          ptx::mbar_arrive(tmem_empty(as));
        }
        for (int cc = 0; cc < tchunks; ++cc) {
          if (((gchunk + static_cast<uint32_t>(cc)) & 1u) != static_cast<uint32_t>(team)) continue;
          const int sub = cc >= nchunks ? 1 : 0;
          const int c = cc - sub * nchunks;
          const int ms = m0 + sub * kBM;
          const uint32_t taddr = taddr0 + static_cast<uint32_t>(sub) * a.tmem_cols;
          const int c0 = c << 6;
          int cp = (cols - c0) >> 4;                 // 16-column pieces in this chunk (<= 4)
          if (cp > 4) cp = 4;
          uint32_t v[64];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < cp) ptx::tmem_ld_32x32b_x16(taddr + c0 + 16 * j, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * j]));
          ptx::tmem_ld_wait();
          ICP_STAMP(a.dbg, 13, lt * 4 + c, et == 0);
          if (cc == last_mine) {                     // last TMEM read of this item (for this team) has landed
            ptx::tc_fence_before();
            ptx::mbar_arrive(tmem_empty(as));
          }
          const int col0 = n_base + c0;
          // A full 64-column chunk that lies in one destination goes out as ONE [128 rows x 128 B] box (128-byte
          // swizzled staging: conflict-free 16-byte writes); tails and chunks straddling n_split use 16-column boxes.
          const bool wide = (cp == 4) && (col0 >= a.n_split || col0 + 64 <= a.n_split);
          // staging buffer is free once this team's previous store has read it
          if (tt == 0) ptx::tma_store_wait_read();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          ICP_STAMP(a.dbg, 8, lt * 4 + c, et == 0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < cp) {
              const int col = col0 + 16 * j;
              uint32_t pk[8];
              convert16(&v[16 * j], pk, col >= a.relu_lo && col < a.relu_hi, a.fmt);
              if (wide) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                  const uint32_t dst = my_stage + static_cast<uint32_t>(row) * 128u +
                                       (static_cast<uint32_t>((2 * j + q) ^ (row & 7)) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                               "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                }
              } else {
                const uint32_t dst = my_stage + static_cast<uint32_t>(j) * 4096u + static_cast<uint32_t>(row) * 32u;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                             : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]),
                             "r"(pk[7]) : "memory");
              }
            }
          }
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (tt == 0) {
            if (wide) {
              if (col0 < a.n_split) ptx::tma_store_2d(&tm_o64, my_stage, a.coff + col0, ms);
              else ptx::tma_store_2d(&tm_o64_2, my_stage, a.coff2 + col0 - a.n_split, ms);
            } else {
              for (int j = 0; j < cp; ++j) {
                const int col = col0 + 16 * j;
                if (col < a.n_split) ptx::tma_store_2d(&tm_o, my_stage + j * 4096u, a.coff + col, ms);
                else ptx::tma_store_2d(&tm_o2, my_stage + j * 4096u, a.coff2 + col - a.n_split, ms);
              }
            }
            ptx::tma_store_commit();
            ICP_STAMP(a.dbg, 9, lt * 4 + c, true);
          }
        }
        gchunk += static_cast<uint32_t>(tchunks);
        ICP_STAMP(a.dbg, 5, lt, et == 0);
      }
      if (tt == 0) ptx::tma_store_wait_all();      // smem must outlive the last store's read
    } else {
      // ---------------- epilogue (warps 2..5 and 7..10) ----------------
      // Each TMEM lane quarter has two warps (w and w+5 share w & 3), each owning half of the 16-column pieces of a
      // 64-column chunk; the next chunk's TMEM loads are in flight while this one is converted and staged.
      const int grp = warp >= 7 ? 1 : 0;                 // column half of each 64-column chunk
      const int et = grp ? static_cast<int>(threadIdx.x) - 224 + 128 : static_cast<int>(threadIdx.x) - 64;   // 0..255
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      uint32_t lt = 0;
      for (int t = first; t < total_tiles; t += step, ++lt) {
        const int m0 = tile_m0(t);
        const int n_base = (t % n_tiles) * a.n_tile;
        // columns of this tile that exist: the weight / shift boxes of a narrower last tile are zero-filled past Cout
        // (the MMAs run at full width), the epilogue simply stops at `cols`
        const int cols = a.cout - n_base < a.n_tile ? a.cout - n_base : a.n_tile;
        const uint32_t as = lt & 1u;
        ICP_STAMP(a.dbg, 6, lt, et == 0);
        ptx::mbar_wait(tmem_full(as), (lt >> 1) & 1u);
        ptx::tc_fence_after();
        ICP_STAMP(a.dbg, 4, lt, et == 0);
        const uint32_t taddr = tmem_base + as * a.tmem_cols + (static_cast<uint32_t>(quarter * 32) << 16);

        // 16-column pieces of the chunk at c0 owned by this thread: [fp, fp + np) of the chunk's cp (<= 4) pieces
        auto geom = [&](int c0, int& cp, int& fp, int& np) {
          cp = (cols - c0) >> 4;
          if (cp > 4) cp = 4;
          const int h0 = (cp + 1) >> 1;
          fp = grp ? h0 : 0;
          np = grp ? cp - h0 : h0;
        };
        auto issue = [&](uint32_t (&v)[32], int c0) {
          int cp, fp, np;
          geom(c0, cp, fp, np);
          if (np > 0) ptx::tmem_ld_32x32b_x16(taddr + c0 + 16 * fp, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          if (np > 1) ptx::tmem_ld_32x32b_x16(taddr + c0 + 16 * fp + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        };
        // The accumulator leaves through shared memory and TMA stores: each thread converts its <=32 columns of one
        // output pixel (TMEM lane), writes them into the 16 KiB staging buffer, and one thread stores the chunk into
        // the destination slice(s).  Rows past the last valid pixel land in the unused tail of the buffer or are
        // clipped by the tensor map.
        auto process = [&](uint32_t (&v)[32], uint32_t (&vn)[32], int c0) {
          int chunk_pieces, fp, np;
          geom(c0, chunk_pieces, fp, np);
          if (c0 + 64 < cols) issue(vn, c0 + 64);
          uint32_t pk[16];
  #pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < np) {
              const int col = n_base + c0 + 16 * (fp + j);
              convert16(&v[16 * j], &pk[8 * j], col >= a.relu_lo && col < a.relu_hi, a.fmt);
            }
          }
          // staging buffer is free once the previous chunk's stores have read it
          ICP_STAMP(a.dbg, 7, lt * 4 + (c0 >> 6), et == 0);
          if (et == 0) ptx::tma_store_wait_read();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          ICP_STAMP(a.dbg, 8, lt * 4 + (c0 >> 6), et == 0);
          // A full 64-column chunk that lies in one destination goes out as ONE [128 rows x 128 B] box (128-byte
          // swizzled staging: conflict-free 16-byte writes, and a quarter of the box rows the TMA unit has to walk
          // compared with four 32-byte-row boxes); tails and chunks straddling n_split use 16-column boxes.
          const int col0 = n_base + c0;
          const bool wide = (chunk_pieces == 4) && (col0 >= a.n_split || col0 + 64 <= a.n_split);
          if (wide) {
  #pragma unroll
            for (int q = 0; q < 4; ++q) {            // 16-byte chunk 4*grp+q of this thread's 128-byte row
              const uint32_t dst = out_stage + static_cast<uint32_t>(row) * 128u +
                                   (static_cast<uint32_t>((4 * grp + q) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                           "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
            }
          } else {
  #pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < np) {
                const uint32_t dst = out_stage + (fp + j) * 4096u + static_cast<uint32_t>(row) * 32u;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[8 * j]), "r"(pk[8 * j + 1]),
                             "r"(pk[8 * j + 2]), "r"(pk[8 * j + 3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u), "r"(pk[8 * j + 4]),
                             "r"(pk[8 * j + 5]), "r"(pk[8 * j + 6]), "r"(pk[8 * j + 7]) : "memory");
              }
            }
          }
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
            if (wide) {
              if (col0 < a.n_split) ptx::tma_store_2d(&tm_o64, out_stage, a.coff + col0, m0);
              else ptx::tma_store_2d(&tm_o64_2, out_stage, a.coff2 + col0 - a.n_split, m0);
            } else {
              for (int j = 0; j < chunk_pieces; ++j) {
                const int col = col0 + 16 * j;
                if (col < a.n_split) ptx::tma_store_2d(&tm_o, out_stage + j * 4096u, a.coff + col, m0);
                else ptx::tma_store_2d(&tm_o2, out_stage + j * 4096u, a.coff2 + col - a.n_split, m0);
              }
            }
            ptx::tma_store_commit();
            ICP_STAMP(a.dbg, 9, lt * 4 + (c0 >> 6), true);
          }
        };
        auto release = [&]() {                       // last TMEM read of this tile has landed: hand the accumulator back
          ptx::tc_fence_before();
          ptx::mbar_arrive(tmem_empty(as));
        };

        uint32_t va[32], vb[32];
        int c0 = 0;
        issue(va, 0);
        while (true) {
          ptx::tmem_ld_wait();
          ICP_STAMP(a.dbg, 13, lt * 4 + (c0 >> 6), et == 0);
          if (c0 + 64 >= cols) release();
          process(va, vb, c0);
          c0 += 64;
          if (c0 >= cols) break;
          ptx::tmem_ld_wait();
          ICP_STAMP(a.dbg, 13, lt * 4 + (c0 >> 6), et == 0);
          if (c0 + 64 >= cols) release();
          process(vb, va, c0);
          c0 += 64;
          if (c0 >= cols) break;
        }
        ICP_STAMP(a.dbg, 5, lt, et == 0);
      }
      if (et == 0) ptx::tma_store_wait_all();      // smem must outlive the last store's read
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (kCluster) ptx::cluster_sync();   // no CTA leaves while its peer can still write its smem / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 2 * a.tmem_cols);
  }
}

inline size_t conv_stage_bytes(int n_tile, int row_bytes) {
  return (static_cast<size_t>(kBM + n_tile) * row_bytes + 1023) & ~size_t(1023);
}
inline size_t conv_smem_bytes(int stages, int n_tile, int row_bytes, bool persistent = false, int teams = 1) {
  size_t ring = static_cast<size_t>(stages) * conv_stage_bytes(n_tile, row_bytes);
  if (!persistent && ring < kOutStageBytes) ring = kOutStageBytes;   // the one-tile kernel stages its epilogue in the ring
  return 1024 + ring + (persistent ? teams * kOutStageBytes : 0) + kOnesBytes + 256 + 1024;
}

}  // namespace icp
