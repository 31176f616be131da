// conv0 (3x3 / stride 2 / VALID, Cin = 3 -> 32, inception_model.py:88) on the tensor core.
//
// K = 3*3*3 = 27 is padded to 32: one 128-pixel tile is a single [128 x 32] x [32 x 32]^T product = two K=16
// tcgen05 MMAs.  TMA cannot gather this operand (3-channel pixels are 6 bytes; im2col boxes need 16-byte
// channel vectors), so 128 builder threads assemble it: each reads the 27 inputs of one output pixel (three
// rows of 9 contiguous values), converts them to 16 bits and writes one 64-byte row of the A tile in the
// SWIZZLE_64B layout the UMMA descriptor expects (16-byte chunk index XOR address bits 7..8, i.e. XOR
// ((row >> 1) & 3)), then fences to the async proxy and arrives on the stage's mbarrier.
//
// * uint8 input: the raw integers 0..255 are exact in fp16/bf16, and the normalisation x = u8 * 2/255 - 1
//   (PyClassification.py:152-155) is folded into the operands on the host: W' = round16(W * scale * 2/255),
//   shift' = shift - 127.5 * sum(W'), i.e. the kernel computes exactly sum W'(u8 - 127.5) + shift.  fp32 input
//   (already normalised) is rounded to 16 bits like every other activation of the tower.
// * The [32 x 32] weight tile (BatchNorm scale folded in) is written into shared memory once per CTA; the
//   shift is added in the epilogue from the constant bank (32 channels = 32 registers per thread).
// * Persistent: one CTA per SM walks the tiles; 8 A stages, TMEM accumulator double-buffered (2 x 32 columns).
// Warp roles: 0..15 = A builders (group g = warp / 4 fills stage g for every 4th tile), 16 = TMEM owner + MMA issuer,
// 17..20 = epilogue (TMEM lane quarter = warp & 3).  MEASURED: with one builder group (128 threads) the kernel took
// 0.184 ms at batch 128 — latency-bound on the 27 loads per pixel; four groups of 128 threads: 0.097 ms.
//
// MEASURED: the FP32-pipe version (tower_kernels.cuh conv0_kernel, 864 FMAs per pixel) took 0.125 ms at batch 128,
// 39 of ~70 TFLOP/s of FP32 issue; this one is bound by the 27 scalar loads per pixel and the 182 MB it writes.
#pragma once
#include "conv_igemm.cuh"
#include "tower_kernels.cuh"

namespace icp {

constexpr int kConv0TcThreads = 672;      // 16 builder warps (4 groups of 128) + MMA warp + 4 epilogue warps
constexpr int kConv0Stages = 8;

struct Conv0TcParams {
  uint16_t w[32 * 32];   // [cout][k] 16-bit, k = (r*3+s)*3+ci for k < 27, zero above; scale (and 2/255 for u8) folded in
  float shift[32];
};

template <typename TIn>
__device__ __forceinline__ uint16_t conv0_tc_cvt(TIn v, int fmt);
template <>
__device__ __forceinline__ uint16_t conv0_tc_cvt<float>(float v, int fmt) {
  return f2h(v, fmt);
}
template <>
__device__ __forceinline__ uint16_t conv0_tc_cvt<uint8_t>(uint8_t v, int fmt) {
  return f2h(static_cast<float>(v), fmt);        // 0..255: exact in both 16-bit formats
}

template <typename TIn>
__global__ void __launch_bounds__(kConv0TcThreads, 1)
conv0_tc_kernel(const TIn* __restrict__ in, int N, int H, int W, const __grid_constant__ Conv0TcParams p,
                uint16_t* __restrict__ out, int OH, int OW, int fmt) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  constexpr uint32_t kABytes0 = 128 * 64;                       // one A stage: 128 rows x 64 B
  const uint32_t w_base = base + kConv0Stages * kABytes0;       // [32 x 64 B] weight tile
  const uint32_t bar_base = w_base + 2048;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kConv0Stages + s); };
  auto tmem_full = [&](int i) { return bar_base + 16u * kConv0Stages + 8u * i; };
  auto tmem_empty = [&](int i) { return bar_base + 16u * kConv0Stages + 16u + 8u * i; };
  const uint32_t tmem_slot = bar_base + 16u * kConv0Stages + 32u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = N * OH * OW;
  const int tiles = (total + kBM - 1) / kBM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kConv0Stages; ++s) {
      ptx::mbar_init(full_bar(s), 128);          // every builder thread arrives
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 16) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < 128) {
    // weight tile: row n = 32 x 16-bit of output channel n, four 16-byte chunks, SWIZZLE_64B
    const int n = threadIdx.x >> 2, c = threadIdx.x & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(&p.w[n * 32 + c * 8]);
    *reinterpret_cast<uint4*>(gbase + (w_base - base) + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = v;
    ptx::fence_proxy_async_smem();
  }
  ptx::pdl_launch();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  ptx::pdl_wait();

  if (warp < 16) {
    // ---------------- A builders: one output pixel (one 64-byte operand row) per thread and tile ----------------
    // Group g fills stages 2g and 2g+1, i.e. local tiles 2g, 2g+1 (mod 8); both tiles' 54 loads are issued before
    // either row is packed, which doubles the loads in flight per thread.
    const int grp = warp >> 2;
    const int row = threadIdx.x & 127;
    const int sw = (row >> 1) & 3;
    uint32_t ph = 0;
    for (int t0 = blockIdx.x + 2 * grp * gridDim.x; t0 < tiles; t0 += 8 * gridDim.x, ph ^= 1u) {
      uint32_t q[2][16];                          // per tile: 32 x 16-bit, k = 0..26 inputs, 27..31 zero
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int i = 0; i < 16; ++i) q[u][i] = 0u;
        const int m = (t0 + u * static_cast<int>(gridDim.x)) * kBM + row;
        if (m < total) {
          const int ox = m % OW;
          const int r2 = m / OW;
          const int oy = r2 % OH;
          const int n = r2 / OH;
          const TIn* ip = in + ((static_cast<size_t>(n) * H + oy * 2) * W + ox * 2) * 3;
          uint16_t h[28];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 9; ++j) h[r * 9 + j] = conv0_tc_cvt<TIn>(__ldg(ip + static_cast<size_t>(r) * W * 3 + j), fmt);
          h[27] = 0;
#pragma unroll
          for (int i = 0; i < 14; ++i) q[u][i] = static_cast<uint32_t>(h[2 * i]) | (static_cast<uint32_t>(h[2 * i + 1]) << 16);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t0 + u * static_cast<int>(gridDim.x) < tiles) {
          const int s = 2 * grp + u;
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          uint8_t* arow = gbase + s * kABytes0 + row * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(arow + ((c ^ sw) << 4)) =
                make_uint4(q[u][4 * c], q[u][4 * c + 1], q[u][4 * c + 2], q[u][4 * c + 3]);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(full_bar(s));
        }
      }
    }
  } else if (warp == 16) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = ptx::make_idesc_f16(fmt, kBM, 32);
    const uint64_t hi64 = ptx::make_smem_desc_rows(0, 64);
    const uint64_t db = hi64 | ((w_base & 0x3FFFFu) >> 4);
    int s = 0;
    uint32_t ph = 0, lt = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
      const uint32_t as = lt & 1u;
      ptx::mbar_wait(tmem_empty(as), ((lt >> 1) & 1u) ^ 1u);
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t da = hi64 | (((base + s * kABytes0) & 0x3FFFFu) >> 4);
        ptx::umma_f16(tmem_base + as * 32u, da, db, idesc, 0u);
        ptx::umma_f16(tmem_base + as * 32u, da + 2u, db + 2u, idesc, 1u);
        ptx::umma_commit(empty_bar(s));
        ptx::umma_commit(tmem_full(as));
      }
      __syncwarp();
      if (++s == kConv0Stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- epilogue: + shift, ReLU, 64 contiguous bytes per pixel ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
      const uint32_t as = lt & 1u;
      ptx::mbar_wait(tmem_full(as), (lt >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + as * 32u + (static_cast<uint32_t>(quarter * 32) << 16);
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      ptx::tmem_ld_32x32b_x16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(tmem_empty(as));
      const int m = t * kBM + row;
      if (m < total) {
        uint4* op = reinterpret_cast<uint4*>(out + static_cast<size_t>(m) * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = g * 8 + 2 * j;
            const float x0 = fmaxf(__uint_as_float(v[c]) + p.shift[c], 0.f);
            const float x1 = fmaxf(__uint_as_float(v[c + 1]) + p.shift[c + 1], 0.f);
            w4[j] = pack2(x0, x1, fmt);
          }
          op[g] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    ptx::tc_fence_after();
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 64);
  }
}

inline size_t conv0_tc_smem_bytes() { return 1024 + kConv0Stages * 128 * 64 + 2048 + 256; }

}  // namespace icp
