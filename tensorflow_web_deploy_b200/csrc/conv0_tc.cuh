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
                uint16_t* __restrict__ out, int OH, int OW, int fmt, int pdl_late) {
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
  const TIn* in_end = in + static_cast<size_t>(N) * H * W * 3;

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
  if (!pdl_late) ptx::pdl_launch();
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
    const int step = 8 * static_cast<int>(gridDim.x);
    auto publish = [&](int u, const uint32_t (&q)[16]) {
      const int s = 2 * grp + u;
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      uint8_t* arow = gbase + s * kABytes0 + row * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(arow + ((c ^ sw) << 4)) = make_uint4(q[4 * c], q[4 * c + 1], q[4 * c + 2], q[4 * c + 3]);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(full_bar(s));
    };
    if constexpr (sizeof(TIn) == 1) {
      // uint8 input.  A run of 9 bytes at any address lies inside three ALIGNED 32-bit words; two funnel shifts bring it to
      // byte 0: 9 word loads per pixel instead of 27 byte loads.  (Requesting the NEXT pair of tiles' words before converting
      // the current pair — software pipelining — was MEASURED slower, 85.6 vs 80.5 us: the registers it holds spill.)
      int t0 = blockIdx.x + 2 * grp * static_cast<int>(gridDim.x);
      int pn[2], py[2], px[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long m = static_cast<long long>(t0 + u * static_cast<int>(gridDim.x)) * kBM + row;
        px[u] = static_cast<int>(m % OW);
        py[u] = static_cast<int>((m / OW) % OH);
        pn[u] = static_cast<int>(m / (static_cast<long long>(OW) * OH));
      }
      const long long delta = static_cast<long long>(step) * kBM;       // pixels between a slot's consecutive tiles
      const int dx = static_cast<int>(delta % OW), dy = static_cast<int>((delta / OW) % OH);
      const int dn = static_cast<int>(delta / (static_cast<long long>(OW) * OH));
      auto request = [&](int tq, uint32_t (&w)[2][9], uint32_t (&sh)[2]) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          sh[u] = 0xFFFFFFFFu;                                    // no pixel: zero row
          const int n = pn[u], oy = py[u], ox = px[u];            // this thread's pixel of tile tq + u * gridDim.x
          px[u] += dx;                                            // ... and of the same slot one iteration later (no divisions)
          if (px[u] >= OW) { px[u] -= OW; ++py[u]; }
          py[u] += dy;
          if (py[u] >= OH) { py[u] -= OH; ++pn[u]; }
          pn[u] += dn;
          if (tq + u * static_cast<int>(gridDim.x) < tiles && n < N) {
            const uint8_t* a = reinterpret_cast<const uint8_t*>(in) + ((static_cast<size_t>(n) * H + oy * 2) * W + ox * 2) * 3;
            uint32_t shs = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r, a += static_cast<size_t>(W) * 3) {
              const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
              const uint32_t* wp = reinterpret_cast<const uint32_t*>(ai & ~uintptr_t(3));
              if (reinterpret_cast<const uint8_t*>(wp) >= reinterpret_cast<const uint8_t*>(in) &&
                  reinterpret_cast<const uint8_t*>(wp + 3) <= reinterpret_cast<const uint8_t*>(in_end)) {
                w[u][3 * r] = __ldg(wp);
                w[u][3 * r + 1] = __ldg(wp + 1);
                w[u][3 * r + 2] = __ldg(wp + 2);
                shs |= static_cast<uint32_t>(ai & 3u) << (8 * r + 3);        // shift in bits, one byte of `sh` per row
              } else {                                                         // first / last bytes of the buffer: byte loads
                uint32_t b[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) b[j] = __ldg(a + j);
                w[u][3 * r] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
                w[u][3 * r + 1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
                w[u][3 * r + 2] = b[8];
              }
            }
            sh[u] = shs;
          }
        }
      };
      // bytes -> 16-bit without the conversion unit (I2F + F2F run at a quarter of the ALU rate; 54 of them per pixel were
      // ~65k of the kernel's ~150k cycles per SM): fp16 0x6400|b IS 1024 + b, minus 1024 exactly b (HSUB2, two values per
      // instruction); bf16: fp32 0x4B000000|b is 2^23 + b, minus 2^23 exactly b, whose upper 16 bits are the bf16 (b < 256).
      auto convert = [&](const uint32_t (&w)[9], uint32_t shs, uint32_t (&q)[16]) {
        uint32_t by[28];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint32_t sb = (shs >> (8 * r)) & 0xFFu;
          const uint32_t v0 = __funnelshift_r(w[3 * r], w[3 * r + 1], sb), v1 = __funnelshift_r(w[3 * r + 1], w[3 * r + 2], sb);
          const uint32_t v2 = w[3 * r + 2] >> sb;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            by[r * 9 + j] = (v0 >> (8 * j)) & 0xFFu;
            by[r * 9 + 4 + j] = (v1 >> (8 * j)) & 0xFFu;
          }
          by[r * 9 + 8] = v2 & 0xFFu;
        }
        by[27] = 0u;
        if (fmt == 0) {
          const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
#pragma unroll
          for (int i = 0; i < 14; ++i) {
            const uint32_t two = 0x64006400u | by[2 * i] | (by[2 * i + 1] << 16);
            const __half2 d = __hsub2(*reinterpret_cast<const __half2*>(&two), k1024);
            q[i] = *reinterpret_cast<const uint32_t*>(&d);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 14; ++i) {
            const float f0 = __uint_as_float(0x4B000000u | by[2 * i]) - 8388608.f;
            const float f1 = __uint_as_float(0x4B000000u | by[2 * i + 1]) - 8388608.f;
            q[i] = (__float_as_uint(f0) >> 16) | (__float_as_uint(f1) & 0xFFFF0000u);
          }
        }
        q[14] = 0u;
        q[15] = 0u;
      };
      uint32_t wc[2][9], sc[2];
      for (; t0 < tiles; t0 += step, ph ^= 1u) {
        request(t0, wc, sc);                       // both tiles' 18 word loads in flight before the first conversion
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (t0 + u * static_cast<int>(gridDim.x) < tiles) {
            uint32_t q[16];
            if (sc[u] == 0xFFFFFFFFu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) q[i] = 0u;
            } else {
              convert(wc[u], sc[u], q);
            }
            publish(u, q);
          }
        }
      }
    } else {
      for (int t0 = blockIdx.x + 2 * grp * gridDim.x; t0 < tiles; t0 += step, ph ^= 1u) {
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
        for (int u = 0; u < 2; ++u)
          if (t0 + u * static_cast<int>(gridDim.x) < tiles) publish(u, q[u]);
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
    if (pdl_late) ptx::pdl_launch();
  } else {
    // ---------------- epilogue: + shift, ReLU, 64 contiguous bytes per pixel ----------------
    // (a second epilogue team — 8 warps, alternate tiles — was MEASURED: 0.098 -> 0.105 ms; this chain is not the bound)
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
