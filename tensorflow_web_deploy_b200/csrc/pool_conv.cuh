// max_pool 3x3/2 VALID fused INTO the operand load of the 1x1 conv that follows it (the stem's pool1 -> conv3,
// /root/reference/inception/slim/inception_model.py:100-103; ops.max_pool ops.py:367-370, ops.conv2d ops.py:229-245).
//
// Unfused, pool1 reads conv2's output (354 MB at batch 128) and writes 87 MB that conv3 reads straight back:
// 72 us + 65 us for 7 GFLOP, both kernels pure HBM / latency.  Here the pooled tensor never exists: 512 builder threads
// take the max over the nine 64-channel input pixels of each output pixel's window (eight threads per pixel, one 16-byte
// channel chunk each) and write the 128-byte rows of the A operand tile in shared memory, in the SWIZZLE_128B layout the UMMA descriptor expects
// (16-byte chunk index XOR (row & 7)); the [Cout x 64] weight tile (BatchNorm scale folded in) is resident; four K = 16
// tcgen05 MMAs per 128-pixel tile after the same K = 16 "bias" MMA (ones tile x (hi, lo) shift tile) the other conv
// kernels start their accumulators with; the epilogue applies ReLU and stores 16-bit NHWC.  max() is exact in the storage
// type and the MMA sequence is the unfused conv3's, so the result is bit-identical to the two-kernel path (the engine
// switches between them by batch size and `test_ragged_batches_and_batch_invariance` holds them equal).
// Structure = conv0_tc.cuh: warps 0..15 A builders (4 groups of 128, each filling two of the 8 stages), 16 = TMEM owner
// + MMA issuer, 17..24 = epilogue (two warps per TMEM lane quarter, splitting the columns).
#pragma once
#include "conv_igemm.cuh"
#include "tower_kernels.cuh"

namespace icp {

constexpr int kPoolConvThreads = 800;
constexpr int kPoolConvStages = 8;
constexpr int kPoolConvMaxCout = 128;

struct PoolConvArgs {
  const uint16_t* in;      // [N, H, W, 64] 16-bit NHWC (dense: pitch 64)
  int N, H, W;
  const uint16_t* w;       // [cout][64] 16-bit K-major, BatchNorm scale folded in
  const uint16_t* bias16;  // [cout][16] 16-bit: (hi, lo) halves of the BatchNorm shift in columns 0, 1 (engine.cu pack_bias)
  int cout;                // multiple of 16, <= 128
  uint16_t* out;           // [N, OH, OW, ldo]
  int OH, OW, ldo, coff;
  int relu, fmt;
  int pdl_late;            // see conv_igemm.cuh
};

__global__ void __launch_bounds__(kPoolConvThreads, 1) pool_conv_kernel(const PoolConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  constexpr uint32_t kStage = 128 * 128;                             // one A stage: 128 rows x 128 B
  const uint32_t w_base = base + kPoolConvStages * kStage;          // [cout x 128 B] weight tile, SWIZZLE_128B
  const uint32_t bias_base = w_base + kPoolConvMaxCout * 128;       // [cout x 32 B] shift tile, SWIZZLE_32B
  const uint32_t ones_base = bias_base + kPoolConvMaxCout * 32;     // [128 x 32 B] ones tile (conv_igemm.cuh)
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPoolConvStages + s); };
  auto tmem_full = [&](int i) { return bar_base + 16u * kPoolConvStages + 8u * i; };
  auto tmem_empty = [&](int i) { return bar_base + 16u * kPoolConvStages + 16u + 8u * i; };
  const uint32_t tmem_slot = bar_base + 16u * kPoolConvStages + 32u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = a.N * a.OH * a.OW;
  const int tiles = (total + kBM - 1) / kBM;

  if (!a.pdl_late) ptx::pdl_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kPoolConvStages; ++s) {
      ptx::mbar_init(full_bar(s), 128);          // every builder thread of the group arrives
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), 256);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 16) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  // weight tile + shift (constants: loaded before the dependency wait)
  for (int i = threadIdx.x; i < a.cout * 8; i += blockDim.x) {
    const int n = i >> 3, c = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w + n * 64 + c * 8));
    *reinterpret_cast<uint4*>(gbase + (w_base - base) + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < a.cout * 2; i += blockDim.x) {
    const int n = i >> 1, c = i & 1;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.bias16 + n * 16 + c * 8));
    *reinterpret_cast<uint4*>(gbase + (bias_base - base) + n * 32 + ((c ^ ((n >> 2) & 1)) << 4)) = v;
  }
  if (threadIdx.x < 128) write_ones_row(gbase + (ones_base - base), static_cast<int>(threadIdx.x), a.fmt);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  ptx::pdl_wait();

  if (warp < 16) {
    // ---------------- A builders ----------------
    // Thread (cl = tid & 7, r0 = tid >> 3 within the group of 128) owns the 16-byte channel chunk cl of rows r0, r0+16, ...,
    // r0+112 of the tile: 8 consecutive lanes read one input pixel's whole 128-byte line, so a warp load touches 4 lines
    // (the first version gave each thread a whole pixel — 32 different lines per load instruction — and was bound by the
    // L1's line rate: the fused kernel took ~220 us against 72 + 62 us for pool1 + conv3).
    const int grp = warp >> 2;
    const int gt = threadIdx.x & 127;
    const int cl = gt & 7;
    const int r0 = gt >> 3;                      // 0..15
    uint32_t ph = 0;
    // group g serves local tiles 2g and 2g+1 (mod 8) -> stages 2g, 2g+1
    for (int t0 = blockIdx.x + 2 * grp * gridDim.x; t0 < tiles; t0 += 8 * gridDim.x, ph ^= 1u) {
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * static_cast<int>(gridDim.x);
        if (t >= tiles) break;
        // first row of this thread: output pixel m -> (n, oy, ox); the next rows are 16 pixels further each
        int m = t * kBM + r0;
        int ox = m % a.OW;
        const int r2 = m / a.OW;
        int oy = r2 % a.OH;
        int n = r2 / a.OH;
        uint4 mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          mx[i] = make_uint4(0u, 0u, 0u, 0u);
          if (m < total) {
            const uint16_t* ip = a.in + ((static_cast<size_t>(n) * a.H + oy * 2) * a.W + ox * 2) * 64 + cl * 8;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int q = 0; q < 3; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(ip + (static_cast<size_t>(r) * a.W + q) * 64));
                mx[i] = (r == 0 && q == 0) ? v : max8_h16(mx[i], v, a.fmt);
              }
            }
          }
          m += 16;
          ox += 16;
          while (ox >= a.OW) {
            ox -= a.OW;
            if (++oy == a.OH) { oy = 0; ++n; }
          }
        }
        const int s = 2 * grp + u;
        ptx::mbar_wait(empty_bar(s), ph ^ 1u);
        uint8_t* st = gbase + s * kStage;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = r0 + 16 * i;
          *reinterpret_cast<uint4*>(st + row * 128 + ((cl ^ (row & 7)) << 4)) = mx[i];
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(full_bar(s));
      }
    }
  } else if (warp == 16) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.cout);
    const uint64_t hi = ptx::make_smem_desc_rows(0, 128);
    const uint64_t db = hi | ((w_base & 0x3FFFFu) >> 4);
    const uint64_t d32 = ptx::make_smem_desc_rows(0, 32);
    const uint64_t d_ones = d32 | ((ones_base & 0x3FFFFu) >> 4);
    const uint64_t d_bias = d32 | ((bias_base & 0x3FFFFu) >> 4);
    int s = 0;
    uint32_t ph = 0, lt = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
      const uint32_t as = lt & 1u;
      ptx::mbar_wait(tmem_empty(as), ((lt >> 1) & 1u) ^ 1u);
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t da = hi | (((base + s * kStage) & 0x3FFFFu) >> 4);
        const uint32_t tacc = tmem_base + as * 128u;
        ptx::umma_f16(tacc, d_ones, d_bias, idesc, 0u);       // accumulator := shift, as in the unfused conv3
        ptx::umma_f16(tacc, da, db, idesc, 1u);
        ptx::umma_f16(tacc, da + 2u, db + 2u, idesc, 1u);
        ptx::umma_f16(tacc, da + 4u, db + 4u, idesc, 1u);
        ptx::umma_f16(tacc, da + 6u, db + 6u, idesc, 1u);
        ptx::umma_commit(empty_bar(s));
        ptx::umma_commit(tmem_full(as));
      }
      __syncwarp();
      if (++s == kPoolConvStages) { s = 0; ph ^= 1u; }
    }
    if (a.pdl_late) ptx::pdl_launch();
  } else {
    // ---------------- epilogue: ReLU, 16-bit NHWC; two warps per lane quarter split the 16-column pieces ----------------
    const int ew = warp - 17;                   // 0..7
    const int quarter = warp & 3;               // TMEM lanes this warp may touch: [32q, 32q+32)
    const int grp = ew >> 2;                    // warps 17..20 -> 0, 21..24 -> 1 (each group covers all four quarters)
    const int row = quarter * 32 + lane;
    const int pieces = a.cout >> 4;
    const int h0 = (pieces + 1) >> 1;
    const int fp = grp ? h0 : 0;
    const int np = grp ? pieces - h0 : h0;      // <= 4
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++lt) {
      const uint32_t as = lt & 1u;
      ptx::mbar_wait(tmem_full(as), (lt >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + as * 128u + (static_cast<uint32_t>(quarter * 32) << 16) + 16u * fp;
      const int m = t * kBM + row;
      uint4* op = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(m) * a.ldo + a.coff + 16 * fp);
      for (int j0 = 0; j0 < np; j0 += 2) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x16(taddr + 16 * j0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        const bool two = j0 + 1 < np;
        if (two) ptx::tmem_ld_32x32b_x16(taddr + 16 * j0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        ptx::tmem_ld_wait();
        if (j0 + 2 >= np) {                       // last TMEM read of this tile
          ptx::tc_fence_before();
          ptx::mbar_arrive(tmem_empty(as));
        }
        uint32_t pk[16];
        convert16(&v[0], &pk[0], a.relu != 0, a.fmt);
        if (two) convert16(&v[16], &pk[8], a.relu != 0, a.fmt);
        if (m < total) {
          op[2 * j0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          op[2 * j0 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if (two) {
            op[2 * j0 + 2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            op[2 * j0 + 3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          }
        }
      }
      if (np <= 0) { ptx::tc_fence_before(); ptx::mbar_arrive(tmem_empty(as)); }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    ptx::tc_fence_after();
    __syncwarp();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

inline size_t pool_conv_smem_bytes() {
  return 1024 + kPoolConvStages * 128 * 128 + kPoolConvMaxCout * 128 + kPoolConvMaxCout * 32 + kOnesBytes + 512;
}

}  // namespace icp
