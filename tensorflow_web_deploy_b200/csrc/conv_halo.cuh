// Halo-resident implicit-GEMM convolution (stride 1, kh x kw > 1x1) for sm_100a.
//
// conv_igemm_kernel fetches every filter tap's A tile from L2 again, so a k_h x k_w conv moves each input
// pixel k_h*k_w times across the L2->SM fabric; ncu shows the stem and the 35x35 convs pinned at that
// fabric's ~9 TB/s (profiles/r01_ncu_conv_igemm_v1_*.csv).  Here the input is read ONCE per tile:
//
//  * Outputs are indexed on the *padded grid*: position p = oh * wp + wc with wp = iw + 2*pad_w columns, so
//    the input pixel feeding output p through tap (r,s) is padded-grid position p + r*wp + s — a constant
//    shift.  Columns wc >= ow are computed and discarded (2..11 % extra MMA work on the maps this kernel is
//    used for).
//  * For each block of ck = row_bytes/2 input channels, ONE 4-D tiled TMA box {ck, wp, rows, 1} starting at
//    (w = -pad_w, h = first_row - pad_h) lands the whole halo strip in shared memory in padded-grid order;
//    the unit's out-of-bounds zero fill IS the SAME padding.  row_bytes in {32, 64, 128} is chosen per layer
//    so that Cin = 32/48/80/96 waste no MMA work (SWIZZLE_32B/64B/128B).
//  * The MMA issuer walks the taps by moving the A descriptor's start address by (r*wp + s) * row_bytes: the
//    UMMA swizzle is a function of the absolute smem address, exactly like the TMA write that produced the
//    strip, so a shifted start reads the shifted pixels.  B tiles ([n_tile x ck] per (channel block, tap))
//    stream through their own small ring.
//
// Epilogue as in conv_igemm_kernel (TMEM -> regs -> scale/shift/ReLU -> 16-bit -> concat slice), with the
// padded-grid position mapped back to (oh, ow) and the discard columns masked.
#pragma once
#include "conv_igemm.cuh"

namespace icp {

struct HaloArgs {
  int oh, ow;           // valid output size (per image)
  int wp;               // padded grid width = iw + 2 * pad_w
  int pad_h, pad_w;
  int kh, kw;
  int chunks;           // ceil(cin / ck)
  int row_bytes;        // 32 / 64 / 128  (ck = row_bytes / 2 channels per block)
  int strip_rows;       // padded-grid rows per strip
  int strip_bytes;      // strip_rows * wp * row_bytes rounded up to 1024
  int tiles_per_image;  // ceil(oh * wp / 128)
  int n_tile;
  int tmem_cols;
  int b_stages;
  int fmt;
  int relu;
  void* out;
  int ldo, coff;
  const float* scale;
  const float* shift;
};

__global__ void __launch_bounds__(kConvThreads)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  const uint32_t b_bytes = static_cast<uint32_t>(a.n_tile) * rb;           // multiple of 512
  const uint32_t b_stride = (b_bytes + 1023u) & ~1023u;
  const uint32_t strip0 = base;
  const uint32_t bring = base + 2u * a.strip_bytes;
  const uint32_t bar_base = bring + a.b_stages * b_stride;
  // barriers: a_full[2], a_empty[2], b_full[S], b_empty[S], tmem_full, then the TMEM slot
  auto a_full = [&](int i) { return bar_base + 8u * i; };
  auto a_empty = [&](int i) { return bar_base + 16u + 8u * i; };
  auto b_full = [&](int i) { return bar_base + 32u + 8u * i; };
  auto b_empty = [&](int i) { return bar_base + 32u + 8u * (a.b_stages + i); };
  const uint32_t tmem_full_bar = bar_base + 32u + 16u * a.b_stages;
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  float* s_scale = reinterpret_cast<float*>(gbase + (bar_base - base) + 256);
  float* s_shift = s_scale + 256;

  const int tile = blockIdx.x % a.tiles_per_image;
  const int img = blockIdx.x / a.tiles_per_image;
  const int n_base = blockIdx.y * a.n_tile;
  const int p0 = tile * kBM;                       // first padded-grid position of this tile
  const int h_first = p0 / a.wp;                   // first padded-grid row held by the strip
  const int taps = a.kh * a.kw;
  const int ck = a.row_bytes / 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(a_full(i), 1); ptx::mbar_init(a_empty(i), 1); }
    for (int i = 0; i < a.b_stages; ++i) { ptx::mbar_init(b_full(i), 1); ptx::mbar_init(b_empty(i), 1); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer: strip(c+1) is requested before chunk c's weights ----------------
      const uint32_t strip_tx = static_cast<uint32_t>(a.strip_rows) * a.wp * rb;
      auto load_strip = [&](int c) {
        const int sa = c & 1;
        ptx::mbar_wait(a_empty(sa), ((c >> 1) & 1) ^ 1u);
        ptx::mbar_expect_tx(a_full(sa), strip_tx);
        ptx::tma_load_4d(strip0 + sa * a.strip_bytes, &tm_a, a_full(sa), c * ck, -a.pad_w, h_first - a.pad_h, img);
      };
      load_strip(0);
      int bi = 0;
      for (int c = 0; c < a.chunks; ++c) {
        if (c + 1 < a.chunks) load_strip(c + 1);
        for (int tap = 0; tap < taps; ++tap, ++bi) {
          const int sb = bi % a.b_stages;
          ptx::mbar_wait(b_empty(sb), ((bi / a.b_stages) & 1) ^ 1u);
          ptx::mbar_expect_tx(b_full(sb), b_bytes);
          ptx::tma_load_2d(bring + sb * b_stride, &tm_b, b_full(sb), (c * taps + tap) * ck, n_base);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.n_tile);
      const int ksteps = a.row_bytes / 32;
      const uint32_t local0 = static_cast<uint32_t>(p0 - h_first * a.wp);   // tile's first position inside the strip
      int bi = 0;
      uint32_t acc = 0;
      for (int c = 0; c < a.chunks; ++c) {
        const int sa = c & 1;
        ptx::mbar_wait(a_full(sa), (c >> 1) & 1);
        const uint32_t strip = strip0 + sa * a.strip_bytes;
        for (int tap = 0; tap < taps; ++tap, ++bi) {
          const int sb = bi % a.b_stages;
          ptx::mbar_wait(b_full(sb), (bi / a.b_stages) & 1);
          ptx::tc_fence_after();
          const int r = tap / a.kw;
          const uint32_t shift = static_cast<uint32_t>(r * a.wp + (tap - r * a.kw));
          const uint32_t a_addr = strip + (local0 + shift) * rb;
          const uint32_t b_addr = bring + sb * b_stride;
          for (int k = 0; k < ksteps; ++k) {
            ptx::umma_f16(tmem_base, ptx::make_smem_desc_rows(a_addr + 32u * k, rb),
                          ptx::make_smem_desc_rows(b_addr + 32u * k, rb), idesc, acc);
            acc = 1;
          }
          ptx::umma_commit(b_empty(sb));
        }
        ptx::umma_commit(a_empty(sa));
      }
      ptx::umma_commit(tmem_full_bar);
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int et = threadIdx.x - 64;
    for (int i = et; i < a.n_tile; i += 128) {
      s_scale[i] = __ldg(a.scale + n_base + i);
      s_shift[i] = __ldg(a.shift + n_base + i);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int quarter = warp & 3;
    const int p = p0 + quarter * 32 + lane;        // padded-grid position of this thread's row
    const int oy = p / a.wp;
    const int ox = p - oy * a.wp;
    const bool valid = (oy < a.oh) && (ox < a.ow);
    uint16_t* orow = reinterpret_cast<uint16_t*>(a.out) +
                     ((static_cast<size_t>(img) * a.oh + oy) * a.ow + ox) * a.ldo + a.coff + n_base;
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(taddr + c0, v);
      ptx::tmem_ld_wait();
      uint32_t q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x0 = fmaf(__uint_as_float(v[2 * j]), s_scale[c0 + 2 * j], s_shift[c0 + 2 * j]);
        float x1 = fmaf(__uint_as_float(v[2 * j + 1]), s_scale[c0 + 2 * j + 1], s_shift[c0 + 2 * j + 1]);
        if (a.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
        q[j] = pack2(x0, x1, a.fmt);
      }
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c0);
        dst[0] = make_uint4(q[0], q[1], q[2], q[3]);
        dst[1] = make_uint4(q[4], q[5], q[6], q[7]);
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

inline size_t halo_smem_bytes(int strip_bytes, int b_stages, int n_tile, int row_bytes) {
  const size_t b_stride = (static_cast<size_t>(n_tile) * row_bytes + 1023) & ~size_t(1023);
  return 1024 + 2 * static_cast<size_t>(strip_bytes) + b_stages * b_stride + 256 + 2048;
}

}  // namespace icp
