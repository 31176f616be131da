// Halo-resident, weight-resident, persistent implicit-GEMM convolution (stride 1, kh x kw > 1x1).
//
// conv_igemm_* fetches every filter tap's A tile and every tile's weights from L2 again, so the stem and
// the 35x35 convs are pinned at the L2->SM fabric (ncu: conv1 moves 3.98 GB across it for 182 MB of input,
// profiles/r01_ncu_conv_igemm_v1_*.csv).  Here both operands are fetched once:
//
//  * Outputs are indexed on the *padded grid*: position p = oh * wp + wc with wp = iw + 2*pad_w columns, so the
//    input pixel feeding output p through tap (r,s) is padded-grid position p + r*wp + s — a constant shift.
//    Columns wc >= ow are computed and discarded (1..11 % extra MMA work on the maps this kernel serves).
//  * A "band" = T consecutive 128-position sub-tiles of one image.  For each block of ck = row_bytes/2 input
//    channels ONE 4-D tiled TMA box {ck, wp, rows, 1} at (w = -pad_w, h = first_row - pad_h) lands the band's
//    whole halo strip in shared memory in padded-grid order; out-of-bounds zero fill IS the SAME padding.
//    row_bytes in {32, 64, 128} (SWIZZLE_32B/64B/128B) is chosen per layer so Cin = 32/48/96 waste no MMA work.
//  * The whole packed weight matrix ([chunk][tap][n_tile x ck]) is loaded into shared memory once per SM and
//    stays there while the persistent CTA walks its bands.
//  * The MMA issuer selects tap (r,s) of sub-tile j by moving the A descriptor's start address to
//    (j*128 + r*wp + s) * row_bytes inside the strip: the UMMA swizzle is a function of the absolute smem
//    address, exactly like the TMA write that produced the strip, so a shifted start reads the shifted pixels.
//  * T accumulators per band, double-buffered in TMEM (2*T*n_tile <= 512 columns): the epilogue of band i
//    overlaps the strip load and MMAs of band i+1.
#pragma once
#include "conv_igemm.cuh"

namespace icp {

struct HaloArgs {
  int n_images;
  int oh, ow;           // valid output size (per image)
  int wp;               // padded grid width = iw + 2 * pad_w
  int pad_h, pad_w;
  int kh, kw;
  int chunks;           // ceil(cin / ck)
  int row_bytes;        // 32 / 64 / 128  (ck = row_bytes / 2 channels per block)
  int T;                // 128-position sub-tiles per band
  int strip_rows;       // padded-grid rows per strip
  int strip_bytes;      // strip_rows * wp * row_bytes (wp * row_bytes is a multiple of 1024)
  int nstrips;          // strips in the shared-memory ring (2..4)
  int bands_per_image;  // ceil(oh * wp / (128 * T))
  int n_tile;           // = cout (<= 256)
  int acc_cols;         // TMEM columns of one accumulator set: T * n_tile
  int tmem_cols;        // power of two >= 2 * acc_cols
  int w_bytes;          // chunks * taps * n_tile * row_bytes (resident weights)
  int fmt;
  int relu;
  void* out;
  int ldo, coff;
  const float* scale;
  const float* shift;
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  const uint32_t strip0 = base;
  const uint32_t wbase = base + static_cast<uint32_t>(a.nstrips) * a.strip_bytes;   // resident weights
  const uint32_t bar_base = wbase + ((static_cast<uint32_t>(a.w_bytes) + 1023u) & ~1023u);
  // barriers: strip_full[4], strip_empty[4], tmem_full[2], tmem_empty[2], w_full, then the TMEM slot
  auto strip_full = [&](int i) { return bar_base + 8u * i; };
  auto strip_empty = [&](int i) { return bar_base + 32u + 8u * i; };
  auto tmem_full = [&](int i) { return bar_base + 64u + 8u * i; };
  auto tmem_empty = [&](int i) { return bar_base + 80u + 8u * i; };
  const uint32_t w_full = bar_base + 96u;
  const uint32_t tmem_slot = bar_base + 104u;
  float* s_scale = reinterpret_cast<float*>(gbase + (bar_base - base) + 128);
  float* s_shift = s_scale + 256;

  const int taps = a.kh * a.kw;
  const int ck = a.row_bytes / 2;
  const int total_bands = a.bands_per_image * a.n_images;
  const int positions = a.oh * a.wp;                 // padded-grid positions holding outputs, per image
  const uint32_t tap_bytes = static_cast<uint32_t>(a.n_tile) * rb;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_b);
    for (int i = 0; i < a.nstrips; ++i) {
      ptx::mbar_init(strip_full(i), 1);
      ptx::mbar_init(strip_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), 128);
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    for (int i = et; i < a.n_tile; i += 128) {
      s_scale[i] = __ldg(a.scale + i);
      s_shift[i] = __ldg(a.shift + i);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      ptx::mbar_expect_tx(w_full, static_cast<uint32_t>(a.w_bytes));
      for (int i = 0; i < a.chunks * taps; ++i) ptx::tma_load_2d(wbase + i * tap_bytes, &tm_b, w_full, i * ck, 0);
      const uint32_t strip_tx = static_cast<uint32_t>(a.strip_rows) * a.wp * rb;
      uint32_t it = 0;
      for (int band = blockIdx.x; band < total_bands; band += gridDim.x) {
        const int img = band / a.bands_per_image;
        const int p0 = (band - img * a.bands_per_image) * kBM * a.T;
        const int h_first = p0 / a.wp;
        for (int c = 0; c < a.chunks; ++c, ++it) {
          const uint32_t s = it % a.nstrips;
          ptx::mbar_wait(strip_empty(s), ((it / a.nstrips) & 1u) ^ 1u);
          ptx::mbar_expect_tx(strip_full(s), strip_tx);
          ptx::tma_load_4d(strip0 + s * a.strip_bytes, &tm_a, strip_full(s), c * ck, -a.pad_w, h_first - a.pad_h, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      // The issuing thread is the bottleneck for narrow layers (an M128 x N32 x K16 MMA occupies the tensor
      // pipe for 16 cycles), so the inner loop carries no division and no 64-bit descriptor construction:
      // descriptors differ only in their 14-bit address field (16-byte units), advanced by table lookups.
      const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.n_tile);
      const int ksteps = a.row_bytes / 32;
      const uint64_t desc_hi = ptx::make_smem_desc_rows(0, rb);          // everything but the address
      uint32_t* tap_off = reinterpret_cast<uint32_t*>(s_shift + 256);     // [taps] A shift of tap (r,s), 16-B units
      for (int tap = 0; tap < taps; ++tap) {
        const int r = tap / a.kw;
        tap_off[tap] = (static_cast<uint32_t>(r * a.wp + (tap - r * a.kw)) * rb) >> 4;
      }
      const uint32_t tap_units = tap_bytes >> 4;
      const uint32_t sub_units = (static_cast<uint32_t>(kBM) * rb) >> 4;
      ptx::mbar_wait(w_full, 0);
      uint32_t it = 0, lb = 0;
      int img = static_cast<int>(blockIdx.x) / a.bands_per_image;
      int bi = static_cast<int>(blockIdx.x) - img * a.bands_per_image;      // band index inside the image
      const int img_step = static_cast<int>(gridDim.x) / a.bands_per_image;
      const int bi_step = static_cast<int>(gridDim.x) - img_step * a.bands_per_image;
      for (int band = blockIdx.x; band < total_bands; band += gridDim.x, ++lb) {
        const int p0 = bi * kBM * a.T;
        const int h_first = p0 / a.wp;
        const uint32_t local_units = (static_cast<uint32_t>(p0 - h_first * a.wp) * rb) >> 4;
        int nsub = (positions - p0 + kBM - 1) / kBM;                       // sub-tiles that hold outputs
        if (nsub > a.T) nsub = a.T;
        const uint32_t as = lb & 1u;
        ptx::mbar_wait(tmem_empty(as), ((lb >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        for (int c = 0; c < a.chunks; ++c, ++it) {
          const uint32_t s = it % a.nstrips;
          ptx::mbar_wait(strip_full(s), (it / a.nstrips) & 1u);
          ptx::tc_fence_after();
          const uint32_t a_units0 = ((strip0 + s * a.strip_bytes) >> 4) + local_units;
          const uint32_t b_units0 = (wbase >> 4) + static_cast<uint32_t>(c * taps) * tap_units;
          for (int j = 0; j < nsub; ++j) {
            const uint32_t tacc = tmem_base + as * a.acc_cols + j * a.n_tile;
            const uint32_t a_units = a_units0 + j * sub_units;
            uint32_t b_units = b_units0;
            for (int tap = 0; tap < taps; ++tap, b_units += tap_units) {
              const uint32_t au = a_units + tap_off[tap];
              for (int k = 0; k < ksteps; ++k)
                ptx::umma_f16(tacc, desc_hi | (au + 2u * k), desc_hi | (b_units + 2u * k), idesc,
                              static_cast<uint32_t>(c | tap | k));
            }
          }
          ptx::umma_commit(strip_empty(s));
        }
        ptx::umma_commit(tmem_full(as));
        bi += bi_step;
        img += img_step;
        if (bi >= a.bands_per_image) { bi -= a.bands_per_image; ++img; }
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t lb = 0;
    for (int band = blockIdx.x; band < total_bands; band += gridDim.x, ++lb) {
      const int img = band / a.bands_per_image;
      const int p0 = (band - img * a.bands_per_image) * kBM * a.T;
      const uint32_t as = lb & 1u;
      ptx::mbar_wait(tmem_full(as), (lb >> 1) & 1u);
      ptx::tc_fence_after();
      for (int j = 0; j < a.T; ++j) {
        if (p0 + j * kBM >= positions) break;
        const int p = p0 + j * kBM + row;
        const int oy = p / a.wp;
        const int ox = p - oy * a.wp;
        const bool valid = (oy < a.oh) && (ox < a.ow);
        uint16_t* orow = reinterpret_cast<uint16_t*>(a.out) +
                         ((static_cast<size_t>(img) * a.oh + oy) * a.ow + ox) * a.ldo + a.coff;
        const uint32_t taddr = tmem_base + as * a.acc_cols + j * a.n_tile + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int c0 = 0; c0 < a.n_tile; c0 += 32) {
          uint32_t v[32];
          const bool two = c0 + 16 < a.n_tile;
          ptx::tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          if (two) ptx::tmem_ld_32x32b_x16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int cc = c0 + 16 * h;
            uint32_t q[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float x0 = fmaf(__uint_as_float(v[16 * h + 2 * i]), s_scale[cc + 2 * i], s_shift[cc + 2 * i]);
              float x1 = fmaf(__uint_as_float(v[16 * h + 2 * i + 1]), s_scale[cc + 2 * i + 1], s_shift[cc + 2 * i + 1]);
              if (a.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
              q[i] = pack2(x0, x1, a.fmt);
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(orow + cc);
              dst[0] = make_uint4(q[0], q[1], q[2], q[3]);
              dst[1] = make_uint4(q[4], q[5], q[6], q[7]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tmem_empty(as));
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

inline size_t halo_smem_bytes(int strip_bytes, int nstrips, int w_bytes) {
  return 1024 + static_cast<size_t>(nstrips) * strip_bytes + ((static_cast<size_t>(w_bytes) + 1023) & ~size_t(1023)) + 128 +
         2048 + 256;   // + per-tap offset table
}

}  // namespace icp
