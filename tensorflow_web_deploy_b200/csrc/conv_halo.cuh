// Halo-resident persistent implicit-GEMM convolution (stride 1, kh x kw > 1x1) for sm_100a.
//
// Measured on B200 (profiles/): an SM pulls at most ~47 B/cycle (~13 TB/s chip-wide) through TMA from L2,
// and the tap-wise im2col kernel moves each input pixel kh*kw times and each weight tile once per 128-pixel
// tile across that path — it, not the tensor pipe, bounds every k>1 layer.  This kernel moves far fewer bytes:
//
//  * Outputs are indexed on a *padded grid*: position p = row * wp + col with wp = iw + 2*pad_w columns, so the
//    input pixel feeding output p through tap (r,s) is padded-grid position p + r*wp + s — a constant shift.
//    Columns col >= ow are computed and discarded.
//  * A "band" = T consecutive 128-position sub-tiles.  For each block of ck = row_bytes/2 input channels ONE
//    4-D tiled TMA box {ck, wp, rows, 1} at (w = -pad_w, h = first_row - pad_h) lands the band's whole halo
//    strip in shared memory in padded-grid order; the unit's out-of-bounds zero fill IS the SAME padding.
//    Every input pixel crosses the fabric ~(1 + halo/T) times instead of kh*kw times.
//  * The MMA issuer selects tap (r,s) of sub-tile j by moving the A descriptor's start address to
//    (j*128 + r*wp + s) * row_bytes inside the strip: the UMMA swizzle is a function of the absolute smem
//    address, exactly like the TMA write that produced the strip, so a shifted start reads the shifted pixels.
//  * Weights: either the whole packed matrix stays resident in shared memory for the life of the persistent CTA
//    (small layers), or [n_tile x ck] tiles per (channel block, tap) stream through a ring fed by a dedicated
//    producer warp; either way one weight tile serves all T sub-tiles of the band.
//  * T accumulators per band in TMEM, double-buffered when 2*T*n_tile <= 512 columns so the epilogue of one band
//    overlaps the MMAs of the next.
//  * "tall" mode (kh == 1): rows of ALL images form one padded grid (no vertical taps, so no vertical padding is
//    needed between images) — bands run across image boundaries and 17x17 / 8x8 maps keep full 128-row tiles.
//    k x 1 filters run as 1 x k on transposed storage: `transpose_out` makes the epilogue store pixel (y,x) at
//    (x,y), so a 1x7 -> 7x1 chain alternates layouts for free.
//
// Warp roles: 0 = strip producer, 1 = TMEM owner + MMA issuer, 2..5 and 7..10 = epilogue, 6 = weight producer.
#pragma once
#include "conv_igemm.cuh"

namespace icp {

constexpr int kHaloThreads = 352;

struct HaloArgs {
  int n_images;
  int oh, ow;           // valid output size (per image)
  int wp;               // padded grid width = iw + 2 * pad_w
  int pad_h, pad_w;
  int kh, kw;
  int chunks;           // ceil(cin / ck)
  int row_bytes;        // 64 / 128  (ck = row_bytes / 2 channels per block)
  int T;                // 128-position sub-tiles per band
  int strip_rows;       // padded-grid rows per strip
  int strip_bytes;      // strip_rows * wp * row_bytes rounded up to 1024
  int nstrips;          // strips in the shared-memory ring (2..4)
  int bands_per_image;  // ceil(oh * wp / (128 * T));  tall mode: unused
  int tall;             // 1: one grid over all images' rows (kh == 1)
  int transpose_out;    // 1: store output pixel (y, x) at (x, y)   (square maps)
  int n_tile, n_tiles;  // output channels per work item, number of N tiles
  int acc_cols;         // TMEM columns of one accumulator set: T * n_tile
  int acc_sets;         // 2 when 2 * acc_cols <= 512, else 1
  int tmem_cols;        // power of two >= acc_sets * acc_cols
  int w_resident;       // 1: the weights of this CTA's N tile live in smem for its whole life (grid is a multiple of n_tiles,
                        //    so a CTA's items all have N tile blockIdx.x % n_tiles)
  int tail_ksteps;      // K = 16 MMAs issued for the LAST channel chunk (its zero-padded K slices are skipped)
  int w_bytes;          // resident: chunks * taps * n_tile * row_bytes; streaming: b_stages * tile pitch
  int b_stages;         // streaming ring depth (<= 8)
  int fmt;
  int relu;
  void* out;
  int ldo, coff;
  int epi_direct;       // 1: epilogue threads store their 32 columns straight from registers (no staging, no barriers)
  long long* dbg;       // optional timeline buffer (CTA 0 only), same rows as the persistent kernel (tools/timeline.py)
  // strip = nbox boxes of box_rows rows (+ one of rem_rows through tm_a_rem): several boxes in flight per strip
  int box_rows, nbox, rem_rows;
  int pdl_late;         // 1: launch_dependents when this CTA's last band's MMAs are issued, not at kernel start (conv_igemm.cuh)
  // kCp (strips copied by cp.async instead of TMA): the input tensor itself
  const uint16_t* in;   // first channel of this layer's input slice
  int ih, iw, ldi;      // input height, width, channels per pixel of the buffer (elements)
};

// kDirect: the epilogue variant is a compile-time choice — carrying both in one kernel cost the direct path 20 %
// (conv1 0.120 -> 0.147 ms when a third variant was added experimentally).
// kStem: the configuration of the two stem layers that use this kernel by default — weights resident, per-image bands,
// no transposed store — as compile-time constants (the streamed / tall / transposed paths are the opt-in ones).
// kCp (stem only): warps 0 and 6 copy the strips with 16-byte cp.async (zero fill outside the image) instead of one TMA box.
// MEASURED: a TMA box of 64-byte rows lands ~one row per 5 cycles per SM, and a band of 4 sub-tiles needs a 7-row strip of a
// 149-wide map = 1043 rows = ~5.2k cycles against ~3.9k of MMA issue — the strip loads, not HBM (2.7 of 6.5 TB/s) or the tensor
// pipe, bounded conv1 / conv2.  The LSU path has no per-row cost: 64 threads x 16 B per instruction.
// RESULT: slower (conv1 0.135 -> 0.215 ms, conv2 0.145 -> 0.231 ms): two warps cannot issue the copies fast enough; kept
// opt-in (ICP_HALO_CP=1) as the measured negative.
template <bool kDirect, bool kStem, bool kCp = false>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a_rem,
                 const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_bias, const HaloArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool f_resident = kStem ? true : a.w_resident != 0;
  const bool f_tall = kStem ? false : a.tall != 0;
  const bool f_transpose = kStem ? false : a.transpose_out != 0;
  const uint32_t rb = static_cast<uint32_t>(a.row_bytes);
  const uint32_t strip0 = base;
  const uint32_t wbase = base + static_cast<uint32_t>(a.nstrips) * a.strip_bytes;   // resident weights or ring
  const uint32_t bar_base = wbase + ((static_cast<uint32_t>(a.w_bytes) + 1023u) & ~1023u);
  // barriers: strip_full[4] strip_empty[4] tmem_full[2] tmem_empty[2] w_full b_full[8] b_empty[8]; TMEM slot
  auto strip_full = [&](int i) { return bar_base + 8u * i; };
  auto strip_empty = [&](int i) { return bar_base + 32u + 8u * i; };
  auto tmem_full = [&](int i) { return bar_base + 64u + 8u * i; };
  auto tmem_empty = [&](int i) { return bar_base + 80u + 8u * i; };
  const uint32_t w_full = bar_base + 96u;
  auto b_full = [&](int i) { return bar_base + 104u + 8u * i; };
  auto b_empty = [&](int i) { return bar_base + 168u + 8u * i; };
  const uint32_t tmem_slot = bar_base + 232u;
  uint32_t* tap_off = reinterpret_cast<uint32_t*>(gbase + (bar_base - base) + 256);   // [taps] A shift of tap (r,s), 16-B units
  int* s_pix = reinterpret_cast<int*>(tap_off + 64);                 // [2][128] per team: output pixel of each sub-tile row, -1 = none
  const uint32_t stage_base = bar_base + 2048u;                      // epilogue staging: 2 teams x 128 rows x 128 B
  const uint32_t ones_base = stage_base + (kDirect ? 0u : 2u * 16384u);   // constant A tile of the bias MMA (after the staged epilogue's buffers)
  const uint32_t bias_base = ones_base + kOnesBytes;                 // resident weights: the [n_tile x 16] shift tile

  const int taps = a.kh * a.kw;
  const int ck = a.row_bytes / 2;
  const uint32_t tap_bytes = static_cast<uint32_t>(a.n_tile) * rb;
  const uint32_t b_stride = (tap_bytes + 1023u) & ~1023u;
  // padded-grid positions that hold outputs: per image, or over the whole batch in tall mode
  const int positions = f_tall ? a.n_images * a.oh * a.wp : a.oh * a.wp;
  const int bands_per_image = f_tall ? (positions + kBM * a.T - 1) / (kBM * a.T) : a.bands_per_image;
  const int total_bands = f_tall ? bands_per_image : bands_per_image * a.n_images;
  const int total_items = total_bands * a.n_tiles;          // work item = (band, N tile), N fastest

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_a_rem);
    ptx::prefetch_tmap(&tm_b);
    ptx::prefetch_tmap(&tm_bias);
    for (int i = 0; i < a.nstrips; ++i) {
      ptx::mbar_init(strip_full(i), kCp ? 64 : 1);
      ptx::mbar_init(strip_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), 256);
    }
    ptx::mbar_init(w_full, 1);
    for (int i = 0; i < a.b_stages; ++i) {
      ptx::mbar_init(b_full(i), 1);
      ptx::mbar_init(b_empty(i), 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp >= 2 && warp <= 5) {
    write_ones_row(gbase + (ones_base - base), static_cast<int>(threadIdx.x) - 64, a.fmt);
    ptx::fence_proxy_async_smem();
  }
  if (!a.pdl_late) ptx::pdl_launch();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  ptx::pdl_wait();            // the predecessor's output (this layer's input) is complete from here on

  // Every role loop runs warp-converged and hands the uniform-datapath instructions (TMA, MMA, commit) to one
  // elected lane: under `if (lane == 0)` the compiler wraps each of them in an ELECT/branch loop, which cost this
  // kernel ~160 cycles per MMA (timeline of conv1: 72 MMAs in 11.6k cycles) against ~32 of tensor-pipe time.
  if (kCp && (warp == 0 || warp == 6)) {
    // ---------------- strip copiers (cp.async) ----------------
    if (warp == 6) {
      if (ptx::elect_one()) {
        const int n_row = (static_cast<int>(blockIdx.x) % a.n_tiles) * a.n_tile;
        ptx::mbar_expect_tx(w_full, static_cast<uint32_t>(a.chunks * taps) * tap_bytes + static_cast<uint32_t>(a.n_tile) * 32u);
        ptx::tma_load_2d(bias_base, &tm_bias, w_full, 0, n_row);
        for (int i = 0; i < a.chunks * taps; ++i) ptx::tma_load_2d(wbase + i * tap_bytes, &tm_b, w_full, i * ck, n_row);
      }
      __syncwarp();
    }
    // 64 threads; thread t owns 16-byte chunk j = t % cpr of pixels pl + ppi * k (k = 0, 1, ...): ppi = 64 / cpr is a multiple
    // of the swizzle period in pixels (8 for both row sizes), so the swizzled chunk slot is a per-thread constant and the
    // destination advances by exactly 1024 bytes per step.
    const int lt = (warp == 0 ? 0 : 32) + lane;
    const int cpr = a.row_bytes >> 4;
    const int ppi = 64 / cpr;
    const int j = lt % cpr, pl = lt / cpr;
    const int sw = (a.row_bytes == 64) ? ((pl >> 1) & 3) : (pl & 7);
    const uint32_t dst_off = static_cast<uint32_t>(pl) * rb + (static_cast<uint32_t>(j ^ sw) << 4);
    const int npix = a.strip_rows * a.wp;
    const int nit = (npix + ppi - 1) / ppi;
    const size_t pix_bytes = static_cast<size_t>(a.ldi) * 2;
    uint32_t s = 0, ph = 0;
    int prev = -1;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int band = item / a.n_tiles;
      const int img = band / bands_per_image;
      const int p0 = (band - img * bands_per_image) * kBM * a.T;
      const int gy0 = p0 / a.wp - a.pad_h;
      for (int c = 0; c < a.chunks; ++c) {
        const char* src0 = reinterpret_cast<const char*>(a.in) + static_cast<size_t>(img) * a.ih * a.iw * pix_bytes +
                           static_cast<size_t>(c) * rb + j * 16;
        ptx::mbar_wait(strip_empty(s), ph ^ 1u);
        uint32_t dst = strip0 + s * a.strip_bytes + dst_off;
        int row = 0, col = pl;
#pragma unroll 4
        for (int k = 0; k < nit; ++k) {
          const int gy = gy0 + row, gx = col - a.pad_w;
          const bool ok = static_cast<unsigned>(gy) < static_cast<unsigned>(a.ih) &&
                          static_cast<unsigned>(gx) < static_cast<unsigned>(a.iw) && row < a.strip_rows;
          const char* src = ok ? src0 + (static_cast<size_t>(gy) * a.iw + gx) * pix_bytes : src0;
          ptx::cp_async16_zfill(dst, src, ok ? 16u : 0u);
          dst += 1024u;
          col += ppi;
          if (col >= a.wp) { col -= a.wp; ++row; }
        }
        ptx::cp_async_commit();
        if (prev >= 0) {                       // the strip issued one step earlier has landed: publish it to the async proxy
          ptx::cp_async_wait_group<1>();
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(strip_full(prev));
        }
        prev = static_cast<int>(s);
        if (++s == static_cast<uint32_t>(a.nstrips)) { s = 0; ph ^= 1u; }
      }
    }
    if (prev >= 0) {
      ptx::cp_async_wait_group<0>();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(strip_full(prev));
    }
  } else if (warp == 0) {
    // ---------------- strip producer ----------------
    const uint32_t strip_tx = static_cast<uint32_t>(a.strip_rows) * a.wp * rb;
    const uint32_t box_bytes = static_cast<uint32_t>(a.box_rows) * a.wp * rb;
    uint32_t s = 0, ph = 0;                // ring position kept incrementally: no division in the loops
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int band = item / a.n_tiles;
      const int img = f_tall ? 0 : band / bands_per_image;
      const int p0 = (band - img * bands_per_image) * kBM * a.T;
      const int h_first = p0 / a.wp;
      for (int c = 0; c < a.chunks; ++c) {
        ptx::mbar_wait(strip_empty(s), ph ^ 1u);
        ICP_STAMP(a.dbg, 0, it, lane == 0);
        ++it;
        if (ptx::elect_one()) {
          // MEASURED (timeline of conv1 / conv2, one 7-row box of 149 x 64 B rows): 7050 cycles from issue to landing =
          // one row per 6.8 cycles = ~9 B/cycle — one box keeps too few requests in flight; split, the boxes overlap.
          ptx::mbar_expect_tx(strip_full(s), strip_tx);
          const uint32_t dst = strip0 + s * a.strip_bytes;
          int h = h_first - a.pad_h;
          for (int b = 0; b < a.nbox; ++b, h += a.box_rows)
            ptx::tma_load_4d(dst + static_cast<uint32_t>(b) * box_bytes, &tm_a, strip_full(s), c * ck, -a.pad_w, h, img);
          if (a.rem_rows) ptx::tma_load_4d(dst + static_cast<uint32_t>(a.nbox) * box_bytes, &tm_a_rem, strip_full(s), c * ck, -a.pad_w, h, img);
        }
        __syncwarp();
        if (++s == static_cast<uint32_t>(a.nstrips)) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 6) {
    // ---------------- weight producer ----------------
    if (f_resident) {
      if (ptx::elect_one()) {
        const int n_row = (static_cast<int>(blockIdx.x) % a.n_tiles) * a.n_tile;     // this CTA's N tile (static)
        ptx::mbar_expect_tx(w_full, static_cast<uint32_t>(a.chunks * taps) * tap_bytes + static_cast<uint32_t>(a.n_tile) * 32u);
        ptx::tma_load_2d(bias_base, &tm_bias, w_full, 0, n_row);
        for (int i = 0; i < a.chunks * taps; ++i) ptx::tma_load_2d(wbase + i * tap_bytes, &tm_b, w_full, i * ck, n_row);
      }
      __syncwarp();
    } else {
      uint32_t sb = 0, phb = 0;
      int n_idx = blockIdx.x % a.n_tiles;
      const int n_step = gridDim.x % a.n_tiles;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int n_base = n_idx * a.n_tile;
        n_idx += n_step;
        if (n_idx >= a.n_tiles) n_idx -= a.n_tiles;
        for (int i = -1; i < a.chunks * taps; ++i) {       // i = -1: the shift tile of the bias MMA
          ptx::mbar_wait(b_empty(sb), phb ^ 1u);
          if (ptx::elect_one()) {
            if (i < 0) {
              ptx::mbar_expect_tx(b_full(sb), static_cast<uint32_t>(a.n_tile) * 32u);
              ptx::tma_load_2d(wbase + sb * b_stride, &tm_bias, b_full(sb), 0, n_base);
            } else {
              ptx::mbar_expect_tx(b_full(sb), tap_bytes);
              ptx::tma_load_2d(wbase + sb * b_stride, &tm_b, b_full(sb), i * ck, n_base);
            }
          }
          __syncwarp();
          if (++sb == static_cast<uint32_t>(a.b_stages)) { sb = 0; phb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // Descriptors differ only in their 14-bit address field (16-byte units), advanced with a per-tap table.
    const uint32_t idesc = ptx::make_idesc_f16(a.fmt, kBM, a.n_tile);
    const int ksteps = a.row_bytes / 32;
    const uint64_t desc_hi = ptx::make_smem_desc_rows(0, rb);          // everything but the address
    const uint64_t desc32 = ptx::make_smem_desc_rows(0, 32);
    const uint64_t d_ones = desc32 | (ones_base >> 4);
    for (int tap = lane; tap < taps; tap += 32) {
      const int r = tap / a.kw;
      tap_off[tap] = (static_cast<uint32_t>(r * a.wp + (tap - r * a.kw)) * rb) >> 4;
    }
    __syncwarp();
    const uint32_t tap_units = tap_bytes >> 4;
    const uint32_t sub_units = (static_cast<uint32_t>(kBM) * rb) >> 4;
    if (f_resident) ptx::mbar_wait(w_full, 0);
    uint32_t s = 0, ph = 0, sb = 0, phb = 0, as = 0, pha = 0;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int band = item / a.n_tiles;
      const int img = f_tall ? 0 : band / bands_per_image;
      const int p0 = (band - img * bands_per_image) * kBM * a.T;
      const int h_first = p0 / a.wp;
      const uint32_t local_units = (static_cast<uint32_t>(p0 - h_first * a.wp) * rb) >> 4;
      int nsub = (positions - p0 + kBM - 1) / kBM;                       // sub-tiles that hold outputs
      if (nsub > a.T) nsub = a.T;
      ptx::mbar_wait(tmem_empty(as), pha ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tacc0 = tmem_base + as * a.acc_cols;
      // every sub-tile's accumulator starts at the BatchNorm shift: ones tile x (hi, lo) shift tile, K = 16
      if (f_resident) {
        if (ptx::elect_one())
          for (int j = 0; j < nsub; ++j) ptx::umma_f16(tacc0 + j * a.n_tile, d_ones, desc32 | (bias_base >> 4), idesc, 0u);
        __syncwarp();
      } else {
        ptx::mbar_wait(b_full(sb), phb);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          for (int j = 0; j < nsub; ++j)
            ptx::umma_f16(tacc0 + j * a.n_tile, d_ones, desc32 | ((wbase + sb * b_stride) >> 4), idesc, 0u);
          ptx::umma_commit(b_empty(sb));
        }
        __syncwarp();
        if (++sb == static_cast<uint32_t>(a.b_stages)) { sb = 0; phb ^= 1u; }
      }
      for (int c = 0; c < a.chunks; ++c) {
        ptx::mbar_wait(strip_full(s), ph);
        ptx::tc_fence_after();
        ICP_STAMP(a.dbg, 2, it, lane == 0);
        const uint32_t a_units0 = ((strip0 + s * a.strip_bytes) >> 4) + local_units;
        if (f_resident) {
          if (ptx::elect_one()) {
            uint32_t b_units = (wbase >> 4) + static_cast<uint32_t>(c * taps) * tap_units;
            const int ks = (c == a.chunks - 1) ? a.tail_ksteps : ksteps;
            if (ks == 2 && ksteps == 2 && nsub == 4) {
              // the stem layers' full band (64-byte rows, four sub-tiles): eight MMAs per tap in straight-line code.
              // MEASURED (timeline, conv1 / conv2): 69 cycles per MMA through the generic loops below against the ~50 the
              // issuing thread manages without a loop branch between MMAs.
              const uint32_t nt = static_cast<uint32_t>(a.n_tile);
              for (int tap = 0; tap < taps; ++tap, b_units += tap_units) {
                const uint32_t au = a_units0 + tap_off[tap];
                const uint64_t db0 = desc_hi | b_units, db1 = desc_hi | (b_units + 2u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  ptx::umma_f16(tacc0 + j * nt, desc_hi | (au + j * sub_units), db0, idesc, 1u);
                  ptx::umma_f16(tacc0 + j * nt, desc_hi | (au + j * sub_units + 2u), db1, idesc, 1u);
                }
              }
            } else
            for (int tap = 0; tap < taps; ++tap, b_units += tap_units) {
              uint32_t au = a_units0 + tap_off[tap];
              uint32_t tacc = tacc0;
              for (int j = 0; j < nsub; ++j, au += sub_units, tacc += a.n_tile) {
                if (ks != ksteps) {
                  // tail chunk: only the K slices that hold channels (unrolled: a counted loop costs the issuing thread ~100
                  // cycles per MMA, MEASURED on the 5x5 48->64 layer, against ~50 for straight-line code)
                  ptx::umma_f16(tacc, desc_hi | au, desc_hi | b_units, idesc, 1u);
                  if (ks > 1) ptx::umma_f16(tacc, desc_hi | (au + 2u), desc_hi | (b_units + 2u), idesc, 1u);
                  if (ks > 2) ptx::umma_f16(tacc, desc_hi | (au + 4u), desc_hi | (b_units + 4u), idesc, 1u);
                } else if (ksteps == 2) {
                  ptx::umma_f16(tacc, desc_hi | au, desc_hi | b_units, idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 2u), desc_hi | (b_units + 2u), idesc, 1u);
                } else {
                  ptx::umma_f16(tacc, desc_hi | au, desc_hi | b_units, idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 2u), desc_hi | (b_units + 2u), idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 4u), desc_hi | (b_units + 4u), idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 6u), desc_hi | (b_units + 6u), idesc, 1u);
                }
              }
            }
            ptx::umma_commit(strip_empty(s));
          }
          __syncwarp();
        } else {
          for (int tap = 0; tap < taps; ++tap) {
            ptx::mbar_wait(b_full(sb), phb);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t b_units = (wbase + sb * b_stride) >> 4;
              uint32_t au = a_units0 + tap_off[tap];
              uint32_t tacc = tacc0;
              for (int j = 0; j < nsub; ++j, au += sub_units, tacc += a.n_tile) {
                if (ksteps == 2) {
                  ptx::umma_f16(tacc, desc_hi | au, desc_hi | b_units, idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 2u), desc_hi | (b_units + 2u), idesc, 1u);
                } else {
                  ptx::umma_f16(tacc, desc_hi | au, desc_hi | b_units, idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 2u), desc_hi | (b_units + 2u), idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 4u), desc_hi | (b_units + 4u), idesc, 1u);
                  ptx::umma_f16(tacc, desc_hi | (au + 6u), desc_hi | (b_units + 6u), idesc, 1u);
                }
              }
              ptx::umma_commit(b_empty(sb));
              if (tap == taps - 1) ptx::umma_commit(strip_empty(s));
            }
            __syncwarp();
            if (++sb == static_cast<uint32_t>(a.b_stages)) { sb = 0; phb ^= 1u; }
          }
        }
        ICP_STAMP(a.dbg, 3, it, lane == 0);
        ++it;
        if (++s == static_cast<uint32_t>(a.nstrips)) { s = 0; ph ^= 1u; }
      }
      if (ptx::elect_one()) ptx::umma_commit(tmem_full(as));
      __syncwarp();
      if (++as == static_cast<uint32_t>(a.acc_sets)) { as = 0; pha ^= 1u; }
    }
    if (a.pdl_late) ptx::pdl_launch();
  } else {
    // ---------------- epilogue (warps 2..5 and 7..10) ----------------
    // TWO independent teams of four warps (one warp per TMEM lane quarter each).  Units = (sub-tile, <= 64-column chunk)
    // are dealt to the teams alternately by a running counter; each team has its own staging buffer, row->pixel table and
    // named barrier, so two dependent chains (TMEM wait -> convert -> staging -> copy-out) are in flight at once.
    // MEASURED before (one 8-warp team, timeline of conv2): ~1600 cycles per 128 x 64 sub-tile, 6500-7100 per band against
    // 3870 of MMA issue: conv1 / conv2 were bound by this chain, not by the tensor pipe or HBM.
    // TMEM -> registers -> ReLU -> 16-bit -> (kDirect: 16-byte stores straight from registers | else: shared staging,
    // 128 rows x 128 B, chunks XOR-swizzled by row -> global with 8 consecutive lanes per pixel = whole 128-byte lines).
    const int team = warp >= 7 ? 1 : 0;
    const int tt = team ? static_cast<int>(threadIdx.x) - 224 : static_cast<int>(threadIdx.x) - 64;   // 0..127 in the team
    const int et = tt + 128 * team;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t bar_id = 1u + static_cast<uint32_t>(team);
    uint32_t as = 0, pha = 0, gunit = 0;
    int lt = 0;
    uint8_t* stage_g = gbase + (stage_base - base) + team * 16384;
    int* my_pix = s_pix + team * 128;
    const int cl = tt & 7;                     // chunk lane of the copy-out: 8 lanes cover one pixel's 128 bytes
    const int r0 = tt >> 3;                    // copy-out rows r0, r0+16, ... r0+112
    const int cps = (a.n_tile + 63) >> 6;      // <= 64-column chunks per sub-tile
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++lt) {
      const int band = item / a.n_tiles;
      const int n_base = (item - band * a.n_tiles) * a.n_tile;
      const int img = f_tall ? 0 : band / bands_per_image;
      const int p0 = (band - img * bands_per_image) * kBM * a.T;
      int nsub = (positions - p0 + kBM - 1) / kBM;
      if (nsub > a.T) nsub = a.T;
      const int nunits = nsub * cps;
      ICP_STAMP(a.dbg, 6, lt, et == 0);
      ptx::mbar_wait(tmem_full(as), pha);
      ptx::tc_fence_after();
      ICP_STAMP(a.dbg, 4, lt, et == 0);
      const uint32_t tbase = tmem_base + as * a.acc_cols + (static_cast<uint32_t>(quarter * 32) << 16);
      int last_mine = -1;
      if constexpr (kDirect) {
        last_mine = nunits - 1;
      } else {
        for (int u = nunits - 1; u >= 0; --u)
          if (((gunit + static_cast<uint32_t>(u)) & 1u) == static_cast<uint32_t>(team)) { last_mine = u; break; }
      }
      if (last_mine < 0) {                       // nothing of this band is mine: hand the accumulator back at once
        ptx::tc_fence_before();
        ptx::mbar_arrive(tmem_empty(as));
      }
      for (int u = 0; u < nunits; ++u) {
        if constexpr (!kDirect) {
          if (((gunit + static_cast<uint32_t>(u)) & 1u) != static_cast<uint32_t>(team)) continue;
        }
        const int j = u / cps;
        int c0 = (u - j * cps) << 6;
        int cp = (a.n_tile - c0) >> 4;           // 16-column pieces in this chunk (<= 4)
        if (cp > 4) cp = 4;
        if constexpr (kDirect) {
          // no barriers in this variant, so the teams bring nothing: both warps of a lane quarter work on the SAME unit and
          // split its pieces (MEASURED conv1, 32 columns: 0.111 ms split against 0.120 ms with alternating sub-tiles)
          const int h0 = (cp + 1) >> 1;
          c0 += team ? 16 * h0 : 0;
          cp = team ? cp - h0 : h0;
        }
        uint32_t v[64];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < cp) ptx::tmem_ld_32x32b_x16(tbase + j * a.n_tile + c0 + 16 * q, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * q]));
        // this thread's row -> output pixel (while the loads are in flight)
        const int p = p0 + j * kBM + row;
        int oy = p / a.wp;                       // tall mode: row over all images
        const int ox = p - oy * a.wp;
        int im = img;
        if (f_tall) { im = oy / a.oh; oy -= im * a.oh; }
        int pix = -1;
        if ((im < a.n_images) && (oy < a.oh) && (ox < a.ow))
          pix = f_transpose ? (im * a.ow + ox) * a.oh + oy : (im * a.oh + oy) * a.ow + ox;
        ptx::tmem_ld_wait();
        ICP_STAMP(a.dbg, 7, lt * a.T + j, et == 0 && c0 == 0);
        if (u == last_mine) {                    // last TMEM read of this band (for this team) has landed
          ptx::tc_fence_before();
          ptx::mbar_arrive(tmem_empty(as));
        }
        if constexpr (kDirect) {
          if (pix >= 0) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) + static_cast<size_t>(pix) * a.ldo + a.coff +
                                                  n_base + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < cp) {
                uint32_t pk[8];
                convert16(&v[16 * q], pk, a.relu != 0, a.fmt);
                dst[2 * q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[2 * q + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
          ICP_STAMP(a.dbg, 9, lt * a.T + j, et == 0);
        } else {
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");       // staging free (this team's previous copy-out finished)
          my_pix[row] = pix;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < cp) {
              uint32_t pk[8];
              convert16(&v[16 * q], pk, a.relu != 0, a.fmt);
              *reinterpret_cast<uint4*>(stage_g + row * 128 + (((2 * q) ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(stage_g + row * 128 + (((2 * q + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          ICP_STAMP(a.dbg, 8, lt * a.T + j, et == 0);
          if (cl < 2 * cp) {                         // lane's 8 columns exist in this chunk
            uint16_t* dst = reinterpret_cast<uint16_t*>(a.out) + a.coff + n_base + c0 + cl * 8;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              int px[4];
              uint4 val[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) px[i] = my_pix[r0 + 16 * (4 * h + i)];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = r0 + 16 * (4 * h + i);
                val[i] = *reinterpret_cast<const uint4*>(stage_g + r * 128 + ((cl ^ (r & 7)) << 4));
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (px[i] >= 0) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(px[i]) * a.ldo) = val[i];
            }
          }
          ICP_STAMP(a.dbg, 9, lt * a.T + j, et == 0);
        }
      }
      gunit += static_cast<uint32_t>(nunits);
      ICP_STAMP(a.dbg, 5, lt, et == 0);
      if (++as == static_cast<uint32_t>(a.acc_sets)) { as = 0; pha ^= 1u; }
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

inline size_t halo_smem_bytes(int strip_bytes, int nstrips, int w_bytes, bool direct = false, int n_tile = 256) {
  // alignment slack, strips, weights (resident or ring), barriers + per-tap offsets + row->pixel tables, 2 x epilogue staging
  // (staged variant only), ones tile, shift tile of the resident-weights bias MMA ([n_tile x 32 B])
  return 1024 + static_cast<size_t>(nstrips) * strip_bytes + ((static_cast<size_t>(w_bytes) + 1023) & ~size_t(1023)) + 2048 +
         (direct ? 0 : 2 * 16384) + kOnesBytes + ((static_cast<size_t>(n_tile) * 32 + 1023) & ~size_t(1023));
}

}  // namespace icp
