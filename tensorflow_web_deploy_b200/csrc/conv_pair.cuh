// CTA-pair, weight-resident, halo-strip implicit GEMM for the stride-1 k x k (k > 1) convolutions of the 35x35 and
// 73x73 stages (ops.conv2d, /root/reference/inception/slim/ops.py:229-245: conv, folded BatchNorm, ReLU).
//
// Why (profiles/r02_pair_kernel.md): at batch 128 the tap-wise im2col kernel is bound by bytes crossing L2 -> SM
// (~45 B/cycle/SM, the chip-wide LTS cap): every input tile crosses once per filter tap and every weight tile once
// per 128-pixel tile.  The r01 halo kernel removed the per-tap re-read of the input but still STREAMED the weights once
// per band, and with the 512 TMEM columns allowing only T = 2 sub-tiles per band at 96 output channels that stream
// alone needs ~41 B/cycle/SM (timeline: 780 cycles per tap against 384 of MMA).  Here the weights never move:
//
//  * Two CTAs of a cluster form a pair (tcgen05 cta_group::2): one MMA instruction covers M = 256 = the 128-row
//    sub-tiles of both CTAs and N = n_tile, and each CTA supplies only HALF of the B operand (n_tile/2 weight rows).
//    Half the weights of a layer fit beside the strips in 227 KB of shared memory for every k x k layer of these
//    stages (96->96 3x3: 83 KB, 48->64 5x5: 102 KB, 80->192 3x3: 69 KB per N tile), so they are loaded ONCE per CTA
//    and stay resident for the life of the persistent kernel.  The only stream is the input strip.
//  * The pair works on the SAME band of two consecutive images (CTA rank = image parity), so both CTAs' strips have the
//    same geometry and one A descriptor serves both.
//  * Input strips as in conv_halo.cuh: outputs live on a padded grid (wp = iw + 2 pad_w columns), the input pixel of
//    output position p under tap (r,s) is strip position p + r*wp + s, and the tap is selected by moving the A
//    descriptor's start address.  Bands are whole output rows (band_rows * wp <= T * 128 positions).  The channel tail
//    (Cin = 80 = 64 + 16, 96 = 64 + 32) travels in narrower rows (SWIZZLE_32B / 64B) so no zero K slices are multiplied.
//  * T accumulators x 2 sets in TMEM: the epilogue of band i (tcgen05.ld -> ReLU -> 16-bit -> stores straight from
//    registers into the NHWC channel slice) overlaps the MMAs of band i+1.
//
// Barriers: strip_full / w_full live in the LEADER (cluster rank 0): both CTAs' TMA loads complete_tx there
// (cp.async.bulk.tensor ... cta_group::2) and the leader expects twice the bytes.  strip_empty / tmem_full are local to
// each CTA and are signalled for both by the leader's tcgen05.commit ... multicast.  tmem_empty lives in the leader and
// collects the epilogue threads of both CTAs.
// Warp roles (per CTA): 0 strip producer, 1 TMEM owner (+ MMA issuer in the leader), 6 weight loader, 2..5 and 7..10 epilogue.
#pragma once
#include "conv_igemm.cuh"

namespace icp {

constexpr int kPairThreads = 352;

struct PairArgs {
  int n_images;
  int oh, ow;           // valid output size (per image)
  int wp;               // padded grid width = iw + 2 * pad_w
  int pad_h, pad_w, kh, kw;
  int chunks;           // K chunks per tap: chunks - 1 of 64 channels (128-byte rows) + the last one
  int tail_rb;          // row bytes of the LAST chunk: 32 / 64 / 128
  int tail_ksteps;      // K = 16 MMAs issued for the last chunk
  int tail_split;       // 1: the last chunk has its own (narrow-row) strip / weight tensor maps
  int T;                // 128-position sub-tiles per band
  int band_rows;        // output rows per band
  int bands_per_image;
  int strip_rows;       // band_rows + kh - 1
  int chunk_pitch;      // bytes between the chunks of one band inside a slot (128-byte-row chunks; multiple of 1024)
  int slot_bytes;       // strip ring slot pitch: ALL chunks of one band (multiple of 1024)
  int nstrips;
  int n_tile, n_tiles;
  int acc_cols;         // T * n_tile
  int acc_sets;
  int tmem_cols;
  int w_main_bytes;     // per CTA: (chunks - tail_split) * taps * (n_tile / 2) * 128
  int w_bytes;          // w_main_bytes + tail tiles, rounded to 1024
  int fmt, relu;
  void* out;
  int ldo, coff;
  long long* dbg;
};

__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & ptx::kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a_tail,
                 const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b_tail,
                 const __grid_constant__ CUtensorMap tm_bias, const PairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;
  const int taps = a.kh * a.kw;
  const int half = a.n_tile >> 1;                       // weight rows this CTA supplies
  const uint32_t trb = static_cast<uint32_t>(a.tail_rb);

  const uint32_t strip0 = base;
  const uint32_t wbase = base + static_cast<uint32_t>(a.nstrips) * a.slot_bytes;
  const uint32_t wtail = wbase + static_cast<uint32_t>(a.w_main_bytes);
  const uint32_t bias_base = wbase + static_cast<uint32_t>(a.w_bytes);        // [half x 16] shift tile (32-byte rows)
  const uint32_t ones_base = bias_base + 4096u;
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto strip_full = [&](int i) { return bar_base + 8u * i; };
  auto strip_empty = [&](int i) { return bar_base + 32u + 8u * i; };
  auto tmem_full = [&](int i) { return bar_base + 64u + 8u * i; };
  auto tmem_empty = [&](int i) { return bar_base + 80u + 8u * i; };
  const uint32_t w_full = bar_base + 96u;
  const uint32_t tmem_slot = bar_base + 104u;
  uint32_t* tap_pos = reinterpret_cast<uint32_t*>(gbase + (bar_base - base) + 128);   // [taps] r * wp + s

  // work: this cluster serves N tile `my_nt`; items = (image pair, band) taken round-robin by the clusters of that N tile
  const int cid = static_cast<int>(blockIdx.x) >> 1;
  const int nclusters = static_cast<int>(gridDim.x) >> 1;
  const int my_nt = cid % a.n_tiles;
  const int first = cid / a.n_tiles;
  const int step = nclusters / a.n_tiles;
  const int n_base = my_nt * a.n_tile;
  const int pairs = (a.n_images + 1) >> 1;
  const int total_items = pairs * a.bands_per_image;

  ptx::pdl_launch();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_a_tail);
    ptx::prefetch_tmap(&tm_b);
    ptx::prefetch_tmap(&tm_b_tail);
    ptx::prefetch_tmap(&tm_bias);
    for (int i = 0; i < a.nstrips; ++i) {
      ptx::mbar_init(strip_full(i), 1);          // the leader's producer arrives (expect_tx); bytes of both CTAs complete here
      ptx::mbar_init(strip_empty(i), 1);         // the leader's commit (multicast) arrives once in each CTA
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tmem_full(i), 1);
      ptx::mbar_init(tmem_empty(i), 2 * kEpiThreads);    // every epilogue thread of both CTAs arrives (in the leader)
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_slot, a.tmem_cols);
    ptx::tmem_relinquish_2sm();
    for (int tap = lane; tap < taps; tap += 32) {
      const int r = tap / a.kw;
      tap_pos[tap] = static_cast<uint32_t>(r * a.wp + (tap - r * a.kw));
    }
  }
  if (warp >= 2 && warp <= 5) {
    write_ones_row(gbase + (ones_base - base), static_cast<int>(threadIdx.x) - 64, a.fmt);
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();          // the peer's barriers exist before anything is signalled at them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == 6) {
    // ---------------- weight loader: this CTA's half of every (chunk, tap) tile + its half of the shift tile, once ----------------
    // (weights are constants: they do not wait for the predecessor kernel)
    if (ptx::elect_one()) {
      if (leader) ptx::mbar_expect_tx(w_full, 2u * (static_cast<uint32_t>(a.w_main_bytes) + static_cast<uint32_t>(a.tail_split ? taps * half * a.tail_rb : 0) +
                                                  static_cast<uint32_t>(half) * 32u));
      const int row0 = n_base + static_cast<int>(crank) * half;
      ptx::tma_load_2d_2sm(bias_base, &tm_bias, w_full, 0, row0);
      const int main_tiles = (a.chunks - a.tail_split) * taps;
      for (int i = 0; i < main_tiles; ++i)
        ptx::tma_load_2d_2sm(wbase + static_cast<uint32_t>(i) * half * 128u, &tm_b, w_full, i * 64, row0);
      if (a.tail_split)
        for (int t = 0; t < taps; ++t)
          ptx::tma_load_2d_2sm(wtail + static_cast<uint32_t>(t) * half * trb, &tm_b_tail, w_full, t * (a.tail_rb >> 1), row0);
    }
    __syncwarp();
  }
  ptx::pdl_wait();              // the predecessor's output (this layer's input) is complete from here on

  if (warp == 0) {
    // ---------------- strip producer (both CTAs; completion is counted on the leader's barrier) ----------------
    uint32_t s = 0, ph = 0;
    int it = 0;
    for (int item = first; item < total_items; item += step) {
      const int ip = item / a.bands_per_image;
      const int band = item - ip * a.bands_per_image;
      const int img = 2 * ip + static_cast<int>(crank);       // >= n_images for the odd tail: the box is zero-filled
      const int y0 = band * a.band_rows;
      // one slot = every channel chunk of the band; one barrier phase per band
      ptx::mbar_wait(strip_empty(s), ph ^ 1u);
      ICP_STAMP(a.dbg, 0, it, lane == 0);
      ++it;
      if (ptx::elect_one()) {
        const uint32_t pos_bytes = static_cast<uint32_t>(a.chunks - a.tail_split) * 128u + (a.tail_split ? trb : 0u);
        if (leader) ptx::mbar_expect_tx(strip_full(s), 2u * static_cast<uint32_t>(a.strip_rows) * a.wp * pos_bytes);
        const uint32_t slot = strip0 + s * a.slot_bytes;
        for (int c = 0; c < a.chunks; ++c) {
          const bool tail = a.tail_split && c == a.chunks - 1;
          tma_load_4d_2sm(slot + static_cast<uint32_t>(c) * a.chunk_pitch, tail ? &tm_a_tail : &tm_a, strip_full(s), c * 64, -a.pad_w,
                          y0 - a.pad_h, img);
        }
      }
      __syncwarp();
      if (++s == static_cast<uint32_t>(a.nstrips)) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    if (leader) {
      // ---------------- MMA issuer (the leader issues for the pair) ----------------
      const uint32_t idesc = ptx::make_idesc_f16(a.fmt, 2 * kBM, a.n_tile);
      const uint64_t d128 = ptx::make_smem_desc_rows(0, 128);
      const uint64_t dtail = ptx::make_smem_desc_rows(0, trb);
      const uint64_t d32 = ptx::make_smem_desc_rows(0, 32);
      const uint64_t d_ones = d32 | ((ones_base & 0x3FFFFu) >> 4);
      const uint64_t d_bias = d32 | ((bias_base & 0x3FFFFu) >> 4);
      ptx::mbar_wait(w_full, 0);                     // both CTAs' weights have landed (also when there is no item: exit safety)
      ptx::tc_fence_after();
      uint32_t s = 0, ph = 0, as = 0, pha = 0;
      int it = 0;
      for (int item = first; item < total_items; item += step) {
        const int ip = item / a.bands_per_image;
        const int band = item - ip * a.bands_per_image;
        const int y0 = band * a.band_rows;
        const int rows = min(a.band_rows, a.oh - y0);
        int nsub = (rows * a.wp + kBM - 1) / kBM;            // sub-tiles that hold outputs
        if (nsub > a.T) nsub = a.T;
        ptx::mbar_wait(tmem_empty(as), pha ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tacc0 = tmem_base + as * a.acc_cols;
        if (ptx::elect_one())
          for (int j = 0; j < nsub; ++j) ptx::umma_f16_2sm(tacc0 + j * a.n_tile, d_ones, d_bias, idesc, 0u);
        __syncwarp();
        ptx::mbar_wait(strip_full(s), ph);              // every chunk of both CTAs' strips has landed
        ptx::tc_fence_after();
        ICP_STAMP(a.dbg, 2, it, lane == 0);
        if (ptx::elect_one()) {
          for (int c = 0; c < a.chunks; ++c) {
            const bool last = c == a.chunks - 1;
            const bool tail = a.tail_split && last;
            const uint32_t rbc = tail ? trb : 128u;
            const uint64_t dh = tail ? dtail : d128;
            const int ksteps = last ? a.tail_ksteps : 4;
            const uint32_t unit = rbc >> 4;                               // 16-byte units per position
            const uint32_t sub_units = static_cast<uint32_t>(kBM) * unit;
            const uint32_t b_tile_units = (static_cast<uint32_t>(half) * rbc) >> 4;
            uint32_t b_units = tail ? ((wtail & 0x3FFFFu) >> 4)
                                    : (((wbase & 0x3FFFFu) >> 4) + static_cast<uint32_t>(c * taps) * b_tile_units);
            const uint32_t a_units0 = ((strip0 + s * a.slot_bytes + static_cast<uint32_t>(c) * a.chunk_pitch) & 0x3FFFFu) >> 4;
            for (int tap = 0; tap < taps; ++tap, b_units += b_tile_units) {
              uint32_t au = a_units0 + tap_pos[tap] * unit;
              uint32_t tacc = tacc0;
              for (int j = 0; j < nsub; ++j, au += sub_units, tacc += a.n_tile) {
                if (ksteps == 4) {
                  ptx::umma_f16_2sm(tacc, dh | au, dh | b_units, idesc, 1u);
                  ptx::umma_f16_2sm(tacc, dh | (au + 2u), dh | (b_units + 2u), idesc, 1u);
                  ptx::umma_f16_2sm(tacc, dh | (au + 4u), dh | (b_units + 4u), idesc, 1u);
                  ptx::umma_f16_2sm(tacc, dh | (au + 6u), dh | (b_units + 6u), idesc, 1u);
                } else {
                  for (int k = 0; k < ksteps; ++k) ptx::umma_f16_2sm(tacc, dh | (au + 2u * k), dh | (b_units + 2u * k), idesc, 1u);
                }
              }
            }
          }
          ptx::umma_commit_2sm(strip_empty(s), static_cast<uint16_t>(3));     // slot s is free in both CTAs
        }
        __syncwarp();
        ICP_STAMP(a.dbg, 3, it, lane == 0);
        ++it;
        if (++s == static_cast<uint32_t>(a.nstrips)) { s = 0; ph ^= 1u; }
        if (ptx::elect_one()) ptx::umma_commit_2sm(tmem_full(as), static_cast<uint16_t>(3));   // both CTAs' epilogues
        __syncwarp();
        if (++as == static_cast<uint32_t>(a.acc_sets)) { as = 0; pha ^= 1u; }
      }
    }
  } else if (warp != 6) {
    // ---------------- epilogue (warps 2..5 and 7..10 of BOTH CTAs; each reads its own CTA's TMEM half) ----------------
    // Two warps per TMEM lane quarter, each owning half of the 16-column pieces; the loads of the next unit are in flight
    // while this one is converted and stored straight from registers (16-byte stores into the NHWC channel slice).
    const int grp = warp >= 7 ? 1 : 0;
    const int et = grp ? static_cast<int>(threadIdx.x) - 224 + 128 : static_cast<int>(threadIdx.x) - 64;   // 0..255
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int pieces = a.n_tile >> 4;
    const int h0 = (pieces + 1) >> 1;
    const int first_piece = grp ? h0 : 0;
    const int my_pieces = grp ? pieces - h0 : h0;          // <= 8
    uint32_t as = 0, pha = 0;
    int lt = 0;
    for (int item = first; item < total_items; item += step, ++lt) {
      const int ip = item / a.bands_per_image;
      const int band = item - ip * a.bands_per_image;
      const int img = 2 * ip + static_cast<int>(crank);
      const int y0 = band * a.band_rows;
      const int rows = min(a.band_rows, a.oh - y0);
      int nsub = (rows * a.wp + kBM - 1) / kBM;
      if (nsub > a.T) nsub = a.T;
      ICP_STAMP(a.dbg, 6, lt, et == 0);
      ptx::mbar_wait(tmem_full(as), pha);
      ptx::tc_fence_after();
      ICP_STAMP(a.dbg, 4, lt, et == 0);
      const uint32_t tbase = tmem_base + as * a.acc_cols + (static_cast<uint32_t>(quarter * 32) << 16) + 16u * first_piece;
      // units = (sub-tile j, pair of pieces q): two 16-column pieces per unit, next unit's loads issued before this one's stores
      const int upj = (my_pieces + 1) >> 1;
      const int nunits = nsub * upj;
      auto issue = [&](uint32_t (&v)[32], int u) {
        const int j = u / upj, q = u - j * upj;
        const uint32_t t0 = tbase + j * a.n_tile + 32u * q;
        ptx::tmem_ld_32x32b_x16(t0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        if (2 * q + 1 < my_pieces) ptx::tmem_ld_32x32b_x16(t0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      };
      auto store = [&](uint32_t (&v)[32], int u) {
        const int j = u / upj, q = u - j * upj;
        const int p = j * kBM + row;                     // position inside the band
        const int ry = p / a.wp;
        const int ox = p - ry * a.wp;
        const bool valid = (img < a.n_images) && (ry < rows) && (ox < a.ow);
        uint32_t pk[16];
        convert16(&v[0], &pk[0], a.relu != 0, a.fmt);
        const bool two = 2 * q + 1 < my_pieces;
        if (two) convert16(&v[16], &pk[8], a.relu != 0, a.fmt);
        if (valid) {
          const size_t pix = (static_cast<size_t>(img) * a.oh + (y0 + ry)) * a.ow + ox;
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) + pix * a.ldo + a.coff + n_base +
                                                16 * (first_piece + 2 * q));
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if (two) {
            dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            dst[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          }
        }
      };
      uint32_t va[32], vb[32];
      if (nunits > 0) issue(va, 0);
      for (int u = 0; u < nunits; u += 2) {
        ptx::tmem_ld_wait();
        if (u + 1 < nunits) issue(vb, u + 1);
        else { ptx::tc_fence_before(); ptx::mbar_arrive_leader(tmem_empty(as)); }   // last TMEM read has landed
        store(va, u);
        if (u + 1 < nunits) {
          ptx::tmem_ld_wait();
          if (u + 2 < nunits) issue(va, u + 2);
          else { ptx::tc_fence_before(); ptx::mbar_arrive_leader(tmem_empty(as)); }
          store(vb, u + 1);
        }
      }
      if (nunits == 0) { ptx::tc_fence_before(); ptx::mbar_arrive_leader(tmem_empty(as)); }
      ICP_STAMP(a.dbg, 5, lt, et == 0);
      if (++as == static_cast<uint32_t>(a.acc_sets)) { as = 0; pha ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();          // no CTA leaves while its peer can still write its smem / barriers / TMEM
  if (warp == 1) {
    ptx::tc_fence_after();
    __syncwarp();
    ptx::tmem_dealloc_2sm(tmem_base, a.tmem_cols);
  }
}

inline size_t pair_smem_bytes(int slot_bytes, int nstrips, int w_bytes) {
  return 1024 + static_cast<size_t>(nstrips) * slot_bytes + w_bytes + 4096 + kOnesBytes + 1024;
}

}  // namespace icp
