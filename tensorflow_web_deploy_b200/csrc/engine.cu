// libinception_b200.so — engine, tower plan and C ABI (include/inception_b200.h).
//
// The tower topology below restates /root/reference/inception/slim/inception_model.py:86-329 as a
// static launch plan: every conv is one tcgen05 implicit-GEMM launch writing its channel slice of the
// block's NHWC output (no concat kernels), pools and the head are HBM-bound kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../../include/inception_b200.h"
#include "conv_igemm.cuh"
#include "conv_halo.cuh"
#include "conv_pair.cuh"
#include "pool_conv.cuh"
#include "conv0_tc.cuh"
#include "tower_kernels.cuh"

namespace icp {

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_tls_error;

struct Fail {
  int code;
  std::string msg;
};
[[noreturn]] static void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Fail{code, buf};
}
#define CK(expr)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) fail(ICP_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                __FILE__, __LINE__);                                                \
  } while (0)

// ------------------------------------------------------------------------------------------ driver API
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
using EncodeIm2colFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static void load_driver_api() {
  if (g_encode_tiled && g_encode_im2col) return;
  cudaDriverEntryPointQueryResult q;
  void* f = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
  if (!f || q != cudaDriverEntryPointSuccess) fail(ICP_E_CUDA, "cuTensorMapEncodeTiled not available");
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f);
  f = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q));
  if (!f || q != cudaDriverEntryPointSuccess) fail(ICP_E_CUDA, "cuTensorMapEncodeIm2col not available");
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f);
  CK(cudaDriverGetVersion(&g_driver_version));
}

static CUtensorMapDataType tm_dtype(int fmt) {
  return fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

// 2-D row-major [rows, cols] 16-bit matrix, box [box_rows x 64 cols], 128-B swizzle.
static CUtensorMapSwizzle tm_swizzle(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static CUtensorMap make_tm_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                              int fmt, int row_bytes = 128) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(row_bytes / 2), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, tm_dtype(fmt), 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, tm_swizzle(row_bytes),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail(ICP_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r,
         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows);
  return m;
}

// Output slice map for the TMA-store epilogue: [rows, ld] 16-bit matrix, box = 128 rows x 16 channels (32 B), no swizzle.
static CUtensorMap make_tm_out(const void* ptr, uint64_t rows, uint64_t ld_elems, int fmt, int box_cols = 16) {
  CUtensorMap m;
  cuuint64_t dims[2] = {ld_elems, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(kBM)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, tm_dtype(fmt), 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(ICP_E_CUDA, "cuTensorMapEncodeTiled(out) failed (%d) rows=%llu ld=%llu", (int)r,
                              (unsigned long long)rows, (unsigned long long)ld_elems);
  return m;
}

// 2-D K-major weight matrix with rows of `row_bytes` per box (halo kernel's B operand).
static CUtensorMap make_tm_2d_rows(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int row_bytes,
                                   int fmt) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(row_bytes / 2), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&m, tm_dtype(fmt), 2, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, tm_swizzle(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(ICP_E_CUDA, "cuTensorMapEncodeTiled(B rows) failed (%d)", (int)r);
  return m;
}

// NHWC activation as a 4-D tiled map whose box is a halo strip: {ck channels, wp columns, rows, 1 image}.
static CUtensorMap make_tm_strip(const void* ptr, int n, int h, int w, int c, int ld, int wp, int rows, int row_bytes,
                                 int fmt) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)wp, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(&m, tm_dtype(fmt), 4, const_cast<void*>(ptr), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, tm_swizzle(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail(ICP_E_CUDA, "cuTensorMapEncodeTiled(strip) failed (%d) n=%d h=%d w=%d c=%d ld=%d wp=%d rows=%d rb=%d", (int)r, n,
         h, w, c, ld, wp, rows, row_bytes);
  return m;
}

// NHWC activation [n,h,w,c] (channel stride ld) in im2col mode: 128 base pixels x 64 channels per load.
static CUtensorMap make_tm_im2col(const void* ptr, int n, int h, int w, int c, int ld, int kh, int kw, int stride,
                                  int pad_h, int pad_w, int fmt, int row_bytes = 128) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
  int lower[2] = {-pad_w, -pad_h};
  int upper[2] = {pad_w - (kw - 1), pad_h - (kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(&m, tm_dtype(fmt), 4, const_cast<void*>(ptr), dims, strides, lower, upper,
                               /*channelsPerPixel=*/row_bytes / 2, /*pixelsPerColumn=*/kBM, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, tm_swizzle(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail(ICP_E_CUDA, "cuTensorMapEncodeIm2col failed (%d) n=%d h=%d w=%d c=%d ld=%d k=%dx%d s=%d pad=%d,%d", (int)r,
         n, h, w, c, ld, kh, kw, stride, pad_h, pad_w);
  // Drivers <= 13.1 mis-set a descriptor bit for im2col maps over tensors smaller than 128 KiB (the
  // same correction CUTLASS applies in cute/atom/copy_traits_sm90_im2col.hpp).
  const size_t bytes = static_cast<size_t>(n) * h * w * ld * 2;
  if (g_driver_version <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  return m;
}

// ------------------------------------------------------------------------------------------ weights blob
struct HostTensor {
  std::vector<int> dims;
  std::vector<float> data;
};
using Blob = std::map<std::string, HostTensor>;

static Blob read_blob(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) fail(ICP_E_IO, "cannot open weight blob %s", path);
  std::unique_ptr<FILE, int (*)(FILE*)> guard(f, fclose);
  char magic[8];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "ICPWv001", 8) != 0) fail(ICP_E_IO, "%s: bad magic", path);
  uint32_t count = 0;
  if (fread(&count, 4, 1, f) != 1) fail(ICP_E_IO, "%s: truncated", path);
  Blob b;
  for (uint32_t i = 0; i < count; ++i) {
    uint16_t ln = 0;
    if (fread(&ln, 2, 1, f) != 1) fail(ICP_E_IO, "%s: truncated", path);
    std::string name(ln, '\0');
    if (fread(&name[0], 1, ln, f) != ln) fail(ICP_E_IO, "%s: truncated", path);
    uint8_t nd = 0;
    if (fread(&nd, 1, 1, f) != 1 || nd > 8) fail(ICP_E_IO, "%s: bad rank for %s", path, name.c_str());
    HostTensor t;
    size_t n = 1;
    for (int d = 0; d < nd; ++d) {
      uint32_t v = 0;
      if (fread(&v, 4, 1, f) != 1) fail(ICP_E_IO, "%s: truncated", path);
      t.dims.push_back(static_cast<int>(v));
      n *= v;
    }
    t.data.resize(n);
    if (fread(t.data.data(), 4, n, f) != n) fail(ICP_E_IO, "%s: truncated in %s", path, name.c_str());
    b.emplace(std::move(name), std::move(t));
  }
  return b;
}

static const HostTensor& blob_get(const Blob& b, const std::string& name, std::initializer_list<int> dims) {
  auto it = b.find(name);
  if (it == b.end()) fail(ICP_E_IO, "weight blob has no tensor '%s'", name.c_str());
  if (it->second.dims != std::vector<int>(dims)) {
    std::string got, want;
    for (int d : it->second.dims) got += std::to_string(d) + ",";
    for (int d : dims) want += std::to_string(d) + ",";
    fail(ICP_E_IO, "tensor '%s' has shape [%s] expected [%s]", name.c_str(), got.c_str(), want.c_str());
  }
  return it->second;
}

static uint16_t host_f2h(float f, int fmt) {
  if (fmt) {  // bf16 round-to-nearest-even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
  }
  __half h = __float2half_rn(f);
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
}

// ------------------------------------------------------------------------------------------ device memory
struct DevicePool {
  std::vector<void*> ptrs;
  size_t total = 0;
  void* alloc(size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) fail(ICP_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(p);
    total += bytes;
    return p;
  }
  ~DevicePool() {
    for (void* p : ptrs) cudaFree(p);
  }
};

struct Act {   // NHWC 16-bit activation for max_batch images
  uint16_t* p = nullptr;
  int h = 0, w = 0, c = 0;   // c = channels in the buffer (row pitch)
};

// ------------------------------------------------------------------------------------------ conv layer
struct ConvLayer {
  std::string name;
  CUtensorMap tm_a, tm_b;
  CUtensorMap tm_a_tail, tm_b_tail;   // last channel block of a tap when <= 32 channels remain: 32/64-byte-row boxes
  CUtensorMap tm_b_half, tm_b_tail_half;   // cluster mode: half of the weight tile's rows per box (multicast by each CTA)
  bool cluster = false;    // 2-CTA clusters sharing every weight tile by multicast (maps built; used for big launches)
  int cluster_min_tiles = 16;   // ... i.e. when m_tiles >= cluster_min_tiles * SMs
  int cluster_size = 2;         // CTAs per cluster of that mode (2 or 8)
  CUtensorMap tm_bias;       // [cout][16] (hi, lo) halves of the BatchNorm shift: B operand of the bias MMA
  CUtensorMap tm_o, tm_o2;   // destination slices (persistent kernel's TMA-store epilogue): 16-column boxes
  CUtensorMap tm_o64, tm_o64_2;   // same, 64-column (128-byte, swizzled) boxes
  ConvArgs args{};
  int cout = 0;
  int n_tiles = 1;
  int pix_per_image = 0;   // oh * ow
  size_t smem = 0;
  long long macs_per_image = 0;
  long long in_elems_per_image = 0;
  bool persistent = true;  // conv_igemm_persistent_kernel (one CTA per SM walks the tiles)
  bool halo = false;       // conv_halo_kernel instead of conv_igemm_kernel
  bool halo_cp = false;    // ... with its strips copied by cp.async (stem configuration)
  HaloArgs hargs{};
  int dual_stages = 0;     // > 0: a dual-tile configuration exists (two M tiles per item share the weight tile)
  size_t dual_smem = 0;
  const uint16_t* d_w = nullptr;      // im2col layers: packed K-major weights [cout][taps * cblocks * ck] and
  const uint16_t* d_bias16 = nullptr; //   the [cout][16] (hi, lo) shift matrix (pool_conv.cuh reads them directly)
  bool pair = false;       // conv_pair_kernel (CTA pairs, resident weights); `alt` serves batches below pair_min_batch
  PairArgs pargs{};
  int pair_min_batch = 0;
  std::shared_ptr<ConvLayer> alt;
  std::shared_ptr<ConvLayer> narrow;   // small-batch twin: same weights, narrow N work items (make_conv)
};

static int g_num_sms = 148;
// Where a kernel issues griddepcontrol.launch_dependents: at its start (0) or when a CTA has issued its last MMAs (1, default).
static bool pdl_late() {
  static const bool v = getenv("ICP_PDL_LATE") ? atoi(getenv("ICP_PDL_LATE")) != 0 : true;
  return v;
}


// Every tower kernel goes through cudaLaunchKernelEx so that it can carry the programmatic-dependent-launch
// attribute (ptx.cuh pdl_wait / pdl_launch): `pdl` is set when the previous operation enqueued on the same stream
// was one of this library's kernels (see PdlTracker) — then its prologue overlaps that kernel's tail.
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster,
                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CK(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}
static int pick_n_tile(int cout) {
  if (cout <= 256) return cout;
  for (int parts = 2; parts <= 16; ++parts)
    if (cout % parts == 0 && (cout / parts) % 16 == 0 && cout / parts <= 256) return cout / parts;
  fail(ICP_E_INVALID, "no UMMA N tiling for cout=%d", cout);
}
static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

// Packs HWIO fp32 weights into the K-major [cout][taps][cblocks*64] 16-bit matrix the B-operand TMA
// reads, zero-filling the channel tail of every tap.
// `scale` (nullable): per-output-channel BatchNorm scale multiplied in before rounding to 16 bits.
static std::vector<uint16_t> pack_weights(const float* w, int kh, int kw, int cin, int cout, int ck, int fmt,
                                          const float* scale = nullptr) {
  const int cb = (cin + ck - 1) / ck;
  const size_t kp = static_cast<size_t>(kh) * kw * cb * ck;
  std::vector<uint16_t> out(static_cast<size_t>(cout) * kp, 0);
  for (int r = 0; r < kh; ++r)
    for (int s = 0; s < kw; ++s)
      for (int ci = 0; ci < cin; ++ci) {
        const float* src = w + ((static_cast<size_t>(r) * kw + s) * cin + ci) * cout;
        const size_t kidx = (static_cast<size_t>(r) * kw + s) * cb * ck + ci;
        for (int co = 0; co < cout; ++co) out[co * kp + kidx] = host_f2h(scale ? src[co] * scale[co] : src[co], fmt);
      }
  return out;
}

static float host_h2f(uint16_t h, int fmt) {
  if (fmt) {
    const uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  }
  __half hh;
  memcpy(&hh, &h, 2);
  return __half2float(hh);
}

// B operand of the bias MMA: [cout][16] 16-bit, K columns 0/1 = shift split into hi + lo halves (the A operand's
// first two K columns are 1), so the accumulator starts at shift with ~2^-17 relative error.
static std::vector<uint16_t> pack_bias(const float* shift, int cout, int fmt) {
  std::vector<uint16_t> out(static_cast<size_t>(cout) * 16, 0);
  for (int co = 0; co < cout; ++co) {
    const uint16_t hi = host_f2h(shift[co], fmt);
    out[co * 16] = hi;
    out[co * 16 + 1] = host_f2h(shift[co] - host_h2f(hi, fmt), fmt);
  }
  return out;
}

// Halo kernel's B operand: K index = ((chunk * taps + tap) * ck + j), channel = chunk*ck + j.
static std::vector<uint16_t> pack_weights_halo(const float* w, int kh, int kw, int cin, int cout, int ck, int fmt,
                                               const float* scale = nullptr) {
  const int chunks = (cin + ck - 1) / ck, taps = kh * kw;
  const size_t kp = static_cast<size_t>(chunks) * taps * ck;
  std::vector<uint16_t> out(static_cast<size_t>(cout) * kp, 0);
  for (int t = 0; t < taps; ++t)
    for (int ci = 0; ci < cin; ++ci) {
      const float* src = w + (static_cast<size_t>(t) * cin + ci) * cout;
      const size_t kidx = (static_cast<size_t>(ci / ck) * taps + t) * ck + (ci % ck);
      for (int co = 0; co < cout; ++co) out[co * kp + kidx] = host_f2h(scale ? src[co] * scale[co] : src[co], fmt);
    }
  return out;
}


// Pair kernel's B operand for channels [c0, c0 + cn): K index = ((chunk * taps + tap) * ck + j), channel = c0 + chunk*ck + j.
static std::vector<uint16_t> pack_weights_range(const float* w, int kh, int kw, int cin, int cout, int c0, int cn, int ck,
                                                int fmt, const float* scale) {
  const int chunks = (cn + ck - 1) / ck, taps = kh * kw;
  const size_t kp = static_cast<size_t>(chunks) * taps * ck;
  std::vector<uint16_t> out(static_cast<size_t>(cout) * kp, 0);
  for (int t = 0; t < taps; ++t)
    for (int ci = 0; ci < cn; ++ci) {
      const float* src = w + (static_cast<size_t>(t) * cin + c0 + ci) * cout;
      const size_t kidx = (static_cast<size_t>(ci / ck) * taps + t) * ck + (ci % ck);
      for (int co = 0; co < cout; ++co) out[co * kp + kidx] = host_f2h(scale ? src[co] * scale[co] : src[co], fmt);
    }
  return out;
}

// CTA-pair, weight-resident halo kernel (conv_pair.cuh).  Returns false when the layer is not eligible.
static bool make_pair_conv(ConvLayer& L, DevicePool& pool, int max_n, const Act& in, int in_coff, int cin, const float* w_hwio,
                           int kh, int kw, int cout, int pad_h, int pad_w, int oh, int ow, const float* scale, const float* shift,
                           bool relu, const Act& out, int out_coff, int fmt) {
  if (cin % 16) return false;
  const int taps = kh * kw, wp = in.w + 2 * pad_w;
  if (wp > 256) return false;
  // N tile: the whole layer up to 128 columns, else an even split (conv4: 192 = 2 x 96)
  int n_tile = 0, n_tiles = 0;
  for (int parts = 1; parts <= 8 && !n_tile; ++parts)
    if (cout % parts == 0 && (cout / parts) % 32 == 0 && cout / parts <= 128) { n_tiles = parts; n_tile = cout / parts; }
  if (!n_tile) return false;
  if (n_tile % 32) return false;                       // each CTA supplies n_tile / 2 rows, a multiple of 16
  const int half = n_tile / 2;
  const int q = cin / 64, rem = cin % 64;
  const int tail_split = (rem == 16 || rem == 32) ? 1 : 0;
  const int chunks = q + (rem ? 1 : 0);
  const int tail_rb = rem == 16 ? 32 : (rem == 32 ? 64 : 128);
  const int tail_ksteps = rem ? rem / 16 : 4;
  const int main_chunks = chunks - tail_split;
  const int w_main_bytes = main_chunks * taps * half * 128;
  const int w_tail_bytes = tail_split ? taps * half * tail_rb : 0;
  const int w_bytes = (w_main_bytes + w_tail_bytes + 1023) & ~1023;
  // sub-tiles per band: as many as TMEM holds with two accumulator sets (<= 4)
  int T = 4;
  while (T > 1 && 2 * T * n_tile > 512) T >>= 1;
  if (2 * T * n_tile > 512) return false;
  int nstrips = 0, slot_bytes = 0, band_rows = 0, chunk_pitch = 0;
  for (; T >= 1; T >>= 1) {
    band_rows = (T * kBM) / wp;
    if (band_rows < 1) continue;
    if (band_rows > oh) band_rows = oh;
    const int positions = T * kBM + (kh - 1) * wp + (kw - 1);            // furthest strip position any MMA row touches
    chunk_pitch = (positions * 128 + 1023) & ~1023;
    slot_bytes = main_chunks * chunk_pitch + (tail_split ? ((positions * tail_rb + 1023) & ~1023) : 0);
    for (nstrips = 4; nstrips >= 2; --nstrips)
      if (pair_smem_bytes(slot_bytes, nstrips, w_bytes) <= 227 * 1024) break;
    if (nstrips >= 2 && band_rows + kh - 1 <= 256) break;
  }
  if (T < 1) return false;
  PairArgs& a = L.pargs;
  a = PairArgs{};
  a.oh = oh; a.ow = ow; a.wp = wp; a.pad_h = pad_h; a.pad_w = pad_w; a.kh = kh; a.kw = kw;
  a.chunks = chunks; a.tail_rb = tail_rb; a.tail_ksteps = tail_ksteps; a.tail_split = tail_split;
  a.T = T; a.band_rows = band_rows; a.bands_per_image = (oh + band_rows - 1) / band_rows; a.strip_rows = band_rows + kh - 1;
  a.chunk_pitch = chunk_pitch; a.slot_bytes = slot_bytes; a.nstrips = nstrips;
  a.n_tile = n_tile; a.n_tiles = n_tiles; a.acc_cols = T * n_tile; a.acc_sets = 2; a.tmem_cols = pow2_cols(2 * T * n_tile);
  a.w_main_bytes = w_main_bytes; a.w_bytes = w_bytes;
  a.fmt = fmt; a.relu = relu ? 1 : 0;
  a.out = out.p; a.ldo = out.c; a.coff = out_coff;
  L.smem = pair_smem_bytes(slot_bytes, nstrips, w_bytes);
  // operands
  const int cmain = tail_split ? q * 64 : cin;
  std::vector<uint16_t> wm = pack_weights_range(w_hwio, kh, kw, cin, cout, 0, cmain, 64, fmt, scale);
  void* d_wm = pool.alloc(wm.size() * 2 + 256);
  if (!wm.empty()) CK(cudaMemcpy(d_wm, wm.data(), wm.size() * 2, cudaMemcpyHostToDevice));
  const uint16_t* a_ptr = in.p + in_coff;
  if (main_chunks > 0) {
    L.tm_b = make_tm_2d_rows(d_wm, cout, static_cast<uint64_t>(main_chunks) * taps * 64, half, 128, fmt);
    L.tm_a = make_tm_strip(a_ptr, max_n, in.h, in.w, cin, in.c, wp, a.strip_rows, 128, fmt);
  }
  L.tm_b_tail = L.tm_b;
  L.tm_a_tail = L.tm_a;
  if (tail_split) {
    const int tck = tail_rb / 2;
    std::vector<uint16_t> wt = pack_weights_range(w_hwio, kh, kw, cin, cout, q * 64, rem, tck, fmt, scale);
    void* d_wt = pool.alloc(wt.size() * 2);
    CK(cudaMemcpy(d_wt, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
    L.tm_b_tail = make_tm_2d_rows(d_wt, cout, static_cast<uint64_t>(taps) * tck, half, tail_rb, fmt);
    L.tm_a_tail = make_tm_strip(a_ptr, max_n, in.h, in.w, cin, in.c, wp, a.strip_rows, tail_rb, fmt);
    if (main_chunks == 0) { L.tm_b = L.tm_b_tail; L.tm_a = L.tm_a_tail; }     // Cin < 64: the narrow chunk is the only one
  }
  std::vector<uint16_t> bp = pack_bias(shift, cout, fmt);
  void* d_bias = pool.alloc(bp.size() * 2);
  CK(cudaMemcpy(d_bias, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice));
  L.tm_bias = make_tm_2d_rows(d_bias, cout, 16, half, 32, fmt);
  L.pair = true;
  L.n_tiles = n_tiles;
  return true;
}

// in: NHWC buffer (pitch in.c) of which channels [in_coff, in_coff+cin) feed the conv.
static ConvLayer make_conv(DevicePool& pool, const std::string& name, int max_n, const Act& in, int in_coff, int cin,
                           const float* w_hwio, int kh, int kw, int cout, int stride, bool same, const float* scale,
                           const float* shift, bool relu, const Act& out, int out_coff, int fmt,
                           const float* floor = nullptr, const Act* out2 = nullptr, int out2_coff = 0,
                           int n_split = 0, bool transpose_out = false, bool allow_pair = true,
                           const ConvLayer* narrow_of = nullptr, int narrow_n_tile = 0) {
  // narrow_of: build the small-batch twin of that layer — same packed weights, output channels cut into narrow_n_tile-wide
  // work items (see the end of this function)
  if (same && stride != 1) fail(ICP_E_INVALID, "%s: SAME padding is only supported with stride 1", name.c_str());
  if (cout % 16) fail(ICP_E_INVALID, "%s: cout=%d is not a multiple of 16", name.c_str(), cout);
  if (cin % 8 || in.c % 8 || in_coff % 8 || out.c % 8 || out_coff % 8)
    fail(ICP_E_INVALID, "%s: channel counts/offsets must be multiples of 8", name.c_str());
  const int pad_h = same ? (kh - 1) / 2 : 0, pad_w = same ? (kw - 1) / 2 : 0;
  const int oh = (in.h + 2 * pad_h - kh) / stride + 1, ow = (in.w + 2 * pad_w - kw) / stride + 1;
  if (oh != out.h || ow != out.w)
    fail(ICP_E_INVALID, "%s: output geometry %dx%d != buffer %dx%d", name.c_str(), oh, ow, out.h, out.w);
  if (out2 && (n_split % 16 || n_split <= 0 || n_split >= cout || out2->h != oh || out2->w != ow || out2->c % 8 ||
               out2_coff % 8))
    fail(ICP_E_INVALID, "%s: bad second destination", name.c_str());
  ConvLayer L;
  L.name = name;
  L.cout = cout;
  L.pix_per_image = oh * ow;
  L.macs_per_image = static_cast<long long>(oh) * ow * cout * kh * kw * cin;
  L.in_elems_per_image = static_cast<long long>(in.h) * in.w * cin;
  // CTA-pair weight-resident kernel (conv_pair.cuh) for the k x k stride-1 layers on maps at least ICP_PAIR_MIN_W wide
  // (ICP_PAIR=0 disables it); batches below ICP_PAIR_MIN_BATCH use the plain layer built next to it.
  {
    static const bool pair_on = getenv("ICP_PAIR") ? atoi(getenv("ICP_PAIR")) != 0 : false;
    static const int pair_min_w = getenv("ICP_PAIR_MIN_W") ? atoi(getenv("ICP_PAIR_MIN_W")) : 30;
    static const int pair_max_w = getenv("ICP_PAIR_MAX_W") ? atoi(getenv("ICP_PAIR_MAX_W")) : 100;
    static const int pair_min_batch = getenv("ICP_PAIR_MIN_BATCH") ? atoi(getenv("ICP_PAIR_MIN_BATCH")) : 8;
    if (pair_on && allow_pair && !narrow_of && !out2 && !floor && !transpose_out && stride == 1 && kh * kw > 1 && ow >= pair_min_w &&
        ow <= pair_max_w &&
        make_pair_conv(L, pool, max_n, in, in_coff, cin, w_hwio, kh, kw, cout, pad_h, pad_w, oh, ow, scale, shift, relu, out,
                       out_coff, fmt)) {
      L.pair_min_batch = pair_min_batch;
      L.alt = std::make_shared<ConvLayer>(make_conv(pool, name, max_n, in, in_coff, cin, w_hwio, kh, kw, cout, stride, same, scale,
                                                    shift, relu, out, out_coff, fmt, floor, out2, out2_coff, n_split,
                                                    transpose_out, false));
      return L;
    }
  }
  // Operand row width: the widest of 128/64/32 bytes (64/32/16 channels per K block) that pads Cin by at most
  // 12 % — Cin = 80 -> 5 x 16, 160 -> 5 x 32, 48 -> 3 x 16, 96 -> 3 x 32, 288 -> 5 x 64 (11 %), the rest x 64.
  // MEASURED: the TMA unit retires ~one box row per 2.7 cycles per SM whatever the row's width, so narrower
  // rows (more rows for the same bytes) made conv4 2.5x SLOWER.  128-byte rows everywhere unless ICP_RB_AUTO.
  static const bool rb_auto = getenv("ICP_RB_AUTO") != nullptr;
  int rb = 128;
  if (rb_auto)
    for (int cand : {128, 64, 32}) {
      const int k = cand / 2, padded = (cin + k - 1) / k * k;
      rb = cand;
      if ((padded - cin) * 100 <= 12 * cin) break;
    }
  const int ck = rb / 2;
  const int cb = (cin + ck - 1) / ck;
  int n_tile = narrow_of ? narrow_n_tile : pick_n_tile(cout);
  L.n_tiles = cout / n_tile;
  // Experiment (ICP_UNEVEN_N=1): full 256-column tiles plus a narrower last one when that needs FEWER tiles than the
  // even split (640 = 256 + 256 + 128 instead of 4 x 160, 704 = 256 + 256 + 192 instead of 4 x 176).  MEASURED: no
  // gain (2.461 vs 2.461 ms/step; 17x17 1x1 768->704 0.0246 vs 0.0232 ms) — at 256 columns the K block is MMA-bound
  // and the ring is one stage shorter.
  static const bool no_uneven = getenv("ICP_UNEVEN_N") == nullptr;
  if (!no_uneven && !narrow_of && getenv("ICP_CONV_V1") == nullptr && kh * kw == 1 && (cout + 255) / 256 < L.n_tiles) {
    n_tile = 256;
    L.n_tiles = (cout + 255) / 256;
  }
  const int num_kb = kh * kw * cb;
  const size_t stage_bytes = conv_stage_bytes(n_tile, rb);
  static const bool conv_v1 = getenv("ICP_CONV_V1") != nullptr;     // one tile per CTA (bring-up version)
  const size_t budget = conv_v1 ? ((n_tile <= 128) ? 100 * 1024 : 200 * 1024) : 204 * 1024;
  int stages = static_cast<int>(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  // Two epilogue teams need a second 16 KiB staging buffer.  MEASURED: where that costs a pipeline stage (the 17x17
  // fused 1x1s, n_tile 176: 5 -> 4 stages) the layer got 10-20 % SLOWER (load-latency-bound), where it does not (conv3,
  // 35x35 fused 1x1s: epilogue-bound) 6-16 % faster — so: two teams only if the ring keeps its depth.  ICP_EPI_TEAMS=1|2 forces.
  int epi_teams = 1;
  {
    int s2 = static_cast<int>((budget - kOutStageBytes) / stage_bytes);
    if (s2 > kMaxStages) s2 = kMaxStages;
    if (s2 == stages) epi_teams = 2;
    static const int force = getenv("ICP_EPI_TEAMS") ? atoi(getenv("ICP_EPI_TEAMS")) : 0;
    if (force == 1) epi_teams = 1;
    if (force == 2) { epi_teams = 2; stages = s2 < 1 ? 1 : s2; }
  }
  if (conv_v1 && stages > num_kb) stages = num_kb;
  if (stages < 1) stages = 1;
  // The persistent kernel is used for every layer: since its role loops and epilogue were tightened it also beats
  // the one-tile-per-CTA kernel on the narrow layers (n_tile < 128: 2.517 vs 2.537 ms/step, conv3 0.066 vs 0.072,
  // 35x35 3x3 64->96 0.035 vs 0.043).  ICP_CONV_V1=1 or ICP_CONV_V1_NARROW=1 select the one-tile kernel for A/B runs.
  static const bool v1_narrow = getenv("ICP_CONV_V1_NARROW") != nullptr;
  const bool persistent = !conv_v1 && !(v1_narrow && n_tile < 128);
  if (!persistent) {
    const size_t b1 = (n_tile <= 128) ? 100 * 1024 : 200 * 1024;
    stages = static_cast<int>(b1 / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > num_kb + 1) stages = num_kb + 1;    // + the bias block
    if (stages < 1) stages = 1;
  }
  L.persistent = persistent;
  static const bool force_im2col = getenv("ICP_FORCE_IM2COL") != nullptr;
  const bool im2col = force_im2col || !(kh == 1 && kw == 1 && stride == 1);
  // Stride-1 convs with a real filter footprint: halo-resident kernel (conv_halo.cuh).  Per-image bands need a
  // map wide enough to fill 128-row tiles; 1 x k filters run in "tall" mode on any map size.
  // ICP_NO_HALO=1 falls back to the im2col kernels for A/B comparison.
  static const bool no_halo = getenv("ICP_NO_HALO") != nullptr;
  // MEASURED (profiles/r01_halo_vs_im2col.md): the halo kernel wins only on the two 147/149-wide stem convs
  // (0.29 vs 0.49 ms, 0.30 vs 0.57 ms); on 73/35-wide maps and in tall mode it is 10-60 % slower than the
  // im2col kernels, so those stay opt-in (ICP_HALO_MIN_W=30, ICP_HALO_TALL=1) until that is understood.
  static const int halo_min_w = getenv("ICP_HALO_MIN_W") ? atoi(getenv("ICP_HALO_MIN_W")) : 100;
  static const bool halo_tall = getenv("ICP_HALO_TALL") != nullptr;
  const bool tall = (kh == 1) && halo_tall;
  // Weight-resident strips with a STATIC N tile per CTA (profiles/r02_pair_kernel.md §3): the layer's output channels are
  // split into n_tiles slices small enough that one slice's weights stay in shared memory for the CTA's whole life (the grid is
  // a multiple of n_tiles, so CTA b only ever sees N tile b % n_tiles); the input strip is the only stream and is re-read once
  // per slice.  Narrow N costs MMA efficiency ((128 + n_tile) x 32 B of operands per MMA against 128 B/cycle of shared memory)
  // but removes both per-tap re-reads that bound the im2col kernel.  ICP_HALO_RES=0 disables, ICP_HALO_RES_MIN_W / _MAX_W bound the maps.
  static const bool halo_res = getenv("ICP_HALO_RES") ? atoi(getenv("ICP_HALO_RES")) != 0 : false;
  static const int res_min_w = getenv("ICP_HALO_RES_MIN_W") ? atoi(getenv("ICP_HALO_RES_MIN_W")) : 30;
  static const int res_max_w = getenv("ICP_HALO_RES_MAX_W") ? atoi(getenv("ICP_HALO_RES_MAX_W")) : 50;
  if (halo_res && allow_pair && !narrow_of && !no_halo && !out2 && !floor && !transpose_out && stride == 1 && kh * kw > 1 && cin % 16 == 0 &&
      ow >= res_min_w && ow <= res_max_w && in.w + 2 * pad_w <= 256) {
    const int rb = cin <= 32 ? 64 : 128, ck = rb / 2;
    const int wp = in.w + 2 * pad_w, chunks = (cin + ck - 1) / ck, taps = kh * kw;
    int best_nt = 0, best_T = 0, best_ns = 0, best_rows = 0, best_sb = 0;
    for (int parts = 1; parts <= 8 && !best_nt; ++parts) {
      if (cout % parts || (cout / parts) % 16) continue;
      const int nt = cout / parts;
      if (nt > 256) continue;
      const int res_bytes = chunks * taps * nt * rb;
      for (int T = 4; T >= 1 && !best_nt; T >>= 1) {
        if (2 * T * nt > 512) continue;
        const int sr = (kh - 1) + 1 + (wp - 1 + (T * kBM - 1) + (kw - 1)) / wp;
        const int sb = ((sr * wp * rb) + 1023) & ~1023;
        if (sr > 256) continue;
        int ns = 4;
        while (ns >= 2 && halo_smem_bytes(sb, ns, res_bytes, true) > 225 * 1024) --ns;
        if (ns < 2) continue;
        best_nt = nt; best_T = T; best_ns = ns; best_rows = sr; best_sb = sb;
      }
    }
    if (best_nt) {
      const int n_tile_r = best_nt, T = best_T;
      const int res_bytes = chunks * taps * n_tile_r * rb;
      std::vector<uint16_t> wp_host = pack_weights_halo(w_hwio, kh, kw, cin, cout, ck, fmt, scale);
      void* d_w = pool.alloc(wp_host.size() * 2);
      CK(cudaMemcpy(d_w, wp_host.data(), wp_host.size() * 2, cudaMemcpyHostToDevice));
      std::vector<uint16_t> bp = pack_bias(shift, cout, fmt);
      void* d_bias = pool.alloc(bp.size() * 2);
      CK(cudaMemcpy(d_bias, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice));
      L.tm_bias = make_tm_2d_rows(d_bias, cout, 16, n_tile_r, 32, fmt);
      L.tm_b = make_tm_2d_rows(d_w, cout, static_cast<uint64_t>(chunks) * taps * ck, n_tile_r, rb, fmt);
      L.tm_a = make_tm_strip(in.p + in_coff, max_n, in.h, in.w, cin, in.c, wp, best_rows, rb, fmt);
      L.tm_a_tail = L.tm_a;
      L.halo = true;
      L.n_tiles = cout / n_tile_r;
      HaloArgs& h = L.hargs;
      h = HaloArgs{};
      h.oh = oh; h.ow = ow; h.wp = wp; h.pad_h = pad_h; h.pad_w = pad_w; h.kh = kh; h.kw = kw;
      h.chunks = chunks; h.row_bytes = rb; h.T = T; h.strip_rows = best_rows; h.strip_bytes = best_sb; h.nstrips = best_ns;
      h.bands_per_image = (oh * wp + kBM * T - 1) / (kBM * T);
      h.n_tile = n_tile_r; h.n_tiles = L.n_tiles;
      h.acc_cols = T * n_tile_r; h.acc_sets = 2; h.tmem_cols = pow2_cols(2 * T * n_tile_r);
      h.w_resident = 1; h.w_bytes = res_bytes; h.b_stages = 1;
      h.box_rows = best_rows; h.nbox = 1; h.rem_rows = 0;
      h.tail_ksteps = (cin - (chunks - 1) * ck + 15) / 16;
      h.fmt = fmt; h.relu = relu ? 1 : 0; h.epi_direct = 1;
      h.out = out.p; h.ldo = out.c; h.coff = out_coff;
      L.smem = halo_smem_bytes(best_sb, best_ns, res_bytes, true);
      L.pair_min_batch = getenv("ICP_HALO_RES_MIN_BATCH") ? atoi(getenv("ICP_HALO_RES_MIN_BATCH")) : 8;
      L.alt = std::make_shared<ConvLayer>(make_conv(pool, name, max_n, in, in_coff, cin, w_hwio, kh, kw, cout, stride, same, scale,
                                                    shift, relu, out, out_coff, fmt, floor, out2, out2_coff, n_split,
                                                    transpose_out, false));
      return L;
    }
  }
  if (!no_halo && !narrow_of && !out2 && !floor && stride == 1 && kh * kw > 1 && in.w + 2 * pad_w <= 256 &&
      (tall || ow >= halo_min_w) && (!transpose_out || (tall && oh == ow))) {
    const int rb = cin <= 32 ? 64 : 128, ck = rb / 2;
    const int wp = in.w + 2 * pad_w;
    const int chunks = (cin + ck - 1) / ck;
    const int taps = kh * kw;
    // sub-tiles per band: double-buffered accumulators when they leave at least two sub-tiles, else one set
    static const int halo_t = getenv("ICP_HALO_T") ? atoi(getenv("ICP_HALO_T")) : 4;    // longest band tried (sub-tiles)
    int T = halo_t, acc_sets = 2;
    while (T > 1 && 2 * T * n_tile > 512) T >>= 1;
    if (2 * T * n_tile > 512 || T == 1) {
      acc_sets = 1;
      T = halo_t;
      while (T > 1 && T * n_tile > 512) T >>= 1;
      if (2 * T * n_tile <= 512) acc_sets = 2;
    }
    const int tile_bytes = n_tile * rb;
    const int tile_pitch = (tile_bytes + 1023) & ~1023;
    const int res_bytes = chunks * taps * tile_bytes;
    int strip_rows = 0, strip_bytes = 0, nstrips = 2, b_stages = 0, w_bytes = 0;
    bool resident = false, ok = false;
    // Experiment (ICP_HALO_PREFER_RESIDENT=1): keep the weights resident even if that means shorter bands (T = 1)
    static const bool prefer_res = getenv("ICP_HALO_PREFER_RESIDENT") != nullptr;
    if (prefer_res && L.n_tiles == 1)
      for (int Tr = T; Tr >= 1 && !ok; Tr >>= 1) {
        const int sr = (kh - 1) + 1 + (wp - 1 + (Tr * kBM - 1) + (kw - 1)) / wp;
        const int sb = ((sr * wp * rb) + 1023) & ~1023;
        if (sr <= 256 && halo_smem_bytes(sb, 2, res_bytes) <= 225 * 1024) {
          T = Tr; strip_rows = sr; strip_bytes = sb; resident = true; w_bytes = res_bytes; b_stages = 1; ok = true;
        }
      }
    for (; T >= 1 && !ok; T >>= 1) {
      strip_rows = (kh - 1) + 1 + (wp - 1 + (T * kBM - 1) + (kw - 1)) / wp;
      strip_bytes = ((strip_rows * wp * rb) + 1023) & ~1023;
      if (strip_rows > 256) continue;
      if (L.n_tiles == 1 && halo_smem_bytes(strip_bytes, 2, res_bytes) <= 225 * 1024) {
        resident = true; w_bytes = res_bytes; b_stages = 1; ok = true;
      } else {
        for (b_stages = 6; b_stages >= 3; --b_stages)
          if (halo_smem_bytes(strip_bytes, 2, b_stages * tile_pitch) <= 225 * 1024) { ok = true; break; }
        resident = false; w_bytes = b_stages * tile_pitch;
      }
      if (ok) break;
    }
    if (ok) {
      acc_sets = (2 * T * n_tile <= 512) ? 2 : 1;
      // narrow tiles store straight from registers (conv1, 32 channels: 0.147 -> 0.117 ms); at 64 channels the 16-byte
      // stores of one warp touch 32 lines per instruction and the staged copy-out is as fast (conv2 0.165 vs 0.169 ms)
      const bool direct = getenv("ICP_HALO_DIRECT") ? atoi(getenv("ICP_HALO_DIRECT")) != 0 : n_tile <= 32;
      // strips in the ring.  MEASURED (timeline, conv1): a strip lands ~5.4k cycles after its issue while the band's MMAs take
      // ~3.3k, so with two strips the band period is their mean (4.4k) — but a THIRD strip (it fits once the direct epilogue's
      // unused staging is not allocated) made conv1 slower, 0.1067 -> 0.1167 ms: these MMAs (N = 32, K = 16: 5 KB of operands
      // for 16 cycles of tensor work) saturate the shared-memory port, and strip writes in flight during them take it away.
      // Streamed-weight layers (opt-in) keep up to 4.  ICP_HALO_MAX_STRIPS overrides.
      static const int max_strips_env = getenv("ICP_HALO_MAX_STRIPS") ? atoi(getenv("ICP_HALO_MAX_STRIPS")) : 0;
      const int max_strips = max_strips_env ? max_strips_env : (resident ? 2 : 4);
      while (nstrips < max_strips && halo_smem_bytes(strip_bytes, nstrips + 1, w_bytes, direct, n_tile) <= 226 * 1024) ++nstrips;
      const size_t smem = halo_smem_bytes(strip_bytes, nstrips, w_bytes, direct, n_tile);
      std::vector<uint16_t> wp_host = pack_weights_halo(w_hwio, kh, kw, cin, cout, ck, fmt, scale);
      void* d_w = pool.alloc(wp_host.size() * 2);
      CK(cudaMemcpy(d_w, wp_host.data(), wp_host.size() * 2, cudaMemcpyHostToDevice));
      std::vector<uint16_t> bp = pack_bias(shift, cout, fmt);
      void* d_bias = pool.alloc(bp.size() * 2);
      CK(cudaMemcpy(d_bias, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice));
      L.tm_bias = make_tm_2d_rows(d_bias, cout, 16, n_tile, 32, fmt);
      const uint64_t kp = static_cast<uint64_t>(chunks) * taps * ck;
      L.halo = true;
      L.tm_b = make_tm_2d_rows(d_w, cout, kp, n_tile, rb, fmt);
      // boxes per strip: box_rows rows each such that every box starts on a 128-byte boundary of the strip (TMA destination
      // alignment; the swizzle is a function of the absolute address, so the boxes tile the strip seamlessly)
      static const int box_rows_env = getenv("ICP_HALO_BOX_ROWS") ? atoi(getenv("ICP_HALO_BOX_ROWS")) : 0;
      int box_rows = strip_rows;
      if (box_rows_env > 0 && box_rows_env < strip_rows && (static_cast<long long>(box_rows_env) * wp * rb) % 128 == 0) box_rows = box_rows_env;
      const int nbox = strip_rows / box_rows, rem_rows = strip_rows - nbox * box_rows;
      const int gn = tall ? 1 : max_n, gh = tall ? max_n * in.h : in.h;     // tall: rows of all images form one grid {C, W, N*H, 1}
      L.tm_a = make_tm_strip(in.p + in_coff, gn, gh, in.w, cin, in.c, wp, box_rows, rb, fmt);
      L.tm_a_tail = rem_rows ? make_tm_strip(in.p + in_coff, gn, gh, in.w, cin, in.c, wp, rem_rows, rb, fmt) : L.tm_a;
      HaloArgs& h = L.hargs;
      h.n_images = 0;
      h.box_rows = box_rows; h.nbox = nbox; h.rem_rows = rem_rows;
      h.oh = oh; h.ow = ow; h.wp = wp; h.pad_h = pad_h; h.pad_w = pad_w; h.kh = kh; h.kw = kw;
      h.chunks = chunks; h.row_bytes = rb; h.T = T; h.strip_rows = strip_rows; h.strip_bytes = strip_bytes;
      h.nstrips = nstrips;
      h.bands_per_image = (oh * wp + kBM * T - 1) / (kBM * T);
      h.tall = tall ? 1 : 0;
      h.transpose_out = transpose_out ? 1 : 0;
      h.n_tile = n_tile; h.n_tiles = L.n_tiles;
      h.acc_cols = T * n_tile; h.acc_sets = acc_sets; h.tmem_cols = pow2_cols(acc_sets * T * n_tile);
      h.w_resident = resident ? 1 : 0; h.w_bytes = w_bytes; h.b_stages = b_stages;
      h.tail_ksteps = rb / 32;                 // this configuration multiplies the zero-padded K slices of the last chunk too
      h.fmt = fmt; h.relu = relu ? 1 : 0;
      h.epi_direct = direct ? 1 : 0;
      h.out = out.p; h.ldo = out.c; h.coff = out_coff;
      h.in = in.p + in_coff; h.ih = in.h; h.iw = in.w; h.ldi = in.c;
      // strips by cp.async instead of TMA (conv_halo.cuh kCp): opt-in (ICP_HALO_CP=1).  MEASURED at batch 128 with the two
      // spare warps as copiers: conv1 0.135 -> 0.215 ms, conv2 0.145 -> 0.231 ms — 64 threads cannot issue a 67 KB strip
      // (65 copies each, ~150 cycles apiece including the index arithmetic) as fast as one TMA box lands it.
      static const bool halo_cp = getenv("ICP_HALO_CP") ? atoi(getenv("ICP_HALO_CP")) != 0 : false;
      L.halo_cp = halo_cp && resident && !tall && !transpose_out && cin % ck == 0 && (in_coff % 8) == 0 && (in.c % 8) == 0;
      L.smem = smem;
      return L;
    }
  }
  if (transpose_out) fail(ICP_E_INVALID, "%s: transposed output needs the halo kernel in tall mode", name.c_str());

  // weights (BatchNorm scale folded in) and the shift tile of the bias MMA
  const void* d_w = narrow_of ? static_cast<const void*>(narrow_of->d_w) : nullptr;
  const void* d_bias = narrow_of ? static_cast<const void*>(narrow_of->d_bias16) : nullptr;
  if (!narrow_of) {
    std::vector<uint16_t> wp = pack_weights(w_hwio, kh, kw, cin, cout, ck, fmt, scale);
    void* dw = pool.alloc(wp.size() * 2);
    CK(cudaMemcpy(dw, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
    std::vector<uint16_t> bp = pack_bias(shift, cout, fmt);
    void* db = pool.alloc(bp.size() * 2);
    CK(cudaMemcpy(db, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice));
    d_w = dw;
    d_bias = db;
  }
  L.tm_bias = make_tm_2d_rows(d_bias, cout, 16, n_tile, 32, fmt);
  L.d_w = static_cast<const uint16_t*>(d_w);
  L.d_bias16 = static_cast<const uint16_t*>(d_bias);

  const uint64_t kp = static_cast<uint64_t>(num_kb) * ck;
  L.tm_b = make_tm_2d(d_w, cout, kp, kp, n_tile, fmt, rb);
  const uint16_t* a_ptr = in.p + in_coff;
  if (im2col)
    L.tm_a = make_tm_im2col(a_ptr, max_n, in.h, in.w, cin, in.c, kh, kw, stride, pad_h, pad_w, fmt, rb);
  else
    L.tm_a = make_tm_2d(a_ptr, static_cast<uint64_t>(max_n) * in.h * in.w, cin, in.c, kBM, fmt, rb);

  // last channel block of every tap: narrower boxes when only 16 / 32 channels remain (conv4: 80 = 64 + 16)
  const int tail_c = cin - (cb - 1) * ck;
  static const bool no_narrow_tail = getenv("ICP_NO_NARROW_TAIL") != nullptr;
  const int tail_rb = (no_narrow_tail || rb != 128) ? rb : (tail_c <= 16 ? 32 : (tail_c <= 32 ? 64 : 128));
  L.tm_a_tail = L.tm_a;
  L.tm_b_tail = L.tm_b;
  if (tail_rb != rb) {
    L.tm_b_tail = make_tm_2d(d_w, cout, kp, kp, n_tile, fmt, tail_rb);
    if (im2col)
      L.tm_a_tail = make_tm_im2col(a_ptr, max_n, in.h, in.w, cin, in.c, kh, kw, stride, pad_h, pad_w, fmt, tail_rb);
    else
      L.tm_a_tail = make_tm_2d(a_ptr, static_cast<uint64_t>(max_n) * in.h * in.w, cin, in.c, kBM, fmt, tail_rb);
  }

  // 2-CTA clusters (weight tiles multicast, each CTA fetching half the rows).  MEASURED at batch 128: conv4 0.1984 ->
  // 0.1895 ms and the K=1024 N=256 GEMM 0.2666 -> 0.2616, but the 35x35 5x5 0.0638 -> 0.0704 and every launch with
  // only a few tiles per SM slower (2.548 vs 2.480 ms/step when used everywhere): the lock-step of the two CTAs and
  // the coarser work items cost more than the weight bytes save unless the weight tile is most of the traffic and
  // there are many waves.  So: wide tiles only, and only when a launch has >= 16 tiles per SM (decided per launch).
  // ICP_NO_CLUSTER=1 / ICP_CLUSTER_ALL=1 for A/B runs.
  static const bool no_cluster = getenv("ICP_NO_CLUSTER") != nullptr;
  static const bool cluster_all = getenv("ICP_CLUSTER_ALL") != nullptr;
  L.cluster = persistent && !no_cluster && num_kb >= 2 && (cluster_all || n_tile >= 160);
  L.cluster_min_tiles = cluster_all ? 0 : 16;
  L.tm_b_half = L.tm_b;
  L.tm_b_tail_half = L.tm_b_tail;
  if (L.cluster) {
    // ICP_CLUSTER_SIZE=8: clusters of 8 where every CTA's share of the weight tile is whole swizzle periods (8 rows).
    // MEASURED (B200, batch 128): only 15 clusters of 8 are co-resident (120 of 148 SMs) and conv4 takes 0.228 ms against 0.194
    // with pairs and 0.202 without clusters (K=1024 N=256 GEMM 0.304 / 0.278 / 0.287; tower 2.309 vs 2.271 ms) — time scaled
    // with the SM count although the L2 reads of the weights fell 4x: what bounds these kernels is each SM's own ~45 B/cycle
    // L2->SM port (148 of them = the ~6.3 kB/cycle "LTS cap"), and a multicast still delivers every byte to every SM.
    static const int cs_env = getenv("ICP_CLUSTER_SIZE") ? atoi(getenv("ICP_CLUSTER_SIZE")) : 2;
    L.cluster_size = (cs_env == 8 && n_tile % 64 == 0) ? 8 : 2;
    const int part = n_tile / L.cluster_size;
    L.tm_b_half = make_tm_2d(d_w, cout, kp, kp, part, fmt, rb);
    L.tm_b_tail_half = tail_rb != rb ? make_tm_2d(d_w, cout, kp, kp, part, fmt, tail_rb) : L.tm_b_half;
  }

  ConvArgs& a = L.args;
  a.m_total = 0;
  a.oh = oh;
  a.ow = ow;
  a.stride = stride;
  a.pad_h = pad_h;
  a.pad_w = pad_w;
  a.kw = kw;
  a.taps = kh * kw;
  a.cblocks = cb;
  a.tail_ksteps = (tail_c + 15) / 16;
  a.tail_rb = tail_rb;
  a.row_bytes = rb;
  a.n_tile = n_tile;
  a.cout = cout;
  a.tmem_cols = pow2_cols(n_tile);
  a.stages = stages;
  a.im2col = im2col ? 1 : 0;
  a.fmt = fmt;
  a.out = out.p;
  a.ldo = out.c;
  a.coff = out_coff;
  static const bool v1_staged = getenv("ICP_V1_STAGED") != nullptr;
  a.epi_staged = v1_staged ? 1 : 0;
  a.epi_teams = persistent ? epi_teams : 1;
  a.pdl_late = 0;          // set per launch (launch_conv)
  a.dbg = nullptr;
  // ReLU column range: everything, nothing, or (fused siblings) the leading run of channels whose floor is 0
  a.relu_lo = 0;
  a.relu_hi = relu ? cout : 0;
  if (floor) {
    int hi = 0;
    while (hi < cout && floor[hi] == 0.0f) ++hi;
    for (int i = hi; i < cout; ++i)
      if (floor[i] == 0.0f) fail(ICP_E_INVALID, "%s: ReLU channels must form one leading run", name.c_str());
    if (hi % 16) fail(ICP_E_INVALID, "%s: ReLU boundary %d is not a multiple of 16", name.c_str(), hi);
    a.relu_hi = hi;
  }
  a.out2 = out2 ? out2->p : nullptr;
  a.ldo2 = out2 ? out2->c : 0;
  a.coff2 = out2_coff;
  a.n_split = out2 ? n_split : (1 << 30);
  L.smem = conv_smem_bytes(stages, n_tile, rb, persistent, persistent ? epi_teams : 1);
  a.dual = 0;
  // Dual-tile items (conv_igemm.cuh): two consecutive M tiles share every weight tile — the weight bytes, the larger half of
  // what the tap-heavy 17x17 layers pull through L2->SM, halve — at the price of the TMEM double buffer (the epilogue of an
  // item no longer overlaps the next item's MMAs), so it is used only for launches with few tiles per SM (launch_conv).
  if (persistent && pow2_cols(n_tile) <= 256) {
    const size_t dstage = (static_cast<size_t>(2 * kBM + n_tile) * rb + 1023) & ~size_t(1023);
    int ds = static_cast<int>((204 * 1024 - kOutStageBytes) / dstage);
    if (ds > kMaxStages) ds = kMaxStages;
    if (ds >= 2) {
      L.dual_stages = ds;
      L.dual_smem = 1024 + static_cast<size_t>(ds) * dstage + 2 * kOutStageBytes + kOnesBytes + 256 + 1024;
    }
  }
  {
    const uint64_t out_rows = static_cast<uint64_t>(max_n) * oh * ow;
    L.tm_o = make_tm_out(out.p, out_rows, out.c, fmt);
    L.tm_o2 = out2 ? make_tm_out(out2->p, out_rows, out2->c, fmt) : L.tm_o;
    L.tm_o64 = make_tm_out(out.p, out_rows, out.c, fmt, 64);
    L.tm_o64_2 = out2 ? make_tm_out(out2->p, out_rows, out2->c, fmt, 64) : L.tm_o64;
  }
  // Small-batch twin.  At batch 1 a 17x17 layer is 3 M tiles and an 8x8 layer one: with 128..256-column work items a handful
  // of CTAs each pull the layer's whole weight matrix (up to 1.5 MB) through one SM's ~47 B/cycle L2 port while 140 SMs idle —
  // the K loop, not the launch, is then most of the layer's ~9 us.  The twin cuts the output channels into <= 64-column items
  // (same packed weights, other tensor-map boxes): 2-4x the CTAs, each streaming 1/2-1/4 of the weights.  launch_conv picks
  // it when the wide layer would occupy less than half the SMs.  ICP_NARROW_NTILE=0 disables, =N sets the item width.
  // MEASURED (u8 in, fp16, CUDA graph + PDL, device ms per batch; wide / 16 / 32 / 64 columns):
  //   n=1 0.452 / 0.374 / 0.370 / 0.371   n=2 0.482 / 0.414 / 0.383 / 0.379   n=4 0.502 / 0.536 / 0.448 / 0.412
  //   n=8 0.551 / 0.657 / 0.531 / 0.484   n=16 0.643 / 0.974 / 0.741 / 0.646
  static const int narrow_nt = getenv("ICP_NARROW_NTILE") ? atoi(getenv("ICP_NARROW_NTILE")) : 64;
  if (!narrow_of && persistent && narrow_nt >= 16 && n_tile > narrow_nt) {
    int nt = 0;
    for (int c = narrow_nt / 16 * 16; c >= 16; c -= 16)
      if (cout % c == 0) { nt = c; break; }
    if (nt)
      L.narrow = std::make_shared<ConvLayer>(make_conv(pool, name, max_n, in, in_coff, cin, w_hwio, kh, kw, cout, stride, same, scale,
                                                       shift, relu, out, out_coff, fmt, floor, out2, out2_coff, n_split,
                                                       transpose_out, false, &L, nt));
  }
  return L;
}

// Clusters of `cs` persistent CTAs (one per SM, `smem` dynamic bytes) the device can hold at once: GPCs whose SM count is
// not a multiple of the cluster size strand SMs.  Cached per (device, cs, smem).
static int max_active_clusters(int cs, size_t smem) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, size_t>, int> cache;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(dev, cs, smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(cs * 64));
  cfg.blockDim = dim3(kPersistThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cs == 8 ? cudaOccupancyMaxActiveClusters(&n, conv_igemm_persistent_kernel<8>, &cfg)
                          : cudaOccupancyMaxActiveClusters(&n, conv_igemm_persistent_kernel<2>, &cfg);
  if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = g_num_sms / cs; }
  if (getenv("ICP_DEBUG_CLUSTERS")) fprintf(stderr, "[icp] max active clusters of %d (smem %zu): %d\n", cs, smem, n);
  cache[key] = n;
  return n;
}

static void launch_conv(const ConvLayer& L, int n, cudaStream_t st, bool pdl = false) {
  if (L.pair) {
    if (n < L.pair_min_batch && L.alt) { launch_conv(*L.alt, n, st, pdl); return; }
    PairArgs a = L.pargs;
    a.n_images = n;
    const int items = ((n + 1) / 2) * a.bands_per_image;             // per N tile
    int nclusters = (g_num_sms / 2) / a.n_tiles * a.n_tiles;
    if (items * a.n_tiles < nclusters) nclusters = items * a.n_tiles;
    launch_k(conv_pair_kernel, 2 * nclusters, kPairThreads, L.smem, st, pdl, 0, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_b_tail, L.tm_bias, a);
    return;
  }
  if (L.halo) {
    if (L.alt && n < L.pair_min_batch) { launch_conv(*L.alt, n, st, pdl); return; }
    HaloArgs h = L.hargs;
    h.n_images = n;
    h.pdl_late = pdl_late() ? 1 : 0;
    const int bands = h.tall ? (n * h.oh * h.wp + kBM * h.T - 1) / (kBM * h.T) : h.bands_per_image * n;
    const int items = bands * h.n_tiles;
    unsigned grid = static_cast<unsigned>(items < g_num_sms ? items : g_num_sms);
    if (h.w_resident && h.n_tiles > 1) grid = grid / h.n_tiles * h.n_tiles;       // static N tile per CTA: b % n_tiles
    const bool stem = h.w_resident && !h.tall && !h.transpose_out;
    if (stem && L.halo_cp) {
      if (h.epi_direct) launch_k(conv_halo_kernel<true, true, true>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
      else launch_k(conv_halo_kernel<false, true, true>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
      return;
    }
    if (h.epi_direct) {
      if (stem) launch_k(conv_halo_kernel<true, true>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
      else launch_k(conv_halo_kernel<true, false>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
    } else {
      if (stem) launch_k(conv_halo_kernel<false, true>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
      else launch_k(conv_halo_kernel<false, false>, grid, kHaloThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_a_tail, L.tm_b, L.tm_bias, h);
    }
    return;
  }
  if (L.narrow) {
    const int mt = (n * L.pix_per_image + kBM - 1) / kBM;
    if (2 * mt * L.n_tiles <= g_num_sms) { launch_conv(*L.narrow, n, st, pdl); return; }
  }
  ConvArgs a = L.args;
  a.m_total = n * L.pix_per_image;
  static const int pdl_point = getenv("ICP_PDL_POINT") ? atoi(getenv("ICP_PDL_POINT")) : 1;   // 2: before the last tile (experiment)
  a.pdl_late = pdl_late() ? pdl_point : 0;
  const int m_tiles = (a.m_total + kBM - 1) / kBM;
  if (L.persistent) {
    if (L.cluster && m_tiles >= L.cluster_min_tiles * g_num_sms) {
      const int cs = L.cluster_size;
      const int items = ((m_tiles + cs - 1) / cs) * L.n_tiles;
      int max_clusters = max_active_clusters(cs, L.smem);
      if (max_clusters > g_num_sms / cs) max_clusters = g_num_sms / cs;
      const unsigned grid = static_cast<unsigned>(cs * (items < max_clusters ? items : max_clusters));
      if (cs == 8)
        launch_k(conv_igemm_persistent_kernel<8>, grid, kPersistThreads, L.smem, st, pdl, 8, L.tm_a, L.tm_b, L.tm_a_tail,
                 L.tm_b_tail, L.tm_b_half, L.tm_b_tail_half, L.tm_bias, L.tm_o, L.tm_o2, L.tm_o64, L.tm_o64_2, a, m_tiles, L.n_tiles);
      else
        launch_k(conv_igemm_persistent_kernel<2>, grid, kPersistThreads, L.smem, st, pdl, 2, L.tm_a, L.tm_b, L.tm_a_tail,
                 L.tm_b_tail, L.tm_b_half, L.tm_b_tail_half, L.tm_bias, L.tm_o, L.tm_o2, L.tm_o64, L.tm_o64_2, a, m_tiles, L.n_tiles);
      return;
    }
    // dual-tile items when the launch has between ICP_DUAL_MIN and ICP_DUAL_MAX tiles per SM (default 1.5 .. 3: the 17x17 stage
    // at batch 128) and the layer has a real filter footprint; ICP_DUAL=0 disables
    static const bool dual_on = getenv("ICP_DUAL") ? atoi(getenv("ICP_DUAL")) != 0 : false;
    static const double dual_min = getenv("ICP_DUAL_MIN") ? atof(getenv("ICP_DUAL_MIN")) : 1.5;
    static const double dual_max = getenv("ICP_DUAL_MAX") ? atof(getenv("ICP_DUAL_MAX")) : 3.0;
    static const int dual_taps = getenv("ICP_DUAL_MIN_TAPS") ? atoi(getenv("ICP_DUAL_MIN_TAPS")) : 2;
    const double per_sm = static_cast<double>(m_tiles) * L.n_tiles / g_num_sms;
    if (dual_on && L.dual_stages > 0 && a.taps >= dual_taps && per_sm >= dual_min && per_sm <= dual_max) {
      a.dual = 1;
      a.epi_teams = 2;
      a.stages = L.dual_stages;
      const int items = ((m_tiles + 1) / 2) * L.n_tiles;
      launch_k(conv_igemm_persistent_kernel<1>, items < g_num_sms ? items : g_num_sms, kPersistThreads, L.dual_smem, st, pdl, 1,
               L.tm_a, L.tm_b, L.tm_a_tail, L.tm_b_tail, L.tm_b_half, L.tm_b_tail_half, L.tm_bias, L.tm_o, L.tm_o2, L.tm_o64,
               L.tm_o64_2, a, m_tiles, L.n_tiles);
      return;
    }
    const int total = m_tiles * L.n_tiles;
    launch_k(conv_igemm_persistent_kernel<1>, total < g_num_sms ? total : g_num_sms, kPersistThreads, L.smem, st, pdl, 1,
             L.tm_a, L.tm_b, L.tm_a_tail, L.tm_b_tail, L.tm_b_half, L.tm_b_tail_half, L.tm_bias, L.tm_o, L.tm_o2, L.tm_o64,
             L.tm_o64_2, a, m_tiles, L.n_tiles);
  } else {
    launch_k(conv_igemm_kernel, dim3(m_tiles, L.n_tiles), kConvThreads, L.smem, st, pdl, 1, L.tm_a, L.tm_b, L.tm_a_tail,
             L.tm_b_tail, L.tm_bias, a);
  }
}

// ------------------------------------------------------------------------------------------ plan ops
struct Op {
  enum Kind { CONV, MAXPOOL, AVGPOOL, FORK, JOIN } kind;
  int lane = 0;                   // independent branches of a block run on their own stream between FORK and JOIN
  ConvLayer conv;                 // CONV
  Act in, out;                    // pools
  int c = 0, in_coff = 0, out_coff = 0;   // pools: channels pooled, channel offsets in `in` / `out`
  int k = 3, stride = 1, pad = 0;
  bool post_relu = false;         // AVGPOOL after the pool branch's 1x1 projection + BN: that layer's ReLU
  std::string name;
};

struct Endpoint {
  Act act;
  int coff = 0, c = 0;      // 4-D taps
  const float* vec = nullptr;   // device float [max_batch, c] taps (logits / predictions / aux_logits)
};

}  // namespace icp

using namespace icp;

// ------------------------------------------------------------------------------------------ JPEG decode (nvJPEG)
// The reference's graph starts at the JPEG string (tf.image.decode_jpeg, PyClassification.py:150-151).  nvJPEG
// decodes on the GPU straight into the staging buffer the crop/resize kernel reads: no host-side pixel buffer,
// no host->device copy of decoded pixels.  Backends are tried in the order hardware engine -> GPU-assisted
// Huffman -> hybrid (CPU Huffman + GPU IDCT); ICP_NVJPEG_BACKEND=<n> pins one (nvjpegBackend_t value).
struct JpegDecoder {
  nvjpegHandle_t h = nullptr;
  nvjpegJpegState_t st = nullptr;
  int backend = -1;
  std::vector<int> candidates;
  void destroy() {
    if (st) nvjpegJpegStateDestroy(st);
    if (h) nvjpegDestroy(h);
    st = nullptr; h = nullptr; backend = -1;
  }
  bool open_next() {      // false when no backend is left
    destroy();
    while (!candidates.empty()) {
      const int b = candidates.front();
      candidates.erase(candidates.begin());
      static const bool interp = getenv("ICP_NVJPEG_NO_INTERP") == nullptr;   // chroma upsampling with interpolation ~ libjpeg "fancy upsampling"
      if (nvjpegCreateEx(static_cast<nvjpegBackend_t>(b), nullptr, nullptr, interp ? NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION : 0, &h) !=
          NVJPEG_STATUS_SUCCESS) { h = nullptr; continue; }
      if (nvjpegJpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) { nvjpegDestroy(h); h = nullptr; st = nullptr; continue; }
      backend = b;
      return true;
    }
    return false;
  }
  void init() {
    if (h || !candidates.empty()) return;
    if (const char* v = getenv("ICP_NVJPEG_BACKEND")) candidates = {atoi(v)};
    else candidates = {NVJPEG_BACKEND_HARDWARE, NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_HYBRID, NVJPEG_BACKEND_DEFAULT};
    open_next();
  }
  ~JpegDecoder() { destroy(); }
};

// ------------------------------------------------------------------------------------------ engine
struct GraphKey {
  int n;
  const void* x;
  bool u8;
  const void* probs;
  int k;
  bool operator<(const GraphKey& o) const {
    return std::tie(n, x, u8, probs, k) < std::tie(o.n, o.x, o.u8, o.probs, o.k);
  }
};

struct icp_engine {
  icp_config cfg{};
  std::string weights_path;
  int fmt = 1;
  int S = 299;
  int C = 0;
  int max_batch = 0;
  mutable std::string err;
  mutable std::mutex mu;            // one call at a time per engine (the reference's tf.Session.run is thread-safe)
  DevicePool pool;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int last_n = 0;

  float* d_input = nullptr;         // [max_batch, S, S, 3] f32
  uint8_t* d_input_u8 = nullptr;    // [max_batch, S, S, 3] u8 (icp_forward_u8); also slot 0 of the async path
  uint8_t* d_slot_u8[2] = {nullptr, nullptr};   // icp_submit_u8 double buffer
  cudaStream_t copy_stream = nullptr;           // H2D of the next batch overlaps the tower of the current one
  cudaEvent_t slot_h2d[2] = {nullptr, nullptr}, slot_done[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  int graph_k = 5;
  uint8_t* d_rgb = nullptr;         // staging for icp_classify_rgb8 / icp_classify_jpeg / the batch calls
  size_t d_rgb_cap = 0;
  JpegDecoder jpeg;                 // created on the first icp_classify_jpeg call
  Conv0Params conv0{};              // weights + folded BN of the Cin=3 stem conv, passed by value (FP32-pipe kernel)
  Conv0TcParams conv0_tc_f{}, conv0_tc_u8{};   // tensor-core conv0: 16-bit weights (+ 2/255 for uint8 input), shift
  bool conv0_tc = true;
  Act conv0_out;
  std::vector<Op> ops;              // everything after conv0 up to the last block (incl. FORK/JOIN markers)
  std::vector<int> launch_ops;      // indices into `ops` of the entries that launch a kernel
  Act final_act;                    // 8x8x2048
  float *d_fc_w = nullptr, *d_fc_b = nullptr;
  float *d_logits = nullptr, *d_probs = nullptr, *d_topk_val = nullptr;
  float *d_feat = nullptr, *d_part = nullptr;       // head: pooled features [n][2048], partial logits [n][8][C<=32]
  int* d_topk_idx = nullptr;
  // aux head
  std::vector<Op> aux_ops;
  Act aux_feat;                     // 1x1x768
  float *d_aux_w = nullptr, *d_aux_b = nullptr, *d_aux_logits = nullptr, *d_aux_feat = nullptr, *d_aux_part = nullptr;
  std::map<std::string, Endpoint> endpoints;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  static constexpr int kLanes = 5;
  cudaStream_t lane_stream[kLanes] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [0] unused (= caller's stream)
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t join_ev[kLanes] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t stem_ops = 0;              // leading entries of `ops` that form the stem (conv1 .. pool2)
  int fuse_pool_op = -1, fuse_conv_op = -1;   // pool1 / conv3: one pool_conv_kernel launch for batches >= fuse_min_batch
  int fuse_min_batch = 16;
  bool last_fused = false;          // the most recent forward did not materialise pool1
  int stem_chunk = 0;               // images per pass of the stem (0 = whole batch at once), see enqueue_tower
  int lanes_max_batch = 1 << 30;    // measured faster at every batch size (0.86 -> 0.61 ms at 1, 3.37 -> 3.23 ms at 128)
  int launches = 0;
  long long conv_macs_per_image = 0;

  ~icp_engine() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    for (int i = 0; i < kLanes; ++i) {
      if (lane_stream[i]) cudaStreamDestroy(lane_stream[i]);
      if (join_ev[i]) cudaEventDestroy(join_ev[i]);
    }
    if (fork_ev) cudaEventDestroy(fork_ev);
    for (int i = 0; i < 2; ++i) {
      if (slot_h2d[i]) cudaEventDestroy(slot_h2d[i]);
      if (slot_done[i]) cudaEventDestroy(slot_done[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    if (d_rgb) cudaFree(d_rgb);
  }
};

namespace icp {

struct Builder {
  icp_engine& e;
  const Blob& blob;
  int fmt;
  int maxn;
  int lane = 0;        // lane of the ops being emitted

  void fork(std::vector<Op>& ops) { Op op; op.kind = Op::FORK; ops.push_back(op); }
  void join(std::vector<Op>& ops) { Op op; op.kind = Op::JOIN; ops.push_back(op); lane = 0; }

  Act alloc_act(int h, int w, int c) {
    Act a;
    a.h = h; a.w = w; a.c = c;
    a.p = static_cast<uint16_t*>(e.pool.alloc(static_cast<size_t>(maxn) * h * w * c * 2));
    return a;
  }
  void bn_fold(const std::string& scope, int cout, std::vector<float>& scale, std::vector<float>& shift) {
    // ops.batch_norm inference branch (ops.py:122-128): no gamma, eps inside the rsqrt
    const HostTensor& beta = blob_get(blob, scope + "/BatchNorm/beta", {cout});
    const HostTensor& mean = blob_get(blob, scope + "/BatchNorm/moving_mean", {cout});
    const HostTensor& var = blob_get(blob, scope + "/BatchNorm/moving_variance", {cout});
    scale.resize(cout);
    shift.resize(cout);
    for (int i = 0; i < cout; ++i) {
      const float r = 1.0f / sqrtf(var.data[i] + 0.001f);
      scale[i] = r;
      shift[i] = beta.data[i] - mean.data[i] * r;
    }
  }
  // conv + BN + ReLU writing channels [coff, coff+cout) of `out`
  void conv(std::vector<Op>& ops, const std::string& scope, const Act& in, int cout, int kh, int kw, int stride,
            bool same, const Act& out, int coff) {
    const HostTensor& w = blob_get(blob, scope + "/weights", {kh, kw, in.c, cout});
    std::vector<float> scale, shift;
    bn_fold(scope, cout, scale, shift);
    Op op;
    op.kind = Op::CONV;
    op.conv = make_conv(e.pool, scope, maxn, in, 0, in.c, w.data.data(), kh, kw, cout, stride, same, scale.data(),
                        shift.data(), true, out, coff, fmt);
    e.conv_macs_per_image += op.conv.macs_per_image;
    op.lane = lane;
    ops.push_back(std::move(op));
  }
  Act conv_new(std::vector<Op>& ops, const std::string& scope, const Act& in, int cout, int kh, int kw, int stride = 1,
               bool same = true) {
    const int ph = same ? (kh - 1) / 2 : 0, pw = same ? (kw - 1) / 2 : 0;
    Act out = alloc_act((in.h + 2 * ph - kh) / stride + 1, (in.w + 2 * pw - kw) / stride + 1, cout);
    conv(ops, scope, in, cout, kh, kw, stride, same, out, 0);
    return out;
  }
  void maxpool(std::vector<Op>& ops, const Act& in, const Act& out, int coff) {
    Op op;
    op.kind = Op::MAXPOOL;
    op.in = in; op.out = out; op.c = in.c; op.out_coff = coff;
    op.lane = lane;
    ops.push_back(std::move(op));
  }
  Act avgpool_new(std::vector<Op>& ops, const Act& in, int k, int stride, bool same) {
    const int pad = same ? (k - 1) / 2 : 0;
    Act out = alloc_act((in.h + 2 * pad - k) / stride + 1, (in.w + 2 * pad - k) / stride + 1, in.c);
    Op op;
    op.kind = Op::AVGPOOL;
    op.in = in; op.out = out; op.c = in.c; op.k = k; op.stride = stride; op.pad = pad;
    op.lane = lane;
    ops.push_back(std::move(op));
    return out;
  }
  void tap(const std::string& name, const Act& a) {
    Endpoint ep;
    ep.act = a; ep.coff = 0; ep.c = a.c;
    e.endpoints[name] = ep;
  }

  // ---- horizontally fused sibling 1x1 convs -------------------------------------------------------
  // All 1x1 convs of a block read the same input, so they run as ONE GEMM over the concatenated output
  // channels: [branch1x1 | branch temps ... | pool-branch projection].  branch1x1's columns land in the
  // block's concat buffer, the rest in a fused temp buffer whose channel slices feed the k>1 convs.
  // The pool branch is projected FIRST (conv + BatchNorm affine, no ReLU) and average-pooled afterwards, with
  // its ReLU applied on the pooled value: avg_pool (exclude-pad divisor) commutes with a 1x1 conv and with a
  // per-channel affine (inception_model.py:122-124 order is pool -> conv -> BN -> relu), and pooling 32..192
  // projected channels instead of 192..2048 input channels removes ~80 % of the pool traffic.
  struct Sib {
    std::string scope;
    int cout;
    bool pool_proj;     // true: keep linear (BN/ReLU deferred to the pool)
  };
  struct Fused {
    Act temp;                      // fused temp buffer (channels after the first destination)
    std::vector<int> temp_off;     // channel offset of sibling i inside `temp` (-1 for the first-destination sibling)
  };
  // sibs[0] goes to (out, out_coff) when `first_to_out`, everything else to the fused temp.
  Fused fused_1x1(std::vector<Op>& ops, const std::string& block, const Act& x, const std::vector<Sib>& sibs,
                  bool first_to_out, const Act& out, int out_coff) {
    int total = 0;
    for (const Sib& sb : sibs) total += sb.cout;
    const int cin = x.c;
    std::vector<float> w(static_cast<size_t>(cin) * total), scale(total), shift(total), floor(total);
    Fused f;
    int col = 0, tcol = 0;
    std::string name = block + "/{";
    for (size_t i = 0; i < sibs.size(); ++i) {
      const Sib& sb = sibs[i];
      const HostTensor& wi = blob_get(blob, sb.scope + "/weights", {1, 1, cin, sb.cout});
      for (int ci = 0; ci < cin; ++ci)
        memcpy(&w[static_cast<size_t>(ci) * total + col], &wi.data[static_cast<size_t>(ci) * sb.cout], sb.cout * 4);
      std::vector<float> sc, sh;
      bn_fold(sb.scope, sb.cout, sc, sh);
      for (int j = 0; j < sb.cout; ++j) {
        // pool projection: the exclude-pad average of a constant is that constant, so the WHOLE BatchNorm affine
        // commutes with the pool and is folded in here (scale into the 16-bit weights, shift into the fp32 accumulator
        // through the bias MMA) exactly as for every other conv; only the ReLU waits for the pooled value.  (r01
        // stored the un-normalised projection in 16 bits and applied the affine after the pool, which amplifies the
        // storage rounding by |mean|/sigma on trained statistics — ADVICE r01.)
        scale[col + j] = sc[j];
        shift[col + j] = sh[j];
        floor[col + j] = sb.pool_proj ? -INFINITY : 0.0f;
      }
      if (i == 0 && first_to_out) {
        f.temp_off.push_back(-1);
      } else {
        f.temp_off.push_back(tcol);
        tcol += sb.cout;
      }
      col += sb.cout;
      name += (i ? "," : "") + sb.scope.substr(block.size() + 1);
      e.conv_macs_per_image += static_cast<long long>(x.h) * x.w * cin * sb.cout;
    }
    name += "}[fused 1x1]";
    f.temp = alloc_act(x.h, x.w, tcol);
    Op op;
    op.kind = Op::CONV;
    if (first_to_out)
      op.conv = make_conv(e.pool, name, maxn, x, 0, cin, w.data(), 1, 1, total, 1, true, scale.data(), shift.data(),
                          true, out, out_coff, fmt, floor.data(), &f.temp, 0, sibs[0].cout);
    else
      op.conv = make_conv(e.pool, name, maxn, x, 0, cin, w.data(), 1, 1, total, 1, true, scale.data(), shift.data(),
                          true, f.temp, 0, fmt, floor.data());
    op.lane = lane;
    ops.push_back(std::move(op));
    return f;
  }
  // conv reading channels [in_coff, in_coff+cin) of a fused temp
  void conv_slice(std::vector<Op>& ops, const std::string& scope, const Act& in, int in_coff, int cin, int cout, int kh,
                  int kw, int stride, bool same, const Act& out, int coff) {
    const HostTensor& w = blob_get(blob, scope + "/weights", {kh, kw, cin, cout});
    std::vector<float> scale, shift;
    bn_fold(scope, cout, scale, shift);
    Op op;
    op.kind = Op::CONV;
    op.conv = make_conv(e.pool, scope, maxn, in, in_coff, cin, w.data.data(), kh, kw, cout, stride, same, scale.data(),
                        shift.data(), true, out, coff, fmt);
    e.conv_macs_per_image += op.conv.macs_per_image;
    op.lane = lane;
    ops.push_back(std::move(op));
  }
  Act conv_slice_new(std::vector<Op>& ops, const std::string& scope, const Act& in, int in_coff, int cin, int cout,
                     int kh, int kw) {
    Act out = alloc_act(in.h, in.w, cout);      // SAME, stride 1
    conv_slice(ops, scope, in, in_coff, cin, cout, kh, kw, 1, true, out, 0);
    return out;
  }
  // 3x3/1 SAME average pool of the projected (conv + BN) pool branch + its deferred ReLU, into the concat slice
  void pool_after_projection(std::vector<Op>& ops, const std::string& scope, const Fused& f, int temp_off, int c,
                             const Act& out, int coff) {
    Op op;
    op.kind = Op::AVGPOOL;
    op.name = scope + "[avg_pool + ReLU after 1x1 + BN]";
    op.in = f.temp; op.in_coff = temp_off; op.out = out; op.out_coff = coff; op.c = c;
    op.k = 3; op.stride = 1; op.pad = 1;
    op.post_relu = true;
    op.lane = lane;
    ops.push_back(std::move(op));
  }

  // mixed_35x35x* (inception_model.py:112-158)
  Act block35(std::vector<Op>& ops, const std::string& s, const Act& x, int pool_c) {
    Act out = alloc_act(x.h, x.w, 64 + 64 + 96 + pool_c);
    Fused f = fused_1x1(ops, s, x, {{s + "/branch1x1/Conv", 64, false}, {s + "/branch5x5/Conv", 48, false},
                                    {s + "/branch3x3dbl/Conv", 64, false}, {s + "/branch_pool/Conv", pool_c, true}},
                        true, out, 0);
    fork(ops);
    lane = 1;
    conv_slice(ops, s + "/branch5x5/Conv_1", f.temp, f.temp_off[1], 48, 64, 5, 5, 1, true, out, 64);
    lane = 0;
    Act t = conv_slice_new(ops, s + "/branch3x3dbl/Conv_1", f.temp, f.temp_off[2], 64, 96, 3, 3);
    conv(ops, s + "/branch3x3dbl/Conv_2", t, 96, 3, 3, 1, true, out, 128);
    lane = 2;
    pool_after_projection(ops, s + "/branch_pool", f, f.temp_off[3], pool_c, out, 224);
    join(ops);
    tap(s, out);
    return out;
  }
  // mixed_17x17x768a (:160-171)
  Act block_red_a(std::vector<Op>& ops, const std::string& s, const Act& x) {
    const int oh = (x.h - 3) / 2 + 1, ow = (x.w - 3) / 2 + 1;
    Act out = alloc_act(oh, ow, 384 + 96 + x.c);
    fork(ops);
    lane = 1;
    conv(ops, s + "/branch3x3/Conv", x, 384, 3, 3, 2, false, out, 0);
    lane = 0;
    Act t = conv_new(ops, s + "/branch3x3dbl/Conv", x, 64, 1, 1);
    t = conv_new(ops, s + "/branch3x3dbl/Conv_1", t, 96, 3, 3);
    conv(ops, s + "/branch3x3dbl/Conv_2", t, 96, 3, 3, 2, false, out, 384);
    lane = 2;
    maxpool(ops, x, out, 480);
    join(ops);
    tap(s, out);
    return out;
  }
  // mixed_17x17x768b..e (:173-247)
  Act block17(std::vector<Op>& ops, const std::string& s, const Act& x, int m) {
    Act out = alloc_act(x.h, x.w, 768);
    Fused f = fused_1x1(ops, s, x, {{s + "/branch1x1/Conv", 192, false}, {s + "/branch7x7/Conv", m, false},
                                    {s + "/branch7x7dbl/Conv", m, false}, {s + "/branch_pool/Conv", 192, true}},
                        true, out, 0);
    fork(ops);
    lane = 1;
    Act t = conv_slice_new(ops, s + "/branch7x7/Conv_1", f.temp, f.temp_off[1], m, m, 1, 7);
    conv(ops, s + "/branch7x7/Conv_2", t, 192, 7, 1, 1, true, out, 192);
    lane = 0;
    t = conv_slice_new(ops, s + "/branch7x7dbl/Conv_1", f.temp, f.temp_off[2], m, m, 7, 1);
    t = conv_new(ops, s + "/branch7x7dbl/Conv_2", t, m, 1, 7);
    t = conv_new(ops, s + "/branch7x7dbl/Conv_3", t, m, 7, 1);
    conv(ops, s + "/branch7x7dbl/Conv_4", t, 192, 1, 7, 1, true, out, 384);
    lane = 2;
    pool_after_projection(ops, s + "/branch_pool", f, f.temp_off[3], 192, out, 576);
    join(ops);
    tap(s, out);
    return out;
  }
  // mixed_17x17x1280a (:266-280)
  Act block_red_b(std::vector<Op>& ops, const std::string& s, const Act& x) {
    const int oh = (x.h - 3) / 2 + 1, ow = (x.w - 3) / 2 + 1;
    Act out = alloc_act(oh, ow, 320 + 192 + x.c);
    Fused f = fused_1x1(ops, s, x, {{s + "/branch3x3/Conv", 192, false}, {s + "/branch7x7x3/Conv", 192, false}}, false,
                        out, 0);
    fork(ops);
    lane = 1;
    conv_slice(ops, s + "/branch3x3/Conv_1", f.temp, f.temp_off[0], 192, 320, 3, 3, 2, false, out, 0);
    lane = 0;
    Act t = conv_slice_new(ops, s + "/branch7x7x3/Conv_1", f.temp, f.temp_off[1], 192, 192, 1, 7);
    t = conv_new(ops, s + "/branch7x7x3/Conv_2", t, 192, 7, 1);
    conv(ops, s + "/branch7x7x3/Conv_3", t, 192, 3, 3, 2, false, out, 320);
    lane = 2;
    maxpool(ops, x, out, 512);
    join(ops);
    tap(s, out);
    return out;
  }
  // mixed_8x8x2048a/b (:282-316); nested concats become adjacent channel slices
  Act block8(std::vector<Op>& ops, const std::string& s, const Act& x) {
    Act out = alloc_act(x.h, x.w, 2048);
    Fused f = fused_1x1(ops, s, x, {{s + "/branch1x1/Conv", 320, false}, {s + "/branch3x3/Conv", 384, false},
                                    {s + "/branch3x3dbl/Conv", 448, false}, {s + "/branch_pool/Conv", 192, true}},
                        true, out, 0);
    fork(ops);
    lane = 1;
    conv_slice(ops, s + "/branch3x3/Conv_1", f.temp, f.temp_off[1], 384, 384, 1, 3, 1, true, out, 320);
    lane = 2;
    conv_slice(ops, s + "/branch3x3/Conv_2", f.temp, f.temp_off[1], 384, 384, 3, 1, 1, true, out, 704);
    lane = 0;
    Act t = conv_slice_new(ops, s + "/branch3x3dbl/Conv_1", f.temp, f.temp_off[2], 448, 384, 3, 3);
    conv(ops, s + "/branch3x3dbl/Conv_2", t, 384, 1, 3, 1, true, out, 1088);
    conv(ops, s + "/branch3x3dbl/Conv_3", t, 384, 3, 1, 1, true, out, 1472);
    lane = 4;
    pool_after_projection(ops, s + "/branch_pool", f, f.temp_off[3], 192, out, 1856);
    join(ops);
    tap(s, out);
    return out;
  }

  void build() {
    const int S = e.S;
    // ---- stem (:86-107) ----
    {
      const HostTensor& w = blob_get(blob, "conv0/weights", {3, 3, 3, 32});
      std::vector<float> scale, shift;
      bn_fold("conv0", 32, scale, shift);
      memcpy(e.conv0.w, w.data.data(), sizeof e.conv0.w);
      memcpy(e.conv0.scale, scale.data(), sizeof e.conv0.scale);
      memcpy(e.conv0.shift, shift.data(), sizeof e.conv0.shift);
      // tensor-core version (conv0_tc.cuh): [cout][32] 16-bit weights with the BN scale folded in; for uint8 input
      // also the 2/255 of the normalisation, whose "- 1" goes into the shift as -127.5 * sum of the ROUNDED weights,
      // so the kernel computes exactly  sum W'(u - 127.5) + shift  on integer inputs
      for (int variant = 0; variant < 2; ++variant) {
        Conv0TcParams& tp = variant ? e.conv0_tc_u8 : e.conv0_tc_f;
        const float in_scale = variant ? 2.0f / 255.0f : 1.0f;
        for (int co = 0; co < 32; ++co) {
          double wsum = 0.0;
          for (int k = 0; k < 32; ++k) {
            const uint16_t h = k < 27 ? host_f2h(w.data[static_cast<size_t>(k) * 32 + co] * scale[co] * in_scale, fmt) : 0;
            tp.w[co * 32 + k] = h;
            wsum += host_h2f(h, fmt);
          }
          tp.shift[co] = variant ? static_cast<float>(shift[co] - 127.5 * wsum) : shift[co];
        }
      }
      e.conv0_tc = getenv("ICP_CONV0_FMA") == nullptr;
      const int o = (S - 3) / 2 + 1;
      e.conv0_out = alloc_act(o, o, 32);
      e.conv_macs_per_image += static_cast<long long>(o) * o * 32 * 27;
      tap("conv0", e.conv0_out);
    }
    std::vector<Op>& ops = e.ops;
    Act x = conv_new(ops, "conv1", e.conv0_out, 32, 3, 3, 1, false);
    tap("conv1", x);
    x = conv_new(ops, "conv2", x, 64, 3, 3, 1, true);
    tap("conv2", x);
    {
      Act p = alloc_act((x.h - 3) / 2 + 1, (x.w - 3) / 2 + 1, x.c);
      maxpool(ops, x, p, 0);
      x = p;
      tap("pool1", x);
    }
    e.fuse_pool_op = static_cast<int>(ops.size()) - 1;
    x = conv_new(ops, "conv3", x, 80, 1, 1, 1, false);
    e.fuse_conv_op = static_cast<int>(ops.size()) - 1;
    tap("conv3", x);
    x = conv_new(ops, "conv4", x, 192, 3, 3, 1, false);
    tap("conv4", x);
    {
      Act p = alloc_act((x.h - 3) / 2 + 1, (x.w - 3) / 2 + 1, x.c);
      maxpool(ops, x, p, 0);
      x = p;
      tap("pool2", x);
    }
    e.stem_ops = ops.size();
    x = block35(ops, "mixed_35x35x256a", x, 32);
    x = block35(ops, "mixed_35x35x288a", x, 64);
    x = block35(ops, "mixed_35x35x288b", x, 64);
    x = block_red_a(ops, "mixed_17x17x768a", x);
    x = block17(ops, "mixed_17x17x768b", x, 128);
    x = block17(ops, "mixed_17x17x768c", x, 160);
    x = block17(ops, "mixed_17x17x768d", x, 160);
    x = block17(ops, "mixed_17x17x768e", x, 192);
    if (e.cfg.enable_aux) {
      // aux_logits (:249-261): avg_pool 5x5/3 VALID -> 1x1 proj 128 -> full-size VALID conv 768 -> fc
      const long long keep = e.conv_macs_per_image;
      Act a = avgpool_new(e.aux_ops, x, 5, 3, false);
      a = conv_new(e.aux_ops, "aux_logits/proj", a, 128, 1, 1);
      a = conv_new(e.aux_ops, "aux_logits/Conv", a, 768, a.h, a.w, 1, false);
      e.conv_macs_per_image = keep;   // aux is not part of the classify_image fetch (F7)
      e.aux_feat = a;
      const HostTensor& w = blob_get(blob, "aux_logits/FC/weights", {768, e.C});
      const HostTensor& b = blob_get(blob, "aux_logits/FC/biases", {e.C});
      e.d_aux_w = static_cast<float*>(e.pool.alloc(w.data.size() * 4));
      e.d_aux_b = static_cast<float*>(e.pool.alloc(b.data.size() * 4));
      e.d_aux_logits = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * e.C * 4));
      e.d_aux_feat = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * 768 * 4));
      e.d_aux_part = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * (768 / kHeadSlice) * kHeadSmallC * 4));
      CK(cudaMemcpy(e.d_aux_w, w.data.data(), w.data.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(e.d_aux_b, b.data.data(), b.data.size() * 4, cudaMemcpyHostToDevice));
      Endpoint ep;
      ep.vec = e.d_aux_logits; ep.c = e.C;
      e.endpoints["aux_logits"] = ep;
    }
    x = block_red_b(ops, "mixed_17x17x1280a", x);
    x = block8(ops, "mixed_8x8x2048a", x);
    x = block8(ops, "mixed_8x8x2048b", x);
    e.final_act = x;
    // ---- head (:318-329) ----
    const HostTensor& w = blob_get(blob, "logits/logits/weights", {x.c, e.C});
    const HostTensor& b = blob_get(blob, "logits/logits/biases", {e.C});
    e.d_fc_w = static_cast<float*>(e.pool.alloc(w.data.size() * 4));
    e.d_fc_b = static_cast<float*>(e.pool.alloc(b.data.size() * 4));
    CK(cudaMemcpy(e.d_fc_w, w.data.data(), w.data.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e.d_fc_b, b.data.data(), b.data.size() * 4, cudaMemcpyHostToDevice));
    e.d_logits = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * e.C * 4));
    e.d_probs = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * e.C * 4));
    e.d_topk_val = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * 16 * 4));
    e.d_feat = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * x.c * 4));
    e.d_part = static_cast<float*>(e.pool.alloc(static_cast<size_t>(maxn) * (x.c / kHeadSlice) * kHeadSmallC * 4));
    e.d_topk_idx = static_cast<int*>(e.pool.alloc(static_cast<size_t>(maxn) * 16 * 4));
    Endpoint lg; lg.vec = e.d_logits; lg.c = e.C;
    Endpoint pr; pr.vec = e.d_probs; pr.c = e.C;
    e.endpoints["logits"] = lg;
    e.endpoints["predictions"] = pr;
  }
};

// Programmatic dependent launch between consecutive kernels of a stream, for batches up to ICP_PDL_MAX_BATCH (0 = never).
// WHERE the predecessor triggers matters.  At its start (ICP_PDL_LATE=0; then the default limit is batch 16): MEASURED at batch 128
// 2.347 ms/step against 2.324 without PDL — the successor's CTAs are placed on SMs as they free up and park there until the
// predecessor has finished, SMs that would otherwise run a kernel of another branch lane.  When a CTA has issued its last MMAs
// (ICP_PDL_LATE=1, the default, every batch size): the successor's launch latency and prologue overlap only the final epilogues —
// MEASURED 2.275 -> 2.257 ms at batch 128, 0.988 -> 0.938 at 32, 0.487 -> 0.477 at 8, 0.370 -> 0.368 at 1.
static thread_local int g_pdl_batch = 0;      // batch of the tower being enqueued
static bool pdl_enabled() {
  static const int max_batch = getenv("ICP_PDL_MAX_BATCH") ? atoi(getenv("ICP_PDL_MAX_BATCH")) : (pdl_late() ? (1 << 30) : 16);
  return g_pdl_batch <= max_batch;
}
// Which streams' most recent operation was one of our kernels (only then may the next launch on that stream be
// programmatic: behind an event wait, a memcpy or a timing event it keeps the full dependency).
struct PdlTracker {
  cudaStream_t s[8] = {};
  bool k[8] = {};
  int find(cudaStream_t st) {
    for (int i = 0; i < 8; ++i)
      if (s[i] == st || (!s[i] && !k[i])) { s[i] = st; return i; }
    return 0;
  }
  bool take(cudaStream_t st) {            // -> launch this kernel with the attribute?  (and note that a kernel follows)
    const int i = find(st);
    const bool r = k[i] && pdl_enabled();
    k[i] = true;
    return r;
  }
  void barrier(cudaStream_t st) { k[find(st)] = false; }     // an event wait / record-for-timing / copy was enqueued
};

static int run_op(const Op& op, int n, int fmt, cudaStream_t st, int out_image0 = 0, bool pdl = false) {
  switch (op.kind) {
    case Op::CONV:
      launch_conv(op.conv, n, st, pdl);
      return 1;
    case Op::MAXPOOL: {
      const unsigned per_image = static_cast<unsigned>(op.out.h) * ((op.out.w + 1) / 2) * (op.c / 8);
      launch_k(maxpool3s2_kernel, dim3((per_image + 255) / 256, n), 256, 0, st, pdl, 1, (const uint16_t*)op.in.p, n, op.in.h,
               op.in.w, op.c, op.in.c, 0, op.out.p + static_cast<size_t>(out_image0) * op.out.h * op.out.w * op.out.c,
               op.out.h, op.out.w, op.out.c, op.out_coff, fmt);
      return 1;
    }
    case Op::AVGPOOL: {
      const long long total = static_cast<long long>(n) * op.out.h * op.out.w * (op.c / 8);
      if (op.post_relu && op.k == 3 && op.stride == 1 && op.pad == 1) {
        // band of rows per CTA: the staged tile fits 100 KiB; more bands only when the batch alone cannot fill the SMs
        const int per_row = (op.in.w + 2) * op.c * 2;
        const int max_rows = 102400 / per_row - 2;
        if (max_rows < 1) fail(ICP_E_INVALID, "avg_pool tile of %d x %d channels does not fit shared memory", op.in.w, op.c);
        int bands = (op.in.h + max_rows - 1) / max_rows;
        while (n * bands < g_num_sms && (op.in.h + bands) / (bands + 1) >= 3) ++bands;
        const int band_rows = (op.in.h + bands - 1) / bands;
        bands = (op.in.h + band_rows - 1) / band_rows;
        const size_t smem = static_cast<size_t>(band_rows + 2) * per_row;
        const int cg = op.c / 8;
        int threads = (op.in.w * cg + 31) & ~31;
        if (threads > 512) threads = 512;
        const unsigned grid = static_cast<unsigned>(n * bands);
#define ICP_POOL_ARGS (const uint16_t*)op.in.p, op.in.h, op.in.w, op.c, op.in.c, op.in_coff, op.out.p, op.out.c, op.out_coff, band_rows, bands, 1, fmt
        if (cg == 4) launch_k(avgpool3_relu_kernel<4>, grid, threads, smem, st, pdl, 1, ICP_POOL_ARGS);
        else if (cg == 8) launch_k(avgpool3_relu_kernel<8>, grid, threads, smem, st, pdl, 1, ICP_POOL_ARGS);
        else if (cg == 24) launch_k(avgpool3_relu_kernel<24>, grid, threads, smem, st, pdl, 1, ICP_POOL_ARGS);
        else launch_k(avgpool3_relu_kernel<0>, grid, threads, smem, st, pdl, 1, ICP_POOL_ARGS);
#undef ICP_POOL_ARGS
      } else {
        launch_k(avgpool_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, pdl, 1, (const uint16_t*)op.in.p, n,
                 op.in.h, op.in.w, op.c, op.in.c, op.in_coff, op.out.p, op.out.h, op.out.w, op.out.c, op.out_coff, op.k,
                 op.stride, op.pad, (const float*)nullptr, (const float*)nullptr, 0, fmt);
      }
      return 1;
    }
  }
  return 0;
}

// Head: pooled features -> (partial) logits -> softmax -> top-k (tower_kernels.cuh); returns the launches made.
static int launch_head(icp_engine& e, const uint16_t* fmap, int hw, int F, int ld, const float* w, const float* b,
                       float* logits, float* probs, float* feat, float* part, int n, int k, cudaStream_t st, bool pdl) {
  if (F % kHeadSlice) fail(ICP_E_INVALID, "head: %d features is not a multiple of %d", F, kHeadSlice);
  const int C = e.C, slices = F / kHeadSlice;
  const bool small = C <= kHeadSmallC;
  const bool chain = pdl_enabled();
  int launches = 2;
  launch_k(head_pool_kernel, dim3(n, slices), 256, 0, st, pdl, 1, fmap, hw, F, ld, feat, w, C, part, e.fmt);
  if (!small) {
    launch_k(head_fc_kernel, dim3((C + 63) / 64, (n + 15) / 16), 256, 0, st, chain, 1, (const float*)feat, n, F, w, b, C, logits);
    ++launches;
  }
  launch_k(head_finish_kernel, n, 256, C * sizeof(float), st, chain, 1, (const float*)(small ? part : nullptr), slices, b, C,
           logits, probs, e.d_topk_val, e.d_topk_idx, probs ? k : 0);
  return launches;
}

// Enqueue the whole tower for n images whose preprocessed fp32 input sits at x_dev.
struct OpTimer {   // optional per-launch CUDA events (icp_profile_forward)
  std::vector<cudaEvent_t>* ev = nullptr;
  cudaStream_t st = nullptr;
  void mark() {
    if (!ev) return;
    cudaEvent_t x;
    CK(cudaEventCreate(&x));
    ev->push_back(x);
    CK(cudaEventRecord(x, st));
  }
};

static int enqueue_tower(icp_engine& e, const void* x_dev, bool x_is_u8, int n, int k, float* probs_dev,
                         cudaStream_t st, OpTimer* tm = nullptr) {
  int launches = 0;
  const int fmt = e.fmt;
  g_pdl_batch = n;
  PdlTracker pdl;
  auto mark = [&](cudaStream_t s_) {
    if (tm) { tm->mark(); pdl.barrier(s_); }
  };
  mark(st);
  // Optional (ICP_STEM_CHUNK=<images>, default off): run the stem in passes over the SAME chunk-sized region of the
  // intermediate buffers so that each producer->consumer pair (<= 88 MB at 32 images) stays in the 126 MB L2
  // instead of making a round trip through HBM.  MEASURED at batch 128: 2.544 ms whole batch, 2.59 / 2.63 / 2.73 ms
  // in passes of 64 / 32 / 16 — the stem kernels are not HBM-bound and the shorter launches quantise worse.
  // Only the first op's input and the last op's output are indexed by the pass's first image (stem endpoints
  // then hold the last pass only; per-op profiling always runs the whole batch).
  const int chunk = (!tm && e.stem_chunk > 0 && n > e.stem_chunk) ? e.stem_chunk : n;
  for (int i0 = 0; i0 < n; i0 += chunk) {
    const int nc = n - i0 < chunk ? n - i0 : chunk;
    const Act& o = e.conv0_out;
    const long long pix = static_cast<long long>(nc) * o.h * o.w;
    const unsigned grid = static_cast<unsigned>((pix + 127) / 128);
    const size_t in_off = static_cast<size_t>(i0) * e.S * e.S * 3;
    // conv0 follows a copy / another stream's work: full dependency (its successors chain programmatically)
    const bool p0 = i0 > 0 && pdl.take(st);
    if (i0 == 0) pdl.take(st);
    if (e.conv0_tc) {
      const unsigned tgrid = grid < static_cast<unsigned>(g_num_sms) ? grid : static_cast<unsigned>(g_num_sms);
      if (x_is_u8)
        launch_k(conv0_tc_kernel<uint8_t>, tgrid, kConv0TcThreads, conv0_tc_smem_bytes(), st, p0, 1,
                 static_cast<const uint8_t*>(x_dev) + in_off, nc, e.S, e.S, e.conv0_tc_u8, o.p, o.h, o.w, fmt, pdl_late() ? 1 : 0);
      else
        launch_k(conv0_tc_kernel<float>, tgrid, kConv0TcThreads, conv0_tc_smem_bytes(), st, p0, 1,
                 static_cast<const float*>(x_dev) + in_off, nc, e.S, e.S, e.conv0_tc_f, o.p, o.h, o.w, fmt, pdl_late() ? 1 : 0);
    } else if (x_is_u8) {
      conv0_kernel<uint8_t><<<grid, 128, 0, st>>>(static_cast<const uint8_t*>(x_dev) + in_off, nc, e.S, e.S, e.conv0,
                                                  o.p, o.h, o.w, fmt);
      pdl.barrier(st);            // conv0_kernel has no griddepcontrol.launch: its successor keeps the full dependency
    } else {
      conv0_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(x_dev) + in_off, nc, e.S, e.S, e.conv0, o.p,
                                                o.h, o.w, fmt);
      pdl.barrier(st);
    }
    CK(cudaGetLastError());
    ++launches;
    mark(st);
    // pool1 -> conv3 as ONE kernel (pool_conv.cuh) for throughput batches: the pooled tensor is never written
    const Op* fc = e.fuse_conv_op >= 0 ? &e.ops[e.fuse_conv_op] : nullptr;
    const bool fuse = !tm && chunk == n && fc && nc >= e.fuse_min_batch && fc->conv.d_w && !fc->conv.halo && !fc->conv.pair &&
                      fc->conv.args.taps == 1 && fc->conv.args.cblocks == 1 && fc->conv.cout <= kPoolConvMaxCout &&
                      e.ops[e.fuse_pool_op].in.c == 64 && e.ops[e.fuse_pool_op].c == 64;
    e.last_fused = fuse;
    for (size_t i = 0; i < e.stem_ops; ++i) {
      if (fuse && static_cast<int>(i) == e.fuse_pool_op) continue;
      if (fuse && static_cast<int>(i) == e.fuse_conv_op) {
        const Op& po = e.ops[e.fuse_pool_op];
        PoolConvArgs pa{};
        pa.in = po.in.p; pa.N = nc; pa.H = po.in.h; pa.W = po.in.w;
        pa.w = fc->conv.d_w; pa.bias16 = fc->conv.d_bias16; pa.cout = fc->conv.cout;
        pa.out = static_cast<uint16_t*>(fc->conv.args.out); pa.OH = po.out.h; pa.OW = po.out.w; pa.ldo = fc->conv.args.ldo;
        pa.coff = fc->conv.args.coff; pa.relu = fc->conv.args.relu_hi > 0 ? 1 : 0; pa.fmt = fmt;
        pa.pdl_late = pdl_late() ? 1 : 0;
        const int tiles = (nc * pa.OH * pa.OW + kBM - 1) / kBM;
        launch_k(pool_conv_kernel, tiles < g_num_sms ? tiles : g_num_sms, kPoolConvThreads, pool_conv_smem_bytes(), st, pdl.take(st), 1, pa);
        ++launches;
        continue;
      }
      launches += run_op(e.ops[i], nc, fmt, st, i + 1 == e.stem_ops ? i0 : 0, pdl.take(st));
      mark(st);
    }
  }
  // Independent branches of an inception block run on side streams between FORK and JOIN markers when one
  // layer cannot fill the GPU (small batches); captured into the same CUDA graph as parallel branches.
  const bool use_lanes = !tm && n <= e.lanes_max_batch;
  bool lane_live[icp_engine::kLanes] = {false, false, false, false, false};
  for (size_t oi = e.stem_ops; oi < e.ops.size(); ++oi) {
    const Op& op = e.ops[oi];
    if (op.kind == Op::FORK) {
      if (use_lanes) CK(cudaEventRecord(e.fork_ev, st));
      continue;
    }
    if (op.kind == Op::JOIN) {
      for (int i = 1; i < icp_engine::kLanes; ++i)
        if (lane_live[i]) {
          CK(cudaEventRecord(e.join_ev[i], e.lane_stream[i]));
          CK(cudaStreamWaitEvent(st, e.join_ev[i], 0));
          // The next block's first kernel keeps its programmatic edge to this stream's last kernel (the block's longest
          // branch runs on this stream) next to the full edges from the side lanes; ICP_PDL_JOIN=0 drops it.
          static const bool pdl_join = getenv("ICP_PDL_JOIN") ? atoi(getenv("ICP_PDL_JOIN")) != 0 : true;
          if (!pdl_join) pdl.barrier(st);
          lane_live[i] = false;
        }
      continue;
    }
    cudaStream_t os = st;
    if (use_lanes && op.lane > 0) {
      os = e.lane_stream[op.lane];
      if (!lane_live[op.lane]) {
        CK(cudaStreamWaitEvent(os, e.fork_ev, 0));
        pdl.barrier(os);
        lane_live[op.lane] = true;
      }
    }
    launches += run_op(op, n, fmt, os, 0, pdl.take(os));
    mark(os);
  }
  if (e.cfg.enable_aux && !tm) {
    for (const Op& op : e.aux_ops) launches += run_op(op, n, fmt, st, 0, pdl.take(st));
    launches += launch_head(e, e.aux_feat.p, 1, 768, e.aux_feat.c, e.d_aux_w, e.d_aux_b, e.d_aux_logits, nullptr,
                            e.d_aux_feat, e.d_aux_part, n, 0, st, pdl.take(st));
  }
  const Act& f = e.final_act;
  launches += launch_head(e, f.p, f.h * f.w, f.c, f.c, e.d_fc_w, e.d_fc_b, e.d_logits, probs_dev ? probs_dev : e.d_probs,
                          e.d_feat, e.d_part, n, k, st, pdl.take(st));
  mark(st);
  return launches;
}

static void run_tower(icp_engine& e, const void* x_dev, bool x_is_u8, int n, int k, float* probs_dev,
                      cudaStream_t st, bool timed) {
  if (n < 1 || n > e.max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e.max_batch);
  if (k < 0 || k > 16 || k > e.C) fail(ICP_E_INVALID, "k=%d outside [0, min(16, %d)]", k, e.C);
  e.last_n = n;
  if (timed) CK(cudaEventRecord(e.ev0, st));
  if (e.cfg.use_cuda_graph) {
    // one captured graph per distinct (batch, input buffer, output buffer, k): replay removes the
    // ~110 launch latencies that dominate small batches
    const GraphKey key{n, x_dev, x_is_u8, probs_dev, k};
    auto it = e.graphs.find(key);
    if (it == e.graphs.end()) {
      if (e.graphs.size() >= 64) {
        for (auto& g : e.graphs) cudaGraphExecDestroy(g.second);
        e.graphs.clear();
      }
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      try {
        e.launches = enqueue_tower(e, x_dev, x_is_u8, n, k, probs_dev, st);
      } catch (...) {
        cudaStreamEndCapture(st, &g);
        if (g) cudaGraphDestroy(g);
        throw;
      }
      CK(cudaStreamEndCapture(st, &g));
      cudaGraphExec_t ge = nullptr;
      CK(cudaGraphInstantiate(&ge, g, 0));
      cudaGraphDestroy(g);
      it = e.graphs.emplace(key, ge).first;
    }
    CK(cudaGraphLaunch(it->second, st));
  } else {
    e.launches = enqueue_tower(e, x_dev, x_is_u8, n, k, probs_dev, st);
  }
  if (timed) CK(cudaEventRecord(e.ev1, st));
}

template <class F>
static int guard(const icp_engine* e, F&& f) {
  std::unique_lock<std::mutex> lock;
  if (e) lock = std::unique_lock<std::mutex>(e->mu);
  try {
    f();
    return ICP_OK;
  } catch (const Fail& x) {
    if (e) e->err = x.msg;
    g_tls_error = x.msg;
    return x.code;
  } catch (const std::bad_alloc&) {
    if (e) e->err = "out of host memory";
    g_tls_error = "out of host memory";
    return ICP_E_NOMEM;
  } catch (const std::exception& x) {
    if (e) e->err = x.what();
    g_tls_error = x.what();
    return ICP_E_INVALID;
  }
}

// Function attributes (opt-in to > 48 KiB of dynamic shared memory) are per device/context, so they are set
// every time a device is selected for building — not once per process (a second engine on another
// device would otherwise launch without the opt-in).
static void set_kernel_attrs(int device) {
  const int big = 227 * 1024;
  CK(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_igemm_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_igemm_persistent_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_igemm_persistent_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_halo_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CK(cudaFuncSetAttribute(pool_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pool_conv_smem_bytes())));
  CK(cudaFuncSetAttribute(conv0_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(conv0_tc_smem_bytes())));
  CK(cudaFuncSetAttribute(conv0_tc_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          static_cast<int>(conv0_tc_smem_bytes())));
  CK(cudaFuncSetAttribute(avgpool3_relu_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(avgpool3_relu_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(avgpool3_relu_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(avgpool3_relu_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, device));
  // ICP_RESERVE_SMS=n: the persistent kernels leave n SMs free (for a co-running NCCL kernel of a multi-GPU run).  MEASURED on
  // 2 GPUs, 30-step runs back to back: 0 reserved 2.265 / 2.279 ms per step, 1 reserved 2.265 / 2.249, 2 reserved 2.274; one GPU
  // 2.253 vs 2.248 — no effect beyond the run-to-run spread, so it is off by default.
  if (const char* v = getenv("ICP_RESERVE_SMS")) {
    const int r = atoi(v);
    if (r > 0 && r < g_num_sms / 2) g_num_sms -= r;
  }
  {
    const int late = pdl_late() ? 1 : 0;
    CK(cudaMemcpyToSymbol(ptx::g_pdl_late, &late, sizeof late));
  }
  if (const char* v = getenv("ICP_MBAR_TIMEOUT_MS")) {
    const long long cycles = atoll(v) * 1900000LL;     // ~1.9 GHz SM clock
    CK(cudaMemcpyToSymbol(ptx::g_mbar_timeout_cycles, &cycles, sizeof cycles));
  }
}

static void require_device(int device) {
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    fail(ICP_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
         err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
  if (device < 0 || device >= count) fail(ICP_E_INVALID, "device %d outside [0, %d)", device, count);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    fail(ICP_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  CK(cudaSetDevice(device));
  load_driver_api();
  set_kernel_attrs(device);
}

}  // namespace icp

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

// CRC-32C (Castagnoli), slice-by-1 table: used by the TF V1 checkpoint reader (tf_checkpoint.py) to verify table blocks.
uint32_t icp_crc32c(const void* data, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);
      table[i] = c;
    }
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

const char* icp_version(void) { return "inception_b200 0.1 (sm_100a, tcgen05 implicit-GEMM)"; }

const char* icp_last_error(const icp_engine* e) { return e ? e->err.c_str() : g_tls_error.c_str(); }

int icp_create(const icp_config* cfg, icp_engine** out) {
  if (out) *out = nullptr;
  std::unique_ptr<icp_engine> e;
  int rc = guard(nullptr, [&] {
    if (!cfg || !out || !cfg->weights_path) fail(ICP_E_INVALID, "icp_create: null argument");
    if (cfg->num_classes < 2 || cfg->max_batch < 1) fail(ICP_E_INVALID, "icp_create: num_classes/max_batch invalid");
    if (cfg->precision != ICP_PRECISION_FP16 && cfg->precision != ICP_PRECISION_BF16)
      fail(ICP_E_INVALID, "icp_create: unknown precision %d", cfg->precision);
    Blob blob = read_blob(cfg->weights_path);   // file errors first: they do not need a GPU
    require_device(cfg->device);
    e.reset(new icp_engine());
    e->cfg = *cfg;
    e->weights_path = cfg->weights_path;
    e->cfg.weights_path = e->weights_path.c_str();
    e->fmt = cfg->precision == ICP_PRECISION_BF16 ? 1 : 0;
    e->S = cfg->image_size > 0 ? cfg->image_size : 299;
    if (e->S < 75) fail(ICP_E_INVALID, "image_size %d too small for the tower", e->S);
    e->C = cfg->num_classes;
    e->max_batch = cfg->max_batch;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    e->d_input = static_cast<float*>(e->pool.alloc(static_cast<size_t>(e->max_batch) * e->S * e->S * 3 * 4));
    e->d_input_u8 = static_cast<uint8_t*>(e->pool.alloc(static_cast<size_t>(e->max_batch) * e->S * e->S * 3));
    e->d_slot_u8[0] = e->d_input_u8;
    e->d_slot_u8[1] = static_cast<uint8_t*>(e->pool.alloc(static_cast<size_t>(e->max_batch) * e->S * e->S * 3));
    CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 1; i < icp_engine::kLanes; ++i) {
      CK(cudaStreamCreateWithFlags(&e->lane_stream[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&e->join_ev[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming));
    if (const char* v = getenv("ICP_LANES_MAX_BATCH")) e->lanes_max_batch = atoi(v);
    if (const char* v = getenv("ICP_STEM_CHUNK")) e->stem_chunk = atoi(v);
    if (const char* v = getenv("ICP_FUSE_POOL_MIN_BATCH")) e->fuse_min_batch = atoi(v);
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&e->slot_h2d[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->slot_done[i], cudaEventDisableTiming));
    }
    e->graph_k = e->C < 5 ? e->C : 5;
    Builder b{*e, blob, e->fmt, e->max_batch};
    b.build();
    for (size_t i = 0; i < e->ops.size(); ++i)
      if (e->ops[i].kind != Op::FORK && e->ops[i].kind != Op::JOIN) e->launch_ops.push_back(static_cast<int>(i));
    CK(cudaDeviceSynchronize());
  });
  if (rc == ICP_OK) *out = e.release();
  return rc;
}

void icp_destroy(icp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  delete e;
}

int icp_forward_topk(icp_engine* e, const float* x_host, int n, int k, float* probs_out_host, float* topk_vals_host,
                     int32_t* topk_idx_host) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!x_host) fail(ICP_E_INVALID, "null input");
    CK(cudaSetDevice(e->cfg.device));
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    const size_t in_bytes = static_cast<size_t>(n) * e->S * e->S * 3 * 4;
    CK(cudaMemcpyAsync(e->d_input, x_host, in_bytes, cudaMemcpyHostToDevice, e->stream));
    run_tower(*e, e->d_input, false, n, k > 0 ? k : e->graph_k, nullptr, e->stream, true);
    const int kk = k > 0 ? k : 0;
    if (probs_out_host)
      CK(cudaMemcpyAsync(probs_out_host, e->d_probs, static_cast<size_t>(n) * e->C * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_vals_host)
      CK(cudaMemcpyAsync(topk_vals_host, e->d_topk_val, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_idx_host)
      CK(cudaMemcpyAsync(topk_idx_host, e->d_topk_idx, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

int icp_forward(icp_engine* e, const float* x_host, int n, float* probs_out_host) {
  return icp_forward_topk(e, x_host, n, 0, probs_out_host, nullptr, nullptr);
}

int icp_forward_device(icp_engine* e, const float* x_dev, int n, float* probs_dev, void* stream) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!x_dev) fail(ICP_E_INVALID, "null input");
    CK(cudaSetDevice(e->cfg.device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : e->stream;
    run_tower(*e, x_dev, false, n, e->graph_k, probs_dev, st, false);
  });
}

int icp_host_register(icp_engine* e, void* host_ptr, size_t bytes) {
  if (!e || !host_ptr || !bytes) return ICP_E_INVALID;
  if (cudaSetDevice(e->cfg.device) != cudaSuccess) { (void)cudaGetLastError(); return ICP_E_CUDA; }
  cudaError_t err = cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    char msg[160];
    snprintf(msg, sizeof msg, "cudaHostRegister(%p, %zu): %s", host_ptr, bytes, cudaGetErrorString(err));
    g_tls_error = msg;
    return ICP_E_CUDA;
  }
  return ICP_OK;
}

int icp_host_unregister(icp_engine* e, void* host_ptr) {
  if (!e || !host_ptr) return ICP_E_INVALID;
  if (cudaSetDevice(e->cfg.device) != cudaSuccess) { (void)cudaGetLastError(); return ICP_E_CUDA; }
  cudaError_t err = cudaHostUnregister(host_ptr);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    return ICP_E_CUDA;
  }
  return ICP_OK;
}

int icp_forward_u8_device(icp_engine* e, const uint8_t* x_dev, int n, float* probs_dev, void* stream) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!x_dev) fail(ICP_E_INVALID, "null input");
    CK(cudaSetDevice(e->cfg.device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : e->stream;
    run_tower(*e, x_dev, true, n, e->graph_k, probs_dev, st, false);
  });
}

static void reserve_rgb(icp_engine* e, size_t bytes) {
  if (bytes <= e->d_rgb_cap) return;
  CK(cudaStreamSynchronize(e->stream));       // earlier work may still read the old buffer
  if (e->d_rgb) cudaFree(e->d_rgb);
  e->d_rgb = nullptr;
  e->d_rgb_cap = 0;
  bytes = (bytes + (1u << 20)) & ~size_t((1u << 20) - 1);
  CK(cudaMalloc(&e->d_rgb, bytes));
  e->d_rgb_cap = bytes;
}

// decoded uint8 [h,w,3] at d_src (device) -> image `slot` of the fp32 network input
static void launch_preprocess(icp_engine* e, const uint8_t* d_src, int h, int w, int slot) {
  const int total = e->S * e->S;
  preprocess_u8_kernel<<<(total + 255) / 256, 256, 0, e->stream>>>(
      d_src, h, w, e->d_input + static_cast<size_t>(slot) * e->S * e->S * 3, e->S);
  CK(cudaGetLastError());
}

static void preprocess_to_input(icp_engine* e, const uint8_t* rgb, int h, int w) {
  if (!rgb || h < 16 || w < 16) fail(ICP_E_INVALID, "image %dx%d invalid", h, w);
  const size_t bytes = static_cast<size_t>(h) * w * 3;
  reserve_rgb(e, bytes);
  CK(cudaMemcpyAsync(e->d_rgb, rgb, bytes, cudaMemcpyHostToDevice, e->stream));
  launch_preprocess(e, e->d_rgb, h, w, 0);
}

// JPEG bytes -> decoded RGB at d_rgb + offset (device), through nvJPEG on the engine's stream; returns h, w.
static void decode_jpeg_to_device(icp_engine* e, const uint8_t* jpg, size_t len, size_t offset, int* oh, int* ow) {
  if (!jpg || len < 4) fail(ICP_E_INVALID, "empty JPEG buffer");
  e->jpeg.init();
  for (;;) {
    if (!e->jpeg.h) fail(ICP_E_CUDA, "nvJPEG: no usable decode backend");
    int ncomp = 0;
    nvjpegChromaSubsampling_t ss;
    int ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegStatus_t r = nvjpegGetImageInfo(e->jpeg.h, jpg, len, &ncomp, &ss, ws, hs);
    if (r != NVJPEG_STATUS_SUCCESS) fail(ICP_E_INVALID, "not a decodable JPEG (nvjpegGetImageInfo status %d)", (int)r);
    const int w = ws[0], h = hs[0];
    if (h < 16 || w < 16 || h > 16384 || w > 16384) fail(ICP_E_INVALID, "image %dx%d invalid", h, w);
    const size_t bytes = static_cast<size_t>(h) * w * 3;
    reserve_rgb(e, offset + bytes);
    nvjpegImage_t img{};
    img.channel[0] = e->d_rgb + offset;
    img.pitch[0] = static_cast<size_t>(w) * 3;
    r = nvjpegDecode(e->jpeg.h, e->jpeg.st, jpg, len, NVJPEG_OUTPUT_RGBI, &img, e->stream);
    if (r == NVJPEG_STATUS_SUCCESS) { *oh = h; *ow = w; return; }
    // this backend cannot take the stream (progressive / unsupported subsampling / no engine): try the next one
    if (!e->jpeg.open_next()) fail(ICP_E_INVALID, "nvjpegDecode failed (status %d) on every backend", (int)r);
  }
}

int icp_preprocess_rgb8(icp_engine* e, const uint8_t* rgb, int h, int w, float* out_host) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!out_host) fail(ICP_E_INVALID, "null output");
    CK(cudaSetDevice(e->cfg.device));
    preprocess_to_input(e, rgb, h, w);
    CK(cudaMemcpyAsync(out_host, e->d_input, static_cast<size_t>(e->S) * e->S * 3 * 4, cudaMemcpyDeviceToHost,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
  });
}

int icp_classify_rgb8(icp_engine* e, const uint8_t* rgb, int h, int w, int k, float* topk_vals, int32_t* topk_idx) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!topk_vals || !topk_idx) fail(ICP_E_INVALID, "null output");
    if (k < 1) fail(ICP_E_INVALID, "k must be >= 1");
    CK(cudaSetDevice(e->cfg.device));
    preprocess_to_input(e, rgb, h, w);
    run_tower(*e, e->d_input, false, 1, k, nullptr, e->stream, true);
    CK(cudaMemcpyAsync(topk_vals, e->d_topk_val, k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(topk_idx, e->d_topk_idx, k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

int icp_classify_jpeg(icp_engine* e, const uint8_t* jpg, size_t len, int k, float* topk_vals, int32_t* topk_idx) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!topk_vals || !topk_idx) fail(ICP_E_INVALID, "null output");
    if (k < 1) fail(ICP_E_INVALID, "k must be >= 1");
    CK(cudaSetDevice(e->cfg.device));
    int h = 0, w = 0;
    decode_jpeg_to_device(e, jpg, len, 0, &h, &w);
    launch_preprocess(e, e->d_rgb, h, w, 0);
    run_tower(*e, e->d_input, false, 1, k, nullptr, e->stream, true);
    CK(cudaMemcpyAsync(topk_vals, e->d_topk_val, k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(topk_idx, e->d_topk_idx, k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

int icp_decode_jpeg(icp_engine* e, const uint8_t* jpg, size_t len, uint8_t* rgb_out_host, size_t cap, int* h_out,
                    int* w_out) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!h_out || !w_out) fail(ICP_E_INVALID, "null output");
    CK(cudaSetDevice(e->cfg.device));
    int h = 0, w = 0;
    decode_jpeg_to_device(e, jpg, len, 0, &h, &w);
    *h_out = h; *w_out = w;
    const size_t bytes = static_cast<size_t>(h) * w * 3;
    if (rgb_out_host) {
      if (cap < bytes) fail(ICP_E_INVALID, "output buffer of %zu bytes is too small for %dx%dx3", cap, h, w);
      CK(cudaMemcpyAsync(rgb_out_host, e->d_rgb, bytes, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
  });
}

int icp_jpeg_backend(icp_engine* e, int* backend) {
  if (!e || !backend) return ICP_E_INVALID;
  *backend = e->jpeg.backend;
  return ICP_OK;
}

// n decoded images of different sizes -> one tower launch at batch n (the micro-batcher's call)
int icp_classify_batch_rgb8(icp_engine* e, const uint8_t* const* rgbs, const int* hs, const int* ws, int n, int k,
                            float* topk_vals, int32_t* topk_idx) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!rgbs || !hs || !ws || !topk_vals || !topk_idx) fail(ICP_E_INVALID, "null argument");
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    if (k < 1) fail(ICP_E_INVALID, "k must be >= 1");
    CK(cudaSetDevice(e->cfg.device));
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
      if (!rgbs[i] || hs[i] < 16 || ws[i] < 16) fail(ICP_E_INVALID, "image %d: %dx%d invalid", i, hs[i], ws[i]);
      total += (static_cast<size_t>(hs[i]) * ws[i] * 3 + 255) & ~size_t(255);
    }
    reserve_rgb(e, total);
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      const size_t bytes = static_cast<size_t>(hs[i]) * ws[i] * 3;
      CK(cudaMemcpyAsync(e->d_rgb + off, rgbs[i], bytes, cudaMemcpyHostToDevice, e->stream));
      launch_preprocess(e, e->d_rgb + off, hs[i], ws[i], i);
      off += (bytes + 255) & ~size_t(255);
    }
    run_tower(*e, e->d_input, false, n, k, nullptr, e->stream, true);
    CK(cudaMemcpyAsync(topk_vals, e->d_topk_val, static_cast<size_t>(n) * k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(topk_idx, e->d_topk_idx, static_cast<size_t>(n) * k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

// same from JPEG bytes: every image is decoded by nvJPEG into the staging buffer, then one tower launch
int icp_classify_batch_jpeg(icp_engine* e, const uint8_t* const* jpgs, const size_t* lens, int n, int k,
                            float* topk_vals, int32_t* topk_idx) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!jpgs || !lens || !topk_vals || !topk_idx) fail(ICP_E_INVALID, "null argument");
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    if (k < 1) fail(ICP_E_INVALID, "k must be >= 1");
    CK(cudaSetDevice(e->cfg.device));
    // sizes first (header parse only), so that the staging buffer is reserved once and never moves mid-batch
    e->jpeg.init();
    if (!e->jpeg.h) fail(ICP_E_CUDA, "nvJPEG: no usable decode backend");
    std::vector<size_t> offs(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
      int ncomp = 0;
      nvjpegChromaSubsampling_t ss;
      int w_[NVJPEG_MAX_COMPONENT] = {0}, h_[NVJPEG_MAX_COMPONENT] = {0};
      if (!jpgs[i] || nvjpegGetImageInfo(e->jpeg.h, jpgs[i], lens[i], &ncomp, &ss, w_, h_) != NVJPEG_STATUS_SUCCESS)
        fail(ICP_E_INVALID, "image %d is not a decodable JPEG", i);
      offs[i] = total;
      total += (static_cast<size_t>(h_[0]) * w_[0] * 3 + 255) & ~size_t(255);
    }
    reserve_rgb(e, total);
    for (int i = 0; i < n; ++i) {
      int h = 0, w = 0;
      decode_jpeg_to_device(e, jpgs[i], lens[i], offs[i], &h, &w);
      launch_preprocess(e, e->d_rgb + offs[i], h, w, i);
    }
    run_tower(*e, e->d_input, false, n, k, nullptr, e->stream, true);
    CK(cudaMemcpyAsync(topk_vals, e->d_topk_val, static_cast<size_t>(n) * k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(topk_idx, e->d_topk_idx, static_cast<size_t>(n) * k * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

int icp_forward_u8(icp_engine* e, const uint8_t* x_host, int n, int k, float* probs_out_host, float* topk_vals_host,
                   int32_t* topk_idx_host) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!x_host) fail(ICP_E_INVALID, "null input");
    CK(cudaSetDevice(e->cfg.device));
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    const size_t in_bytes = static_cast<size_t>(n) * e->S * e->S * 3;
    CK(cudaMemcpyAsync(e->d_input_u8, x_host, in_bytes, cudaMemcpyHostToDevice, e->stream));
    run_tower(*e, e->d_input_u8, true, n, k > 0 ? k : e->graph_k, nullptr, e->stream, true);
    const int kk = k > 0 ? k : 0;
    if (probs_out_host)
      CK(cudaMemcpyAsync(probs_out_host, e->d_probs, static_cast<size_t>(n) * e->C * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_vals_host)
      CK(cudaMemcpyAsync(topk_vals_host, e->d_topk_val, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_idx_host)
      CK(cudaMemcpyAsync(topk_idx_host, e->d_topk_idx, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  });
}

static int submit_u8_impl(icp_engine* e, const uint8_t* x_host, int n, int k, float* probs_out_host,
                          float* topk_vals_host, int32_t* topk_idx_host, float* probs_dev, int slot) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!x_host) fail(ICP_E_INVALID, "null input");
    if (slot < 0 || slot > 1) fail(ICP_E_INVALID, "slot must be 0 or 1");
    if (e->slot_busy[slot]) fail(ICP_E_INVALID, "slot %d has an uncollected request", slot);
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    CK(cudaSetDevice(e->cfg.device));
    const size_t in_bytes = static_cast<size_t>(n) * e->S * e->S * 3;
    // the slot's input buffer is free once the tower that last read it has finished
    CK(cudaStreamWaitEvent(e->copy_stream, e->slot_done[slot], 0));
    CK(cudaMemcpyAsync(e->d_slot_u8[slot], x_host, in_bytes, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(e->slot_h2d[slot], e->copy_stream));
    CK(cudaStreamWaitEvent(e->stream, e->slot_h2d[slot], 0));
    const int kk = k > 0 ? k : 0;
    run_tower(*e, e->d_slot_u8[slot], true, n, kk ? kk : e->graph_k, probs_dev, e->stream, false);
    if (probs_out_host && !probs_dev)
      CK(cudaMemcpyAsync(probs_out_host, e->d_probs, static_cast<size_t>(n) * e->C * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_vals_host)
      CK(cudaMemcpyAsync(topk_vals_host, e->d_topk_val, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    if (kk && topk_idx_host)
      CK(cudaMemcpyAsync(topk_idx_host, e->d_topk_idx, static_cast<size_t>(n) * kk * 4, cudaMemcpyDeviceToHost,
                         e->stream));
    CK(cudaEventRecord(e->slot_done[slot], e->stream));
    e->slot_busy[slot] = true;
  });
}

int icp_submit_u8(icp_engine* e, const uint8_t* x_host, int n, int k, float* probs_out_host, float* topk_vals_host,
                  int32_t* topk_idx_host, int slot) {
  return submit_u8_impl(e, x_host, n, k, probs_out_host, topk_vals_host, topk_idx_host, nullptr, slot);
}

int icp_submit_u8_device_out(icp_engine* e, const uint8_t* x_host, int n, float* probs_dev, int slot) {
  if (!probs_dev) return ICP_E_INVALID;
  return submit_u8_impl(e, x_host, n, 0, nullptr, nullptr, nullptr, probs_dev, slot);
}

int icp_collect(icp_engine* e, int slot) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (slot < 0 || slot > 1) fail(ICP_E_INVALID, "slot must be 0 or 1");
    if (!e->slot_busy[slot]) fail(ICP_E_INVALID, "slot %d has no request in flight", slot);
    CK(cudaSetDevice(e->cfg.device));
    e->slot_busy[slot] = false;
    CK(cudaEventSynchronize(e->slot_done[slot]));
  });
}

int icp_num_ops(icp_engine* e, int* num_ops) {
  if (!e || !num_ops) return ICP_E_INVALID;
  *num_ops = static_cast<int>(e->launch_ops.size()) + 2;   // conv0 + plan + head
  return ICP_OK;
}

int icp_op_info(icp_engine* e, int i, char* name, int name_cap, int* kind, double* flops_per_image,
                double* bytes_per_image) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    const int nops = static_cast<int>(e->launch_ops.size()) + 2;
    if (i < 0 || i >= nops) fail(ICP_E_INVALID, "op index %d outside [0, %d)", i, nops);
    std::string nm;
    int kd = 0;
    double fl = 0, by = 0;
    if (i == 0) {
      const Act& o = e->conv0_out;
      nm = "conv0"; kd = ICP_OP_CONV0;
      fl = 2.0 * o.h * o.w * 32 * 27;
      by = static_cast<double>(e->S) * e->S * 3 * 4 + static_cast<double>(o.h) * o.w * 32 * 2;
    } else if (i == nops - 1) {
      const Act& f = e->final_act;
      nm = "head"; kd = ICP_OP_HEAD;
      fl = 2.0 * f.c * e->C;
      by = static_cast<double>(f.h) * f.w * f.c * 2;
    } else {
      const Op& op = e->ops[e->launch_ops[i - 1]];
      if (op.kind == Op::CONV) {
        const ConvLayer& L = op.conv;
        nm = L.name; kd = ICP_OP_CONV_IGEMM;
        fl = 2.0 * static_cast<double>(L.macs_per_image);
        const double k = static_cast<double>(L.macs_per_image) / (static_cast<double>(L.pix_per_image) * L.cout);
        // algorithmic traffic: input read once + output written once (weights amortised over the batch)
        by = static_cast<double>(L.in_elems_per_image) * 2 + static_cast<double>(L.pix_per_image) * L.cout * 2;
        (void)k;
      } else {
        nm = !op.name.empty() ? op.name : (op.kind == Op::MAXPOOL ? "max_pool" : "avg_pool");
        kd = op.kind == Op::MAXPOOL ? ICP_OP_MAXPOOL : ICP_OP_AVGPOOL;
        fl = 0;
        by = static_cast<double>(op.in.h) * op.in.w * op.c * 2 + static_cast<double>(op.out.h) * op.out.w * op.c * 2;
      }
    }
    if (name && name_cap > 0) snprintf(name, name_cap, "%s", nm.c_str());
    if (kind) *kind = kd;
    if (flops_per_image) *flops_per_image = fl;
    if (bytes_per_image) *bytes_per_image = by;
  });
}

int icp_profile_forward(icp_engine* e, int n, int reps, float* ms_per_op) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!ms_per_op || reps < 1) fail(ICP_E_INVALID, "bad argument");
    if (n < 1 || n > e->max_batch) fail(ICP_E_INVALID, "batch %d outside [1, %d]", n, e->max_batch);
    CK(cudaSetDevice(e->cfg.device));
    const int nops = static_cast<int>(e->launch_ops.size()) + 2;
    std::vector<double> acc(nops, 0.0);
    for (int r = 0; r < reps; ++r) {
      std::vector<cudaEvent_t> ev;
      OpTimer tm{&ev, e->stream};
      cudaError_t err = cudaSuccess;
      try {
        enqueue_tower(*e, e->d_input, false, n, e->graph_k, nullptr, e->stream, &tm);
        err = cudaStreamSynchronize(e->stream);
      } catch (...) {
        for (cudaEvent_t x : ev) cudaEventDestroy(x);
        throw;
      }
      if (err == cudaSuccess && static_cast<int>(ev.size()) == nops + 1) {
        for (int i = 0; i < nops; ++i) {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
          acc[i] += ms;
        }
      }
      for (cudaEvent_t x : ev) cudaEventDestroy(x);
      if (err != cudaSuccess) fail(ICP_E_CUDA, "profile run failed: %s", cudaGetErrorString(err));
    }
    for (int i = 0; i < nops; ++i) ms_per_op[i] = static_cast<float>(acc[i] / reps);
  });
}

int icp_endpoint_shape(icp_engine* e, const char* name, int dims[4]) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!name || !dims) fail(ICP_E_INVALID, "null argument");
    auto it = e->endpoints.find(name);
    if (it == e->endpoints.end()) fail(ICP_E_INVALID, "unknown end point '%s'", name);
    const Endpoint& ep = it->second;
    dims[0] = e->last_n;
    dims[1] = ep.vec ? 1 : ep.act.h;
    dims[2] = ep.vec ? 1 : ep.act.w;
    dims[3] = ep.c;
  });
}

int icp_endpoint(icp_engine* e, const char* name, float* out_host) {
  if (!e) return ICP_E_INVALID;
  return guard(e, [&] {
    if (!name || !out_host) fail(ICP_E_INVALID, "null argument");
    if (e->last_n < 1) fail(ICP_E_INVALID, "no forward call has run yet");
    CK(cudaSetDevice(e->cfg.device));
    auto it = e->endpoints.find(name);
    if (it == e->endpoints.end()) fail(ICP_E_INVALID, "unknown end point '%s'", name);
    const Endpoint& ep = it->second;
    if (e->last_fused && strcmp(name, "pool1") == 0)
      fail(ICP_E_INVALID, "end point 'pool1' was not materialised: at batch >= %d pool1 is fused into conv3's operand load "
                          "(ICP_FUSE_POOL_MIN_BATCH)", e->fuse_min_batch);
    CK(cudaStreamSynchronize(e->stream));
    if (ep.vec) {
      CK(cudaMemcpy(out_host, ep.vec, static_cast<size_t>(e->last_n) * ep.c * 4, cudaMemcpyDeviceToHost));
      return;
    }
    const size_t pixels = static_cast<size_t>(e->last_n) * ep.act.h * ep.act.w;
    float* tmp = nullptr;
    CK(cudaMalloc(&tmp, pixels * ep.c * 4));
    const size_t total = pixels * ep.c;
    h16_to_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, e->stream>>>(ep.act.p, tmp, pixels, ep.c,
                                                                                          ep.act.c, ep.coff, e->fmt);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(out_host, tmp, total * 4, cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(tmp);
    if (err != cudaSuccess) fail(ICP_E_CUDA, "end point copy failed: %s", cudaGetErrorString(err));
  });
}

int icp_last_forward_ms(icp_engine* e, float* ms) {
  if (!e || !ms) return ICP_E_INVALID;
  *ms = e->last_ms;
  return ICP_OK;
}

int icp_launches_per_forward(icp_engine* e, int* launches) {
  if (!e || !launches) return ICP_E_INVALID;
  *launches = e->launches;
  return ICP_OK;
}

void* icp_stream(icp_engine* e) { return e ? e->stream : nullptr; }

int icp_conv2d_bn_relu(int device, int precision, const float* x, int n, int h, int w, int cin, const float* wgt,
                       int kh, int kw, int cout, int stride, int same, const float* scale, const float* shift,
                       int relu, float* y, int reps, float* ms_out) {
  return guard(nullptr, [&] {
    if (!x || !wgt || !scale || !shift || !y) fail(ICP_E_INVALID, "null argument");
    if (n < 1 || h < 1 || w < 1 || cin < 1 || cout < 1 || kh < 1 || kw < 1 || stride < 1)
      fail(ICP_E_INVALID, "bad geometry");
    if (precision != ICP_PRECISION_FP16 && precision != ICP_PRECISION_BF16) fail(ICP_E_INVALID, "bad precision");
    require_device(device);
    const int fmt = precision == ICP_PRECISION_BF16 ? 1 : 0;
    DevicePool pool;
    cudaStream_t st = nullptr;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{st};
    const int ph = same ? (kh - 1) / 2 : 0, pw = same ? (kw - 1) / 2 : 0;
    const int oh = (h + 2 * ph - kh) / stride + 1, ow = (w + 2 * pw - kw) / stride + 1;
    if (oh < 1 || ow < 1) fail(ICP_E_INVALID, "empty output");
    Act in, out;
    in.h = h; in.w = w; in.c = cin;
    out.h = oh; out.w = ow; out.c = cout;
    const size_t in_elems = static_cast<size_t>(n) * h * w * cin, out_elems = static_cast<size_t>(n) * oh * ow * cout;
    float* d_xf = static_cast<float*>(pool.alloc(in_elems * 4));
    in.p = static_cast<uint16_t*>(pool.alloc(in_elems * 2));
    out.p = static_cast<uint16_t*>(pool.alloc(out_elems * 2));
    float* d_yf = static_cast<float*>(pool.alloc(out_elems * 4));
    CK(cudaMemcpyAsync(d_xf, x, in_elems * 4, cudaMemcpyHostToDevice, st));
    f32_to_h16_kernel<<<static_cast<unsigned>((in_elems + 255) / 256), 256, 0, st>>>(d_xf, in.p, in_elems, fmt);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(out.p, 0xFF, out_elems * 2, st));   // NaN pattern: unwritten outputs are visible
    ConvLayer L = make_conv(pool, "conv2d_bn_relu", n, in, 0, cin, wgt, kh, kw, cout, stride, same != 0, scale, shift,
                            relu != 0, out, 0, fmt);
    const char* tl_path = getenv("ICP_TIMELINE");
    long long* d_dbg = nullptr;
    if (tl_path && (L.persistent || L.halo || L.pair)) {
      d_dbg = static_cast<long long*>(pool.alloc(14 * 512 * 8));
      CK(cudaMemsetAsync(d_dbg, 0, 14 * 512 * 8, st));
      L.args.dbg = d_dbg;
      L.hargs.dbg = d_dbg;
      L.pargs.dbg = d_dbg;
    }
    launch_conv(L, n, st);
    if (d_dbg) {
      std::vector<long long> hd(14 * 512);
      CK(cudaMemcpyAsync(hd.data(), d_dbg, hd.size() * 8, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (FILE* f = fopen(tl_path, "wb")) { fwrite(hd.data(), 8, hd.size(), f); fclose(f); }
      L.args.dbg = nullptr;
      L.hargs.dbg = nullptr;
      L.pargs.dbg = nullptr;
    }
    h16_to_f32_kernel<<<static_cast<unsigned>((out_elems + 255) / 256), 256, 0, st>>>(out.p, d_yf,
                                                                                     static_cast<size_t>(n) * oh * ow,
                                                                                     cout, cout, 0, fmt);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y, d_yf, out_elems * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (reps > 0 && ms_out) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      for (int i = 0; i < 3; ++i) launch_conv(L, n, st);
      CK(cudaEventRecord(a, st));
      for (int i = 0; i < reps; ++i) launch_conv(L, n, st);
      CK(cudaEventRecord(b, st));
      cudaError_t err = cudaStreamSynchronize(st);
      float ms = 0.f;
      if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, a, b);
      cudaEventDestroy(a);
      cudaEventDestroy(b);
      if (err != cudaSuccess) fail(ICP_E_CUDA, "timing failed: %s", cudaGetErrorString(err));
      *ms_out = ms / reps;
    }
  });
}

int icp_pool2d(int device, int precision, int kind, const float* x, int n, int h, int w, int c, int k, int stride,
               int same, float* y) {
  return guard(nullptr, [&] {
    if (!x || !y) fail(ICP_E_INVALID, "null argument");
    if (c % 8) fail(ICP_E_INVALID, "channels must be a multiple of 8");
    if (precision != ICP_PRECISION_FP16 && precision != ICP_PRECISION_BF16) fail(ICP_E_INVALID, "bad precision");
    require_device(device);
    const int fmt = precision == ICP_PRECISION_BF16 ? 1 : 0;
    if (kind == 0) { k = 3; stride = 2; same = 0; }
    const int pad = same ? (k - 1) / 2 : 0;
    const int oh = (h + 2 * pad - k) / stride + 1, ow = (w + 2 * pad - k) / stride + 1;
    if (oh < 1 || ow < 1) fail(ICP_E_INVALID, "empty output");
    DevicePool pool;
    const size_t in_elems = static_cast<size_t>(n) * h * w * c, out_elems = static_cast<size_t>(n) * oh * ow * c;
    float* d_xf = static_cast<float*>(pool.alloc(in_elems * 4));
    uint16_t* d_x = static_cast<uint16_t*>(pool.alloc(in_elems * 2));
    uint16_t* d_y = static_cast<uint16_t*>(pool.alloc(out_elems * 2));
    float* d_yf = static_cast<float*>(pool.alloc(out_elems * 4));
    CK(cudaMemcpy(d_xf, x, in_elems * 4, cudaMemcpyHostToDevice));
    f32_to_h16_kernel<<<static_cast<unsigned>((in_elems + 255) / 256), 256>>>(d_xf, d_x, in_elems, fmt);
    const long long total = static_cast<long long>(n) * oh * ow * (c / 8);
    if (kind == 0)
      maxpool3s2_kernel<<<dim3((static_cast<unsigned>(oh) * ((ow + 1) / 2) * (c / 8) + 255) / 256, n), 256>>>(d_x, n, h, w, c, c, 0, d_y, oh, ow, c, 0,
                                                                             fmt);
    else
      avgpool_kernel<<<static_cast<unsigned>((total + 255) / 256), 256>>>(d_x, n, h, w, c, c, 0, d_y, oh, ow, c, 0, k,
                                                                          stride, pad, nullptr, nullptr, 0, fmt);
    h16_to_f32_kernel<<<static_cast<unsigned>((out_elems + 255) / 256), 256>>>(d_y, d_yf,
                                                                               static_cast<size_t>(n) * oh * ow, c, c, 0,
                                                                               fmt);
    CK(cudaGetLastError());
    CK(cudaMemcpy(y, d_yf, out_elems * 4, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
