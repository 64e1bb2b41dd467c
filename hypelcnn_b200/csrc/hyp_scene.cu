// Scene-side kernels: per-band min / max of the resident scene cube and the patch gather that replaces the reference's
// per-pixel window slice (importer/InMemoryImporter.py:27-38 -> common/common_nn_ops.py:169-176,
// loader/GRSS2018DataLoader.py:10-44; normalisation common/common_nn_ops.py:55-78).  HBM-bound byte movers.
#include <algorithm>

#include "hyp_common.cuh"

namespace hyp {

// ------------------------------------------------------------------------------------------
// scene min / max(value - min): two passes over the cube, channel-coalesced
template <typename T>
__global__ void scene_min_kernel(const T* __restrict__ cube, int64_t pixels, int C, int pixels_per_block,
                                 unsigned int* __restrict__ min_bits) {
  // float order-preserving encoding so atomicMin on uint works for non-negative and negative values
  const int64_t p0 = (int64_t)blockIdx.x * pixels_per_block;
  const int64_t p1 = min(pixels, p0 + pixels_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mn = INFINITY;
    for (int64_t px = p0; px < p1; px++) mn = fminf(mn, (float)cube[px * C + c]);
    unsigned int b = __float_as_uint(mn);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    atomicMin(&min_bits[c], b);
  }
}
template <typename T>
__global__ void scene_max_kernel(const T* __restrict__ cube, int64_t pixels, int C, int pixels_per_block,
                                 const float* __restrict__ mn, unsigned int* __restrict__ max_bits) {
  const int64_t p0 = (int64_t)blockIdx.x * pixels_per_block;
  const int64_t p1 = min(pixels, p0 + pixels_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mx = -INFINITY;
    const float lo = mn[c];
    for (int64_t px = p0; px < p1; px++) mx = fmaxf(mx, (float)cube[px * C + c] - lo);
    unsigned int b = __float_as_uint(mx);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    atomicMax(&max_bits[c], b);
  }
}
__global__ void decode_ordered_kernel(const unsigned int* __restrict__ bits, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  unsigned int b = bits[c];
  b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
  out[c] = __uint_as_float(b);
}

// ------------------------------------------------------------------------------------------
// patch gather.  numpy "symmetric" padding == reflect with the edge repeated.
__device__ __forceinline__ int reflect_sym(int i, int n) {
  if (i < 0) i = -i - 1;
  if (i >= n) i = 2 * n - 1 - i;
  return i;
}

struct GatherArgs {
  const void* casi;
  int casi_u16;
  int Hc, Wc, C;
  const float* cmin;
  const float* cmax;
  const float* lidar;
  int Hl, Wl;
  const float* lminmax;
  int nb, mode;
  const int32_t* xy;
  int64_t N;
  float* out;
  int out_ld;
};

// one block per patch; threads sweep (pixel, channel) with channel fastest -> coalesced
__global__ void gather_kernel(const GatherArgs p) {
  const int64_t n = blockIdx.x;
  const int S = 2 * p.nb + 1;
  const int x = p.xy[2 * n], y = p.xy[2 * n + 1];
  int bx = x, by = y;
  if (p.mode == HYP_GATHER_GRSS2018) {  // loader/GRSS2018DataLoader.py:23-29 (int() truncation)
    bx = x / 2 + p.nb - p.nb / 2;
    by = y / 2 + p.nb - p.nb / 2;
  }
  const int per_pixel = p.out_ld;
  const int total = S * S * per_pixel;
  float* op = p.out + (size_t)n * total;
  const float lmin = p.lminmax ? p.lminmax[0] : 0.f, lmax = p.lminmax ? p.lminmax[1] : 1.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int pix = i / per_pixel, c = i - pix * per_pixel;
    const int py = pix / S, px = pix - py * S;
    float v = 0.f;
    if (c < p.C) {
      const int oy = (p.mode == HYP_GATHER_GRSS2018) ? py / 2 : py;
      const int ox = (p.mode == HYP_GATHER_GRSS2018) ? px / 2 : px;
      const int ry = reflect_sym(by + oy - p.nb, p.Hc), rx = reflect_sym(bx + ox - p.nb, p.Wc);
      const size_t off = ((size_t)ry * p.Wc + rx) * p.C + c;
      if (p.casi_u16) {
        const unsigned short raw = reinterpret_cast<const unsigned short*>(p.casi)[off];
        if (p.cmin) {
          const unsigned short sh = (unsigned short)(raw - (unsigned short)p.cmin[c]);
          v = __fdiv_rn((float)sh, p.cmax[c]);
        } else {
          v = (float)raw;
        }
      } else {
        const float raw = reinterpret_cast<const float*>(p.casi)[off];
        v = p.cmin ? __fdiv_rn(__fsub_rn(raw, p.cmin[c]), p.cmax[c]) : raw;
      }
    } else if (c == p.C && p.lidar) {
      const int ry = reflect_sym(y + py - p.nb, p.Hl), rx = reflect_sym(x + px - p.nb, p.Wl);
      const float raw = p.lidar[(size_t)ry * p.Wl + rx];
      v = p.lminmax ? __fdiv_rn(__fsub_rn(raw, lmin), lmax) : raw;
    }
    op[i] = v;
  }
}


// ------------------------------------------------------------------------------------------
// gather_rows_kernel: the gather for scenes whose pixel spectra are whole 16-byte chunks (C % 8 == 0 for uint16,
// C % 4 == 0 for fp32 -- every scene of the reference: 144, 48, 64 bands).
//
//   * a block owns one patch at a time (persistent, grid-stride over the target list);
//   * loads: thread (chunk, row slot) loads ONE 16-byte chunk (8 uint16 / 4 fp32 bands) of up to four window pixels with
//     128-bit read-only loads; the loads of the block's NEXT patch are issued before the current one is streamed out
//     (software pipeline: the registers carry them across the store phase);
//   * convert: the thread normalises its chunks and puts the fp32 values into the patch image in shared memory.  A thread keeps the same chunk for the whole kernel, so its bands' constants live
//     in registers; neighbouring lanes hold neighbouring PIXELS of the same chunk, i.e. shared-memory addresses one
//     output row (C + 1 floats, odd for every scene of the reference) apart: conflict-free scalar stores;
//   * store: the patch image leaves as one flat stream of 16-byte streaming stores.  A patch of (2n+1)^2 * (C+1)
//     floats starts at any 4-byte phase of a 16-byte line, so the image sits in shared memory at the same phase
//     (`shift`) and only the first / last vector of a patch is written element-wise.
//
// Arithmetic is bit-identical to gather_kernel / the numpy reference for every valid scene: uint16 bands are
// (raw - min), an exact integer (computed in fp32 through the 2^23 exponent trick: both operands are integers below
// 2^16; casi_min must not exceed the band's minimum, which hyp_scene_minmax guarantees), then an IEEE-correct division
// by max.  For integer max in [1, 65535] (what hyp_scene_minmax returns for uint16 cubes) the division is
// q = a * r, q' = fma(fma(-q, b, a), r, q) with r = RN(1 / b): Markstein's correction, exact for these operands
// (integers below 2^16, no all-ones significand; tests/test_gpu_parity.py checks every dividend for eight divisors);
// a chunk with any other max takes __fdiv_rn.
template <typename T>
struct ChunkOf;
template <>
struct ChunkOf<unsigned short> { static constexpr int N = 8; };
template <>
struct ChunkOf<float> { static constexpr int N = 4; };

constexpr int GATHER_THREADS = 256;

template <typename T>
__device__ __forceinline__ void gather_convert(const uint4 raw, const float (&bias)[ChunkOf<T>::N],
                                               const float (&cmx)[ChunkOf<T>::N], const float (&rcp)[ChunkOf<T>::N],
                                               bool normalize, bool fast, float (&v)[ChunkOf<T>::N]) {
  constexpr int EPC = ChunkOf<T>::N;
  if (sizeof(T) == 2) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < EPC; j++) {
      // 0x4B000000 | r16 is the float 2^23 + r16: one byte permute builds it, one subtraction removes 2^23 + min
      const uint32_t bits = __byte_perm(w[j >> 1], 0x4B000000u, (j & 1) ? 0x7632 : 0x7610);
      const float a = __uint_as_float(bits) - bias[j];
      const float q = a * rcp[j];
      v[j] = normalize ? __fmaf_rn(__fmaf_rn(-q, cmx[j], a), rcp[j], q) : a;
    }
    if (!fast && normalize) {  // a divisor outside the proven range: rare, per thread
#pragma unroll
      for (int j = 0; j < EPC; j++) {
        const uint32_t bits = __byte_perm(w[j >> 1], 0x4B000000u, (j & 1) ? 0x7632 : 0x7610);
        v[j] = __fdiv_rn(__uint_as_float(bits) - bias[j], cmx[j]);
      }
    }
  } else {
    const float f[4] = {__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w)};
#pragma unroll
    for (int j = 0; j < EPC; j++) v[j] = normalize ? __fdiv_rn(__fsub_rn(f[j & 3], bias[j]), cmx[j]) : f[j & 3];
  }
}

// The loads of one patch that a thread keeps in flight: GATHER_PF window pixels of its chunk + one LiDAR sample.
constexpr int GATHER_PF = 4;
struct GatherLoads {
  uint4 raw[GATHER_PF];
  float lidar;
};

// issue this thread's loads for the patch at target (x, y); window pixels beyond GATHER_PF * rows_per_iter (large
// windows of wide scenes) are fetched later, inside the convert step
template <typename T>
__device__ __forceinline__ void gather_issue(const GatherArgs& p, const T* __restrict__ casi, int x, int y, int chunk, int slot,
                                             int rows_per_iter, bool active, GatherLoads& L) {
  const int S = 2 * p.nb + 1, npix = S * S;
  const bool half_res = p.mode == HYP_GATHER_GRSS2018;
  int bx = x, by = y;
  if (half_res) {  // loader/GRSS2018DataLoader.py:23-29 (int() truncation)
    bx = x / 2 + p.nb - p.nb / 2;
    by = y / 2 + p.nb - p.nb / 2;
  }
  if (active) {
    const int dy = rows_per_iter / S, dx = rows_per_iter - dy * S;  // one step of rows_per_iter pixels in (row, column)
    int py = slot / S, px = slot - py * S;
#pragma unroll
    for (int u = 0; u < GATHER_PF; u++) {
      if (slot + u * rows_per_iter < npix) {
        const int ry = reflect_sym(by + (half_res ? py / 2 : py) - p.nb, p.Hc);
        const int rx = reflect_sym(bx + (half_res ? px / 2 : px) - p.nb, p.Wc);
        L.raw[u] = __ldg(reinterpret_cast<const uint4*>(casi + ((size_t)ry * p.Wc + rx) * p.C) + chunk);
      }
      py += dy;
      px += dx;
      if (px >= S) { px -= S; py++; }
    }
  }
  // LiDAR: one thread per window pixel, the last threads first (they idle in the chunk loads when 256 is not a multiple
  // of the chunk count)
  const int lpix = GATHER_THREADS - 1 - (int)threadIdx.x;
  if (p.lidar && lpix < npix) {
    const int py = lpix / S, px = lpix - py * S;
    L.lidar = __ldg(p.lidar + (size_t)reflect_sym(y + py - p.nb, p.Hl) * p.Wl + reflect_sym(x + px - p.nb, p.Wl));
  }
}

template <typename T>
__global__ void __launch_bounds__(GATHER_THREADS, 4) gather_rows_kernel(const GatherArgs p) {
  constexpr int EPC = ChunkOf<T>::N;
  extern __shared__ float4 gather_smem4[];
  float* const sm = reinterpret_cast<float*>(gather_smem4);
  const int S = 2 * p.nb + 1, ld = p.out_ld, total = S * S * ld, npix = S * S;
  const int nchunks = p.C / EPC;
  const int rows_per_iter = GATHER_THREADS / nchunks;
  // neighbouring lanes: neighbouring window pixels of one chunk
  const int chunk = (int)threadIdx.x / rows_per_iter, slot = (int)threadIdx.x - chunk * rows_per_iter;
  const bool active = chunk < nchunks;
  const bool half_res = p.mode == HYP_GATHER_GRSS2018;
  const bool normalize = p.cmin != nullptr;
  const int n_out_c = p.C + (p.lidar ? 1 : 0);
  // this thread's bands: bias = what is subtracted from the raw value (+ 2^23 for the uint16 exponent trick)
  float bias[EPC], cmx[EPC], rcp[EPC];
  bool fast = sizeof(T) == 2;
#pragma unroll
  for (int j = 0; j < EPC; j++) {
    const int c = (active ? chunk : 0) * EPC + j;
    const float mn = normalize ? __ldg(p.cmin + c) : 0.f;
    cmx[j] = normalize ? __ldg(p.cmax + c) : 1.f;
    bias[j] = sizeof(T) == 2 ? 8388608.f + (float)(unsigned short)mn : mn;
    rcp[j] = __frcp_rn(cmx[j]);
    fast = fast && cmx[j] >= 1.f && cmx[j] <= 65535.f && cmx[j] == floorf(cmx[j]);  // see the header comment
  }
  const float lmin = p.lminmax ? __ldg(p.lminmax) : 0.f, lmax = p.lminmax ? __ldg(p.lminmax + 1) : 1.f;
  const T* const casi = reinterpret_cast<const T*>(p.casi);

  // software pipeline over the block's patches: the loads of patch n + grid are issued before patch n is streamed out,
  // so their latency is covered by the store phase and the two barriers
  int64_t n = blockIdx.x;
  GatherLoads L;
  int x = 0, y = 0;
  if (n < p.N) {
    x = __ldg(p.xy + 2 * n); y = __ldg(p.xy + 2 * n + 1);
    gather_issue<T>(p, casi, x, y, chunk, slot, rows_per_iter, active, L);
  }
  for (; n < p.N; n += gridDim.x) {
    float* const gout = p.out + (size_t)n * total;
    const int shift = (int)((reinterpret_cast<uintptr_t>(gout) >> 2) & 3);
    float* const img = sm + shift;
    // ---- convert: the loads in flight -> fp32 patch image in shared memory
    if (active) {
#pragma unroll
      for (int u = 0; u < GATHER_PF; u++) {
        const int at = slot + u * rows_per_iter;
        if (at < npix) {
          float v[EPC];
          gather_convert<T>(L.raw[u], bias, cmx, rcp, normalize, fast, v);
          float* const dst = img + at * ld + chunk * EPC;
#pragma unroll
          for (int j = 0; j < EPC; j++) dst[j] = v[j];
        }
      }
      if (GATHER_PF * rows_per_iter < npix) {  // window pixels beyond the prefetch depth
        int bx = x, by = y;
        if (half_res) { bx = x / 2 + p.nb - p.nb / 2; by = y / 2 + p.nb - p.nb / 2; }
        for (int at = slot + GATHER_PF * rows_per_iter; at < npix; at += rows_per_iter) {
          const int py = at / S, px = at - py * S;
          const int ry = reflect_sym(by + (half_res ? py / 2 : py) - p.nb, p.Hc);
          const int rx = reflect_sym(bx + (half_res ? px / 2 : px) - p.nb, p.Wc);
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(casi + ((size_t)ry * p.Wc + rx) * p.C) + chunk);
          float v[EPC];
          gather_convert<T>(raw, bias, cmx, rcp, normalize, fast, v);
          float* const dst = img + at * ld + chunk * EPC;
#pragma unroll
          for (int j = 0; j < EPC; j++) dst[j] = v[j];
        }
      }
    }
    {  // LiDAR channel and zero padding channels
      const int lpix = GATHER_THREADS - 1 - (int)threadIdx.x;
      if (lpix < npix) {
        if (p.lidar) img[lpix * ld + p.C] = p.lminmax ? __fdiv_rn(__fsub_rn(L.lidar, lmin), lmax) : L.lidar;
        for (int c = n_out_c; c < ld; c++) img[lpix * ld + c] = 0.f;
      }
      for (int pix = lpix + GATHER_THREADS; pix < npix; pix += GATHER_THREADS) {  // windows of more than 256 pixels
        if (p.lidar) {
          const int py = pix / S, px = pix - py * S;
          const float raw = __ldg(p.lidar + (size_t)reflect_sym(y + py - p.nb, p.Hl) * p.Wl + reflect_sym(x + px - p.nb, p.Wl));
          img[pix * ld + p.C] = p.lminmax ? __fdiv_rn(__fsub_rn(raw, lmin), lmax) : raw;
        }
        for (int c = n_out_c; c < ld; c++) img[pix * ld + c] = 0.f;
      }
    }
    __syncthreads();
    // ---- next patch's loads go out now
    const int64_t nn = n + gridDim.x;
    if (nn < p.N) {
      x = __ldg(p.xy + 2 * nn); y = __ldg(p.xy + 2 * nn + 1);
      gather_issue<T>(p, casi, x, y, chunk, slot, rows_per_iter, active, L);
    }
    // ---- stream the image out, 16 bytes per store
    float4* const gvec = reinterpret_cast<float4*>(gout - shift);
    const int nvec = (shift + total + 3) >> 2;
    for (int v = threadIdx.x; v < nvec; v += GATHER_THREADS) {
      const float4 val = gather_smem4[v];
      const int e0 = 4 * v - shift;  // patch element of val.x
      if (e0 >= 0 && e0 + 4 <= total) {
        __stcs(gvec + v, val);
      } else {
        const float f[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (e0 + t >= 0 && e0 + t < total) __stcs(gout + e0 + t, f[t]);
      }
    }
    __syncthreads();
  }
}

}  // namespace hyp

using namespace hyp;

extern "C" {

int hyp_scene_minmax(const void* cube, int dtype, int H, int W, int C, float* min_out, float* max_out, void* stream) {
  HYP_CHECK_ARG(cube && min_out && max_out && H > 0 && W > 0 && C > 0, "bad argument");
  HYP_CHECK_ARG(dtype == HYP_DT_F32 || dtype == HYP_DT_U16, "dtype must be f32 or u16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* bits = nullptr;
  HYP_CUDA(cudaMallocAsync(&bits, 2 * (size_t)C * sizeof(unsigned int), st));
  HYP_CUDA(cudaMemsetAsync(bits, 0xff, (size_t)C * sizeof(unsigned int), st));
  HYP_CUDA(cudaMemsetAsync(bits + C, 0x00, (size_t)C * sizeof(unsigned int), st));
  const int64_t pixels = (int64_t)H * W;
  const int ppb = (int)std::max<int64_t>(16, cdiv(pixels, 148 * 8));
  const unsigned grid = (unsigned)cdiv(pixels, ppb);
  const int threads = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
  if (dtype == HYP_DT_U16)
    scene_min_kernel<unsigned short><<<grid, threads, 0, st>>>((const unsigned short*)cube, pixels, C, ppb, bits);
  else
    scene_min_kernel<float><<<grid, threads, 0, st>>>((const float*)cube, pixels, C, ppb, bits);
  HYP_LAUNCHED();
  decode_ordered_kernel<<<(unsigned)cdiv(C, 128), 128, 0, st>>>(bits, C, min_out);
  HYP_LAUNCHED();
  if (dtype == HYP_DT_U16)
    scene_max_kernel<unsigned short><<<grid, threads, 0, st>>>((const unsigned short*)cube, pixels, C, ppb, min_out,
                                                               bits + C);
  else
    scene_max_kernel<float><<<grid, threads, 0, st>>>((const float*)cube, pixels, C, ppb, min_out, bits + C);
  HYP_LAUNCHED();
  decode_ordered_kernel<<<(unsigned)cdiv(C, 128), 128, 0, st>>>(bits + C, C, max_out);
  HYP_LAUNCHED();
  HYP_CUDA(cudaFreeAsync(bits, st));
  return HYP_OK;
}


int hyp_gather_patches(const void* casi, int casi_dtype, int Hc, int Wc, int C_hsi, const float* casi_min,
                       const float* casi_max, const float* lidar, int Hl, int Wl, const float* lidar_minmax,
                       int neighborhood, int mode, const int32_t* targets_xy, int64_t N, float* out, int out_ld,
                       void* stream) {
  HYP_CHECK_ARG(casi && targets_xy && out, "null argument");
  HYP_CHECK_ARG(casi_dtype == HYP_DT_F32 || casi_dtype == HYP_DT_U16, "casi dtype must be f32 or u16");
  HYP_CHECK_ARG(Hc > 0 && Wc > 0 && C_hsi > 0 && neighborhood >= 0 && N >= 0, "bad shape");
  HYP_CHECK_ARG((casi_min == nullptr) == (casi_max == nullptr), "casi_min and casi_max go together");
  HYP_CHECK_ARG(mode == HYP_GATHER_SAME_RES || mode == HYP_GATHER_GRSS2018, "unknown gather mode");
  HYP_CHECK_ARG(!lidar || (Hl > 0 && Wl > 0), "bad lidar shape");
  HYP_CHECK_ARG(mode != HYP_GATHER_GRSS2018 || lidar, "GRSS2018 mode needs the LiDAR raster");
  HYP_CHECK_ARG(out_ld >= C_hsi + (lidar ? 1 : 0), "out_ld too small");
  HYP_CHECK_ARG(neighborhood <= Hc && neighborhood <= Wc, "neighborhood larger than the scene");
  HYP_CHECK_ARG(N <= INT32_MAX, "too many targets for one call");
  if (N == 0) return HYP_OK;
  GatherArgs a;
  a.casi = casi; a.casi_u16 = casi_dtype == HYP_DT_U16; a.Hc = Hc; a.Wc = Wc; a.C = C_hsi;
  a.cmin = casi_min; a.cmax = casi_max; a.lidar = lidar; a.Hl = Hl; a.Wl = Wl; a.lminmax = lidar_minmax;
  a.nb = neighborhood; a.mode = mode; a.xy = targets_xy; a.N = N; a.out = out; a.out_ld = out_ld;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = 2 * neighborhood + 1;
  // algorithmic bytes: every patch element written once (fp32) + every window element read once (scene dtype)
  const double gather_bytes = (double)N * S * S * (4.0 * out_ld + (a.casi_u16 ? 2.0 : 4.0) * C_hsi + (lidar ? 4.0 : 0.0));
  const int epc = a.casi_u16 ? 8 : 4;
  const size_t smem = ((size_t)S * S * out_ld + 8) * sizeof(float);
  const char* force_v1 = getenv("HYP_GATHER_SCALAR");  // measurement knob: the element-wise kernel
  const bool rows_ok = C_hsi % epc == 0 && C_hsi / epc <= GATHER_THREADS && (reinterpret_cast<uintptr_t>(casi) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 3) == 0 && smem <= 200 * 1024 && !(force_v1 && force_v1[0] == '1');
  if (!rows_ok) {
    PROF("gather_kernel", gather_bytes, (gather_kernel<<<(unsigned)N, 256, 0, st>>>(a)));
    return HYP_OK;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  // persistent grid: as many blocks as stay resident (registers / shared memory decide), each walks the target list
  auto launch = [&](auto kernel) -> int {
    HYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HYP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 1;
    HYP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, GATHER_THREADS, smem));
    const unsigned grid = (unsigned)std::min<int64_t>(N, (int64_t)sms * std::max(1, per_sm));
    PROF("gather_rows_kernel", gather_bytes, (kernel<<<grid, GATHER_THREADS, smem, st>>>(a)));
    return HYP_OK;
  };
  return a.casi_u16 ? launch(gather_rows_kernel<unsigned short>) : launch(gather_rows_kernel<float>);
}

}  // extern "C"
