// HBM-bound kernels of the tensor-core path.  Activations live position-major:
//   tensor[pos][b][ld]   (pos = h*P + w of the patch, b = sample, ld = channels padded to 4)
// so every spatial tap of a k x k conv is a plain row-offset into the same matrix, and they
// come as an fp32 value plane followed by the GEMM operand planes of the model's operand format (hyp_tc.cuh OP_*):
//   OP_TF32X3 : the value plane doubles as the TF32 "hi" operand (the tensor core ignores its 13 low mantissa
//               bits); the second plane holds the TF32 rounding of what it drops
//   OP_F16X3  : two fp16 planes (hi, lo) of plane_elems elements each, in the bytes of the second fp32 plane
//   OP_BF16   : one bf16 plane there
// `lo` below is always the start of that second region, `op_plane` the element stride between 16-bit planes.
#pragma once
#include "hyp_kernels.cuh"
#include "hyp_tc.cuh"

namespace hyp {
namespace tc {

__device__ __forceinline__ float tf32_lo(float a) {
  return tf32_rna(a - __uint_as_float(__float_as_uint(a) & 0xffffe000u));
}

// operand planes of one element / of 4 consecutive elements at index `idx` of a tensor whose second region starts at lo
__device__ __forceinline__ void tc_store_operand(float* lo, size_t op_plane, int op, int64_t idx, float v) {
  if (op == OP_TF32X3) lo[idx] = tf32_lo(v);
  else store_op16(reinterpret_cast<uint16_t*>(lo) + idx, op_plane, op, v);
}
__device__ __forceinline__ void tc_store_operand4(float* lo, size_t op_plane, int op, int64_t idx, float v0, float v1, float v2,
                                                  float v3) {
  if (op == OP_TF32X3)
    *reinterpret_cast<float4*>(lo + idx) = make_float4(tf32_lo(v0), tf32_lo(v1), tf32_lo(v2), tf32_lo(v3));
  else
    store_op16x4(reinterpret_cast<uint16_t*>(lo) + idx, op_plane, op, v0, v1, v2, v3);
}

// view of x [B][P0][P0][C0] (NHWC patches as the importer hands them over): channels [c0, c0 + C), window cropped
// by `crop` pixels per side to P x P  -> planes [P*P][B][ld]
__global__ void __launch_bounds__(256) tc_prep_input_kernel(const float* __restrict__ x, int B, int P0, int C0, int c0, int crop,
                                                            int P, int C, int ld, float* __restrict__ hi, float* __restrict__ lo,
                                                            int op, size_t op_plane) {
  // one warp per (sample, pixel) row of C contiguous channels: the index arithmetic is per row, not per element
  const int PP = P * P, lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * PP;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t bp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bp < rows; bp += wstride) {
    const int b = (int)(bp / PP), pos = (int)(bp - (int64_t)b * PP);
    const int ph = pos / P, pw = pos - ph * P;
    const float* src = x + (((int64_t)b * P0 + ph + crop) * P0 + pw + crop) * C0 + c0;
    const int64_t o = ((int64_t)pos * B + b) * ld;
    // 4 channels per lane: the source row (C floats at an arbitrary 4-byte phase) is read element-wise, the planes
    // are written 16 / 8 bytes at a time (rows of the planes are 16-byte aligned; columns C..ld-1 are padding)
    for (int c = 4 * lane; c < C; c += 128) {
      const float v0 = __ldg(src + c), v1 = c + 1 < C ? __ldg(src + c + 1) : 0.f;
      const float v2 = c + 2 < C ? __ldg(src + c + 2) : 0.f, v3 = c + 3 < C ? __ldg(src + c + 3) : 0.f;
      *reinterpret_cast<float4*>(hi + o + c) = make_float4(v0, v1, v2, v3);
      tc_store_operand4(lo, op_plane, op, o + c, v0, v1, v2, v3);
    }
  }
}
__global__ void tc_fill_kernel(float* __restrict__ p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// position-major [PP][B][ld] -> dense [B][PP][C] (API outputs, tests)
// the same from the fp16 (hi, lo) operand planes of a tensor that keeps no value plane
__global__ void tc_extract_planes_kernel(const uint16_t* __restrict__ op, size_t plane, int ld, int B, int PP, int C,
                                         float* __restrict__ dst) {
  const int64_t total = (int64_t)B * PP * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bp = i / C;
    const int c = (int)(i - bp * C);
    const int b = (int)(bp / PP), pos = (int)(bp - (int64_t)b * PP);
    const int64_t at = ((int64_t)pos * B + b) * ld + c;
    dst[i] = __half2float(__ushort_as_half(op[at])) + __half2float(__ushort_as_half(op[at + plane]));
  }
}
__global__ void tc_extract_kernel(const float* __restrict__ src, int ld, int B, int PP, int C, float* __restrict__ dst) {
  const int64_t total = (int64_t)B * PP * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bp = i / C;
    const int c = (int)(i - bp * C);
    const int b = (int)(bp / PP), pos = (int)(bp - (int64_t)b * PP);
    dst[i] = src[((int64_t)pos * B + b) * ld + c];
  }
}

// ------------------------------------------------------------------------------------------
// weight packing: dst[r][k] (two planes) = src[r1*sr1 + r0*sr0 + k1*sk1 + k0*sk0] or 0, with
// r = r1*RD + r0, k = k1*KD + k0, valid iff r0 < RV and k0 < KV.
struct PackJob {
  int64_t src_off;  // element offset inside params
  int64_t dst_off;  // element offset inside one plane of the packed buffer
  int32_t rows, cols, ld;
  int32_t RD, RV, KD, KV;
  int32_t sr1, sr0, sk1, sk0;
  int32_t pad;
};
// One CTA packs one 32 x 32 (rows x cols) tile of one job; tile_first[j] = first tile of job j (prefix sums, njobs + 1
// entries).  The source is read along whichever of (r0, k0) is contiguous in params (sr0 == 1: a transpose through
// shared memory), the two planes are always written along k.
__global__ void __launch_bounds__(256) tc_pack_weights_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ tile_first,
                                                              int njobs, const float* __restrict__ params,
                                                              float* __restrict__ hi, float* __restrict__ lo, int op,
                                                              size_t op_plane, float w_scale) {
  __shared__ float tile[32][33];
  int jlo = 0, jhi = njobs - 1;  // last job with tile_first <= blockIdx.x
  while (jlo < jhi) {
    const int mid = (jlo + jhi + 1) >> 1;
    if (__ldg(&tile_first[mid]) <= (int)blockIdx.x) jlo = mid; else jhi = mid - 1;
  }
  const PackJob J = jobs[jlo];
  const int t = (int)blockIdx.x - __ldg(&tile_first[jlo]);
  const int ctiles = (J.cols + 31) >> 5;
  const int r_t = (t / ctiles) << 5, k_t = (t % ctiles) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool transpose = J.sr0 == 1;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int rr = transpose ? tx : ty + 8 * i, kk = transpose ? ty + 8 * i : tx;
    const int r = r_t + rr, k = k_t + kk;
    const int r1 = r / J.RD, r0 = r - r1 * J.RD;
    const int k1 = k / J.KD, k0 = k - k1 * J.KD;
    float v = 0.f;
    if (r < J.rows && k < J.cols && r0 < J.RV && k0 < J.KV)
      v = params[J.src_off + (int64_t)r1 * J.sr1 + (int64_t)r0 * J.sr0 + (int64_t)k1 * J.sk1 + (int64_t)k0 * J.sk0];
    tile[rr][kk] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int rr = ty + 8 * i, r = r_t + rr, k = k_t + tx;
    if (r < J.rows && k < J.cols) {
      const float v = tile[rr][tx];
      const int64_t o = J.dst_off + (int64_t)r * J.ld + k;
      if (op == OP_TF32X3) {
        hi[o] = v;
        lo[o] = tf32_lo(v);
      } else {  // 16-bit packs hold the operand planes only (hi = start of the pack), scaled into the fp16 sweet spot
        store_op16(reinterpret_cast<uint16_t*>(hi) + o, op_plane, op, v * w_scale);
      }
    }
  }
}

// Row sweep of the 2-D element-wise kernels: rows go in groups of G; block row y takes groups y, y + gridDim.y, ... so the
// whole grid moves through the tensor as one narrow front, ascending or (rev) descending.  Consecutive passes over the
// same tensors alternate the direction: a pass then starts on the rows its predecessor touched last, which are still
// in the 126 MB L2.
__device__ __forceinline__ int64_t tc_sweep_groups(int64_t rows, int G) { return (rows + G - 1) / G; }
__device__ __forceinline__ int64_t tc_sweep_row(int64_t i, int64_t ng, int rev, int G) { return (rev ? ng - 1 - i : i) * G; }

__device__ __forceinline__ void tc_load_ch4(const float* __restrict__ src, int c0, int C, float (&v)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = (c0 + k < C) ? src[c0 + k] : 0.f;
}

struct TcApplyArgs {
  const float* z;  // [rows][ldz] pre-BN
  int ldz;
  const float *mean, *rstd, *beta;
  float *hi, *lo;  // output: value plane, operand region [rows][ldo]
  int ldo;
  int op;          // OP_*
  size_t op_plane;
  int64_t rows;
  int C, act;
  float alpha, keep;
  uint64_t seed;
  uint32_t stream_id;
  // residual sources [rows][ld]; idx == NULL: identity channel map.  res = the source's fp32 value plane, or NULL
  // when it keeps none: the value is then rebuilt from its fp16 (hi, lo) operand planes res_op / res_plane
  // pat: 0 generic idx table, 2 = tf.repeat by 2 (idx[j] = j / 2), 3 = strided gather with step 2 (idx[j] = 2 j): the
  // two regular maps are read with vector loads instead of one indexed scalar load per element
  const float* res0;
  const int* idx0;
  int ld0, has0, pat0;
  const uint16_t* res0_op;
  size_t res0_plane;
  const float* res1;
  const int* idx1;
  int ld1, has1, pat1;
  const uint16_t* res1_op;
  size_t res1_plane;
};

// out = dropout(act(bn(z))) + res0[:, idx0] + res1[:, idx1]; VEC channels per thread
template <int VEC>
struct TcApplyItem {
  int64_t m;
  int c0;
  float v[VEC], r[VEC];
};
// hi + lo of the fp16 operand planes = the value to 22 significand bits
__device__ __forceinline__ float tc_value_of_planes(const uint16_t* op, size_t plane, int64_t idx) {
  return __half2float(__ushort_as_half(op[idx])) + __half2float(__ushort_as_half(op[idx + plane]));
}
__device__ __forceinline__ float tc_half_bits(uint32_t w, int hi) { return __half2float(__ushort_as_half((unsigned short)(hi ? w >> 16 : w & 0xffffu))); }
template <int VEC>
__device__ __forceinline__ void tc_bn_apply_residual(const float* res, const uint16_t* op, size_t plane, const int* idx, int ld,
                                                     int C, int pat, TcApplyItem<VEC>& it) {
  float add[VEC];
  if (VEC == 4 && !idx) {  // identity: the same 4 channels
    if (res) {
      const float4 t = *reinterpret_cast<const float4*>(res + it.m * ld + it.c0);
      add[0] = t.x; add[1 % VEC] = t.y; add[2 % VEC] = t.z; add[3 % VEC] = t.w;
    } else {
      const uint2 h = *reinterpret_cast<const uint2*>(op + it.m * ld + it.c0);
      const uint2 l = *reinterpret_cast<const uint2*>(op + plane + it.m * ld + it.c0);
      add[0] = tc_half_bits(h.x, 0) + tc_half_bits(l.x, 0); add[1 % VEC] = tc_half_bits(h.x, 1) + tc_half_bits(l.x, 1);
      add[2 % VEC] = tc_half_bits(h.y, 0) + tc_half_bits(l.y, 0); add[3 % VEC] = tc_half_bits(h.y, 1) + tc_half_bits(l.y, 1);
    }
  } else if (VEC == 4 && pat == 2 && it.c0 + 4 <= C) {  // repeat by 2: source channels c0 / 2, c0 / 2 + 1
    float a, b;
    if (res) {
      const float2 t = *reinterpret_cast<const float2*>(res + it.m * ld + it.c0 / 2);
      a = t.x; b = t.y;
    } else {
      const uint32_t h = *reinterpret_cast<const uint32_t*>(op + it.m * ld + it.c0 / 2);
      const uint32_t l = *reinterpret_cast<const uint32_t*>(op + plane + it.m * ld + it.c0 / 2);
      a = tc_half_bits(h, 0) + tc_half_bits(l, 0); b = tc_half_bits(h, 1) + tc_half_bits(l, 1);
    }
    add[0] = a; add[1 % VEC] = a; add[2 % VEC] = b; add[3 % VEC] = b;
  } else if (VEC == 4 && pat == 3 && it.c0 + 4 <= C) {  // stride 2: source channels 2 c0, 2 c0 + 2, 2 c0 + 4, 2 c0 + 6
    if (res) {
      const float4 t = *reinterpret_cast<const float4*>(res + it.m * ld + 2 * it.c0);
      const float4 u = *reinterpret_cast<const float4*>(res + it.m * ld + 2 * it.c0 + 4);
      add[0] = t.x; add[1 % VEC] = t.z; add[2 % VEC] = u.x; add[3 % VEC] = u.z;
    } else {
      const uint4 h = *reinterpret_cast<const uint4*>(op + it.m * ld + 2 * it.c0);
      const uint4 l = *reinterpret_cast<const uint4*>(op + plane + it.m * ld + 2 * it.c0);
      add[0] = tc_half_bits(h.x, 0) + tc_half_bits(l.x, 0); add[1 % VEC] = tc_half_bits(h.y, 0) + tc_half_bits(l.y, 0);
      add[2 % VEC] = tc_half_bits(h.z, 0) + tc_half_bits(l.z, 0); add[3 % VEC] = tc_half_bits(h.w, 0) + tc_half_bits(l.w, 0);
    }
  } else {
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const int64_t at = it.m * ld + (idx ? idx[min(it.c0 + j, C - 1)] : it.c0 + j);
      add[j] = res ? res[at] : tc_value_of_planes(op, plane, at);
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; j++) it.r[j] += add[j];
}
// issue every global load of one item (z and the residual sources) ...
template <int VEC>
__device__ __forceinline__ void tc_bn_apply_load(const TcApplyArgs& p, int64_t m, int cg, TcApplyItem<VEC>& it) {
  it.m = m;
  it.c0 = cg * VEC;
  if (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p.z + it.m * p.ldz + it.c0);
    it.v[0] = t.x; it.v[1 % VEC] = t.y; it.v[2 % VEC] = t.z; it.v[3 % VEC] = t.w;
  } else {
    it.v[0] = p.z[it.m * p.ldz + it.c0];
  }
#pragma unroll
  for (int j = 0; j < VEC; j++) it.r[j] = 0.f;
  if (p.has0) tc_bn_apply_residual<VEC>(p.res0, p.res0_op, p.res0_plane, p.idx0, p.ld0, p.C, p.pat0, it);
  if (p.has1) tc_bn_apply_residual<VEC>(p.res1, p.res1_op, p.res1_plane, p.idx1, p.ld1, p.C, p.pat1, it);
}
// ... then normalise, activate, drop out, add the residuals and write the two planes
template <int VEC>
__device__ __forceinline__ void tc_bn_apply_finish(const TcApplyArgs& p, TcApplyItem<VEC>& it) {
#pragma unroll
  for (int j = 0; j < VEC; j++) {
    const int c = min(it.c0 + j, p.C - 1);  // columns C..ldo-1 of the last group are padding (their value is unused)
    float y = (it.v[j] - p.mean[c]) * p.rstd[c] + p.beta[c];
    y = act_fwd(y, p.act, p.alpha);
    if (p.keep < 1.f) y = (philox_uniform(p.seed, p.stream_id, (uint64_t)(it.m * p.C + c)) < p.keep) ? y / p.keep : 0.f;
    it.v[j] = y + it.r[j];
  }
  if (VEC == 4) {
    if (p.hi) *reinterpret_cast<float4*>(p.hi + it.m * p.ldo + it.c0) = make_float4(it.v[0], it.v[1 % VEC], it.v[2 % VEC], it.v[3 % VEC]);
    tc_store_operand4(p.lo, p.op_plane, p.op, it.m * p.ldo + it.c0, it.v[0], it.v[1 % VEC], it.v[2 % VEC], it.v[3 % VEC]);
  } else {
    if (p.hi) p.hi[it.m * p.ldo + it.c0] = it.v[0];
    tc_store_operand(p.lo, p.op_plane, p.op, it.m * p.ldo + it.c0, it.v[0]);
  }
}
// two items per thread and iteration, all loads issued before the first dependent instruction (four items measured
// 35 % slower: the item array goes to local memory)
template <int VEC>
__global__ void __launch_bounds__(256) tc_bn_apply_kernel(const TcApplyArgs p) {
  const int cq = (p.C + VEC - 1) / VEC;
  const int64_t total = p.rows * cq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // item i = (row m, channel group cg); the walk advances (m, cg) by the constant stride with a carry instead of
  // dividing a 64-bit index per item (the emulated 64-bit division was most of this kernel's instructions)
  const int64_t dm = stride / cq;
  const int dc = (int)(stride - dm * cq);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t m = i / cq;
  int cg = (int)(i - m * cq);
  for (; i < total; i += 2 * stride) {
    TcApplyItem<VEC> a, b;
    const bool two = i + stride < total;
    int64_t m2 = m + dm;
    int cg2 = cg + dc;
    if (cg2 >= cq) { cg2 -= cq; m2++; }
    tc_bn_apply_load<VEC>(p, m, cg, a);
    if (two) tc_bn_apply_load<VEC>(p, m2, cg2, b);
    tc_bn_apply_finish<VEC>(p, a);
    if (two) tc_bn_apply_finish<VEC>(p, b);
    m = m2 + dm;
    cg = cg2 + dc;
    if (cg >= cq) { cg -= cq; m++; }
  }
}

struct TcBnBwdArgs {
  const float* gout;  // [rows][ldg] gradient w.r.t. the layer's output tensor
  int ldg;
  const float* z;     // [rows][ldz]
  int ldz;
  const float *mean, *rstd, *beta;
  int64_t rows;
  int C, act;
  float alpha, keep;
  uint64_t seed;
  uint32_t stream_id;
  float* part;        // [row blocks][2][C]   (reduce)
  const float *s1, *s2;  // [C] means of g_y and g_y*zhat   (apply)
  float *gz_hi, *gz_lo;  // [rows][ldgz]                   (apply)  OP_TF32X3: two fp32 planes
  int ldgz;
  int op;                // 16-bit formats: gz_hi = start of the operand planes, gz_plane = their element stride
  size_t gz_plane;
  const unsigned int* gmax_bits;  // OP_F16X3: bits of an upper bound of |gz| over the layer (bn_bwd_finalize)
  float* gz_scale_out;   // OP_F16X3: [2] = (power-of-two scale applied to gz before the fp16 split, its inverse)
  int gcols;          // columns of gz to write
  int fpad, f, R;     // level layers: gz column j = slot*fpad + n holds channel q*f + jt*ft + n with
  int nt, ft;         //   q = R-1 - slot/nt, jt = slot % nt (n < min(ft, f - jt*ft)); fpad == 0: identity
};

// OP_F16X3: gz is multiplied by a power of two before the fp16 split so that its largest possible magnitude lands just
// below 2^15 (fp16 tops out at 65504; remainders of values >= 2^-3 stay normal fp16 numbers).  The bound comes from the
// backward statistics pass (tc_bn_bwd_finalize8_kernel); the inverse goes to the dgrad / wgrad epilogues.
__device__ __forceinline__ float tc_gz_scale(const TcBnBwdArgs& p) {
  if (p.op != OP_F16X3 || !p.gmax_bits) return 1.f;
  const float bound = __uint_as_float(__ldg(p.gmax_bits));  // non-negative floats order like their bit patterns
  if (!(bound > 0.f) || !(bound < 3e38f)) return 1.f;
  int e;
  frexpf(bound, &e);  // bound < 2^e
  return ldexpf(1.f, max(-100, min(100, 15 - e)));
}
__device__ __forceinline__ void tc_store_gz4(const TcBnBwdArgs& p, int64_t idx, float scale, float v0, float v1, float v2, float v3) {
  if (p.op == OP_TF32X3) {
    *reinterpret_cast<float4*>(p.gz_hi + idx) = make_float4(v0, v1, v2, v3);
    *reinterpret_cast<float4*>(p.gz_lo + idx) = make_float4(tf32_lo(v0), tf32_lo(v1), tf32_lo(v2), tf32_lo(v3));
  } else {
    store_op16x4(reinterpret_cast<uint16_t*>(p.gz_hi) + idx, p.gz_plane, p.op, v0 * scale, v1 * scale, v2 * scale, v3 * scale);
  }
}
__device__ __forceinline__ void tc_store_gz(const TcBnBwdArgs& p, int64_t idx, float scale, float v) {
  if (p.op == OP_TF32X3) {
    p.gz_hi[idx] = v;
    p.gz_lo[idx] = tf32_lo(v);
  } else {
    store_op16(reinterpret_cast<uint16_t*>(p.gz_hi) + idx, p.gz_plane, p.op, v * scale);
  }
}
__device__ __forceinline__ void tc_publish_gz_scale(const TcBnBwdArgs& p, float scale) {
  if (p.gz_scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    p.gz_scale_out[0] = scale;
    p.gz_scale_out[1] = 1.f / scale;
  }
}

// ------------------------------------------------------------------------------------------
// Vectorised backward passes (statistics, gz, residual pushes).  A block is TX column groups (4 channels,
// one float4 each) x TY = 256 / TX row lanes; every thread walks its rows with 4 independent
// loads in flight, so a warp reads TX * 16 contiguous bytes of several rows per step and there
// is no per-element index division.  Rows are 16-byte aligned (ld % 4 == 0); the channel tail
// (C % 4 != 0) is masked per element.
struct Gy4 { float g[4], zh[4]; };
// DROP = false: the layer has no dropout (every conv layer): the Philox recomputation of the mask is compiled out,
// which is most of the kernels' registers
template <bool DROP>
__device__ __forceinline__ Gy4 tc_bn_gy4(const TcBnBwdArgs& p, int64_t m, int c0, const float4 gv, const float4 zv,
                                         const float (&mean)[4], const float (&rstd)[4], const float (&beta)[4]) {
  Gy4 r;
  const float g_in[4] = {gv.x, gv.y, gv.z, gv.w}, z_in[4] = {zv.x, zv.y, zv.z, zv.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float zhat = (z_in[k] - mean[k]) * rstd[k];
    const float y = zhat + beta[k];
    float g = g_in[k];
    if (DROP && p.keep < 1.f) g = (philox_uniform(p.seed, p.stream_id, (uint64_t)(m * p.C + c0 + k)) < p.keep) ? g / p.keep : 0.f;
    if (p.act == ACT_LRELU) {
      g = (y > 0.f) ? g : g * p.alpha;
    } else if (p.act == ACT_SIGMOID) {
      const float sg = 1.f / (1.f + __expf(-y));
      g = g * sg * (1.f - sg);
    }
    r.g[k] = g;
    r.zh[k] = zhat;
  }
  return r;
}

template <int TX, bool DROP>
__global__ void __launch_bounds__(256) tc_bn_bwd_reduce_v4_kernel(const TcBnBwdArgs p, int rev) {
  constexpr int TY = 256 / TX;
  __shared__ float sh[4][TY][TX * 4 + 4];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int c0 = (blockIdx.x * TX + tx) * 4;
  const int64_t r1 = p.rows, ng = tc_sweep_groups(p.rows, 4 * TY);
  float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
  float mg[4] = {0.f, 0.f, 0.f, 0.f}, mz[4] = {0.f, 0.f, 0.f, 0.f};  // max |g_y|, max |zhat| (bound of |gz|, OP_F16X3)
  if (c0 < p.C) {
    float mean[4], rstd[4], beta[4];
    tc_load_ch4(p.mean, c0, p.C, mean);
    tc_load_ch4(p.rstd, c0, p.C, rstd);
    tc_load_ch4(p.beta, c0, p.C, beta);
    for (int64_t gi = blockIdx.y; gi < ng; gi += gridDim.y) {
      const int64_t r = tc_sweep_row(gi, ng, rev, 4 * TY) + ty;
      float4 gv[4], zv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int64_t rr = r + u * TY;
        if (rr < r1) {
          gv[u] = *reinterpret_cast<const float4*>(p.gout + rr * p.ldg + c0);
          zv[u] = *reinterpret_cast<const float4*>(p.z + rr * p.ldz + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int64_t rr = r + u * TY;
        if (rr < r1) {
          const Gy4 y = tc_bn_gy4<DROP>(p, rr, c0, gv[u], zv[u], mean, rstd, beta);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            a1[k] += y.g[k];
            a2[k] += y.g[k] * y.zh[k];
            mg[k] = fmaxf(mg[k], fabsf(y.g[k]));
            mz[k] = fmaxf(mz[k], fabsf(y.zh[k]));
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    sh[0][ty][tx * 4 + k] = a1[k]; sh[1][ty][tx * 4 + k] = a2[k];
    sh[2][ty][tx * 4 + k] = mg[k]; sh[3][ty][tx * 4 + k] = mz[k];
  }
  __syncthreads();
  // TX * 4 columns x (2 sums + 2 maxima), reduced over the TY row lanes by the first TX * 16 threads
  for (int o = threadIdx.x; o < TX * 16; o += 256) {
    const int which = o / (TX * 4), col = o % (TX * 4);
    const int c = blockIdx.x * TX * 4 + col;
    if (c < p.C) {
      float a = 0.f;
      if (which < 2) {
#pragma unroll
        for (int i = 0; i < TY; i++) a += sh[which][i][col];
      } else {
#pragma unroll
        for (int i = 0; i < TY; i++) a = fmaxf(a, sh[which][i][col]);
      }
      p.part[((size_t)blockIdx.y * 4 + which) * p.C + c] = a;
    }
  }
}

// gz = rstd * (g_y - mean(g_y) - zhat * mean(g_y * zhat)) -> (value, lo) planes.  Level layers (fpad > 0,
// f % 4 == 0) write slot-ordered columns: gz column j = slot * fpad + n holds channel (R-1-slot) * f + n.
// row[c0 .. c0 + 3] for any c0 (the row itself is 16-byte aligned)
__device__ __forceinline__ float4 tc_load4_shifted(const float* __restrict__ row, int c0) {
  const int base = c0 & ~3, sh = c0 & 3;
  const float4 a = *reinterpret_cast<const float4*>(row + base);
  if (sh == 0) return a;
  const float4 b = *reinterpret_cast<const float4*>(row + base + 4);
  if (sh == 1) return make_float4(a.y, a.z, a.w, b.x);
  if (sh == 2) return make_float4(a.z, a.w, b.x, b.y);
  return make_float4(a.w, b.x, b.y, b.z);
}

// SHIFT: the source channels of a thread's 4 columns may start at any 4-byte phase (level kernels with f % 4 != 0)
template <int TX, bool SHIFT, bool DROP>
__global__ void __launch_bounds__(256, 2) tc_bn_bwd_apply_v4_kernel(const TcBnBwdArgs p, int rev) {
  constexpr int TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int j0 = (blockIdx.x * TX + tx) * 4;  // first gz column of this thread
  const float gscale = tc_gz_scale(p);
  tc_publish_gz_scale(p, gscale);
  if (j0 >= p.gcols) return;
  int c0 = j0;
  int nv = min(4, p.C - j0);  // valid values of this thread's 4 columns (<= 0: padding only)
  if (p.fpad) {
    const int slot = j0 / p.fpad, n = j0 - slot * p.fpad;
    const int jt = slot % p.nt;
    nv = slot < p.R * p.nt ? min(4, min(p.ft, p.f - jt * p.ft) - n) : 0;  // a slot's tail columns are padding
    c0 = (p.R - 1 - slot / p.nt) * p.f + jt * p.ft + n;
  }
  const int64_t r1 = p.rows, ng = tc_sweep_groups(p.rows, 4 * TY);
  if (nv <= 0) {  // padding columns of the slot layout: zeros
    for (int64_t gi = blockIdx.y; gi < ng; gi += gridDim.y) {
      const int64_t r = tc_sweep_row(gi, ng, rev, 4 * TY) + ty;
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (r + u * TY < r1) tc_store_gz4(p, (r + u * TY) * p.ldgz + j0, 1.f, 0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  float mean[4], rstd[4], beta[4], s1[4], s2[4];
  tc_load_ch4(p.mean, c0, p.C, mean);
  tc_load_ch4(p.rstd, c0, p.C, rstd);
  tc_load_ch4(p.beta, c0, p.C, beta);
  tc_load_ch4(p.s1, c0, p.C, s1);
  tc_load_ch4(p.s2, c0, p.C, s2);
  const bool vec = !SHIFT || (c0 & 3) == 0;  // level slots with f % 4 == 0 keep the 16-byte alignment of the source row
  for (int64_t gi = blockIdx.y; gi < ng; gi += gridDim.y) {
    const int64_t r = tc_sweep_row(gi, ng, rev, 4 * TY) + ty;
    float4 gv[4], zv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int64_t rr = r + u * TY;
      if (rr < r1) {
        if (vec) {
          gv[u] = *reinterpret_cast<const float4*>(p.gout + rr * p.ldg + c0);
          zv[u] = *reinterpret_cast<const float4*>(p.z + rr * p.ldz + c0);
        } else {
          // source channels at an arbitrary 4-byte phase (level kernels with f % 4 != 0): two aligned 16-byte loads
          // and a shift instead of four 4-byte loads per tensor.  The second load may run a few floats past the row
          // (into the next row / the padding behind the tensor inside the workspace); those values are never used.
          gv[u] = tc_load4_shifted(p.gout + rr * p.ldg, c0);
          zv[u] = tc_load4_shifted(p.z + rr * p.ldz, c0);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int64_t rr = r + u * TY;
      if (rr < r1) {
        const Gy4 y = tc_bn_gy4<DROP>(p, rr, c0, gv[u], zv[u], mean, rstd, beta);
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = k < nv ? rstd[k] * (y.g[k] - s1[k] - y.zh[k] * s2[k]) : 0.f;
        tc_store_gz4(p, rr * p.ldgz + j0, gscale, v[0], v[1], v[2], v[3]);
      }
    }
  }
}

// residual backward, 4 source channels per thread.  MODE 0: identity (gsrc += gout); MODE 1: generic
// contiguous ranges [lo[c], hi[c]) of the consumer's channels; MODE 2: tf.repeat by 2 (source channel c feeds
// consumer channels 2c, 2c+1: two float4 loads, pairwise sums); MODE 3: gather with stride `step` (consumer channel
// j reads source channel step*j: only every step-th source channel gets a gradient).
template <int TX, int MODE>
__global__ void __launch_bounds__(256) tc_resid_bwd_v4_kernel(const float* __restrict__ gout, int ldg, float* __restrict__ gsrc,
                                                              int lds, int Csrc, const int* __restrict__ lo,
                                                              const int* __restrict__ hi, int64_t rows, int accumulate,
                                                              int rev, int step) {
  constexpr int TY = 256 / TX;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int c0 = (blockIdx.x * TX + tx) * 4;
  if (c0 >= Csrc) return;
  const int64_t r1 = rows, ng = tc_sweep_groups(rows, 2 * TY);
  int jl[4] = {0, 0, 0, 0}, jh[4] = {0, 0, 0, 0};
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (c0 + k < Csrc) { jl[k] = lo[c0 + k]; jh[k] = hi[c0 + k]; }
  }
  for (int64_t gi = blockIdx.y; gi < ng; gi += gridDim.y) {
    const int64_t r = tc_sweep_row(gi, ng, rev, 2 * TY) + ty;
    float4 acc[2], old[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int64_t rr = r + u * TY;
      if (rr < r1) {
        if (MODE == 0) {
          acc[u] = *reinterpret_cast<const float4*>(gout + rr * ldg + c0);
        } else if (MODE == 2) {
          const float4 p = *reinterpret_cast<const float4*>(gout + rr * ldg + 2 * c0);
          const float4 q = *reinterpret_cast<const float4*>(gout + rr * ldg + 2 * c0 + 4);
          acc[u] = make_float4(p.x + p.y, p.z + p.w, q.x + q.y, q.z + q.w);
        } else if (MODE == 3) {
          const float* gp = gout + rr * ldg;
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; k++) v[k] = ((c0 + k) % step == 0 && c0 + k < Csrc) ? gp[(c0 + k) / step] : 0.f;
          acc[u] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          const float* gp = gout + rr * ldg;
          float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int k = 0; k < 4; k++)
            for (int j = jl[k]; j < jh[k]; j++) v[k] += gp[j];
          acc[u] = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (accumulate) old[u] = *reinterpret_cast<const float4*>(gsrc + rr * lds + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int64_t rr = r + u * TY;
      if (rr < r1) {
        float4 v = acc[u];
        if (accumulate) { v.x += old[u].x; v.y += old[u].y; v.z += old[u].z; v.w += old[u].w; }
        *reinterpret_cast<float4*>(gsrc + rr * lds + c0) = v;
      }
    }
  }
}

// per-row-block partial sums [nrows][2][ld] -> mean / rstd / moving statistics.  8 channels per block
// (one 32-byte sector per partial row) x FIN_LANES row lanes: the kernel is pure load latency (a few thousand
// partial rows, a few dozen channels), so the rows are spread over as many threads as a block holds and every
// thread keeps four loads in flight; the cross-lane sum is a shuffle tree + one shared-memory round.
constexpr int FIN_LANES = 128;
// E = entries (rows of ld floats) per partial record; the sums are entries 0 and 1
template <int E = 2>
__device__ __forceinline__ void tc_fin_reduce(const float* __restrict__ part, int nrows, size_t ld, int c, bool valid,
                                              double& a1, double& a2, double (*sh)[8][2]) {
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  a1 = a2 = 0;
  if (valid) {
    int r = ty;
    for (; r + FIN_LANES < nrows; r += 2 * FIN_LANES) {
      const float u0 = part[((size_t)r * E + 0) * ld + c], u1 = part[((size_t)r * E + 1) * ld + c];
      const float v0 = part[((size_t)(r + FIN_LANES) * E + 0) * ld + c], v1 = part[((size_t)(r + FIN_LANES) * E + 1) * ld + c];
      a1 += (double)u0 + (double)v0;
      a2 += (double)u1 + (double)v1;
    }
    if (r < nrows) {
      a1 += (double)part[((size_t)r * E + 0) * ld + c];
      a2 += (double)part[((size_t)r * E + 1) * ld + c];
    }
  }
  // lanes 8, 16 of a warp hold the same channel: fold the 4 row lanes of a warp, then the 32 warps
  a1 += __shfl_xor_sync(0xffffffffu, a1, 8);  a2 += __shfl_xor_sync(0xffffffffu, a2, 8);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 16); a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) < 8) { sh[warp][tx][0] = a1; sh[warp][tx][1] = a2; }
  __syncthreads();
  if (threadIdx.x < 8) {
    a1 = a2 = 0;
    for (int w = 0; w < FIN_LANES / 4; w++) { a1 += sh[w][tx][0]; a2 += sh[w][tx][1]; }
  }
}
__global__ void __launch_bounds__(8 * FIN_LANES) tc_bn_finalize8_kernel(const float* __restrict__ part, int nrows, int ld, int C,
                                                              double count, float eps, float decay,
                                                              float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                              int update_moving) {
  __shared__ double sh[FIN_LANES / 4][8][2];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  double a1, a2;
  tc_fin_reduce(part, nrows, (size_t)ld, c, c < C, a1, a2, sh);
  if (threadIdx.x < 8 && c < C) {
    const double mean = a1 / count;
    double var = a2 / count - mean * mean;
    if (var < 0) var = 0;
    const float meanf = (float)mean, varf = (float)var;
    mean_out[c] = meanf;
    const float x = varf + eps;
    float r = rsqrtf(x);
    r = r * (1.5f - 0.5f * x * r * r);
    rstd_out[c] = r;
    if (update_moving) {
      const double unbiased = var * (count / fmax(count - 1.0, 1.0));
      moving_mean[c] = moving_mean[c] * decay + meanf * (1.f - decay);
      moving_var[c] = moving_var[c] * decay + (float)unbiased * (1.f - decay);
    }
  }
}
// partial records of the backward statistics pass: [nblocks][4][C] = (sum g_y, sum g_y zhat, max |g_y|, max |zhat|).
// gmax_bits (nullable, zeroed by the caller): atomic maximum over channels of rstd * (max|g_y| + |s1| + max|zhat| |s2|),
// an upper bound of |gz| that sets the layer's fp16 scale (OP_F16X3).
__global__ void __launch_bounds__(8 * FIN_LANES) tc_bn_bwd_finalize8_kernel(const float* __restrict__ part, int nblocks, int C, double rows,
                                                                  float* __restrict__ s1, float* __restrict__ s2,
                                                                  float* __restrict__ gbeta, int bias_mode,
                                                                  const float* __restrict__ rstd, unsigned int* gmax_bits) {
  __shared__ double sh[FIN_LANES / 4][8][2];
  __shared__ float shm[FIN_LANES / 4][8][2];
  const int c = blockIdx.x * 8 + (threadIdx.x & 7);
  double a1, a2;
  tc_fin_reduce<4>(part, nblocks, (size_t)C, c, c < C, a1, a2, sh);
  float mg = 0.f, mz = 0.f;
  if (gmax_bits) {  // block-uniform
    if (c < C)
      for (int r = threadIdx.x >> 3; r < nblocks; r += FIN_LANES) {
        mg = fmaxf(mg, part[((size_t)r * 4 + 2) * C + c]);
        mz = fmaxf(mz, part[((size_t)r * 4 + 3) * C + c]);
      }
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 8));  mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, 8));
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 16)); mz = fmaxf(mz, __shfl_xor_sync(0xffffffffu, mz, 16));
    if ((threadIdx.x & 31) < 8) { shm[threadIdx.x >> 5][threadIdx.x & 7][0] = mg; shm[threadIdx.x >> 5][threadIdx.x & 7][1] = mz; }
    __syncthreads();
    if (threadIdx.x < 8)
      for (int w = 0; w < FIN_LANES / 4; w++) { mg = fmaxf(mg, shm[w][threadIdx.x][0]); mz = fmaxf(mz, shm[w][threadIdx.x][1]); }
  }
  if (threadIdx.x < 8 && c < C) {
    // bias layers have no batch statistics to differentiate through: gz = g_y, d bias = sum g_y
    const float m1 = bias_mode ? 0.f : (float)(a1 / rows), m2 = bias_mode ? 0.f : (float)(a2 / rows);
    s1[c] = m1;
    s2[c] = m2;
    gbeta[c] = (float)a1;
    if (gmax_bits) {
      const float bound = fabsf(rstd[c]) * (mg + fabsf(m1) + mz * fabsf(m2));
      if (bound == bound) atomicMax(gmax_bits, __float_as_uint(fminf(bound, 3e38f)));
    }
  }
}

// ------------------------------------------------------------------------------------------
// tf.nn.local_response_normalization with TF's defaults (CONCNNModel.py:37,41): depth_radius 5, bias 1, alpha 1,
// beta 0.5:  y[c] = x[c] / sqrt(1 + sum_{|j-c|<=5} x[j]^2) over the channel axis.  One warp per row (pixel), the row
// in shared memory.  Forward keeps the pre-LRN row in `u` (backward recomputes the sums from it) and rewrites the
// activation planes in place.  Backward rewrites the gradient in place:
//   dx[k] = g[k] / sqrt(s[k]) - x[k] * sum_{|i-k|<=5} g[i] x[i] s[i]^-1.5
constexpr int LRN_RADIUS = 5;
__global__ void __launch_bounds__(256) tc_lrn_fwd_kernel(float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ u,
                                                         int ld, int C, int64_t rows, int op, size_t op_plane) {
  extern __shared__ float lrn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* xs = lrn_sm + warp * C;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
    for (int c = lane; c < C; c += 32) {
      const float v = hi[r * ld + c];
      xs[c] = v;
      u[r * ld + c] = v;
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      float s = 1.f;
      const int j0 = max(0, c - LRN_RADIUS), j1 = min(C - 1, c + LRN_RADIUS);
      for (int j = j0; j <= j1; j++) s += xs[j] * xs[j];
      const float y = xs[c] / sqrtf(s);
      hi[r * ld + c] = y;
      tc_store_operand(lo, op_plane, op, r * ld + c, y);
    }
    __syncwarp();
  }
}
__global__ void __launch_bounds__(256) tc_lrn_bwd_kernel(float* __restrict__ g, const float* __restrict__ u, int ldg, int ldu, int C,
                                                         int64_t rows) {
  extern __shared__ float lrn_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* xs = lrn_sm + warp * 3 * C;
  float* ts = xs + C;
  float* gi = ts + C;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < rows; r += (int64_t)gridDim.x * nwarps) {
    for (int c = lane; c < C; c += 32) xs[c] = u[r * ldu + c];
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      float s = 1.f;
      const int j0 = max(0, c - LRN_RADIUS), j1 = min(C - 1, c + LRN_RADIUS);
      for (int j = j0; j <= j1; j++) s += xs[j] * xs[j];
      const float gv = g[r * ldg + c];
      const float rs = 1.f / sqrtf(s);
      gi[c] = gv * rs;
      ts[c] = gv * xs[c] * rs / s;
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) {
      float t = 0.f;
      const int j0 = max(0, c - LRN_RADIUS), j1 = min(C - 1, c + LRN_RADIUS);
      for (int j = j0; j <= j1; j++) t += ts[j];
      g[r * ldg + c] = gi[c] - xs[c] * t;
    }
    __syncwarp();
  }
}

// softmax cross-entropy per row (one warp per row) on padded logits + gradient
__global__ void tc_ce_loss_kernel(const float* __restrict__ logits, int ld, const uint8_t* __restrict__ labels,
                                  int64_t B, int classes, float* __restrict__ ce_out, float* __restrict__ glogits,
                                  int ldg, float gscale) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* lp = logits + row * ld;
  float mx = -INFINITY;
  for (int c = lane; c < classes; c += 32) mx = fmaxf(mx, lp[c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < classes; c += 32) se += expf(lp[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int lab = labels[row];
  if (lane == 0) ce_out[row] = lse - lp[lab];
  if (glogits) {
    for (int c = lane; c < classes; c += 32) {
      const float sm = expf(lp[c] - lse);
      glogits[row * ldg + c] = (sm - (c == lab ? 1.f : 0.f)) * gscale;
    }
  }
}

// sum (recon - x)^2: recon [B][ldr] with feature pos*C + c; x planes position-major [PP][B][ldx]
__global__ void __launch_bounds__(256) tc_mse_kernel(const float* __restrict__ recon, int ldr, const float* __restrict__ xin, int ldx,
                                                     int B, int PP, int C, double* __restrict__ acc, float* __restrict__ grecon,
                                                     int ldg, float gscale) {
  __shared__ float sh[32];
  // one warp per (sample, pixel): C contiguous features of recon against C contiguous channels of the x plane
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * PP;
  const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
  float s = 0.f;
  for (int64_t bp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bp < rows; bp += wstride) {
    const int b = (int)(bp / PP), pos = (int)(bp - (int64_t)b * PP);
    const float* r = recon + (int64_t)b * ldr + (int64_t)pos * C;
    const float* xv = xin + ((int64_t)pos * B + b) * ldx;
    float* g = grecon ? grecon + (int64_t)b * ldg + (int64_t)pos * C : nullptr;
    for (int c = lane; c < C; c += 32) {
      const float d = r[c] - xv[c];
      s += d * d;
      if (g) g[c] = 2.f * d * gscale;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(acc, (double)s);
  }
}

}  // namespace tc
}  // namespace hyp
