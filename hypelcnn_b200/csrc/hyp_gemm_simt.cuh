// fp32 FFMA implicit-GEMM kernels (HYP_PRECISION_FP32 path and the fallback for shapes the
// tcgen05 path does not take).  One "segment" = one tap of one conv kernel of a layer:
//   forward : Z[m, col0+n]   = sum_seg sum_c A[shift(m,seg), c]        * w_seg[c][n]
//   dgrad   : G[m, c]       += sum_seg sum_n gZ[shift(m,-seg), col0+n] * wt_seg[n][c]
//   wgrad   : gw_seg[c][n]  += sum_m  A[shift(m,seg), c]               * gZ[m, col0+n]
// A 1x1 conv / FC layer is the one-segment case (dy=dx=0, P=1 for FC).
// Rows m enumerate (sample, h, w) of a PxP patch; SAME zero padding = masked rows.
#pragma once
#include "hyp_common.cuh"

namespace hyp {

constexpr int GEMM_BK = 16;
constexpr int GEMM_BM = 128;
constexpr int GEMM_NT = 256;

struct Seg {
  const float* w;   // [K_in, width] row-major, ld = ldw   (slice of the TF-layout variable)
  const float* wt;  // [width, K_in] row-major, ld = ldwt  (transposed copy, rebuilt every step)
  float* gw;        // gradient of w, same layout as w
  int ldw, ldwt;
  int col0;         // first output channel of this conv inside the level's channel concat
  int width;        // output channels of this conv
  int dy, dx;       // tap offset relative to the kernel centre
};

struct RowGemmArgs {
  const float* A;
  int lda, M, P;
  const Seg* segs;
  int nseg;
  int mode;  // 0 forward (K = Kfix, B = seg.w), 1 dgrad (K = seg.width, B = seg.wt, shift negated)
  int Kfix;
  float* C;
  int ldc, c_col0, N;
  int accumulate;
  double* stats;  // nullable: [2][stats_ld] column sums / sums of squares of the written tile
  int stats_ld;
  int a_vec, b_vec;
};

struct WgradArgs {
  const float* A;  // layer input [M, Cin]
  int lda, M, P, Cin;
  const float* G;  // gZ [M, ldg]
  int ldg;
  const Seg* segs;
  int nseg;
  int rows_per_split;
  int a_vec, g_vec;
};

// thread (ty,tx) of a 16x16 layout owns rows {ty*4+i, 64+ty*4+i} and TN columns
template <int BN, int TN>
__device__ __forceinline__ int tile_col(int tx, int j) {
  if (TN == 8) return (j < 4) ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4);
  return tx * TN + j;
}
__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }

template <int BN, int TN>
__device__ __forceinline__ void tile_fma(const float (*As)[GEMM_BM + 4], const float (*Bs)[BN + 4], int ty, int tx,
                                         float (&acc)[8][TN]) {
#pragma unroll
  for (int kk = 0; kk < GEMM_BK; kk++) {
    float a[8], b[TN];
    const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    if (TN == 8) {
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][BN / 2 + tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
    } else if (TN == 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      b[0] = b0.x; b[1 % TN] = b0.y; b[2 % TN] = b0.z; b[3 % TN] = b0.w;
    } else if (TN == 2) {
      const float2 b0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
      b[0] = b0.x; b[1 % TN] = b0.y;
    } else {
      b[0] = Bs[kk][tx];
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------------
template <int BN, int TN>
__global__ void __launch_bounds__(GEMM_NT) rowgemm_kernel(const RowGemmArgs p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr int EB = BN / 16;  // B elements per thread per chunk
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  __shared__ float s_sum[BN], s_sq[BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // A loader: 2 threads per row, 8 consecutive k each
  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int am = m0 + a_row;
  int ah = 0, aw = 0;
  if (p.P > 1) {
    const int pix = am % (p.P * p.P);
    ah = pix / p.P;
    aw = pix - ah * p.P;
  }
  // B loader: 16 threads per k row
  const int b_k = tid >> 4, b_n = (tid & 15) * EB;

  int nchunks = 0;
  for (int s = 0; s < p.nseg; s++) nchunks += ((p.mode ? p.segs[s].width : p.Kfix) + BK - 1) / BK;

  float ra[8], rb[EB];
  int seg = 0, k0 = 0;

  auto load = [&]() {
    const Seg S = p.segs[seg];
    const int dy = p.mode ? -S.dy : S.dy, dx = p.mode ? -S.dx : S.dx;
    const int K = p.mode ? S.width : p.Kfix;
    const int a_col0 = p.mode ? S.col0 : 0;
    const float* B = p.mode ? S.wt : S.w;
    const int ldb = p.mode ? S.ldwt : S.ldw;
    bool ok = am < p.M;
    if (p.P > 1) ok = ok && (unsigned)(ah + dy) < (unsigned)p.P && (unsigned)(aw + dx) < (unsigned)p.P;
    const int kk = k0 + a_k;
    if (ok) {
      const float* ap = p.A + (size_t)(am + dy * p.P + dx) * p.lda + a_col0 + kk;
      if (p.a_vec && kk + 7 < K) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(ap));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
        ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
        ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) ra[i] = (kk + i < K) ? __ldg(ap + i) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) ra[i] = 0.f;
    }
    const int bk = k0 + b_k;
    const int bn = n0 + b_n;
    if (bk < K) {
      const float* bp = B + (size_t)bk * ldb + bn;
      if (EB >= 4 && p.b_vec && bn + EB - 1 < p.N) {
#pragma unroll
        for (int i = 0; i < EB; i += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(bp + i));
          rb[i] = v.x; rb[(i + 1) % EB] = v.y; rb[(i + 2) % EB] = v.z; rb[(i + 3) % EB] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < EB; i++) rb[i] = (bn + i < p.N) ? __ldg(bp + i) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < EB; i++) rb[i] = 0.f;
    }
    k0 += BK;
    if (k0 >= K) { k0 = 0; seg++; }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) As[buf][a_k + i][a_row] = ra[i];
#pragma unroll
    for (int i = 0; i < EB; i++) Bs[buf][b_k][b_n + i] = rb[i];
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  load();
  stash(0);
  __syncthreads();
  for (int c = 0; c < nchunks; c++) {
    const int buf = c & 1;
    if (c + 1 < nchunks) load();
    tile_fma<BN, TN>(As[buf], Bs[buf], ty, tx, acc);
    if (c + 1 < nchunks) stash(buf ^ 1);
    __syncthreads();
  }

  // epilogue
  if (p.stats) {
    if (tid < BN) { s_sum[tid] = 0.f; s_sq[tid] = 0.f; }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < TN; j++) {
    const int cn = tile_col<BN, TN>(tx, j);
    const int n = n0 + cn;
    float cs = 0.f, cq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int m = m0 + tile_row(ty, i);
      if (m < p.M && n < p.N) {
        float* cp = p.C + (size_t)m * p.ldc + p.c_col0 + n;
        float v = acc[i][j];
        if (p.accumulate) v += *cp;
        *cp = v;
        cs += v;
        cq += v * v;
      }
    }
    if (p.stats && n < p.N) {
      atomicAdd(&s_sum[cn], cs);
      atomicAdd(&s_sq[cn], cq);
    }
  }
  if (p.stats) {
    __syncthreads();
    if (tid < BN && n0 + tid < p.N) {
      atomicAdd(&p.stats[p.c_col0 + n0 + tid], (double)s_sum[tid]);
      atomicAdd(&p.stats[p.stats_ld + p.c_col0 + n0 + tid], (double)s_sq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// grid: x = c-tiles * n-tiles, y = segment, z = row split.  gw must be zeroed beforehand.
template <int BN, int TN>
__global__ void __launch_bounds__(GEMM_NT) wgrad_kernel(const WgradArgs p) {
  constexpr int BM = GEMM_BM, BK = GEMM_BK;
  constexpr int EB = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const Seg S = p.segs[blockIdx.y];
  const int ntn = (S.width + BN - 1) / BN;
  const int c0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
  const int row_begin = blockIdx.z * p.rows_per_split;
  const int row_end = min(p.M, row_begin + p.rows_per_split);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int l_k = tid >> 4;                 // row inside the chunk for both loaders
  const int a_c = (tid & 15) * 8;           // 8 consecutive input channels
  const int b_n = (tid & 15) * EB;
  const int PP = p.P * p.P;

  float ra[8], rb[EB];
  int mrow = row_begin;
  auto load = [&]() {
    const int m = mrow + l_k;
    bool ok = m < row_end;
    int src = m;
    if (p.P > 1 && ok) {
      const int pix = m % PP;
      const int h = pix / p.P, w = pix - h * p.P;
      ok = (unsigned)(h + S.dy) < (unsigned)p.P && (unsigned)(w + S.dx) < (unsigned)p.P;
      src = m + S.dy * p.P + S.dx;
    }
    const int c = c0 + a_c;
    if (ok && c < p.Cin) {
      const float* ap = p.A + (size_t)src * p.lda + c;
      if (p.a_vec && c + 7 < p.Cin) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(ap));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
        ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
        ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) ra[i] = (c + i < p.Cin) ? __ldg(ap + i) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) ra[i] = 0.f;
    }
    const int n = n0 + b_n;
    if (m < row_end && n < S.width) {
      const float* gp = p.G + (size_t)m * p.ldg + S.col0 + n;
      if (EB >= 4 && p.g_vec && n + EB - 1 < S.width) {
#pragma unroll
        for (int i = 0; i < EB; i += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(gp + i));
          rb[i] = v.x; rb[(i + 1) % EB] = v.y; rb[(i + 2) % EB] = v.z; rb[(i + 3) % EB] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < EB; i++) rb[i] = (n + i < S.width) ? __ldg(gp + i) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < EB; i++) rb[i] = 0.f;
    }
    mrow += BK;
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) As[buf][l_k][a_c + i] = ra[i];
#pragma unroll
    for (int i = 0; i < EB; i++) Bs[buf][l_k][b_n + i] = rb[i];
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const int nchunks = (row_end - row_begin + BK - 1) / BK;
  if (nchunks <= 0) return;
  load();
  stash(0);
  __syncthreads();
  for (int c = 0; c < nchunks; c++) {
    const int buf = c & 1;
    if (c + 1 < nchunks) load();
    tile_fma<BN, TN>(As[buf], Bs[buf], ty, tx, acc);
    if (c + 1 < nchunks) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int c = c0 + tile_row(ty, i);
    if (c >= p.Cin) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int n = n0 + tile_col<BN, TN>(tx, j);
      if (n < S.width) atomicAdd(S.gw + (size_t)c * S.ldw + n, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// batched 2-D transpose: dst[c][r] = src[r][c]; tilemap[b] = {job, tile_r, tile_c}
struct TransposeJob {
  const float* src;
  float* dst;
  int rows, cols;
};
__global__ void transpose_kernel(const TransposeJob* jobs, const int4* tilemap) {
  __shared__ float t[32][33];
  const int4 tm = tilemap[blockIdx.x];
  const TransposeJob J = jobs[tm.x];
  const int r0 = tm.y * 32, c0 = tm.z * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    t[ty + i][tx] = (r < J.rows && c < J.cols) ? J.src[(size_t)r * J.cols + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < J.rows && c < J.cols) J.dst[(size_t)c * J.rows + r] = t[tx][ty + i];
  }
}

// host-side dispatch on the output width
inline int launch_rowgemm(const RowGemmArgs& a, cudaStream_t s, const char* role, double flops) {
  const dim3 block(GEMM_NT);
  const int gm = (int)cdiv(a.M, GEMM_BM);
  char tag[64];
  const int bn = a.N > 64 ? 128 : (a.N > 32 ? 64 : (a.N > 16 ? 32 : 16));
  snprintf(tag, sizeof(tag), "rowgemm_kernel<%d>/%s", bn, role);
  g_prof.begin(s, tag, flops, 0.0);
  if (a.N > 64) {
    rowgemm_kernel<128, 8><<<dim3(gm, (unsigned)cdiv(a.N, 128)), block, 0, s>>>(a);
  } else if (a.N > 32) {
    rowgemm_kernel<64, 4><<<dim3(gm, 1), block, 0, s>>>(a);
  } else if (a.N > 16) {
    rowgemm_kernel<32, 2><<<dim3(gm, 1), block, 0, s>>>(a);
  } else {
    rowgemm_kernel<16, 1><<<dim3(gm, 1), block, 0, s>>>(a);
  }
  g_prof.end(s);
  HYP_LAUNCHED();
  return HYP_OK;
}

// all segments of one launch share `width`
inline int launch_wgrad(const WgradArgs& a, int width, int ksplit, cudaStream_t s, double flops) {
  const dim3 block(GEMM_NT);
  const int ctiles = (int)cdiv(a.Cin, GEMM_BM);
  char tag[64];
  const int bn = width > 64 ? 128 : (width > 32 ? 64 : (width > 16 ? 32 : 16));
  snprintf(tag, sizeof(tag), "wgrad_kernel<%d>/wgrad", bn);
  g_prof.begin(s, tag, flops, 0.0);
  if (width > 64) {
    wgrad_kernel<128, 8><<<dim3(ctiles * (unsigned)cdiv(width, 128), a.nseg, ksplit), block, 0, s>>>(a);
  } else if (width > 32) {
    wgrad_kernel<64, 4><<<dim3(ctiles, a.nseg, ksplit), block, 0, s>>>(a);
  } else if (width > 16) {
    wgrad_kernel<32, 2><<<dim3(ctiles, a.nseg, ksplit), block, 0, s>>>(a);
  } else {
    wgrad_kernel<16, 1><<<dim3(ctiles, a.nseg, ksplit), block, 0, s>>>(a);
  }
  g_prof.end(s);
  HYP_LAUNCHED();
  return HYP_OK;
}

}  // namespace hyp
