// HBM-bound kernels around the GEMMs: BatchNorm (training statistics, apply, backward),
// LeakyReLU / sigmoid / dropout / channel-resampled residuals, losses, Adam, argmax +
// confusion, scene min-max and the patch gather.
#pragma once
#include "hyp_common.cuh"

namespace hyp {

// ------------------------------------------------------------------------------------------
// BatchNorm statistics -> (mean, rstd); slim batch_norm: biased variance to normalise,
// Bessel-corrected variance into the moving average.
// stats: double [2][C] = (sum z, sum z^2) over `rows` rows.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double rows, float eps, float decay,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int is_training,
                                   int update_moving) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (is_training) {
    const double mean = stats[c] / rows;
    double var = stats[C + c] / rows - mean * mean;
    if (var < 0) var = 0;
    const float meanf = (float)mean, varf = (float)var;
    mean_out[c] = meanf;
    rstd_out[c] = rsqrtf(varf + eps);
    // rsqrtf is 2 ulp; refine once so the result matches 1/sqrt in fp32 to <= 1 ulp
    {
      const float x = varf + eps;
      float r = rstd_out[c];
      r = r * (1.5f - 0.5f * x * r * r);
      rstd_out[c] = r;
    }
    if (update_moving) {
      const double unbiased = var * (rows / fmax(rows - 1.0, 1.0));
      moving_mean[c] = moving_mean[c] * decay + meanf * (1.f - decay);
      moving_var[c] = moving_var[c] * decay + (float)unbiased * (1.f - decay);
    }
  } else {
    mean_out[c] = moving_mean[c];
    const float x = moving_var[c] + eps;
    float r = rsqrtf(x);
    r = r * (1.5f - 0.5f * x * r * r);
    rstd_out[c] = r;
  }
}

__device__ __forceinline__ float act_fwd(float y, int act, float alpha) {
  if (act == ACT_LRELU) return fmaxf(y, alpha * y);
  if (act == ACT_SIGMOID) return 1.f / (1.f + __expf(-y));
  return y;
}

struct ApplyArgs {
  const float* z;       // [rows, C] pre-BN
  const float* mean;    // [C]
  const float* rstd;    // [C]
  const float* beta;    // [C]
  float* out;           // [rows, C]
  int64_t rows;
  int C;
  int act;
  float alpha;
  float keep;           // dropout keep_prob; 1 -> no dropout
  uint64_t seed;
  uint32_t stream_id;
  const float* res0;    // residual sources [rows, C0] (nullable)
  const int* idx0;      // nullable -> identity
  int C0;
  const float* res1;
  const int* idx1;
  int C1;
};

// out = dropout(act(bn(z))) + res0[:, idx0] + res1[:, idx1]
__global__ void bn_apply_fwd_kernel(const ApplyArgs p) {
  const int64_t total = p.rows * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / p.C;
    const int c = (int)(i - m * p.C);
    float v = (p.z[i] - p.mean[c]) * p.rstd[c] + p.beta[c];
    v = act_fwd(v, p.act, p.alpha);
    if (p.keep < 1.f) v = (philox_uniform(p.seed, p.stream_id, (uint64_t)i) < p.keep) ? v / p.keep : 0.f;
    if (p.res0) v += p.res0[m * p.C0 + (p.idx0 ? p.idx0[c] : c)];
    if (p.res1) v += p.res1[m * p.C1 + (p.idx1 ? p.idx1[c] : c)];
    p.out[i] = v;
  }
}

struct BnBwdArgs {
  const float* gout;  // [rows, C] gradient w.r.t. the layer's output tensor
  const float* z;     // [rows, C] pre-BN
  const float* mean;
  const float* rstd;
  const float* beta;
  int64_t rows;
  int C;
  int act;
  float alpha;
  float keep;
  uint64_t seed;
  uint32_t stream_id;
  double* sums;        // [2][C]: sum g_y, sum g_y*zhat           (reduce)
  const float* s1;     // [C] mean g_y                            (apply)
  const float* s2;     // [C] mean g_y*zhat                       (apply)
  float* gz;           // [rows, C]                               (apply)
};

__device__ __forceinline__ float bn_gy(const BnBwdArgs& p, int64_t i, int c, float& zhat) {
  zhat = (p.z[i] - p.mean[c]) * p.rstd[c];
  const float y = zhat + p.beta[c];
  float g = p.gout[i];
  if (p.keep < 1.f) g = (philox_uniform(p.seed, p.stream_id, (uint64_t)i) < p.keep) ? g / p.keep : 0.f;
  if (p.act == ACT_LRELU) {
    g = (y > 0.f) ? g : g * p.alpha;
  } else if (p.act == ACT_SIGMOID) {
    const float s = 1.f / (1.f + __expf(-y));
    g = g * s * (1.f - s);
  }
  return g;
}

// block = 32 columns x 8 row lanes; grid = (ceil(C/32), row chunks)
__global__ void bn_bwd_reduce_kernel(const BnBwdArgs p, int rows_per_block) {
  __shared__ float sh1[8][33], sh2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min((int64_t)p.rows, r0 + rows_per_block);
  float a1 = 0.f, a2 = 0.f;
  if (c < p.C) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float zhat;
      const float g = bn_gy(p, r * p.C + c, c, zhat);
      a1 += g;
      a2 += g * zhat;
    }
  }
  sh1[ty][tx] = a1;
  sh2[ty][tx] = a2;
  __syncthreads();
  if (ty == 0 && c < p.C) {
#pragma unroll
    for (int i = 1; i < 8; i++) { a1 += sh1[i][tx]; a2 += sh2[i][tx]; }
    atomicAdd(&p.sums[c], (double)a1);
    atomicAdd(&p.sums[p.C + c], (double)a2);
  }
}

// sums (double) -> s1, s2 = means (float); gbeta = sum g_y
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, int C, double rows, float* __restrict__ s1,
                                       float* __restrict__ s2, float* __restrict__ gbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  s1[c] = (float)(sums[c] / rows);
  s2[c] = (float)(sums[C + c] / rows);
  gbeta[c] = (float)sums[c];
}

// gz = rstd * (g_y - mean(g_y) - zhat * mean(g_y * zhat))
__global__ void bn_bwd_apply_kernel(const BnBwdArgs p) {
  const int64_t total = p.rows * p.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    float zhat;
    const float g = bn_gy(p, i, c, zhat);
    p.gz[i] = p.rstd[c] * (g - p.s1[c] - zhat * p.s2[c]);
  }
}

// residual backward: gsrc[m, c'] (+)= sum_{j in [lo[c'], hi[c'])} gout[m, j]   (lo == NULL: identity)
__global__ void resid_bwd_kernel(const float* __restrict__ gout, int Cout, float* __restrict__ gsrc, int Csrc,
                                 const int* __restrict__ lo, const int* __restrict__ hi, int64_t rows,
                                 int accumulate) {
  const int64_t total = rows * Csrc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / Csrc;
    const int c = (int)(i - m * Csrc);
    float v = 0.f;
    if (lo) {
      for (int j = lo[c]; j < hi[c]; j++) v += gout[m * Cout + j];
    } else {
      v = gout[m * Cout + c];
    }
    gsrc[i] = accumulate ? gsrc[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------
// softmax cross-entropy per row (one warp per row) + gradient (softmax - onehot) * gscale
__global__ void ce_loss_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, int64_t B,
                               int classes, float* __restrict__ ce_out, float* __restrict__ glogits, float gscale) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* lp = logits + row * classes;
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
      glogits[row * classes + c] = (sm - (c == lab ? 1.f : 0.f)) * gscale;
    }
  }
}

// sum (recon - x)^2 -> acc (double); grecon = 2 (recon - x) * gscale
__global__ void mse_kernel(const float* __restrict__ recon, const float* __restrict__ x, int64_t n,
                           double* __restrict__ acc, float* __restrict__ grecon, float gscale) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = recon[i] - x[i];
    s += d * d;
    if (grecon) grecon[i] = 2.f * d * gscale;
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

// loss_out = {mean ce + mse, mean ce, mse}; per_sample[b] = ce[b] + mse (nullable)
__global__ void loss_finalize_kernel(const float* __restrict__ ce, int64_t B, const double* __restrict__ mse_acc,
                                     double mse_count, float* __restrict__ loss_out,
                                     float* __restrict__ per_sample) {
  __shared__ double sh[256];
  const double mse = (mse_acc && mse_count > 0) ? *mse_acc / mse_count : 0.0;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    s += (double)ce[i];
    if (per_sample) per_sample[i] = ce[i] + (float)mse;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss_out) {
    const double mce = sh[0] / (double)B;
    loss_out[0] = (float)(mce + mse);
    loss_out[1] = (float)mce;
    loss_out[2] = (float)mse;
  }
}

// ------------------------------------------------------------------------------------------
// TF1 ApplyAdam on a flat buffer
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_t, float one_minus_b1, float one_minus_b2,
                            float eps, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * one_minus_b1;
    const float vi = v[i] + (gi * gi - v[i]) * one_minus_b2;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// tf MomentumOptimizer (use_nesterov = False): accum = momentum * accum + g;  p -= lr * accum
__global__ void momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ accum, int64_t n,
                                float lr, float momentum, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = momentum * accum[i] + g[i] * gscale;
    accum[i] = a;
    p[i] = p[i] - lr * a;
  }
}

// ------------------------------------------------------------------------------------------
// Training-time augmentation of a batch of patches (common/common_nn_ops.py:397-440), one random draw per
// SAMPLE as in the reference's tf.data map: rot90 by k in {0,1,2} (counter-clockwise, tf.image.rot90), random
// left-right / up-down flips, and a per-channel spectral offset delta[c] ~ U(-s, 0) added to every pixel.
// The geometric part is an index permutation, the spectral part one fp32 add.  choices [B][4] (uint8: k,
// flip_lr, flip_ud, 0) and deltas [B][C] are written so that a caller (the parity tests) can replay the draw.
struct AugmentArgs {
  const float* in;
  float* out;
  int64_t B;
  int P, C;
  int do_rot, do_flip;
  float spectral;  // s; 0 = off
  uint64_t seed;
  uint8_t* choices;  // nullable
  float* deltas;     // nullable
};
__global__ void augment_kernel(const AugmentArgs a) {
  const int64_t b = blockIdx.x;
  const int P = a.P, C = a.C;
  const int k = a.do_rot ? min(2, (int)(philox_uniform(a.seed, 0x41u, (uint64_t)b * 4 + 0) * 3.f)) : 0;
  const int flr = a.do_flip ? (philox_uniform(a.seed, 0x41u, (uint64_t)b * 4 + 1) < 0.5f) : 0;
  const int fud = a.do_flip ? (philox_uniform(a.seed, 0x41u, (uint64_t)b * 4 + 2) < 0.5f) : 0;
  if (threadIdx.x == 0 && a.choices) {
    a.choices[b * 4 + 0] = (uint8_t)k; a.choices[b * 4 + 1] = (uint8_t)flr; a.choices[b * 4 + 2] = (uint8_t)fud;
    a.choices[b * 4 + 3] = 0;
  }
  const float* src = a.in + b * P * P * C;
  float* dst = a.out + b * P * P * C;
  for (int i = threadIdx.x; i < P * P * C; i += blockDim.x) {
    const int pos = i / C, c = i - pos * C;
    int r = pos / P, q = pos - r * P;             // output pixel (row, col)
    if (fud) r = P - 1 - r;                       // undo tf.image.flip_up_down
    if (flr) q = P - 1 - q;                       // undo tf.image.flip_left_right
    int sr = r, sq = q;                           // undo rot90: R1[i,j] = in[j, P-1-i], R2[i,j] = in[P-1-i, P-1-j]
    if (k == 1) { sr = q; sq = P - 1 - r; }
    else if (k == 2) { sr = P - 1 - r; sq = P - 1 - q; }
    float v = src[(sr * P + sq) * C + c];
    if (a.spectral > 0.f) {
      const float d = -a.spectral * philox_uniform(a.seed, 0x42u, (uint64_t)b * C + c);
      if (a.deltas && pos == 0) a.deltas[b * C + c] = d;
      v = v + d;
    }
    dst[i] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Shadow GAN generator, forward (gan/shadow_data_models.py:43-90): per spectrum x[C] seven 1-filter conv1d layers
// (SAME, kernel sizes C, C/2, C/4, C/8, C/4, C/2, C; bias; leaky_relu 0.1) with the dense residual pattern
// net_i = conv(net_{i-1}) + net_{i-1} + net_{i-2}; the last layer is tanh without residual; encoder-only stops
// after net4.  One warp per spectrum: the three live activations sit in shared memory (C floats each), the <= 239
// weights of the model in shared memory per block.  `rows` spectra of `C` bands at stride ld_in / ld_out; extra
// channels (the LiDAR band of a patch pixel, create_gan_struct gan/gan_utilities.py:31-35) are copied through.
// clip (create_inference_for_matrix_input, gan/wrappers/gan_common.py:282-304): keep the input unless the
// generated mean is lower (is_shadow) / higher (deshadow) than the input mean.
struct GanGenArgs {
  const float* in;
  float* out;
  int64_t rows;
  int C, ld_in, ld_out, copy_extra;  // copy_extra: channels after the C bands copied unchanged
  const float* weights;              // net1 w[K1] b, net2 w[K2] b, ... in layer order
  int nlayers;                       // 7 full generator, 4 encoder only
  int clip, is_shadow;
  float* nets;                       // nullable (training): [rows][nlayers + 1][C] = net0 (input) .. net_nlayers
};
constexpr int GAN_MAX_C = 512;
__global__ void __launch_bounds__(128) gan_generator_fwd_kernel(const GanGenArgs a) {
  extern __shared__ float sm[];
  const int C = a.C;
  int K[7];
  K[0] = C; K[1] = C / 2; K[2] = C / 4; K[3] = C / 8; K[4] = C / 4; K[5] = C / 2; K[6] = C;
  int nw = 0;
  for (int l = 0; l < a.nlayers; l++) nw += K[l] + 1;
  float* w = sm;                                 // [nw]
  float* act = sm + ((nw + 3) & ~3);             // [warps][3][C]
  for (int i = threadIdx.x; i < nw; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* buf = act + warp * 3 * C;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < a.rows; r += (int64_t)gridDim.x * nwarps) {
    const float* xin = a.in + r * a.ld_in;
    float* xout = a.out + r * a.ld_out;
    float* p2 = buf;          // net_{i-2}
    float* p1 = buf + C;      // net_{i-1}
    float* cur = buf + 2 * C;
    float in_sum = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = xin[c]; p1[c] = v; p2[c] = 0.f; in_sum += v; }
    if (a.nets)
      for (int c = lane; c < C; c += 32) a.nets[(r * (a.nlayers + 1)) * C + c] = p1[c];
    __syncwarp();
    const float* wl = w;
    for (int l = 0; l < a.nlayers; l++) {
      const int k = K[l], left = (k - 1) / 2;  // SAME: total pad k-1, the extra one on the right (App. A.12)
      const float bias = wl[k];
      const bool last = l == 6;
      for (int c = lane; c < C; c += 32) {
        float s = bias;
        const int t0 = max(0, left - c), t1 = min(k, C + left - c);
        for (int t = t0; t < t1; t++) s += wl[t] * p1[c + t - left];
        if (last) s = tanhf(s);
        else {
          s = fmaxf(s, 0.1f * s);
          s += p1[c];
          if (l > 0) s += p2[c];              // net1 = conv + net0 only
        }
        cur[c] = s;
        if (a.nets) a.nets[(r * (a.nlayers + 1) + l + 1) * C + c] = s;
      }
      __syncwarp();
      float* t = p2; p2 = p1; p1 = cur; cur = t;
      wl += k + 1;
    }
    // p1 = result
    bool keep_generated = true;
    if (a.clip) {
      float gs = 0.f;
      for (int c = lane; c < C; c += 32) gs += p1[c];
      gs = warp_sum(gs);
      in_sum = warp_sum(in_sum);
      keep_generated = a.is_shadow ? (gs < in_sum) : (gs > in_sum);  // means over the same C bands
    }
    for (int c = lane; c < C; c += 32) xout[c] = keep_generated ? p1[c] : xin[c];
    for (int c = lane; c < a.copy_extra; c += 32) xout[C + c] = xin[C + c];
    __syncwarp();
  }
}

// Generator backward (full 7-layer generator): from the saved layer outputs nets [rows][8][C] and dL/dnet7 to dL/dnet0
// and the weight / bias gradients.  One warp per spectrum; activations and the running gradients of net0..net7 in
// shared memory; per-block weight-gradient accumulation in shared memory, one global atomicAdd per weight per block.
//   pre_l = conv(net_{l-1}; w_l) + b_l;  act_l = lrelu_0.1(pre_l) (tanh for l = 7);
//   net_l = act_l + net_{l-1} + net_{l-2}   (l = 1: + net_0 only; l = 7: no residual)
struct GanGenBwdArgs {
  const float* nets;
  const float* gout;      // dL/dnet7 [rows][C]; nullable: the decoder (net5..net7) is then skipped
  const float* gout_enc;  // dL/dnet4 (the encoder output, create_only_encoder=True) [rows][C], nullable, added
  int64_t rows;
  int C;
  const float* weights;
  float* gin;       // nullable
  float* gweights;  // += (same order as weights)
};
__global__ void __launch_bounds__(128) gan_generator_bwd_kernel(const GanGenBwdArgs a) {
  extern __shared__ float sm[];
  const int C = a.C;
  int K[7];
  K[0] = C; K[1] = C / 2; K[2] = C / 4; K[3] = C / 8; K[4] = C / 4; K[5] = C / 2; K[6] = C;
  int woff[8];
  woff[0] = 0;
  for (int l = 0; l < 7; l++) woff[l + 1] = woff[l] + K[l] + 1;
  const int nw = woff[7], nwp = (nw + 3) & ~3;
  float* w = sm;                 // [nwp]
  float* gw = sm + nwp;          // [nwp] block accumulator
  float* per_warp = gw + nwp;    // [warps][(8 + 8 + 1) * C]
  for (int i = threadIdx.x; i < nw; i += blockDim.x) { w[i] = a.weights[i]; gw[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* nets = per_warp + warp * 17 * C;  // [8][C]
  float* G = nets + 8 * C;                 // [8][C]
  float* dpre = G + 8 * C;                 // [C]
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < a.rows; r += (int64_t)gridDim.x * nwarps) {
    for (int i = lane; i < 8 * C; i += 32) { nets[i] = a.nets[r * 8 * C + i]; G[i] = 0.f; }
    if (a.gout)
      for (int c = lane; c < C; c += 32) G[7 * C + c] = a.gout[r * C + c];
    if (a.gout_enc)
      for (int c = lane; c < C; c += 32) G[4 * C + c] = a.gout_enc[r * C + c];
    __syncwarp();
    for (int l = a.gout ? 7 : 4; l >= 1; l--) {
      const int k = K[l - 1], left = (k - 1) / 2;
      const float* wl = w + woff[l - 1];
      const float* in = nets + (l - 1) * C;
      float* Gl = G + l * C;
      float* Gin = G + (l - 1) * C;
      float bsum = 0.f;
      for (int c = lane; c < C; c += 32) {
        float d;
        if (l == 7) {
          const float y = nets[7 * C + c];
          d = Gl[c] * (1.f - y * y);
        } else {
          const float act = nets[l * C + c] - in[c] - (l > 1 ? nets[(l - 2) * C + c] : 0.f);
          d = Gl[c] * (act > 0.f ? 1.f : 0.1f);
        }
        dpre[c] = d;
        bsum += d;
      }
      __syncwarp();
      if (l < 7) {  // residual paths
        for (int c = lane; c < C; c += 32) {
          Gin[c] += Gl[c];
          if (l > 1) G[(l - 2) * C + c] += Gl[c];
        }
      }
      bsum = warp_sum(bsum);
      if (lane == 0) atomicAdd(&gw[woff[l - 1] + k], bsum);
      // dW_l[t] = sum_c dpre[c] * in[c + t - left]
      for (int t = lane; t < k; t += 32) {
        float s = 0.f;
        const int c0 = max(0, left - t), c1 = min(C, C + left - t);
        for (int c = c0; c < c1; c++) s += dpre[c] * in[c + t - left];
        atomicAdd(&gw[woff[l - 1] + t], s);
      }
      // dL/dnet_{l-1}[c'] += sum_t w[t] * dpre[c' - t + left]
      for (int cp = lane; cp < C; cp += 32) {
        float s = 0.f;
        const int t0 = max(0, cp + left - (C - 1)), t1 = min(k, cp + left + 1);
        for (int t = t0; t < t1; t++) s += wl[t] * dpre[cp - t + left];
        Gin[cp] += s;
      }
      __syncwarp();
    }
    if (a.gin)
      for (int c = lane; c < C; c += 32) a.gin[r * C + c] = G[c];
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw; i += blockDim.x) atomicAdd(&a.gweights[i], gw[i]);
}

// ------------------------------------------------------------------------------------------
// Shadow GAN discriminator (gan/shadow_data_models.py:93-123): flatten -> FC C->C -> FC C->C (leaky_relu 0.1) -> FC
// C->C/2 linear.  Weights [W1 (C x C, input-major), b1, W2, b2, W3 (C x C/2), b3] in shared memory (C <= 64),
// one warp per spectrum.  Forward keeps h1, h2 for the backward pass.
struct GanDiscArgs {
  const float* x;     // [rows][C]
  int64_t rows;
  int C;
  const float* weights;
  float* h;           // [rows][2][C]   (forward: out; backward: in)
  float* out;         // forward: [rows][C/2]
  const float* gout;  // backward: [rows][C/2]
  float* gin;         // backward, nullable: [rows][C]
  float* gweights;    // backward, nullable: += weight gradients
};
__device__ __forceinline__ int gan_disc_nweights(int C) { return C * C + C + C * C + C + C * (C / 2) + C / 2; }
__global__ void __launch_bounds__(256) gan_discriminator_fwd_kernel(const GanDiscArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, H = C / 2, nw = gan_disc_nweights(C);
  float* w = sm;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) w[i] = a.weights[i];
  __syncthreads();
  const float *W1 = w, *b1 = W1 + C * C, *W2 = b1 + C, *b2 = W2 + C * C, *W3 = b2 + C, *b3 = W3 + C * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* v = sm + ((nw + 3) & ~3) + warp * 3 * C;  // x, h1, h2
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < a.rows; r += (int64_t)gridDim.x * nwarps) {
    for (int c = lane; c < C; c += 32) v[c] = a.x[r * C + c];
    __syncwarp();
    for (int j = lane; j < C; j += 32) {
      float s = b1[j];
      for (int i = 0; i < C; i++) s += v[i] * W1[i * C + j];
      s = fmaxf(s, 0.1f * s);
      v[C + j] = s;
      a.h[(r * 2 + 0) * C + j] = s;
    }
    __syncwarp();
    for (int j = lane; j < C; j += 32) {
      float s = b2[j];
      for (int i = 0; i < C; i++) s += v[C + i] * W2[i * C + j];
      s = fmaxf(s, 0.1f * s);
      v[2 * C + j] = s;
      a.h[(r * 2 + 1) * C + j] = s;
    }
    __syncwarp();
    for (int j = lane; j < H; j += 32) {
      float s = b3[j];
      for (int i = 0; i < C; i++) s += v[2 * C + i] * W3[i * H + j];
      a.out[r * H + j] = s;
    }
    __syncwarp();
  }
}
__global__ void __launch_bounds__(256) gan_discriminator_bwd_kernel(const GanDiscArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, H = C / 2, nw = gan_disc_nweights(C), nwp = (nw + 3) & ~3;
  // weights with padded rows (the backward products walk a ROW per lane: stride C+1 / H+1 keeps lanes on distinct banks)
  const int LW = C + 1, LH = H + 1;
  const int wpad = ((2 * C * LW + C * LH) + 3) & ~3;
  float* W1 = sm;
  float* W2 = W1 + C * LW;
  float* W3 = W2 + C * LW;
  float* gw = sm + wpad;  // block accumulator in the dense layout of `weights` (only when gweights)
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    W1[(i / C) * LW + i % C] = a.weights[i];
    W2[(i / C) * LW + i % C] = a.weights[C * C + C + i];
  }
  for (int i = threadIdx.x; i < C * H; i += blockDim.x) W3[(i / H) * LH + i % H] = a.weights[2 * (C * C + C) + i];
  if (a.gweights)
    for (int i = threadIdx.x; i < nw; i += blockDim.x) gw[i] = 0.f;
  __syncthreads();
  float *gW1 = gw, *gb1 = gW1 + C * C, *gW2 = gb1 + C, *gb2 = gW2 + C * C, *gW3 = gb2 + C, *gb3 = gW3 + C * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* v = sm + wpad + (a.gweights ? nwp : 0) + warp * 6 * C;  // x, h1, h2, d2 (dpre2), d1 (dpre1), dout
  float *xs = v, *h1 = v + C, *h2 = v + 2 * C, *d2 = v + 3 * C, *d1 = v + 4 * C, *dout_s = v + 5 * C;
  for (int64_t r = (int64_t)blockIdx.x * nwarps + warp; r < a.rows; r += (int64_t)gridDim.x * nwarps) {
    for (int c = lane; c < C; c += 32) {
      xs[c] = a.x[r * C + c];
      h1[c] = a.h[(r * 2 + 0) * C + c];
      h2[c] = a.h[(r * 2 + 1) * C + c];
    }
    for (int j = lane; j < H; j += 32) dout_s[j] = a.gout[r * H + j];
    __syncwarp();
    // layer 3 (linear): dh2[i] = sum_j W3[i][j] dout[j];  dpre2 = dh2 * lrelu'(h2)
    for (int i = lane; i < C; i += 32) {
      float s = 0.f;
      for (int j = 0; j < H; j++) s += W3[i * LH + j] * dout_s[j];
      d2[i] = s * (h2[i] > 0.f ? 1.f : 0.1f);
    }
    if (a.gweights) {
      for (int j = lane; j < H; j += 32) {
        atomicAdd(&gb3[j], dout_s[j]);
        for (int i = 0; i < C; i++) atomicAdd(&gW3[i * H + j], h2[i] * dout_s[j]);
      }
    }
    __syncwarp();
    for (int i = lane; i < C; i += 32) {
      float s = 0.f;
      for (int j = 0; j < C; j++) s += W2[i * LW + j] * d2[j];
      d1[i] = s * (h1[i] > 0.f ? 1.f : 0.1f);
    }
    if (a.gweights) {
      for (int j = lane; j < C; j += 32) {
        atomicAdd(&gb2[j], d2[j]);
        for (int i = 0; i < C; i++) atomicAdd(&gW2[i * C + j], h1[i] * d2[j]);
      }
    }
    __syncwarp();
    if (a.gin) {
      for (int i = lane; i < C; i += 32) {
        float s = 0.f;
        for (int j = 0; j < C; j++) s += W1[i * LW + j] * d1[j];
        a.gin[r * C + i] = s;
      }
    }
    if (a.gweights) {
      for (int j = lane; j < C; j += 32) {
        atomicAdd(&gb1[j], d1[j]);
        for (int i = 0; i < C; i++) atomicAdd(&gW1[i * C + j], xs[i] * d1[j]);
      }
    }
    __syncwarp();
  }
  if (a.gweights) {
    __syncthreads();
    for (int i = threadIdx.x; i < nw; i += blockDim.x) atomicAdd(&a.gweights[i], gw[i]);
  }
}

// GAN loss terms with their gradients.  mode 0 (tfgan least squares): sum 0.5 * scale * (a - target)^2, grad scale * (a -
// target);  mode 1 (absolute_difference): sum scale * |a - b|, grad scale * sign(a - b).  The caller passes scale =
// weight / numel (the losses are means).  grad (nullable) is overwritten or accumulated; loss_acc += the sum.
__global__ void gan_loss_grad_kernel(int mode, const float* __restrict__ av, const float* __restrict__ bv, float target,
                                     float scale, int64_t n, float* __restrict__ grad, int accumulate,
                                     double* __restrict__ loss_acc) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g;
    if (mode == 0) {
      const float d = av[i] - target;
      s += 0.5f * scale * d * d;
      g = scale * d;
    } else if (mode == 1) {
      const float d = av[i] - bv[i];
      s += scale * fabsf(d);
      g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
    } else {  // mode 2 (tfgan wasserstein_*): sum scale * a, grad scale (scale carries the sign)
      s += scale * av[i];
      g = scale;
    }
    if (grad) grad[i] = accumulate ? grad[i] + g : g;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, (double)s);
  }
}
// slim l2_regularizer(scale): loss += scale * sum(w^2) / 2, grad += scale * w
__global__ void gan_l2_reg_kernel(const float* __restrict__ w, float* __restrict__ g, int64_t n, float scale,
                                  double* __restrict__ loss_acc) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    s += 0.5f * scale * v * v;
    if (g) g[i] += scale * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0 && loss_acc) atomicAdd(loss_acc, (double)s);
  }
}

// ------------------------------------------------------------------------------------------
// CUT / DCLGAN patch feature discriminator (gan/shadow_data_models.py:126-149): the C-band encoder embedding is cut
// into slices of ps = C / patch_count bands (the last slice is ragged when ps does not divide C — 64 bands, 6 patches
// give 7 slices, the last 4 wide); every slice has its own 4-layer MLP  in -> ps -> ps/4 -> ps/2 -> E, leaky_relu(0.1)
// after every layer (the arg_scope default also applies to the last), and the [rows, E] output of a slice is divided
// by its norm over the WHOLE batch (tf.math.l2_normalize with axis=None, epsilon 1e-12).
// A CTA owns 128 rows of one slice: the slice's weights and the per-row activations ([dim][129] so that both the
// row-per-thread MLP and the transposed loads / weight-gradient reductions are bank-conflict free) sit in shared memory.
// forward writes the un-normalised z and adds sum z^2 to sumsq[slice]; the division happens in the consumers.
struct GanFeatArgs {
  const float* x;        // [rows][C]
  int64_t rows;
  int C, ps, E;
  const float* weights;  // per slice: W1 [in][ps] b1 | W2 [ps][ps/4] b2 | W3 [ps/4][ps/2] b3 | W4 [ps/2][E] b4
  float* z;              // [rows][slices][E]   forward: out, backward: in
  float* sumsq;          // [slices]            forward: +=, backward: in
  const float* gf;       // backward: dL/d(normalised embedding) [rows][slices][E]
  const float* dot;      // backward: [slices] sum over the batch of gf * z
  float* gin;            // backward, nullable: [rows][C]
  float* gweights;       // backward, nullable: +=
};
constexpr int FD_ROWS = 128, FD_PITCH = 129;
__host__ __device__ inline int featdisc_slice_weights(int ps, int E) {
  return ps * ps + ps + ps * (ps / 4) + ps / 4 + (ps / 4) * (ps / 2) + ps / 2 + (ps / 2) * E + E;
}
// forward of one CTA's rows; returns with a0..a4 in act (a_l at row offset aoff[l]) and all threads synchronised
__device__ __forceinline__ void featdisc_forward_block(const GanFeatArgs& a, const float* w, float* act, const int* dim,
                                                      const int* aoff, int slice, int64_t r0) {
  const int t = threadIdx.x, in_p = dim[0];
  for (int idx = t; idx < FD_ROWS * in_p; idx += FD_ROWS) {
    const int rr = idx / in_p, i = idx - rr * in_p;
    const int64_t r = r0 + rr;
    act[i * FD_PITCH + rr] = r < a.rows ? a.x[r * a.C + slice * a.ps + i] : 0.f;
  }
  __syncthreads();
  const float* wl = w;
  for (int l = 0; l < 4; l++) {
    const int ni = dim[l], no = dim[l + 1];
    const float* ain = act + aoff[l] * FD_PITCH;
    float* aout = act + aoff[l + 1] * FD_PITCH;
    const float* bias = wl + ni * no;
    for (int j = 0; j < no; j++) {
      float s = bias[j];
      for (int i = 0; i < ni; i++) s += wl[i * no + j] * ain[i * FD_PITCH + t];
      aout[j * FD_PITCH + t] = fmaxf(s, 0.1f * s);
    }
    wl += ni * no + no;
  }
  __syncthreads();
}
#define FEATDISC_SETUP()                                                                        \
  const int slice = blockIdx.y, ps = a.ps, E = a.E;                                             \
  const int dim[5] = {min(ps, a.C - slice * ps), ps, ps / 4, ps / 2, E};                        \
  const int aoff[5] = {0, ps, 2 * ps, 2 * ps + ps / 4, 2 * ps + ps / 4 + ps / 2};               \
  const int nrows_act = 2 * ps + ps / 4 + ps / 2 + E;                                           \
  const int n_full = featdisc_slice_weights(ps, E), nw = n_full - (ps - dim[0]) * ps;           \
  float* w = sm;                                                                                \
  float* act = sm + ((n_full + 3) & ~3);                                                        \
  for (int i = threadIdx.x; i < nw; i += FD_ROWS) w[i] = a.weights[(size_t)slice * n_full + i]; \
  const int64_t r0 = (int64_t)blockIdx.x * FD_ROWS;                                             \
  const int t = threadIdx.x;                                                                    \
  const int64_t r = r0 + t;                                                                     \
  const int slices = gridDim.y

__global__ void __launch_bounds__(FD_ROWS) gan_featdisc_fwd_kernel(const GanFeatArgs a) {
  extern __shared__ float sm[];
  FEATDISC_SETUP();
  (void)nrows_act;
  featdisc_forward_block(a, w, act, dim, aoff, slice, r0);
  float s2 = 0.f;
  if (r < a.rows) {
    for (int e = 0; e < E; e++) {
      const float v = act[(aoff[4] + e) * FD_PITCH + t];
      a.z[(r * slices + slice) * E + e] = v;
      s2 += v * v;
    }
  }
  s2 = warp_sum(s2);
  if ((t & 31) == 0) atomicAdd(&a.sumsq[slice], s2);
}

__global__ void __launch_bounds__(FD_ROWS) gan_featdisc_bwd_kernel(const GanFeatArgs a) {
  extern __shared__ float sm[];
  FEATDISC_SETUP();
  float* dact = act + nrows_act * FD_PITCH;  // gradients, same row offsets as act
  featdisc_forward_block(a, w, act, dim, aoff, slice, r0);
  {  // through z / max(|z|_batch, 1e-6): d z = gf / n - z (sum gf z) / n^3  (no second term when the clamp is active)
    const float ss = a.sumsq[slice];
    const bool clamped = ss < 1e-12f;
    const float inv = rsqrtf(fmaxf(ss, 1e-12f));
    const float k = clamped ? 0.f : a.dot[slice] * inv * inv * inv;
    for (int e = 0; e < E; e++) {
      float d = 0.f;
      if (r < a.rows) {
        const size_t o = (r * slices + slice) * E + e;
        d = a.gf[o] * inv - a.z[o] * k;
      }
      dact[(aoff[4] + e) * FD_PITCH + t] = d;
    }
  }
  int woff[5];
  woff[0] = 0;
  for (int l = 0; l < 4; l++) woff[l + 1] = woff[l] + dim[l] * dim[l + 1] + dim[l + 1];
  float* gwg = a.gweights ? a.gweights + (size_t)slice * n_full : nullptr;
  const int lane = t & 31;
  for (int l = 4; l >= 1; l--) {
    const int ni = dim[l - 1], no = dim[l];
    const float* wl = w + woff[l - 1];
    const float* al = act + aoff[l] * FD_PITCH;
    const float* ain = act + aoff[l - 1] * FD_PITCH;
    float* dl = dact + aoff[l] * FD_PITCH;
    float* din = dact + aoff[l - 1] * FD_PITCH;
    for (int j = 0; j < no; j++) dl[j * FD_PITCH + t] *= al[j * FD_PITCH + t] > 0.f ? 1.f : 0.1f;
    if (l > 1 || a.gin) {
      for (int i = 0; i < ni; i++) {
        float s = 0.f;
        for (int j = 0; j < no; j++) s += wl[i * no + j] * dl[j * FD_PITCH + t];
        din[i * FD_PITCH + t] = s;
      }
    }
    if (gwg) {
      __syncthreads();
      for (int idx = t; idx < ni * no + no; idx += FD_ROWS) {
        float s = 0.f;
        if (idx < ni * no) {
          const int i = idx / no, j = idx - i * no;
          for (int tt = 0; tt < FD_ROWS; tt++) {
            const int q = (tt + lane) & (FD_ROWS - 1);
            s += ain[i * FD_PITCH + q] * dl[j * FD_PITCH + q];
          }
        } else {
          const int j = idx - ni * no;
          for (int tt = 0; tt < FD_ROWS; tt++) s += dl[j * FD_PITCH + ((tt + lane) & (FD_ROWS - 1))];
        }
        atomicAdd(&gwg[woff[l - 1] + idx], s);
      }
    }
  }
  if (a.gin) {
    __syncthreads();
    const int in_p = dim[0];
    for (int idx = t; idx < FD_ROWS * in_p; idx += FD_ROWS) {
      const int rr = idx / in_p, i = idx - rr * in_p;
      if (r0 + rr < a.rows) a.gin[(r0 + rr) * a.C + slice * ps + i] = dact[i * FD_PITCH + rr];
    }
  }
}

// PatchNCE (gan/wrappers/cut_wrapper.py:360-420): per sample logits[i][j] = <f_gen[i], f_real[j]> / tau over the
// slices, labels = eye(slices) flattened, ONE softmax over all slices^2 logits, loss_b = -sum_i log_softmax[i][i],
// mean over the batch.  fused_grad = 1 reproduces the gradient of TensorFlow's fused SoftmaxCrossEntropyWithLogits
// kernel, softmax - labels, which is what the reference trains with although its labels sum to `slices` (the exact
// derivative, slices * softmax - labels, is fused_grad = 0).  One warp per sample; f = z * rsqrt(max(sumsq, 1e-12)).
struct GanNceArgs {
  const float* zg;
  const float* zr;   // [rows][S][E]
  const float* ssg;
  const float* ssr;  // [S]
  int64_t rows;
  int S, E;
  float inv_tau, scale;
  int fused_grad;
  float* gfg;
  float* gfr;        // nullable (both): dL/d f_gen, dL/d f_real  [rows][S][E]
  float* dotg;
  float* dotr;       // [S] += sum_b gf * z
  double* loss_acc;  // += scale * sum_b loss_b
};
constexpr int NCE_MAX_S = 16, NCE_MAX_E = 8;
__global__ void __launch_bounds__(128) gan_patchnce_kernel(const GanNceArgs a) {
  __shared__ float fg_s[4][NCE_MAX_S * NCE_MAX_E], fr_s[4][NCE_MAX_S * NCE_MAX_E], gl_s[4][NCE_MAX_S * NCE_MAX_S];
  __shared__ float dot_s[2][NCE_MAX_S];
  __shared__ float loss_s;
  const int S = a.S, E = a.E, SE = S * E, SS = S * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 2 * NCE_MAX_S) dot_s[threadIdx.x / NCE_MAX_S][threadIdx.x % NCE_MAX_S] = 0.f;
  if (threadIdx.x == 0) loss_s = 0.f;
  __syncthreads();
  float* fg = fg_s[warp];
  float* fr = fr_s[warp];
  float* gl = gl_s[warp];
  float loss_w = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * 4 + warp; r < a.rows; r += (int64_t)gridDim.x * 4) {
    for (int i = lane; i < SE; i += 32) {
      const int s = i / E;
      fg[i] = a.zg[r * SE + i] * rsqrtf(fmaxf(a.ssg[s], 1e-12f));
      fr[i] = a.zr[r * SE + i] * rsqrtf(fmaxf(a.ssr[s], 1e-12f));
    }
    __syncwarp();
    float mx = -INFINITY, diag = 0.f;
    for (int idx = lane; idx < SS; idx += 32) {
      const int i = idx / S, j = idx - i * S;
      float l = 0.f;
      for (int e = 0; e < E; e++) l += fg[i * E + e] * fr[j * E + e];
      l *= a.inv_tau;
      gl[idx] = l;
      mx = fmaxf(mx, l);
      if (i == j) diag += l;
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int idx = lane; idx < SS; idx += 32) se += __expf(gl[idx] - mx);
    se = warp_sum(se);
    diag = warp_sum(diag);
    const float logz = mx + __logf(se);
    loss_w += (float)S * logz - diag;
    if (a.gfg) {
      const float ksm = a.fused_grad ? 1.f : (float)S;
      for (int idx = lane; idx < SS; idx += 32) {
        const int i = idx / S, j = idx - i * S;
        gl[idx] = (ksm * __expf(gl[idx] - logz) - (i == j ? 1.f : 0.f)) * a.scale * a.inv_tau;
      }
      __syncwarp();
      for (int idx = lane; idx < SE; idx += 32) {
        const int s = idx / E, e = idx - s * E;
        float gg = 0.f, gr = 0.f;
        for (int q = 0; q < S; q++) {
          gg += gl[s * S + q] * fr[q * E + e];   // d/d f_gen[s][e]
          gr += gl[q * S + s] * fg[q * E + e];   // d/d f_real[s][e]
        }
        a.gfg[r * SE + idx] = gg;
        a.gfr[r * SE + idx] = gr;
        atomicAdd(&dot_s[0][s], gg * a.zg[r * SE + idx]);
        atomicAdd(&dot_s[1][s], gr * a.zr[r * SE + idx]);
      }
    }
    __syncwarp();
  }
  if (lane == 0) atomicAdd(&loss_s, loss_w);
  __syncthreads();
  if (a.gfg && threadIdx.x < 2 * NCE_MAX_S) {
    const int k = threadIdx.x / NCE_MAX_S, s = threadIdx.x % NCE_MAX_S;
    if (s < S) atomicAdd(k ? &a.dotr[s] : &a.dotg[s], dot_s[k][s]);
  }
  if (threadIdx.x == 0 && a.loss_acc) atomicAdd(a.loss_acc, (double)loss_s * (double)a.scale);
}

// tf.argmax (first maximum) + confusion[label, pred] += 1
__global__ void argmax_confusion_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels,
                                        int64_t B, int classes, uint8_t* __restrict__ pred,
                                        int32_t* __restrict__ confusion) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* lp = logits + i * classes;
  float best = lp[0];
  int bi = 0;
  for (int c = 1; c < classes; c++) {
    const float v = lp[c];
    if (v > best) { best = v; bi = c; }
  }
  if (pred) pred[i] = (uint8_t)bi;
  if (confusion && labels) atomicAdd(&confusion[(int)labels[i] * classes + bi], 1);
}

__global__ void scatter_class_map_kernel(const uint8_t* __restrict__ pred, const int32_t* __restrict__ xy,
                                         int64_t N, int H, int W, uint8_t* __restrict__ map) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int x = xy[2 * i], y = xy[2 * i + 1];
  if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) map[(size_t)y * W + x] = pred[i];
}

__global__ void dropout_mask_kernel(uint64_t seed, uint32_t stream_id, int64_t n, float keep,
                                    uint8_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_uniform(seed, stream_id, (uint64_t)i) < keep ? 1 : 0;
}

}  // namespace hyp
