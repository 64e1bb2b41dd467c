// Fused CycleGAN train-step kernels (reference: gan/wrappers/cycle_gan_wrapper.py:189-333 — cyclegan_model_with_identity /
// cyclegan_loss_with_identity over gan/shadow_data_models.py:43-123).  The models are tiny (generator: 239 scalars,
// discriminator: 10 400 at 64 bands) and a train iteration at the reference's batch of 32 is pure launch latency when it
// is chained from ~40 small kernels; here ONE kernel computes every gradient of a train op:
//
//   hyp_gan_cycle_generator_step     : G(x), F(y), F(G x), G(F y), D_Y(G x), D_X(F y), the LSGAN / cycle / "identity"
//                                      losses, backward through the (frozen) discriminators and all four generator
//                                      applications  ->  d loss / d [G | F]
//   hyp_gan_cycle_discriminator_step : the fakes G(x), F(y), tfgan's tensor pool (device resident, the host only draws
//                                      the slot), D_Y / D_X on real and fake, LSGAN discriminator loss, L2 regulariser
//                                      ->  d loss / d [D_Y | D_X]
//
// One warp owns one (x, y) pair from the first forward to the last weight gradient: every activation it needs later
// stays in its slice of shared memory, the weights of both generators and both discriminators sit in shared memory per
// block, weight gradients are summed per block in shared memory and leave with one atomic per weight and block.
// The arithmetic (operation order inside each layer) is that of the per-op kernels in hyp_kernels.cuh, which stay the
// reference implementation the fused kernels are tested against.
#include <algorithm>

#include "hyp_common.cuh"

namespace hyp {

constexpr int GS_WARPS = 4;
constexpr int GS_MAX_C = 64;

__host__ __device__ inline int gs_gen_nweights(int C) { return (C + C / 2 + C / 4 + C / 8 + C / 4 + C / 2 + C) + 7; }
__host__ __device__ inline int gs_disc_nweights(int C) { return C * C + C + C * C + C + C * (C / 2) + C / 2; }

// ---- generator, one warp per spectrum -------------------------------------------------------------------------
// nets [8][C]: net0 = input (already in place) .. net7 = output.  w: layer weights w[K] b, in layer order.
__device__ __forceinline__ void gs_gen_forward(const float* __restrict__ w, float* nets, int C, int lane) {
  int K[7] = {C, C / 2, C / 4, C / 8, C / 4, C / 2, C};
  const float* wl = w;
#pragma unroll 1
  for (int l = 0; l < 7; l++) {
    const int k = K[l], left = (k - 1) / 2;  // SAME: total pad k-1, the extra one on the right
    const float bias = wl[k];
    const float* p1 = nets + l * C;
    const float* p2 = nets + (l > 0 ? l - 1 : 0) * C;
    float* cur = nets + (l + 1) * C;
    for (int c = lane; c < C; c += 32) {
      float s = bias;
      const int t0 = max(0, left - c), t1 = min(k, C + left - c);
      for (int t = t0; t < t1; t++) s += wl[t] * p1[c + t - left];
      if (l == 6) s = tanhf(s);
      else {
        s = fmaxf(s, 0.1f * s);
        s += p1[c];
        if (l > 0) s += p2[c];  // net1 = conv + net0 only
      }
      cur[c] = s;
    }
    __syncwarp();
    wl += k + 1;
  }
}
// G [8][C]: running gradients of net0..net7, G[7] preset to dL/dnet7, the rest is zeroed here; result dL/dnet0 in G[0].
// gw: block accumulator of the weight gradients (shared memory, layout of w).
__device__ __forceinline__ void gs_gen_backward(const float* __restrict__ w, float* gw, const float* nets, float* G, float* dpre,
                                                int C, int lane) {
  int K[7] = {C, C / 2, C / 4, C / 8, C / 4, C / 2, C};
  int woff[8];
  woff[0] = 0;
#pragma unroll
  for (int l = 0; l < 7; l++) woff[l + 1] = woff[l] + K[l] + 1;
  for (int i = lane; i < 7 * C; i += 32) G[i] = 0.f;
  __syncwarp();
#pragma unroll 1
  for (int l = 7; l >= 1; l--) {
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
    for (int t = lane; t < k; t += 32) {  // dW_l[t] = sum_c dpre[c] * in[c + t - left]
      float s = 0.f;
      const int c0 = max(0, left - t), c1 = min(C, C + left - t);
      for (int c = c0; c < c1; c++) s += dpre[c] * in[c + t - left];
      atomicAdd(&gw[woff[l - 1] + t], s);
    }
    for (int cp = lane; cp < C; cp += 32) {  // dL/dnet_{l-1}[c'] += sum_t w[t] * dpre[c' - t + left]
      float s = 0.f;
      const int t0 = max(0, cp + left - (C - 1)), t1 = min(k, cp + left + 1);
      for (int t = t0; t < t1; t++) s += wl[t] * dpre[cp - t + left];
      Gin[cp] += s;
    }
    __syncwarp();
  }
}

// ---- discriminator, one warp per spectrum ---------------------------------------------------------------------
// Weights in shared memory with padded rows (row i of W1 / W2 at stride C + 1, of W3 at stride C / 2 + 1): the forward
// walks a column per lane, the backward a row per lane, both conflict free.
struct GsDisc {
  const float *W1, *b1, *W2, *b2, *W3, *b3;
  int LW, LH;
};
__device__ __forceinline__ int gs_disc_smem_floats(int C) { return 2 * C * (C + 1) + C * (C / 2 + 1) + 2 * C + C / 2; }
__device__ __forceinline__ GsDisc gs_disc_load(float* dst, const float* __restrict__ w, int C) {
  const int H = C / 2, LW = C + 1, LH = H + 1;
  float* W1 = dst;
  float* W2 = W1 + C * LW;
  float* W3 = W2 + C * LW;
  float* b1 = W3 + C * LH;
  float* b2 = b1 + C;
  float* b3 = b2 + C;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
    W1[(i / C) * LW + i % C] = w[i];
    W2[(i / C) * LW + i % C] = w[C * C + C + i];
  }
  for (int i = threadIdx.x; i < C * H; i += blockDim.x) W3[(i / H) * LH + i % H] = w[2 * (C * C + C) + i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { b1[i] = w[C * C + i]; b2[i] = w[2 * C * C + C + i]; }
  for (int i = threadIdx.x; i < H; i += blockDim.x) b3[i] = w[2 * (C * C + C) + C * H + i];
  GsDisc d{W1, b1, W2, b2, W3, b3, LW, LH};
  return d;
}
// v [3][C] = x (in place), h1, h2;  out [H]
__device__ __forceinline__ void gs_disc_forward(const GsDisc& d, float* v, float* out, int C, int lane) {
  const int H = C / 2;
  for (int j = lane; j < C; j += 32) {
    float s = d.b1[j];
    for (int i = 0; i < C; i++) s += v[i] * d.W1[i * d.LW + j];
    v[C + j] = fmaxf(s, 0.1f * s);
  }
  __syncwarp();
  for (int j = lane; j < C; j += 32) {
    float s = d.b2[j];
    for (int i = 0; i < C; i++) s += v[C + i] * d.W2[i * d.LW + j];
    v[2 * C + j] = fmaxf(s, 0.1f * s);
  }
  __syncwarp();
  for (int j = lane; j < H; j += 32) {
    float s = d.b3[j];
    for (int i = 0; i < C; i++) s += v[2 * C + i] * d.W3[i * d.LH + j];
    out[j] = s;
  }
  __syncwarp();
}
// v [3][C] = x, h1, h2 of the forward; dout [H]; scratch d2 [C], d1 [C].  gin (nullable, shared memory [C]) = dL/dx;
// gw (nullable): block accumulator in the dense layout of the weight buffer.
__device__ __forceinline__ void gs_disc_backward(const GsDisc& d, const float* v, const float* dout, float* d2, float* d1,
                                                 float* gin, float* gw, int C, int lane) {
  const int H = C / 2;
  const float *xs = v, *h1 = v + C, *h2 = v + 2 * C;
  float *gW1 = gw, *gb1 = gW1 + C * C, *gW2 = gb1 + C, *gb2 = gW2 + C * C, *gW3 = gb2 + C, *gb3 = gW3 + C * H;
  for (int i = lane; i < C; i += 32) {
    float s = 0.f;
    for (int j = 0; j < H; j++) s += d.W3[i * d.LH + j] * dout[j];
    d2[i] = s * (h2[i] > 0.f ? 1.f : 0.1f);
  }
  if (gw) {
    for (int j = lane; j < H; j += 32) {
      atomicAdd(&gb3[j], dout[j]);
      for (int i = 0; i < C; i++) atomicAdd(&gW3[i * H + j], h2[i] * dout[j]);
    }
  }
  __syncwarp();
  for (int i = lane; i < C; i += 32) {
    float s = 0.f;
    for (int j = 0; j < C; j++) s += d.W2[i * d.LW + j] * d2[j];
    d1[i] = s * (h1[i] > 0.f ? 1.f : 0.1f);
  }
  if (gw) {
    for (int j = lane; j < C; j += 32) {
      atomicAdd(&gb2[j], d2[j]);
      for (int i = 0; i < C; i++) atomicAdd(&gW2[i * C + j], h1[i] * d2[j]);
    }
  }
  __syncwarp();
  if (gin) {
    for (int i = lane; i < C; i += 32) {
      float s = 0.f;
      for (int j = 0; j < C; j++) s += d.W1[i * d.LW + j] * d1[j];
      gin[i] = s;
    }
  }
  if (gw) {
    for (int j = lane; j < C; j += 32) {
      atomicAdd(&gb1[j], d1[j]);
      for (int i = 0; i < C; i++) atomicAdd(&gW1[i * C + j], xs[i] * d1[j]);
    }
  }
  __syncwarp();
}

// ---- generator step -----------------------------------------------------------------------------------------
struct GsGenStepArgs {
  const float *x, *y;      // [rows][C] lit / shadowed spectra
  int64_t rows;
  int C;
  const float *wG, *wF;    // generators x -> y, y -> x
  const float *wDY, *wDX;  // discriminators of the y / x domain (frozen here)
  float w_cyc, w_id;
  float *gG, *gF;          // += weight gradients
  double* loss_acc;        // [4]: [0] += total, [1] += LSGAN generator loss, [2] += cycle term, [3] += identity term
  float *gen_y, *gen_x, *rec_x, *rec_y;  // nullable [rows][C]: G(x), F(y), F(G x), G(F y)
};
__global__ void __launch_bounds__(GS_WARPS * 32) gan_cycle_gstep_kernel(const GsGenStepArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, H = C / 2, ng = gs_gen_nweights(C), ngp = (ng + 3) & ~3;
  float* wG = sm;
  float* wF = wG + ngp;
  float* gwG = wF + ngp;
  float* gwF = gwG + ngp;
  float* dsm = gwF + ngp;
  const int dfl = (gs_disc_smem_floats(C) + 3) & ~3;
  for (int i = threadIdx.x; i < ng; i += blockDim.x) { wG[i] = a.wG[i]; wF[i] = a.wF[i]; gwG[i] = 0.f; gwF[i] = 0.f; }
  const GsDisc DY = gs_disc_load(dsm, a.wDY, C);
  const GsDisc DX = gs_disc_load(dsm + dfl, a.wDX, C);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: four generator applications [4][8][C], gradient workspace [8][C], dpre [C], discriminator [3][C] + d2, d1
  // [2][C] + dout [C], input gradients of G(x), F(y) [2][C]
  float* pw = dsm + 2 * dfl + warp * (32 + 8 + 1 + 3 + 2 + 1 + 2) * C;
  float *n_gx = pw, *n_fy = pw + 8 * C, *n_rx = pw + 16 * C, *n_ry = pw + 24 * C;
  float *G = pw + 32 * C, *dpre = G + 8 * C, *dv = dpre + C, *d2 = dv + 3 * C, *d1 = d2 + C, *dout = d1 + C;
  float *g_gx = dout + C, *g_fy = g_gx + C;
  const float s_gan = 1.f / ((float)a.rows * H), s_cyc = a.w_cyc / ((float)a.rows * C), s_id = 2.f * a.w_id / ((float)a.rows * C);
  float l_gan = 0.f, l_cyc = 0.f, l_id = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * GS_WARPS + warp; r < a.rows; r += (int64_t)gridDim.x * GS_WARPS) {
    const float* x = a.x + r * C;
    const float* y = a.y + r * C;
    for (int c = lane; c < C; c += 32) { n_gx[c] = x[c]; n_fy[c] = y[c]; }
    __syncwarp();
    gs_gen_forward(wG, n_gx, C, lane);  // G(x)
    gs_gen_forward(wF, n_fy, C, lane);  // F(y)
    for (int c = lane; c < C; c += 32) { n_rx[c] = n_gx[7 * C + c]; n_ry[c] = n_fy[7 * C + c]; }
    __syncwarp();
    gs_gen_forward(wF, n_rx, C, lane);  // F(G x)
    gs_gen_forward(wG, n_ry, C, lane);  // G(F y)
    const float *gx = n_gx + 7 * C, *fy = n_fy + 7 * C, *rx = n_rx + 7 * C, *ry = n_ry + 7 * C;
    for (int c = lane; c < C; c += 32) {
      if (a.gen_y) a.gen_y[r * C + c] = gx[c];
      if (a.gen_x) a.gen_x[r * C + c] = fy[c];
      if (a.rec_x) a.rec_x[r * C + c] = rx[c];
      if (a.rec_y) a.rec_y[r * C + c] = ry[c];
    }
    // least-squares generator loss through the frozen discriminators: D_Y(G x) -> 1, D_X(F y) -> 1
    for (int side = 0; side < 2; side++) {
      const float* fake = side == 0 ? gx : fy;
      for (int c = lane; c < C; c += 32) dv[c] = fake[c];
      __syncwarp();
      gs_disc_forward(side == 0 ? DY : DX, dv, dout, C, lane);
      for (int j = lane; j < H; j += 32) {
        const float d = dout[j] - 1.f;
        l_gan += 0.5f * s_gan * d * d;
        dout[j] = s_gan * d;
      }
      __syncwarp();
      gs_disc_backward(side == 0 ? DY : DX, dv, dout, d2, d1, side == 0 ? g_gx : g_fy, nullptr, C, lane);
    }
    // cycle consistency |F(G x) - x|, |G(F y) - y| (tfgan counts the term twice, weight / 2 each), "identity"
    // |x - G(x)|, |y - F(y)| (the generator applied to its own domain's input, cycle_gan_wrapper.py:308-311,325-328)
    for (int c = lane; c < C; c += 32) {
      const float dx = rx[c] - x[c], dy = ry[c] - y[c];
      l_cyc += s_cyc * (fabsf(dx) + fabsf(dy));
      G[7 * C + c] = dx > 0.f ? s_cyc : (dx < 0.f ? -s_cyc : 0.f);
      dpre[c] = dy > 0.f ? s_cyc : (dy < 0.f ? -s_cyc : 0.f);   // parked until the second backward
      if (a.w_id != 0.f) {
        const float ix = gx[c] - x[c], iy = fy[c] - y[c];
        l_id += s_id * (fabsf(ix) + fabsf(iy));
        g_gx[c] += ix > 0.f ? s_id : (ix < 0.f ? -s_id : 0.f);
        g_fy[c] += iy > 0.f ? s_id : (iy < 0.f ? -s_id : 0.f);
      }
    }
    __syncwarp();
    for (int c = lane; c < C; c += 32) dout[c] = dpre[c];  // dL/d G(F y)
    __syncwarp();
    gs_gen_backward(wF, gwF, n_rx, G, dpre, C, lane);       // F applied to G(x): dL/d G(x) +=
    for (int c = lane; c < C; c += 32) { g_gx[c] += G[c]; G[7 * C + c] = dout[c]; }
    __syncwarp();
    gs_gen_backward(wG, gwG, n_ry, G, dpre, C, lane);       // G applied to F(y): dL/d F(y) +=
    for (int c = lane; c < C; c += 32) { g_fy[c] += G[c]; G[7 * C + c] = g_gx[c]; }
    __syncwarp();
    gs_gen_backward(wG, gwG, n_gx, G, dpre, C, lane);       // G applied to x
    for (int c = lane; c < C; c += 32) G[7 * C + c] = g_fy[c];
    __syncwarp();
    gs_gen_backward(wF, gwF, n_fy, G, dpre, C, lane);       // F applied to y
  }
  l_gan = warp_sum(l_gan); l_cyc = warp_sum(l_cyc); l_id = warp_sum(l_id);
  if (lane == 0 && a.loss_acc) {
    atomicAdd(a.loss_acc + 0, (double)l_gan + (double)l_cyc + (double)l_id);
    atomicAdd(a.loss_acc + 1, (double)l_gan);
    atomicAdd(a.loss_acc + 2, (double)l_cyc);
    atomicAdd(a.loss_acc + 3, (double)l_id);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng; i += blockDim.x) { atomicAdd(&a.gG[i], gwG[i]); atomicAdd(&a.gF[i], gwF[i]); }
}

// ---- discriminator step ---------------------------------------------------------------------------------------
struct GsDisStepArgs {
  const float *x, *y;
  int64_t rows;
  int C;
  const float *wG, *wF, *wDY, *wDX;
  float reg;               // slim l2_regularizer scale on the first two FC weight matrices of each discriminator
  float *gDY, *gDX;        // += weight gradients
  double* loss_acc;        // [4]: [0] += total, [1] += LSGAN discriminator loss, [2] += regularisation
  // tfgan tensor pool: pool_y / pool_x [slots][rows][C].  mode 0: no pool (the fresh fakes are used); 1: store the
  // fresh fakes into `slot` and use them (the pool is still filling); 2: use the tensors stored in `slot` and replace
  // them by the fresh ones.  The two domains draw independently.
  float *pool_y, *pool_x;
  int mode_y, slot_y, mode_x, slot_x;
};
// Weight gradients: a warp that accumulated its rank-1 updates h (x) d with shared-memory atomics would issue C * C of
// them per layer and pass, one after the other (measured: most of the step at batch 32).  Instead every warp only keeps
// the vectors of its four discriminator passes (x, h1, h2, d1, d2, dout) and, once per round, ALL threads of the block
// sum the outer products, each thread owning whole weight elements — no atomics inside the block.
constexpr int GS_SAVE_PER_PASS = 6;  // vectors of C floats kept per pass: x, h1, h2, d2, d1, dout
__global__ void __launch_bounds__(GS_WARPS * 32) gan_cycle_dstep_kernel(const GsDisStepArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, H = C / 2, ng = gs_gen_nweights(C), ngp = (ng + 3) & ~3, nd = gs_disc_nweights(C), ndp = (nd + 3) & ~3;
  float* wG = sm;
  float* wF = wG + ngp;
  float* gwY = wF + ngp;   // dense layout of the weight buffer
  float* gwX = gwY + ndp;
  float* dsm = gwX + ndp;
  const int dfl = (gs_disc_smem_floats(C) + 3) & ~3;
  for (int i = threadIdx.x; i < ng; i += blockDim.x) { wG[i] = a.wG[i]; wF[i] = a.wF[i]; }
  for (int i = threadIdx.x; i < nd; i += blockDim.x) { gwY[i] = 0.f; gwX[i] = 0.f; }
  const GsDisc DY = gs_disc_load(dsm, a.wDY, C);
  const GsDisc DX = gs_disc_load(dsm + dfl, a.wDX, C);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* nets = dsm + 2 * dfl + warp * 8 * C;                                   // generator nets [8][C]
  float* save = dsm + 2 * dfl + GS_WARPS * 8 * C;                               // [warp][side][pass][6][C]
  const int pass_fl = GS_SAVE_PER_PASS * C;
  const float s_gan = 1.f / ((float)a.rows * H);
  float l_gan = 0.f;
  const int64_t stride = (int64_t)gridDim.x * GS_WARPS;
  for (int64_t r0 = (int64_t)blockIdx.x * GS_WARPS; r0 < a.rows; r0 += stride) {  // block-uniform rounds
    const int64_t r = r0 + warp;
    const bool valid = r < a.rows;
    if (valid) {
      for (int side = 0; side < 2; side++) {  // 0: D_Y on (y real, G(x) fake); 1: D_X on (x real, F(y) fake)
        const float* real = (side == 0 ? a.y : a.x) + r * C;
        const float* src = (side == 0 ? a.x : a.y) + r * C;
        const GsDisc& D = side == 0 ? DY : DX;
        for (int c = lane; c < C; c += 32) nets[c] = src[c];
        __syncwarp();
        gs_gen_forward(side == 0 ? wG : wF, nets, C, lane);
        const int mode = side == 0 ? a.mode_y : a.mode_x;
        float* slot = (side == 0 ? a.pool_y : a.pool_x);
        if (slot) slot += ((size_t)(side == 0 ? a.slot_y : a.slot_x) * a.rows + r) * C;
        for (int pass = 0; pass < 2; pass++) {  // real -> 1, fake -> 0 (least_squares_discriminator_loss)
          float* dv = save + ((warp * 2 + side) * 2 + pass) * pass_fl;  // x, h1, h2 | d2 | d1 | dout
          float *d2 = dv + 3 * C, *d1 = dv + 4 * C, *dout = dv + 5 * C;
          for (int c = lane; c < C; c += 32) {
            float v;
            if (pass == 0) {
              v = real[c];
            } else {
              const float fresh = nets[7 * C + c];
              v = fresh;
              if (mode == 2) v = slot[c];
              if (mode != 0) slot[c] = fresh;
            }
            dv[c] = v;
          }
          __syncwarp();
          gs_disc_forward(D, dv, dout, C, lane);
          const float target = pass == 0 ? 1.f : 0.f;
          for (int j = lane; j < H; j += 32) {
            const float d = dout[j] - target;
            l_gan += 0.5f * s_gan * d * d;
            dout[j] = s_gan * d;
          }
          __syncwarp();
          gs_disc_backward(D, dv, dout, d2, d1, nullptr, nullptr, C, lane);
        }
      }
    }
    __syncthreads();
    // ---- outer products of this round's rows: thread e owns weight elements e, e + blockDim, ...
    const int nvalid = (int)min((int64_t)GS_WARPS, a.rows - r0);
    for (int side = 0; side < 2; side++) {
      float* gw = side == 0 ? gwY : gwX;
      for (int e = threadIdx.x; e < nd; e += blockDim.x) {
        // dense layout: W1 [C][C] | b1 [C] | W2 [C][C] | b2 [C] | W3 [C][H] | b3 [H];  vectors per pass: x 0, h1 1, h2 2, d2 3, d1 4, dout 5
        int e2 = e, in_vec, out_vec, i, j;
        bool bias = false;
        if (e2 < C * C + C) { in_vec = 0; out_vec = 4; bias = e2 >= C * C; i = e2 / C; j = bias ? e2 - C * C : e2 - i * C; }
        else if ((e2 -= C * C + C) < C * C + C) { in_vec = 1; out_vec = 3; bias = e2 >= C * C; i = e2 / C; j = bias ? e2 - C * C : e2 - i * C; }
        else { e2 -= C * C + C; in_vec = 2; out_vec = 5; bias = e2 >= C * H; i = e2 / H; j = bias ? e2 - C * H : e2 - i * H; }
        float acc = 0.f;
        for (int w = 0; w < nvalid; w++)
#pragma unroll
          for (int pass = 0; pass < 2; pass++) {
            const float* dv = save + ((w * 2 + side) * 2 + pass) * pass_fl;
            acc += (bias ? 1.f : dv[in_vec * C + i]) * dv[out_vec * C + j];
          }
        gw[e] += acc;
      }
    }
    __syncthreads();
  }
  l_gan = warp_sum(l_gan);
  if (lane == 0 && a.loss_acc) {
    atomicAdd(a.loss_acc + 0, (double)l_gan);
    atomicAdd(a.loss_acc + 1, (double)l_gan);
  }
  __syncthreads();
  // block 0 adds the L2 regulariser on the two hidden layers' weight matrices (scale * sum w^2 / 2, grad scale * w)
  float l_reg = 0.f;
  for (int i = threadIdx.x; i < nd; i += blockDim.x) {
    float gy = gwY[i], gx = gwX[i];
    const bool regd = i < C * C || (i >= C * C + C && i < 2 * C * C + C);
    if (blockIdx.x == 0 && regd && a.reg != 0.f) {
      const float wy = a.wDY[i], wx = a.wDX[i];
      gy += a.reg * wy;
      gx += a.reg * wx;
      l_reg += 0.5f * a.reg * (wy * wy + wx * wx);
    }
    atomicAdd(&a.gDY[i], gy);
    atomicAdd(&a.gDX[i], gx);
  }
  if (blockIdx.x == 0 && a.loss_acc) {
    l_reg = warp_sum(l_reg);
    __shared__ float reg_part[GS_WARPS];
    if (lane == 0) reg_part[warp] = l_reg;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < GS_WARPS; i++) t += reg_part[i];
      atomicAdd(a.loss_acc + 0, (double)t);
      atomicAdd(a.loss_acc + 2, (double)t);
    }
  }
}

static size_t gs_gstep_smem(int C) {
  const int ngp = (gs_gen_nweights(C) + 3) & ~3;
  const int dfl = (2 * C * (C + 1) + C * (C / 2 + 1) + 2 * C + C / 2 + 3) & ~3;
  return (size_t)(4 * ngp + 2 * dfl + GS_WARPS * 49 * C) * sizeof(float);
}
static size_t gs_dstep_smem(int C) {
  const int ngp = (gs_gen_nweights(C) + 3) & ~3, ndp = (gs_disc_nweights(C) + 3) & ~3;
  const int dfl = (2 * C * (C + 1) + C * (C / 2 + 1) + 2 * C + C / 2 + 3) & ~3;
  return (size_t)(2 * ngp + 2 * ndp + 2 * dfl + GS_WARPS * (8 + 4 * 6) * C) * sizeof(float);
}

}  // namespace hyp

using namespace hyp;

extern "C" {

int hyp_gan_cycle_generator_step(const float* x, const float* y, int64_t rows, int bands, const float* w_g, const float* w_f,
                                 const float* w_dy, const float* w_dx, float cycle_weight, float identity_weight,
                                 float* grad_g, float* grad_f, double* loss_acc, float* gen_y, float* gen_x, float* rec_x,
                                 float* rec_y, void* stream) {
  HYP_CHECK_ARG(x && y && w_g && w_f && w_dy && w_dx && grad_g && grad_f, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= GS_MAX_C && bands % 8 == 0 && rows >= 0, "bands must be a multiple of 8 in 8..64");
  if (rows == 0) return HYP_OK;
  GsGenStepArgs a;
  a.x = x; a.y = y; a.rows = rows; a.C = bands; a.wG = w_g; a.wF = w_f; a.wDY = w_dy; a.wDX = w_dx;
  a.w_cyc = cycle_weight; a.w_id = identity_weight; a.gG = grad_g; a.gF = grad_f; a.loss_acc = loss_acc;
  a.gen_y = gen_y; a.gen_x = gen_x; a.rec_x = rec_x; a.rec_y = rec_y;
  const size_t smem = gs_gstep_smem(bands);
  static bool attr = false;
  if (!attr) {
    HYP_CUDA(cudaFuncSetAttribute(gan_cycle_gstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  HYP_CHECK_ARG(smem <= 200 * 1024, "bands too large for the fused step's shared memory");
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, GS_WARPS), 148);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_cycle_gstep_kernel", 4.0 * rows * 6 * bands, (gan_cycle_gstep_kernel<<<grid, GS_WARPS * 32, smem, st>>>(a)));
  return HYP_OK;
}

int hyp_gan_cycle_discriminator_step(const float* x, const float* y, int64_t rows, int bands, const float* w_g,
                                     const float* w_f, const float* w_dy, const float* w_dx, float reg_scale, float* grad_dy,
                                     float* grad_dx, double* loss_acc, float* pool_y, int pool_mode_y, int pool_slot_y,
                                     float* pool_x, int pool_mode_x, int pool_slot_x, void* stream) {
  HYP_CHECK_ARG(x && y && w_g && w_f && w_dy && w_dx && grad_dy && grad_dx, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= GS_MAX_C && bands % 8 == 0 && rows >= 0, "bands must be a multiple of 8 in 8..64");
  HYP_CHECK_ARG(pool_mode_y >= 0 && pool_mode_y <= 2 && pool_mode_x >= 0 && pool_mode_x <= 2, "pool mode is 0, 1 or 2");
  HYP_CHECK_ARG((pool_mode_y == 0 || pool_y) && (pool_mode_x == 0 || pool_x), "pool mode without a pool");
  HYP_CHECK_ARG(pool_slot_y >= 0 && pool_slot_x >= 0, "negative pool slot");
  if (rows == 0) return HYP_OK;
  GsDisStepArgs a;
  a.x = x; a.y = y; a.rows = rows; a.C = bands; a.wG = w_g; a.wF = w_f; a.wDY = w_dy; a.wDX = w_dx; a.reg = reg_scale;
  a.gDY = grad_dy; a.gDX = grad_dx; a.loss_acc = loss_acc;
  a.pool_y = pool_mode_y ? pool_y : nullptr; a.mode_y = pool_mode_y; a.slot_y = pool_slot_y;
  a.pool_x = pool_mode_x ? pool_x : nullptr; a.mode_x = pool_mode_x; a.slot_x = pool_slot_x;
  const size_t smem = gs_dstep_smem(bands);
  static bool attr = false;
  if (!attr) {
    HYP_CUDA(cudaFuncSetAttribute(gan_cycle_dstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr = true;
  }
  HYP_CHECK_ARG(smem <= 220 * 1024, "bands too large for the fused step's shared memory");
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, GS_WARPS), 148);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_cycle_dstep_kernel", 4.0 * rows * 4 * bands, (gan_cycle_dstep_kernel<<<grid, GS_WARPS * 32, smem, st>>>(a)));
  return HYP_OK;
}

}  // extern "C"
