// Tensor-core engine (HYP_PRECISION_3XTF32): the same layer plan as the FFMA engine, executed
// with the tcgen05 segment-GEMM of hyp_tc.cuh on position-major activations.
//
// Included by hyp_engine.cu after hyp_model / Layer / Tensor are defined.
//
// Layer kinds
//   rowwise : 1x1 conv or FC fed by an FC          z[r, :]    = a[r, :] W                (one K segment)
//   level   : the k x k convs of one spatial level z[p, b, :] = sum_taps a[p+tap, b, :] W_tap
//             all kernel sizes share the A tile; accumulator columns are "slots"
//             [k_max | ... | 3x3 | 1x1] (fpad columns each) so a tap of ring r = max(|dy|,|dx|)
//             multiplies the first (R - r) slots only
//   flatten : FC fed by a conv tensor              z[b, :]    = sum_pos a[pos, b, :] W_pos
#pragma once
#include <algorithm>
#include <cstdlib>

#include "hyp_tc_kernels.cuh"

namespace hyp {
namespace tc {

struct TcTensor {
  int PP, C, Cp;
  // fp32 value plane (bytes in workspace) and, behind it, the operand region.  3xF16 models drop the value plane of
  // the HYPELCNN-internal tensors (has_value = false): residual readers rebuild the value from the fp16 (hi, lo)
  // planes — 22 significand bits, the precision every GEMM of the mode sees anyway — which saves 4 of the 8 bytes
  // tc_bn_apply writes per element.
  size_t a_off = 0, o_off = 0;
  bool has_value = true;
  size_t plane_elems = 0;
  size_t g_off = 0;
};

struct TcLaunch {
  CUtensorMap tmA, tmB;
  CUtensorMap tmO;        // 2-D view of the output for the epilogue's TMA slabs (make_out_map)
  bool tma_out = false;
  int out_cols = 0;
  int64_t out_rows = 0;
  int tile0 = 0, ntiles = 0;
  int b_rows = 0, bn = 0;
  bool mn = false;
  int cg = 1;  // CTA group size: 2 = tiles come in pairs (2i, 2i+1) that share their B operand
  bool windowed = false;  // schedule: keep the tile order's locality (load-aware dealing inside windows) instead of global LPT
};

struct TcLayer {
  int kind = 0;  // 0 rowwise, 1 level, 2 flatten
  int R = 1, f = 0, fpad = 0, Gp = 0, Kp = 0, Cq = 0;
  // level slots: kernel q = R-1 - slot / nt (largest kernel first), filters [jt * ft, jt * ft + width) with jt = slot % nt;
  // nt > 1 only when a kernel has more than 256 filters (one accumulator holds at most 256 columns)
  int nt = 1, ft = 0, NS = 1;
  int slot_q(int s) const { return R - 1 - s / nt; }
  int slot_ch0(int s) const { return slot_q(s) * f + (s % nt) * ft; }   // first channel in the level's output tensor
  int slot_w(int s) const { return std::min(ft, f - (s % nt) * ft); }
  size_t z_off = 0, mean_off = 0, rstd_off = 0, s1_off = 0, s2_off = 0;
  size_t u_off = 0;  // LRN layers: the activation before local response normalisation (backward needs it)
  int64_t wf_off = 0, wd_off = 0;  // packed forward / dgrad weights: element offsets inside a pack plane
  int wf_rows = 0, wf_ld = 0, wd_rows = 0, wd_ld = 0;
  TcLaunch fwd, dg, wg;
  int stats_rows = 0;
  size_t gzs_off = 0;  // OP_F16X3: [0] bits of the |gz| bound, [1] the power-of-two gz scale, [2] its inverse (4 floats)
};

struct TcState {
  std::vector<TcTensor> tt;
  std::vector<TcLayer> tl;
  size_t gz_off = 0, gz_plane_elems = 0, part_off = 0, bpart_off = 0, pack_off = 0, pack_plane_elems = 0;
  size_t ce_off = 0, mse_off = 0, dbg_off = 0, ws_bytes = 0;
  std::vector<PackJob> jobs;
  PackJob* jobs_dev = nullptr;
  int* job_tiles_dev = nullptr;          // first 32x32 tile of every job (+ total)
  std::vector<int> job_tile_first;
  int job_blocks = 1;
  TcSeg* segs_dev = nullptr;
  TcTile* tiles_dev = nullptr;
  size_t segs_cap = 0, tiles_cap = 0;
  int planned_B = -1;
  bool pack_cleared = false;
  // operand format of every GEMM of the model (hyp_tc.cuh OP_*), K elements per 128-byte block, and the power of two
  // the packed weights are multiplied by before a 16-bit split (undone in the forward / dgrad epilogues)
  int op = OP_TF32X3, kbe = 32;
  float w_scale = 1.f;
  size_t gzs_off = 0, gzs_bytes = 0;
  int rk(int v) const { return (int)align_up((size_t)v, (size_t)kbe); }                  // K padding
  int rc(int v) const { return (int)align_up((size_t)v, (size_t)(op == OP_TF32X3 ? 4 : 8)); }  // row pitch: 16-byte rows
  size_t op_bytes(size_t elems) const { return (size_t)op_planes(op) * op_esize(op) * elems; }   // operand-only buffers
};

inline int r16(int v) { return (int)align_up((size_t)v, 16); }
inline int r32(int v) { return (int)align_up((size_t)v, 32); }
inline int r4(int v) { return (int)align_up((size_t)v, 4); }

}  // namespace tc
}  // namespace hyp

// ---------------------------------------------------------------------------------------------
namespace hyp {
namespace tc {

static float* tc_plane0(const hyp_model& m, int t) {  // fp32 value plane, nullptr where it is not kept
  return m.tc->tt[t].has_value ? reinterpret_cast<float*>(m.ws + m.tc->tt[t].a_off) : nullptr;
}
static float* tc_plane1(const hyp_model& m, int t) { return reinterpret_cast<float*>(m.ws + m.tc->tt[t].o_off); }
static float* tc_grad(const hyp_model& m, int t) { return reinterpret_cast<float*>(m.ws + m.tc->tt[t].g_off); }
// first GEMM operand plane of an activation tensor: the value plane itself (3xTF32) or the 16-bit planes that live in
// the second region
static void* tc_operand(const hyp_model& m, int t) {
  return m.tc->op == OP_TF32X3 ? static_cast<void*>(tc_plane0(m, t)) : static_cast<void*>(tc_plane1(m, t));
}

// HYP_SWEEP_ZIGZAG=0: every element-wise pass sweeps ascending (diagnostic: the L2 carry-over between passes is lost)
static inline int tc_zigzag() {
  static const int v = [] { const char* e = getenv("HYP_SWEEP_ZIGZAG"); return (e && e[0] == '0') ? 0 : 1; }();
  return v;
}

// launch shape of the vectorised elementwise kernels: TX column groups of 4 channels per block row,
// 256 / TX row lanes, `rblocks` block rows sweeping the tensor's rows together (tc_sweep_row)
struct EwGrid { int TX, gx, rblocks; };
static inline EwGrid ew_grid2(int cols, int64_t rows) {
  EwGrid g;
  const int c4 = (int)cdiv(cols, 4);
  g.TX = c4 > 16 ? 32 : (c4 > 8 ? 16 : 8);
  g.gx = (int)cdiv(c4, g.TX);
  const int TY = 256 / g.TX;
  int64_t rb = std::max<int64_t>(1, std::min<int64_t>(cdiv(rows, 4 * TY), cdiv(tc_sm_count() * 8, g.gx)));
  g.rblocks = (int)rb;
  return g;
}

// ---- static layout: workspace offsets, packed-weight offsets, pack jobs ------------------------
static int tc_layout(hyp_model& m) {
  m.tc = new TcState();
  TcState& S = *m.tc;
  S.op = m.d.precision_mode == HYP_PRECISION_3XF16 ? OP_F16X3 : (m.d.precision_mode == HYP_PRECISION_BF16 ? OP_BF16 : OP_TF32X3);
  S.kbe = op_kbe(S.op);
  // weights are O(1e-2): x 2^6 puts the fp16 remainders of everything above 2e-3 into the normal range (smaller weights
  // keep an absolute error below 2^-31, far under the scale of the tensor) and leaves room up to |w| = 1023
  S.w_scale = S.op == OP_F16X3 ? 64.f : 1.f;
  const size_t Bm = (size_t)m.d.max_batch;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    const size_t o = off;
    off += bytes;
    return o;
  };
  S.tt.resize(m.tensors.size());
  for (size_t t = 0; t < m.tensors.size(); t++) {
    TcTensor& T = S.tt[t];
    T.PP = m.tensors[t].rows_per_sample;
    T.C = m.tensors[t].C;
    T.Cp = S.rc(T.C);
    T.plane_elems = align_up(Bm * T.PP * T.Cp, 64);
    // value planes stay for: 3xTF32 (the plane IS an operand) and bf16 (8-bit planes cannot stand in for the value),
    // DUALCNN / CONCNN (LRN and the two-input FC read them), the model inputs (MSE target), logits and reconstruction
    T.has_value = S.op != OP_F16X3 || m.d.kind != HYP_MODEL_HYPELCNN || m.tensors[t].external || (int)t == m.logits_t ||
                  (int)t == m.recon_t;
    T.a_off = take((T.has_value ? 2 : 1) * T.plane_elems * sizeof(float));
    T.o_off = T.a_off + (T.has_value ? T.plane_elems * sizeof(float) : 0);
    if (m.tensors[t].needs_grad) T.g_off = take(T.plane_elems * sizeof(float));
  }
  S.tl.resize(m.layers.size());
  int64_t pk = 0;
  auto take_pack = [&](int64_t elems) {
    pk = (int64_t)align_up((size_t)pk, 64);
    const int64_t o = pk;
    pk += elems;
    return o;
  };
  size_t gz_max = 0, part_max = 0, bpart_max = 0, dbg_max = 0;
  for (size_t li = 0; li < m.layers.size(); li++) {
    Layer& L = m.layers[li];
    TcLayer& T = S.tl[li];
    const int P = L.P;
    const TcTensor& tin = S.tt[L.in_t];
    const TcTensor& tout = S.tt[L.out_t];
    const size_t rows_out = Bm * tout.PP;
    T.z_off = take(rows_out * tout.Cp * sizeof(float));
    if (L.share == 2) T.z_off = S.tl[li - 1].z_off;  // second part of a two-input FC accumulates into the first part's z
    if (L.lrn) T.u_off = take(rows_out * tout.Cp * sizeof(float));
    T.mean_off = take(L.Cout * sizeof(float));
    T.rstd_off = take(L.Cout * sizeof(float));
    T.s1_off = take(L.Cout * sizeof(float));
    T.s2_off = take(L.Cout * sizeof(float));
    dbg_max = std::max(dbg_max, rows_out * (size_t)tout.C);
    dbg_max = std::max(dbg_max, Bm * tin.PP * (size_t)tin.C);
    const bool flatten = L.is_fc && tin.PP > 1;
    const bool level = !L.is_fc && L.ksizes.size() > 0 && !(L.ksizes.size() == 1 && L.ksizes[0] == 1);
    T.kind = flatten ? 2 : (level ? 1 : 0);
    if (T.kind == 0) {
      const int Cin = tin.C;
      T.Kp = S.rk(Cin);
      T.Gp = tout.Cp;
      T.wf_rows = r16(L.Cout); T.wf_ld = T.Kp;
      T.wd_rows = r16(Cin); T.wd_ld = S.rk(L.Cout);
      T.wf_off = take_pack((int64_t)T.wf_rows * T.wf_ld);
      T.wd_off = take_pack((int64_t)T.wd_rows * T.wd_ld);
      PackJob j{};
      j.src_off = L.w_off[0]; j.dst_off = T.wf_off; j.rows = T.wf_rows; j.cols = T.wf_ld; j.ld = T.wf_ld;
      j.RD = T.wf_rows; j.RV = L.Cout; j.KD = T.wf_ld; j.KV = Cin; j.sr1 = 0; j.sr0 = 1; j.sk1 = 0; j.sk0 = L.Cout;
      S.jobs.push_back(j);
      if (m.tensors[L.in_t].needs_grad) {
        PackJob d{};
        d.src_off = L.w_off[0]; d.dst_off = T.wd_off; d.rows = T.wd_rows; d.cols = T.wd_ld; d.ld = T.wd_ld;
        d.RD = T.wd_rows; d.RV = Cin; d.KD = T.wd_ld; d.KV = L.Cout; d.sr1 = 0; d.sr0 = L.Cout; d.sk1 = 0; d.sk0 = 1;
        S.jobs.push_back(d);
      }
      T.stats_rows = (int)cdiv((int64_t)rows_out, 128);
    } else if (T.kind == 1) {
      const int Cin = tin.C;
      T.R = (int)L.ksizes.size();
      for (int q = 0; q < T.R; q++)
        if (L.ksizes[q] != 2 * q + 1) return fail(HYP_E_UNSUPPORTED, "tc engine: level kernels must be 1,3,5,...");
      T.f = L.f;
      T.nt = (int)cdiv(L.f, 256);
      T.ft = (int)cdiv(L.f, T.nt);
      T.NS = T.R * T.nt;
      T.fpad = r16(T.ft);
      T.Kp = S.rk(Cin);
      T.Gp = S.rk(T.NS * T.fpad);
      T.Cq = r16(Cin);
      // taps reach max(|dy|, |dx|) <= h: the largest kernel's radius, at most P - 1 (a 5x5 kernel on a 3x3 patch still
      // connects pixel 0 to pixel 2); tap index = (dy + h) * TW + (dx + h)
      const int h = std::min(T.R - 1, P - 1), TW = 2 * h + 1;
      const int ntaps = TW * TW;
      T.wf_rows = ntaps * T.NS * T.fpad; T.wf_ld = T.Kp;
      T.wd_rows = ntaps * T.Cq; T.wd_ld = T.Gp;
      T.wf_off = take_pack((int64_t)T.wf_rows * T.wf_ld);
      T.wd_off = take_pack((int64_t)T.wd_rows * T.wd_ld);
      for (int dy = -h; dy <= h; dy++)
        for (int dx = -h; dx <= h; dx++) {
          const int tap = (dy + h) * TW + (dx + h);
          const int ring = std::max(std::abs(dy), std::abs(dx));
          for (int slot = 0; slot < T.NS; slot++) {
            const int q = T.slot_q(slot);
            if (ring > q) continue;
            const int k = 2 * q + 1;
            const int64_t src = L.w_off[q] + (int64_t)((dy + q) * k + (dx + q)) * Cin * L.f + (slot % T.nt) * T.ft;
            PackJob j{};
            j.src_off = src; j.dst_off = T.wf_off + (int64_t)((tap * T.NS + slot) * T.fpad) * T.Kp;
            j.rows = T.fpad; j.cols = T.Kp; j.ld = T.Kp;
            j.RD = T.fpad; j.RV = T.slot_w(slot); j.KD = T.Kp; j.KV = Cin; j.sr0 = 1; j.sk0 = L.f;
            S.jobs.push_back(j);
            if (m.tensors[L.in_t].needs_grad) {
              PackJob d{};
              d.src_off = src; d.dst_off = T.wd_off + (int64_t)(tap * T.Cq) * T.Gp + slot * T.fpad;
              d.rows = T.Cq; d.cols = T.fpad; d.ld = T.Gp;
              d.RD = T.Cq; d.RV = Cin; d.KD = T.fpad; d.KV = T.slot_w(slot); d.sr0 = L.f; d.sk0 = 1;
              S.jobs.push_back(d);
            }
          }
        }
      T.stats_rows = (int)(tout.PP * cdiv((int64_t)Bm, 128));
    } else {
      const int Ct = tin.C;
      T.Kp = S.rk(Ct);
      T.Gp = tout.Cp;
      T.Cq = r16(Ct);
      T.wf_rows = r16(L.Cout); T.wf_ld = tin.PP * T.Kp;
      T.wd_rows = tin.PP * T.Cq; T.wd_ld = S.rk(L.Cout);
      T.wf_off = take_pack((int64_t)T.wf_rows * T.wf_ld);
      T.wd_off = take_pack((int64_t)T.wd_rows * T.wd_ld);
      PackJob j{};
      j.src_off = L.w_off[0]; j.dst_off = T.wf_off; j.rows = T.wf_rows; j.cols = T.wf_ld; j.ld = T.wf_ld;
      j.RD = T.wf_rows; j.RV = L.Cout; j.KD = T.Kp; j.KV = Ct; j.sr0 = 1; j.sk1 = Ct * L.Cout; j.sk0 = L.Cout;
      S.jobs.push_back(j);
      PackJob d{};
      d.src_off = L.w_off[0]; d.dst_off = T.wd_off; d.rows = T.wd_rows; d.cols = T.wd_ld; d.ld = T.wd_ld;
      d.RD = T.Cq; d.RV = Ct; d.KD = T.wd_ld; d.KV = L.Cout; d.sr1 = Ct * L.Cout; d.sr0 = L.Cout; d.sk0 = 1;
      S.jobs.push_back(d);
      T.stats_rows = (int)cdiv((int64_t)Bm, 128);
    }
    gz_max = std::max(gz_max, rows_out * (size_t)T.Gp);
    part_max = std::max(part_max, (size_t)T.stats_rows * 2 * L.Cout);
    bpart_max = std::max(bpart_max, (size_t)(ew_grid2(L.Cout, (int64_t)rows_out).rblocks + 1) * 4 * L.Cout);
  }
  S.gz_plane_elems = align_up(gz_max, 64);
  S.gz_off = take(S.op_bytes(S.gz_plane_elems));
  S.gzs_bytes = m.layers.size() * 4 * sizeof(float);  // one contiguous array: cleared with one memset per backward
  S.gzs_off = take(S.gzs_bytes);
  for (size_t li = 0; li < m.layers.size(); li++) S.tl[li].gzs_off = S.gzs_off + li * 4 * sizeof(float);
  S.part_off = take(part_max * sizeof(float));
  S.bpart_off = take(bpart_max * sizeof(float));
  S.pack_plane_elems = align_up((size_t)pk, 64);
  S.pack_off = take(S.op_bytes(S.pack_plane_elems));
  S.ce_off = take(Bm * sizeof(float));
  S.mse_off = take(256);
  S.dbg_off = take(dbg_max * sizeof(float));
  S.ws_bytes = align_up(off, 256);
  S.job_tile_first.assign(1, 0);
  for (const PackJob& j : S.jobs)
    S.job_tile_first.push_back(S.job_tile_first.back() + (int)(cdiv(j.rows, 32) * cdiv(j.cols, 32)));
  S.job_blocks = S.job_tile_first.back();
  return HYP_OK;
}

static void tc_destroy(hyp_model& m) {
  if (!m.tc) return;
  if (m.tc->jobs_dev) cudaFree(m.tc->jobs_dev);
  if (m.tc->job_tiles_dev) cudaFree(m.tc->job_tiles_dev);
  if (m.tc->segs_dev) cudaFree(m.tc->segs_dev);
  if (m.tc->tiles_dev) cudaFree(m.tc->tiles_dev);
  delete m.tc;
  m.tc = nullptr;
}

static int tc_bind(hyp_model& m) {
  TcState& S = *m.tc;
  if (!S.jobs_dev) {
    HYP_CUDA(cudaMalloc(&S.jobs_dev, S.jobs.size() * sizeof(PackJob)));
    HYP_CUDA(cudaMemcpy(S.jobs_dev, S.jobs.data(), S.jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    HYP_CUDA(cudaMalloc(&S.job_tiles_dev, S.job_tile_first.size() * sizeof(int)));
    HYP_CUDA(cudaMemcpy(S.job_tiles_dev, S.job_tile_first.data(), S.job_tile_first.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  HYP_CUDA(cudaMemset(m.ws + S.pack_off, 0, S.op_bytes(S.pack_plane_elems)));
  for (size_t li = 0; li < m.layers.size(); li++)
    if (m.layers[li].bias_mode) {  // no normaliser: z + biases is the (mean 0, rstd 1, beta = biases) case of the BN kernels
      const int C = m.layers[li].Cout;
      HYP_CUDA(cudaMemset(m.ws + S.tl[li].mean_off, 0, C * sizeof(float)));
      tc_fill_kernel<<<(unsigned)cdiv(C, 256), 256>>>(reinterpret_cast<float*>(m.ws + S.tl[li].rstd_off), C, 1.f);
      HYP_LAUNCHED();
    }
  S.planned_B = -1;
  return HYP_OK;
}

// ---- per-batch-size plan: tensor maps, tiles, segments ------------------------------------------
struct PlanBuf {
  std::vector<TcSeg> segs;
  std::vector<TcTile> tiles;
};

static TcTile blank_tile() {
  TcTile t;
  memset(&t, 0, sizeof(t));
  return t;
}
static void schedule_tiles(std::vector<TcTile>& tiles, size_t begin, const std::vector<TcSeg>& segs, int cg, bool windowed);
static void finish_launch(PlanBuf& pb, TcLaunch& l) {
  schedule_tiles(pb.tiles, (size_t)l.tile0, pb.segs, l.cg, l.windowed || !l.mn);
  l.ntiles = (int)pb.tiles.size() - l.tile0;
}
// the launch's output as a 2-D fp32 view [rows][ld] with `cols` valid columns: its one-column-block tiles accumulate
// through the TMA unit (tc_gemm_kernel, store_slab)
static int out_view(TcLaunch& l, const float* base, int cols, int64_t rows, int ld) {
  const int rc = make_out_map(&l.tmO, base, (uint64_t)cols, (uint64_t)rows, (uint64_t)ld);
  if (rc) return rc;
  l.tma_out = true; l.out_cols = cols; l.out_rows = rows;
  return HYP_OK;
}

// wgrad (MN-major) launches run as CTA pairs over two adjacent 128-row tiles of the M side when their count is even
// (the pair shares its gz columns); B rows reserved per stage: boxes of kb columns covering n / cg columns, per CTA
static inline int wg_cg(int mt) { return (mt % 2 == 0) ? 2 : 1; }
// K steps the last K block of a K-major segment needs when its K range holds `k_elems` real elements (TcSeg.ks_last)
static inline int tc_ks_last(int64_t k_elems, int64_t nk, int kb) {
  if (cdiv(k_elems, kb) != nk) return 0;  // whole padding blocks behind the data: not expected, keep every step
  const int rem = (int)(k_elems % kb);
  return rem ? (int)cdiv(rem, kb / 4) : 4;
}
static inline int wg_brows(int n_mma, int cg, int kb) { return cg * (int)cdiv(n_mma / cg, kb) * kb; }

// K-major launches run as CTA pairs (cta_group::2): consecutive tiles (2i, 2i+1) must share their
// segment list.  A run of tiles that differ only in their row block is closed with a phantom
// tile (no valid rows; its A box is out of bounds and reads zeros) when its length is odd.
constexpr int TC_CG_KMAJOR = 2;
static void close_pair_run(std::vector<TcTile>& tiles, size_t run_begin, int oob_a1) {
  if ((tiles.size() - run_begin) % TC_CG_KMAJOR == 0) return;
  TcTile t = tiles.back();
  t.m_valid = 0;
  t.a1_add = oob_a1;
  tiles.push_back(t);
}


// ---- static tile schedule ------------------------------------------------------------------------
// The GEMM kernel is persistent: CTA group g walks tile slots g, g + G, g + 2G, ...  Tiles of one
// launch differ in cost (taps per position, ring-dependent widths, positions per wgrad chunk), so
// the launch's tile array is re-laid-out here: slot i * G + g holds the i-th unit of group g.
//   windowed = true  (K-major level launches): units keep their locality order (batch-pair major: a
//       wave of CTAs shares its A rows through L2); inside every window of 2G units the costs are
//       sorted and dealt out heaviest first to the least loaded group (2 units per group and window).
//   windowed = false (wgrad): plain LPT, longest unit first onto the least loaded group; groups that
//       end up with fewer units get empty tiles.
static double tile_cost(const TcTile& t, const std::vector<TcSeg>& segs) {
  double c = 1500.0;  // epilogue
  for (int i = 0; i < t.seg_count; i++) {
    const TcSeg& sg = segs[t.seg_begin + i];
    c += (double)sg.nk * (256.0 + (double)sg.n_mma * std::max(1, sg.nsets));
  }
  return c;
}
// the assignment itself, on plain (cost, unit) pairs: lists[g] = the units of group g in execution order
static std::vector<std::vector<int>> assign_units(std::vector<std::pair<double, int>> cost, int G, bool windowed) {
  const int units = (int)cost.size();
  std::vector<std::vector<int>> lists(G);
  if (windowed) {
    // Inside every window the sorted units go, heaviest first, to the group with the smallest load accumulated so
    // far (at most ceil(window / G) units per group and window).  The windows are ASSIGNED last to first — the
    // partially filled last window hands one unit to only some groups, and the full windows processed after it
    // make up for that — but EXECUTED in their original order.  HYP_TC_SCHED=snake: the plain heaviest-with-
    // lightest snake per window; =forward: assign first to last (measured 3 % / 0.5 % slower steps).
    static const char* mode = getenv("HYP_TC_SCHED") ? getenv("HYP_TC_SCHED") : "reverse";
    static const bool snake = !strcmp(mode, "snake"), forward = !strcmp(mode, "forward");
    const int W = 2 * G, nwin = (int)cdiv(units, W);
    std::vector<double> load(G, 0.0);
    std::vector<std::vector<std::vector<int>>> wl(nwin, std::vector<std::vector<int>>(G));
    for (int wi = 0; wi < nwin; wi++) {
      const int w = (snake || forward) ? wi : nwin - 1 - wi;
      const int w0 = w * W, w1 = std::min(units, w0 + W);
      std::stable_sort(cost.begin() + w0, cost.begin() + w1,
                       [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
      const int cap = (int)cdiv(w1 - w0, G);
      std::vector<int> taken(G, 0);
      for (int j = w0; j < w1; j++) {
        const int k = j - w0;
        int g = k < G ? k : 2 * G - 1 - k;
        if (!snake) {
          g = -1;
          for (int c = 0; c < G; c++)
            if (taken[c] < cap && (g < 0 || load[c] < load[g])) g = c;
        }
        taken[g]++;
        load[g] += cost[j].first;
        wl[w][g].push_back(cost[j].second);
      }
    }
    for (int w = 0; w < nwin; w++)
      for (int g = 0; g < G; g++) lists[g].insert(lists[g].end(), wl[w][g].begin(), wl[w][g].end());
    if (getenv("HYP_TC_SCHED_DEBUG")) {
      double mx = 0, sum = 0;
      for (double l : load) { mx = std::max(mx, l); sum += l; }
      fprintf(stderr, "[tc_sched] units=%d G=%d predicted max/avg load = %.3f\n", units, G, mx / (sum / G));
    }
  } else {
    std::stable_sort(cost.begin(), cost.end(),
                     [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
    std::vector<double> load(G, 0.0);
    for (const auto& cu : cost) {
      int best = 0;
      for (int g = 1; g < G; g++)
        if (load[g] < load[best]) best = g;
      load[best] += cu.first;
      lists[best].push_back(cu.second);
    }
    if (getenv("HYP_TC_SCHED_DEBUG")) {
      double mx = 0, sum = 0;
      for (double l : load) { mx = std::max(mx, l); sum += l; }
      fprintf(stderr, "[tc_sched] LPT units=%d G=%d predicted max/avg load = %.3f (largest unit / avg load %.3f)\n", units, G,
              mx / (sum / G), cost[0].first / (sum / G));
    }
  }
  return lists;
}
static void schedule_tiles(std::vector<TcTile>& tiles, size_t begin, const std::vector<TcSeg>& segs, int cg, bool windowed) {
  const int units = (int)((tiles.size() - begin) / cg);
  const int G = tc_sm_count() / cg;
  if (units <= G) return;
  std::vector<std::pair<double, int>> cost(units);
  bool uniform = true;
  for (int u = 0; u < units; u++) {
    cost[u] = {tile_cost(tiles[begin + (size_t)u * cg], segs), u};
    uniform = uniform && cost[u].first == cost[0].first;
  }
  if (uniform) return;
  const std::vector<std::vector<int>> lists = assign_units(cost, G, windowed);
  size_t depth = 0;
  for (const auto& l : lists) depth = std::max(depth, l.size());
  std::vector<TcTile> out(depth * G * cg, blank_tile());
  for (int g = 0; g < G; g++)
    for (size_t i = 0; i < lists[g].size(); i++)
      for (int r = 0; r < cg; r++) out[(i * G + g) * cg + r] = tiles[begin + (size_t)lists[g][i] * cg + r];
  tiles.resize(begin);
  tiles.insert(tiles.end(), out.begin(), out.end());
}
static thread_local int g_map_op = OP_TF32X3;  // operand format of the plan being built (tc_plan is single-threaded per model)
static int map4(CUtensorMap* mp, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                uint64_t plane, uint32_t b0, uint32_t b1, bool mn) {
  const uint64_t dims[4] = {d0, d1, d2, (uint64_t)op_planes(g_map_op)};
  const uint64_t strides[3] = {s1, s2, plane};
  const uint32_t box[4] = {b0, b1, 1, 1};
  return make_map(mp, base, dims, strides, box, mn, g_map_op);
}

// split `total` output columns into the fewest tiles of <= 256, each a multiple of 16 wide
static void n_tiling(int total, int& ntn, int& nw) {
  ntn = (int)cdiv(total, 256);
  nw = r16((int)cdiv(total, ntn));
}

// ---- level wgrad tap groups ---------------------------------------------------------------------
// The taps of one ring multiply the same slot prefix, and neighbouring taps read almost the same activation rows.  A
// group of neighbouring taps of a ring shares ONE A tile per (source position, K block) and multiplies it with the
// group's gz tiles (B sets, TcSeg.nsets).  h = tap radius, R kernels, nt N tiles per kernel, NS slots of fpad columns,
// ngroups slot groups per position (sets only when there is one), PP positions, max_sets taps per group at most.
struct TapGroup { std::vector<std::pair<int, int>> t; int ring; };
static std::vector<TapGroup> level_tap_groups(int h, int R, int nt, int NS, int fpad, int ngroups, int PP, int max_sets) {
  std::vector<TapGroup> groups;
  for (int ring = 0; ring <= h; ring++) {
    // the ring's taps in perimeter order (neighbours in the list are neighbours in the window)
    std::vector<std::pair<int, int>> per;
    if (ring == 0) per.push_back({0, 0});
    else {
      for (int dx = -ring; dx < ring; dx++) per.push_back({-ring, dx});
      for (int dy = -ring; dy < ring; dy++) per.push_back({dy, ring});
      for (int dx = ring; dx > -ring; dx--) per.push_back({ring, dx});
      for (int dy = ring; dy > -ring; dy--) per.push_back({dy, -ring});
    }
    const int slots = std::min(NS, (R - ring) * nt);
    int gsz = 1;
    if (ngroups == 1 && slots > 0 && PP < 255)  // (positions travel as bytes in TcSeg.b2x, 255 = out of bounds)
      gsz = std::max(1, std::min(std::min(max_sets, TC_MAX_COLS / (slots * fpad)), TC_MAX_CB / slots));
    for (size_t i = 0; i < per.size(); i += gsz) {
      TapGroup g;
      g.ring = ring;
      for (size_t j = i; j < std::min(per.size(), i + gsz); j++) g.t.push_back(per[j]);
      groups.push_back(g);
    }
  }
  return groups;
}
// gz (output) position that tap (dy, dx) pairs with source position q of a P x P patch; 255 = outside the patch (an
// out-of-bounds TMA coordinate: the set contributes zeros)
static inline int tap_output_position(int P, int q, int dy, int dx) {
  const int ph = q / P - dy, pw = q % P - dx;
  return (ph >= 0 && ph < P && pw >= 0 && pw < P) ? ph * P + pw : 255;
}

static int tc_plan(hyp_model& m, int64_t B) {
  TcState& S = *m.tc;
  if (S.planned_B == B) return HYP_OK;
  PlanBuf pb;
  const int nbt = (int)cdiv(B, 128);
  const int CG = TC_CG_KMAJOR;
  const int KB = S.kbe, esz = op_esize(S.op);  // K elements per 128-byte block, bytes per operand element
  g_map_op = S.op;
  // packed weights / gz hold operand planes only; element offsets -> pointers
  auto pk = [&](int64_t off) -> const void* { return m.ws + S.pack_off + (size_t)off * esz; };
  const void* gz0 = m.ws + S.gz_off;
  int rc;
  for (size_t li = 0; li < m.layers.size(); li++) {
    Layer& L = m.layers[li];
    TcLayer& T = S.tl[li];
    const int P = L.P, h = std::min(T.R - 1, P - 1), TW = 2 * h + 1;  // tap radius / tap-table pitch (see tc_layout)
    // taps sorted by ring so that the MMA N of a tile's segments never increases
    std::vector<std::pair<int, int>> taps;
    if (T.kind == 1)
      for (int ring = 0; ring <= h; ring++)
        for (int dy = -h; dy <= h; dy++)
          for (int dx = -h; dx <= h; dx++)
            if (std::max(std::abs(dy), std::abs(dx)) == ring) taps.push_back({dy, dx});
    const TcTensor& tin = S.tt[L.in_t];
    const TcTensor& tout = S.tt[L.out_t];
    const bool need_dgrad = m.tensors[L.in_t].needs_grad;
    const void* a0 = tc_operand(m, L.in_t);
    const int64_t rows_in = B * tin.PP, rows_out = B * tout.PP;
    if (T.kind == 0) {
      const int Cin = tin.C, Cout = L.Cout;
      // ---------------- forward ----------------
      int ntn, nw;
      n_tiling(Cout, ntn, nw);
      if ((rc = map4(&T.fwd.tmA, a0, Cin, rows_in, 1, tin.Cp, (uint64_t)rows_in * tin.Cp, tin.plane_elems, KB, 128, false))) return rc;
      if ((rc = map4(&T.fwd.tmB, pk(T.wf_off), T.wf_ld, T.wf_rows, 1, T.wf_ld, (uint64_t)T.wf_rows * T.wf_ld,
                     S.pack_plane_elems, KB, nw / CG, false))) return rc;
      T.fwd.mn = false; T.fwd.cg = CG; T.fwd.b_rows = nw; T.fwd.bn = nw; T.fwd.tile0 = (int)pb.tiles.size();
      const int seg0 = (int)pb.segs.size();
      for (int j = 0; j < ntn; j++) {
        TcSeg s{};
        s.b1 = j * nw; s.nk = T.Kp / KB; s.n_mma = r16(std::min(nw, Cout - j * nw)); s.nb = 1;
        s.ks_last = tc_ks_last(Cin, s.nk, KB);
        pb.segs.push_back(s);
      }
      const int nrt = (int)cdiv(rows_out, 128);
      for (int rt2 = 0; rt2 < nrt; rt2 += CG)
        for (int j = 0; j < ntn; j++) {
          const size_t run = pb.tiles.size();
          for (int rt = rt2; rt < std::min(nrt, rt2 + CG); rt++) {
            TcTile t = blank_tile();
            t.seg_begin = seg0 + j; t.seg_count = 1; t.total_kb = T.Kp / KB;
            t.m_valid = (int)std::min<int64_t>(128, rows_out - (int64_t)rt * 128);
            t.ncb = 1; t.ld_out = tout.Cp; t.stats_row = rt; t.a1_add = rt * 128;
            t.cb[0].out_off = (int64_t)rt * 128 * tout.Cp + j * nw;
            t.cb[0].width = std::min(nw, Cout - j * nw);
            t.cb[0].stats_col = j * nw;
            pb.tiles.push_back(t);
          }
          close_pair_run(pb.tiles, run, nrt * 128);
        }
      finish_launch(pb, T.fwd);
      T.stats_rows = nrt;
      // ---------------- dgrad ----------------
      if (need_dgrad) {
        n_tiling(Cin, ntn, nw);
        if ((rc = map4(&T.dg.tmA, gz0, Cout, rows_out, 1, T.Gp, (uint64_t)rows_out * T.Gp, S.gz_plane_elems, KB, 128, false))) return rc;
        if ((rc = map4(&T.dg.tmB, pk(T.wd_off), T.wd_ld, T.wd_rows, 1, T.wd_ld, (uint64_t)T.wd_rows * T.wd_ld,
                       S.pack_plane_elems, KB, nw / CG, false))) return rc;
        T.dg.mn = false; T.dg.cg = CG; T.dg.b_rows = nw; T.dg.bn = nw; T.dg.tile0 = (int)pb.tiles.size();
        const int dseg0 = (int)pb.segs.size();
        for (int j = 0; j < ntn; j++) {
          TcSeg s{};
          s.b1 = j * nw; s.nk = T.wd_ld / KB; s.n_mma = r16(std::min(nw, Cin - j * nw)); s.nb = 1;
          s.ks_last = tc_ks_last(Cout, s.nk, KB);
          pb.segs.push_back(s);
        }
        for (int rt2 = 0; rt2 < nrt; rt2 += CG)
          for (int j = 0; j < ntn; j++) {
            const size_t run = pb.tiles.size();
            for (int rt = rt2; rt < std::min(nrt, rt2 + CG); rt++) {
              TcTile t = blank_tile();
              t.seg_begin = dseg0 + j; t.seg_count = 1; t.total_kb = T.wd_ld / KB;
              t.m_valid = (int)std::min<int64_t>(128, rows_out - (int64_t)rt * 128);
              t.ncb = 1; t.ld_out = tin.Cp; t.a1_add = rt * 128;
              t.cb[0].out_off = (int64_t)rt * 128 * tin.Cp + j * nw;
              t.cb[0].width = std::min(nw, Cin - j * nw);
              pb.tiles.push_back(t);
            }
            close_pair_run(pb.tiles, run, nrt * 128);
          }
        finish_launch(pb, T.dg);
        if ((rc = out_view(T.dg, tc_grad(m, L.in_t), Cin, rows_in, tin.Cp))) return rc;
      }
      // ---------------- wgrad ----------------
      {
        n_tiling(Cout, ntn, nw);
        const int mt = (int)cdiv(Cin, 128);
        if ((rc = map4(&T.wg.tmA, a0, Cin, rows_in, 1, tin.Cp, (uint64_t)rows_in * tin.Cp, tin.plane_elems, KB, KB, true))) return rc;
        if ((rc = map4(&T.wg.tmB, gz0, Cout, rows_out, 1, T.Gp, (uint64_t)rows_out * T.Gp, S.gz_plane_elems, KB, KB, true))) return rc;
        T.wg.mn = true; T.wg.bn = 32; T.wg.b_rows = 0; T.wg.tile0 = (int)pb.tiles.size();
        T.wg.cg = wg_cg(mt);
        const int kblocks = (int)cdiv(rows_out, KB);
        int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(2 * tc_sm_count(), mt * ntn), std::max(1, kblocks / 4)));
        {  // nudge the split so that the units fill the persistent grid evenly (units / groups just below an integer)
          const double groups = (double)tc_sm_count() / T.wg.cg, units0 = (double)mt * ntn / T.wg.cg;
          auto ratio = [&](int k) { const double per = units0 * k / groups; return std::ceil(per) / per; };
          int bestk = ksplit;
          for (int k = std::max(1, ksplit / 2); k <= std::min(2 * ksplit, std::max(1, kblocks / 4)); k++)
            if (ratio(k) < ratio(bestk) - 1e-9) bestk = k;
          if (ratio(ksplit) > 1.03) ksplit = bestk;
        }
        const int kb_per = (int)cdiv(kblocks, ksplit);
        // K piece outermost: the (im, j) tiles of one row range run in the same wave and share their A / gz rows
        // through L2, so every activation and gradient row crosses HBM once
        for (int kb0 = 0; kb0 < kblocks; kb0 += kb_per)
          for (int j = 0; j < ntn; j++)
            for (int im = 0; im < mt; im++) {  // adjacent M tiles are consecutive: CTA pairs share (j, kb0)
              const int width = std::min(nw, Cout - j * nw);
              TcSeg s{};
              s.a1 = kb0 * KB; s.b0 = j * nw; s.b1 = kb0 * KB;
              s.nk = std::min(kb_per, kblocks - kb0); s.n_mma = r16(width); s.nb = (int)cdiv(s.n_mma, KB);
              T.wg.b_rows = std::max(T.wg.b_rows, wg_brows(s.n_mma, T.wg.cg, KB));
              TcTile t = blank_tile();
              t.a0_add = im * 128;
              t.seg_begin = (int)pb.segs.size(); t.seg_count = 1; t.total_kb = s.nk;
              t.m_valid = std::min(128, Cin - im * 128);
              t.ncb = 1; t.ld_out = Cout;
              t.cb[0].out_off = L.w_off[0] + (int64_t)im * 128 * Cout + j * nw;
              t.cb[0].width = width;
              pb.segs.push_back(s);
              pb.tiles.push_back(t);
            }
        finish_launch(pb, T.wg);
      }
    } else if (T.kind == 1) {
      const int Cin = tin.C, PP = tin.PP, R = T.R, NS = T.NS, nt = T.nt, fpad = T.fpad, f = T.f;
      const int spg = std::min(std::min(256 / fpad, TC_MAX_CB), NS);  // slots per accumulator group
      const int ngroups = (int)cdiv(NS, spg);
      // ---------------- forward ----------------
      if ((rc = map4(&T.fwd.tmA, a0, Cin, B, PP, tin.Cp, (uint64_t)B * tin.Cp, tin.plane_elems, KB, 128, false))) return rc;
      if ((rc = map4(&T.fwd.tmB, pk(T.wf_off), T.wf_ld, T.wf_rows, 1, T.wf_ld, (uint64_t)T.wf_rows * T.wf_ld,
                     S.pack_plane_elems, KB, fpad / CG, false))) return rc;
      T.fwd.mn = false; T.fwd.cg = CG; T.fwd.b_rows = spg * fpad; T.fwd.bn = fpad; T.fwd.tile0 = (int)pb.tiles.size();
      // one segment list per (position, slot group); tiles are emitted batch-pair major so that a wave of
      // CTAs works on few batch rows x all positions (the A rows stay in L2 across the taps that reuse
      // them), heaviest positions first (static round-robin over persistent CTAs stays balanced)
      struct Run { int seg0, nseg, tkb, p, g; };
      std::vector<Run> runs;
      for (int p = 0; p < PP; p++) {
        const int ph = p / P, pw = p % P;
        for (int g = 0; g < ngroups; g++) {
          const int s0 = g * spg, s1 = std::min(NS, s0 + spg);
          const int seg0 = (int)pb.segs.size();
          for (auto& tp : taps) {
            const int dy = tp.first, dx = tp.second, ring = std::max(std::abs(dy), std::abs(dx));
            if (ph + dy < 0 || ph + dy >= P || pw + dx < 0 || pw + dx >= P) continue;
            const int nbx = std::min(s1, (R - ring) * nt) - s0;
            if (nbx <= 0) continue;
            const int tap = (dy + h) * TW + (dx + h);
            TcSeg s{};
            s.a2 = p + dy * P + dx; s.b1 = tap * NS * fpad; s.nk = T.Kp / KB; s.n_mma = nbx * fpad; s.nb = nbx;
            s.ks_last = tc_ks_last(Cin, s.nk, KB);
            pb.segs.push_back(s);
          }
          const int nseg = (int)pb.segs.size() - seg0;
          runs.push_back({seg0, nseg, nseg * (T.Kp / KB), p, g});
        }
      }
      std::stable_sort(runs.begin(), runs.end(), [](const Run& a, const Run& b) { return a.tkb > b.tkb; });
      for (int bt2 = 0; bt2 < nbt; bt2 += CG)
        for (const Run& rn : runs) {
          const int s0 = rn.g * spg, s1 = std::min(NS, s0 + spg), p = rn.p;
          const size_t run = pb.tiles.size();
          for (int bt = bt2; bt < std::min(nbt, bt2 + CG); bt++) {
            TcTile t = blank_tile();
            t.seg_begin = rn.seg0; t.seg_count = rn.nseg; t.total_kb = rn.tkb;
            t.m_valid = (int)std::min<int64_t>(128, B - (int64_t)bt * 128);
            t.ld_out = tout.Cp; t.stats_row = p * nbt + bt; t.a1_add = bt * 128; t.b1_add = s0 * fpad;
            t.ncb = s1 - s0;
            for (int s = s0; s < s1; s++) {
              TcColBlock& cb = t.cb[s - s0];
              cb.tcol = (s - s0) * fpad; cb.width = T.slot_w(s);
              cb.out_off = ((int64_t)p * B + (int64_t)bt * 128) * tout.Cp + T.slot_ch0(s);
              cb.stats_col = T.slot_ch0(s);
            }
            pb.tiles.push_back(t);
          }
          close_pair_run(pb.tiles, run, nbt * 128);
        }
      finish_launch(pb, T.fwd);
      T.stats_rows = PP * nbt;
      // ---------------- dgrad ----------------
      if (need_dgrad) {
        int ntn, nw;
        n_tiling(Cin, ntn, nw);
        if ((rc = map4(&T.dg.tmA, gz0, T.Gp, B, PP, T.Gp, (uint64_t)B * T.Gp, S.gz_plane_elems, KB, 128, false))) return rc;
        if ((rc = map4(&T.dg.tmB, pk(T.wd_off), T.wd_ld, T.wd_rows, 1, T.wd_ld, (uint64_t)T.wd_rows * T.wd_ld,
                       S.pack_plane_elems, KB, nw / CG, false))) return rc;
        T.dg.mn = false; T.dg.cg = CG; T.dg.b_rows = nw; T.dg.bn = nw; T.dg.tile0 = (int)pb.tiles.size();
        struct DRun { int seg0, nseg, tkb, p, j; };
        std::vector<DRun> druns;
        for (int p = 0; p < PP; p++) {
          const int ph = p / P, pw = p % P;
          for (int j = 0; j < ntn; j++) {
            const int seg0 = (int)pb.segs.size();
            int tkb = 0;
            for (auto& tp : taps) {
              const int dy = tp.first, dx = tp.second, ring = std::max(std::abs(dy), std::abs(dx));
              if (ph - dy < 0 || ph - dy >= P || pw - dx < 0 || pw - dx >= P) continue;
              const int tap = (dy + h) * TW + (dx + h);
              TcSeg s{};
              s.a2 = p - (dy * P + dx); s.b1 = tap * T.Cq + j * nw;
              s.nk = (int)cdiv((R - ring) * nt * fpad, KB); s.n_mma = r16(std::min(nw, Cin - j * nw)); s.nb = 1;
              s.ks_last = tc_ks_last((R - ring) * nt * fpad, s.nk, KB);
              tkb += s.nk;
              pb.segs.push_back(s);
            }
            druns.push_back({seg0, (int)pb.segs.size() - seg0, tkb, p, j});
          }
        }
        std::stable_sort(druns.begin(), druns.end(), [](const DRun& a, const DRun& b) { return a.tkb > b.tkb; });
        for (int bt2 = 0; bt2 < nbt; bt2 += CG)
          for (const DRun& rn : druns) {
            const size_t run = pb.tiles.size();
            for (int bt = bt2; bt < std::min(nbt, bt2 + CG); bt++) {
              TcTile t = blank_tile();
              t.seg_begin = rn.seg0; t.seg_count = rn.nseg; t.total_kb = rn.tkb;
              t.m_valid = (int)std::min<int64_t>(128, B - (int64_t)bt * 128);
              t.ncb = 1; t.ld_out = tin.Cp; t.a1_add = bt * 128;
              t.cb[0].out_off = ((int64_t)rn.p * B + (int64_t)bt * 128) * tin.Cp + rn.j * nw;
              t.cb[0].width = std::min(nw, Cin - rn.j * nw);
              pb.tiles.push_back(t);
            }
            close_pair_run(pb.tiles, run, nbt * 128);
          }
        finish_launch(pb, T.dg);
        if ((rc = out_view(T.dg, tc_grad(m, L.in_t), Cin, rows_in, tin.Cp))) return rc;
      }
      // ---------------- wgrad ----------------
      {
        if ((rc = map4(&T.wg.tmA, a0, Cin, B, PP, tin.Cp, (uint64_t)B * tin.Cp, tin.plane_elems, KB, KB, true))) return rc;
        if ((rc = map4(&T.wg.tmB, gz0, T.Gp, B, PP, T.Gp, (uint64_t)B * T.Gp, S.gz_plane_elems, KB, KB, true))) return rc;
        T.wg.mn = true; T.wg.bn = 32; T.wg.b_rows = 0; T.wg.tile0 = (int)pb.tiles.size();
        const int mt = (int)cdiv(Cin, 128);
        T.wg.cg = wg_cg(mt);
        const int nkb = (int)cdiv(B, KB);
        int64_t pairs = 0;
        for (auto& tp : taps) pairs += (int64_t)(P - std::abs(tp.first)) * (P - std::abs(tp.second));
        // Batch slices: every (tap, position) pair re-reads a[p + tap] and gz[p]; the whole level (activations +
        // gradients, both planes) is several hundred MB, so the K range (batch rows) is cut into slices whose
        // operands fit L2 and all tiles of a slice run in the same waves -> each row crosses HBM about once.
        // The slices accumulate into the weight gradient with the epilogue's atomics (split-K).
        const double level_bytes = (double)B * PP * (tin.Cp + T.Gp) * (double)(op_planes(S.op) * esz);
        int nslice = 1;
        static const double slice_bytes = getenv("HYP_WG_SLICE_MB") ? atof(getenv("HYP_WG_SLICE_MB")) * 1e6 : 128e6;
        while (nslice < nkb && level_bytes / nslice > slice_bytes) nslice *= 2;
        const int slice_kb = (int)cdiv(nkb, nslice);
        nslice = (int)cdiv(nkb, slice_kb);
        // tile = (slice, chunk of source positions, tap group, slot group, M tile).  Order: slice, then position chunk,
        // then tap group: the groups of one chunk reuse its activation rows back to back and the slice's gz rows stay
        // L2-resident across chunks (every position is in every other's 7x7 neighbourhood).
        // Tap groups (level_tap_groups above), two taps by default: the activation traffic per tap, which paces these
        // launches together with the MMA stream, drops by the group size.
        // Measured (C2, 4096 patches, ms per launch of the three levels): 1 set 0.60 / 0.86 / 0.38, 2 sets 0.55 / 0.73 /
        // 0.35, 3 sets 0.56 / 0.78 / 0.37, 4 sets 0.58 / 0.76 / 0.37 -- beyond two, the stage ring shrinks to two stages and
        // the zero-filled border sets cost more MMAs than the shared tile saves.
        static const int max_sets = getenv("HYP_WG_TAP_SETS") ? std::max(1, std::min(TC_MAX_SETS, atoi(getenv("HYP_WG_TAP_SETS")))) : 2;
        const std::vector<TapGroup> groups = level_tap_groups(h, R, nt, NS, fpad, ngroups, PP, max_sets);
        const int cgw = T.wg.cg;
        const int pc = (int)std::min<int64_t>(PP, std::max<int64_t>(1, cdiv((int64_t)PP * groups.size() * mt * ngroups / cgw,
                                                                        4 * (tc_sm_count() / cgw))));
        for (int sl = 0; sl < nslice; sl++) {
          const int kb0 = sl * slice_kb, kbn = std::min(slice_kb, nkb - kb0);
          for (int pc0 = 0; pc0 < PP; pc0 += pc)
            for (const TapGroup& tg : groups) {
              const int ring = tg.ring, nset = (int)tg.t.size();
              // source (activation) positions of the chunk that feed at least one tap of the group
              std::vector<int> qs;
              for (int q = pc0; q < std::min(PP, pc0 + pc); q++) {
                bool any = false;
                for (auto& tp : tg.t) any = any || tap_output_position(P, q, tp.first, tp.second) != 255;
                if (any) qs.push_back(q);
              }
              if (qs.empty()) continue;
              for (int g = 0; g < ngroups; g++) {
                const int s0 = g * spg, s1 = std::min(std::min(NS, s0 + spg), (R - ring) * nt);
                if (s1 <= s0) continue;
                const int ncols = (s1 - s0) * fpad;
                for (int im = 0; im < mt; im++) {  // adjacent M tiles are consecutive: CTA pairs share the gz columns
                  TcTile t = blank_tile();
                  t.seg_begin = (int)pb.segs.size();
                  t.a0_add = im * 128;
                  for (int q : qs) {
                    TcSeg s{};
                    s.a1 = kb0 * KB; s.a2 = q;
                    s.b0 = s0 * fpad; s.b1 = kb0 * KB;
                    s.nk = kbn; s.n_mma = r16(ncols); s.nb = (int)cdiv(s.n_mma, KB);
                    s.nsets = nset;
                    for (int j = 0; j < nset; j++) {
                      const int pos = tap_output_position(P, q, tg.t[j].first, tg.t[j].second);  // 255 >= PP: out of bounds
                      if (j == 0) s.b2 = pos; else s.b2x |= pos << (8 * (j - 1));
                    }
                    T.wg.b_rows = std::max(T.wg.b_rows, nset * wg_brows(s.n_mma, T.wg.cg, KB));
                    t.total_kb += s.nk;
                    pb.segs.push_back(s);
                  }
                  t.seg_count = (int)pb.segs.size() - t.seg_begin;
                  t.m_valid = std::min(128, Cin - im * 128);
                  t.ld_out = f;
                  t.ncb = nset * (s1 - s0);
                  for (int j = 0; j < nset; j++) {
                    const int dy = tg.t[j].first, dx = tg.t[j].second;
                    for (int sidx = s0; sidx < s1; sidx++) {
                      const int q = T.slot_q(sidx), k = 2 * q + 1;
                      TcColBlock& cb = t.cb[j * (s1 - s0) + (sidx - s0)];
                      cb.tcol = j * r16(ncols) + (sidx - s0) * fpad; cb.width = T.slot_w(sidx);
                      cb.out_off = L.w_off[q] + (int64_t)((dy + q) * k + (dx + q)) * Cin * f + (int64_t)im * 128 * f +
                                   (sidx % nt) * T.ft;
                    }
                  }
                  pb.tiles.push_back(t);
                }
              }
            }
        }
        T.wg.windowed = true;
        finish_launch(pb, T.wg);
      }
    } else {
      const int Ct = tin.C, PP = tin.PP, Cout = L.Cout;
      int ntn, nw;
      n_tiling(Cout, ntn, nw);
      // ---------------- forward ----------------
      if ((rc = map4(&T.fwd.tmA, a0, Ct, B, PP, tin.Cp, (uint64_t)B * tin.Cp, tin.plane_elems, KB, 128, false))) return rc;
      if ((rc = map4(&T.fwd.tmB, pk(T.wf_off), T.wf_ld, T.wf_rows, 1, T.wf_ld, (uint64_t)T.wf_rows * T.wf_ld,
                     S.pack_plane_elems, KB, nw / CG, false))) return rc;
      T.fwd.mn = false; T.fwd.cg = CG; T.fwd.b_rows = nw; T.fwd.bn = nw; T.fwd.tile0 = (int)pb.tiles.size();
      const int fseg0 = (int)pb.segs.size();
      for (int j = 0; j < ntn; j++)
        for (int pos = 0; pos < PP; pos++) {
          TcSeg s{};
          s.a2 = pos; s.b0 = pos * T.Kp; s.b1 = j * nw; s.nk = T.Kp / KB;
          s.n_mma = r16(std::min(nw, Cout - j * nw)); s.nb = 1;
          s.ks_last = tc_ks_last(Ct, s.nk, KB);
          pb.segs.push_back(s);
        }
      for (int bt2 = 0; bt2 < nbt; bt2 += CG)
        for (int j = 0; j < ntn; j++) {
          const size_t run = pb.tiles.size();
          for (int bt = bt2; bt < std::min(nbt, bt2 + CG); bt++) {
            TcTile t = blank_tile();
            t.seg_begin = fseg0 + j * PP; t.seg_count = PP; t.total_kb = PP * (T.Kp / KB);
            t.m_valid = (int)std::min<int64_t>(128, B - (int64_t)bt * 128);
            t.ncb = 1; t.ld_out = tout.Cp; t.stats_row = bt; t.a1_add = bt * 128;
            t.cb[0].out_off = (int64_t)bt * 128 * tout.Cp + j * nw;
            t.cb[0].width = std::min(nw, Cout - j * nw);
            t.cb[0].stats_col = j * nw;
            pb.tiles.push_back(t);
          }
          close_pair_run(pb.tiles, run, nbt * 128);
        }
      finish_launch(pb, T.fwd);
      T.stats_rows = nbt;
      // ---------------- dgrad ----------------
      if (need_dgrad) {
        if ((rc = map4(&T.dg.tmA, gz0, Cout, B, 1, T.Gp, (uint64_t)B * T.Gp, S.gz_plane_elems, KB, 128, false))) return rc;
        int dtn, dw;  // column tiles over the flattened tensor's channels
        n_tiling(Ct, dtn, dw);
        if ((rc = map4(&T.dg.tmB, pk(T.wd_off), T.wd_ld, T.wd_rows, 1, T.wd_ld, (uint64_t)T.wd_rows * T.wd_ld,
                       S.pack_plane_elems, KB, dw / CG, false))) return rc;
        T.dg.mn = false; T.dg.cg = CG; T.dg.b_rows = dw; T.dg.bn = dw; T.dg.tile0 = (int)pb.tiles.size();
        const int dseg0 = (int)pb.segs.size();
        for (int pos = 0; pos < PP; pos++)
          for (int jd = 0; jd < dtn; jd++) {
            TcSeg s{};
            s.b1 = pos * T.Cq + jd * dw; s.nk = T.wd_ld / KB; s.n_mma = r16(std::min(dw, Ct - jd * dw)); s.nb = 1;
            s.ks_last = tc_ks_last(Cout, s.nk, KB);
            pb.segs.push_back(s);
          }
        for (int bt2 = 0; bt2 < nbt; bt2 += CG)
          for (int pos = 0; pos < PP; pos++)
            for (int jd = 0; jd < dtn; jd++) {
              const size_t run = pb.tiles.size();
              for (int bt = bt2; bt < std::min(nbt, bt2 + CG); bt++) {
                TcTile t = blank_tile();
                t.seg_begin = dseg0 + pos * dtn + jd; t.seg_count = 1; t.total_kb = T.wd_ld / KB;
                t.m_valid = (int)std::min<int64_t>(128, B - (int64_t)bt * 128);
                t.ncb = 1; t.ld_out = tin.Cp; t.a1_add = bt * 128;
                t.cb[0].out_off = ((int64_t)pos * B + (int64_t)bt * 128) * tin.Cp + jd * dw;
                t.cb[0].width = std::min(dw, Ct - jd * dw);
                pb.tiles.push_back(t);
              }
              close_pair_run(pb.tiles, run, nbt * 128);
            }
        finish_launch(pb, T.dg);
        if ((rc = out_view(T.dg, tc_grad(m, L.in_t), Ct, rows_in, tin.Cp))) return rc;
      }
      // ---------------- wgrad ----------------
      {
        if ((rc = map4(&T.wg.tmA, a0, Ct, B, PP, tin.Cp, (uint64_t)B * tin.Cp, tin.plane_elems, KB, KB, true))) return rc;
        if ((rc = map4(&T.wg.tmB, gz0, Cout, B, 1, T.Gp, (uint64_t)B * T.Gp, S.gz_plane_elems, KB, KB, true))) return rc;
        T.wg.mn = true; T.wg.bn = 32; T.wg.b_rows = 0; T.wg.tile0 = (int)pb.tiles.size();
        const int mt = (int)cdiv(Ct, 128);
        T.wg.cg = wg_cg(mt);
        // K (= batch) pieces: the split that fills the persistent grid most evenly (the epilogue is atomic, so pieces
        // are free to land in any order); at least 8 K blocks per piece
        const int kblocks = (int)cdiv(B, KB);
        const double groups = (double)tc_sm_count() / T.wg.cg, units0 = (double)PP * ntn * mt / T.wg.cg;
        int ksplit = 1;
        double best = 1e30;
        for (int k = 1; k <= std::max(1, std::min(8, kblocks / 8)); k++) {
          const double per = units0 * k / groups, ratio = std::ceil(per) / per;
          if (ratio < best - 1e-9) { best = ratio; ksplit = k; }
        }
        const int kb_per = (int)cdiv(kblocks, ksplit);
        for (int kb0 = 0; kb0 < kblocks; kb0 += kb_per)
          for (int pos = 0; pos < PP; pos++)
            for (int j = 0; j < ntn; j++)
              for (int im = 0; im < mt; im++) {
                const int width = std::min(nw, Cout - j * nw);
                TcSeg s{};
                s.a1 = kb0 * KB; s.a2 = pos; s.b0 = j * nw; s.b1 = kb0 * KB;
                s.nk = std::min(kb_per, kblocks - kb0); s.n_mma = r16(width);
                s.nb = (int)cdiv(s.n_mma, KB);
                T.wg.b_rows = std::max(T.wg.b_rows, wg_brows(s.n_mma, T.wg.cg, KB));
                TcTile t = blank_tile();
                t.a0_add = im * 128;
                t.seg_begin = (int)pb.segs.size(); t.seg_count = 1; t.total_kb = s.nk;
                t.m_valid = std::min(128, Ct - im * 128); t.ncb = 1; t.ld_out = Cout;
                t.cb[0].out_off = L.w_off[0] + ((int64_t)pos * Ct + im * 128) * Cout + j * nw;
                t.cb[0].width = width;
                pb.segs.push_back(s);
                pb.tiles.push_back(t);
              }
        finish_launch(pb, T.wg);
      }
    }
  }
  if ((rc = tc_finalize_tiles(pb.tiles, pb.segs))) return rc;
  // upload
  if (pb.segs.size() > S.segs_cap) {
    if (S.segs_dev) cudaFree(S.segs_dev);
    S.segs_cap = pb.segs.size() + pb.segs.size() / 4;
    HYP_CUDA(cudaMalloc(&S.segs_dev, S.segs_cap * sizeof(TcSeg)));
  }
  if (pb.tiles.size() > S.tiles_cap) {
    if (S.tiles_dev) cudaFree(S.tiles_dev);
    S.tiles_cap = pb.tiles.size() + pb.tiles.size() / 4;
    HYP_CUDA(cudaMalloc(&S.tiles_dev, S.tiles_cap * sizeof(TcTile)));
  }
  HYP_CUDA(cudaDeviceSynchronize());  // earlier launches may still read the old tables
  HYP_CUDA(cudaMemcpy(S.segs_dev, pb.segs.data(), pb.segs.size() * sizeof(TcSeg), cudaMemcpyHostToDevice));
  HYP_CUDA(cudaMemcpy(S.tiles_dev, pb.tiles.data(), pb.tiles.size() * sizeof(TcTile), cudaMemcpyHostToDevice));
  S.planned_B = (int)B;
  return HYP_OK;
}

// out_scale / out_scale_ptr: what the epilogue multiplies the accumulators by (undoes the operand scales of the 16-bit
// formats: 1 / w_scale in forward, 1 / (w_scale * gz scale) in dgrad, 1 / gz scale in wgrad)
static int tc_run(hyp_model& m, const TcLaunch& l, float* out, float* stats, int stats_ld, int epi, const char* tag,
                  double flops, cudaStream_t st, const char* scope = nullptr, float out_scale = 1.f,
                  const float* out_scale_ptr = nullptr) {
  if (l.ntiles == 0) return HYP_OK;
  TcState& S = *m.tc;
  TcParams p{};
  p.segs = S.segs_dev; p.tiles = S.tiles_dev + l.tile0; p.out = out; p.stats = stats; p.stats_ld = stats_ld;
  p.op = S.op; p.out_scale = out_scale; p.out_scale_ptr = out_scale_ptr;
  p.epi = epi; p.b_rows = l.b_rows; p.bn = l.bn; p.chunk_kb = 0; p.stages = 0;
  static const bool per_layer = getenv("HYP_PROF_LAYERS") != nullptr;
  std::string full = tag;
  if (per_layer && scope) full += std::string("/") + scope;
  static const bool timing_on = getenv("HYP_TC_TIMING") != nullptr;
  if (timing_on && !per_layer && scope) full += std::string("/") + scope;
  g_tc_timing_tag = full.c_str();
  g_prof.begin(st, full.c_str(), flops, 0.0);
  // dgrad accumulations (EPI_ATOMIC) leave through the TMA unit as f32 add reductions; HYP_TC_TMA_STORE=0: red.global
  static const bool tma_off = getenv("HYP_TC_TMA_STORE") && getenv("HYP_TC_TMA_STORE")[0] == '0';
  const CUtensorMap* tmO = (l.tma_out && !tma_off && epi == EPI_ATOMIC) ? &l.tmO : nullptr;
  p.out_cols = l.out_cols; p.out_rows = l.out_rows;
  const int rc = l.mn ? (l.cg == 2 ? launch_tc<true, 2>(l.tmA, l.tmB, p, l.ntiles, st, tmO)
                                   : launch_tc<true, 1>(l.tmA, l.tmB, p, l.ntiles, st, tmO))
                      : (l.cg == 2 ? launch_tc<false, 2>(l.tmA, l.tmB, p, l.ntiles, st, tmO)
                                   : launch_tc<false, 1>(l.tmA, l.tmB, p, l.ntiles, st, tmO));
  g_prof.end(st);
  return rc;
}

#define TC_PROF(name, bytes, launch)              \
  do {                                            \
    g_prof.begin(st, name, 0.0, (double)(bytes)); \
    launch;                                       \
    g_prof.end(st);                               \
    HYP_LAUNCHED();                               \
  } while (0)

static inline int tc_grid(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(total, 256), 148 * 16)); }

static int tc_forward(hyp_model& m, const float* x, int64_t B, bool training, bool update_moving, uint64_t seed,
                      cudaStream_t st) {
  TcState& S = *m.tc;
  int rc = tc_plan(m, B);
  if (rc) return rc;
  const int nl = training ? (int)m.layers.size() : m.last_eval_layer + 1;
  float* pack0 = reinterpret_cast<float*>(m.ws + S.pack_off);
  {
    for (size_t t = 0; t < m.tensors.size(); t++) {
      if (!m.tensors[t].external) continue;
      const Tensor& tn = m.tensors[t];
      const TcTensor& tx = S.tt[t];
      TC_PROF("tc_prep_input_kernel", 12.0 * B * tx.PP * tx.C,
              (tc_prep_input_kernel<<<tc_grid(B * tx.PP * cdiv(tx.C, 4) * 8), 256, 0, st>>>(
                  x, (int)B, m.d.patch, m.d.channels, tn.x_c0, tn.x_crop, tn.P, tx.C, tx.Cp, tc_plane0(m, (int)t),
                  tc_plane1(m, (int)t), S.op, tx.plane_elems)));
    }
    TC_PROF("tc_pack_weights_kernel", 12.0 * S.pack_plane_elems,
            (tc_pack_weights_kernel<<<(unsigned)S.job_blocks, 256, 0, st>>>(
                S.jobs_dev, S.job_tiles_dev, (int)S.jobs.size(), m.params, pack0, pack0 + S.pack_plane_elems, S.op,
                S.pack_plane_elems, S.w_scale)));
  }
  const float keep_prob = m.keep_prob;
  float* part = reinterpret_cast<float*>(m.ws + S.part_off);
  for (int li = 0; li < nl; li++) {
    Layer& L = m.layers[li];
    TcLayer& T = S.tl[li];
    const TcTensor& tout = S.tt[L.out_t];
    const int64_t rows = B * tout.PP;
    float* Z = reinterpret_cast<float*>(m.ws + T.z_off);
    rc = tc_run(m, T.fwd, Z, (training && !L.bias_mode) ? part : nullptr, L.Cout, L.share == 2 ? EPI_ACCUM : EPI_STORE,
                "tc_gemm_kernel/fwd", layer_flops(L, B), st, L.scope.c_str(), 1.f / S.w_scale);
    if (rc) return rc;
    if (L.share == 1) continue;  // the second part adds its product, then bias / activation run once
    float* mean = reinterpret_cast<float*>(m.ws + T.mean_off);
    float* rstd = reinterpret_cast<float*>(m.ws + T.rstd_off);
    if (L.bias_mode) {
      // no normaliser: mean = 0, rstd = 1 (set at bind), beta = the layer's biases
    } else if (training) {
      TC_PROF("tc_bn_finalize_kernel", 8.0 * T.stats_rows * L.Cout,
              (tc_bn_finalize8_kernel<<<(unsigned)cdiv(L.Cout, 8), 8 * FIN_LANES, 0, st>>>(
                  part, T.stats_rows, L.Cout, L.Cout, (double)rows, m.d.bn_eps, m.d.bn_decay, m.state + L.mm_off,
                  m.state + L.mm_off + L.Cout, mean, rstd, update_moving ? 1 : 0)));
    } else {
      TC_PROF("bn_finalize_kernel", 32.0 * L.Cout,
              (bn_finalize_kernel<<<(unsigned)cdiv(L.Cout, 128), 128, 0, st>>>(
                  nullptr, L.Cout, (double)rows, m.d.bn_eps, m.d.bn_decay, m.state + L.mm_off,
                  m.state + L.mm_off + L.Cout, mean, rstd, 0, 0)));
    }
    TcApplyArgs p{};
    p.z = Z; p.ldz = tout.Cp; p.mean = mean; p.rstd = rstd; p.beta = m.params + L.beta_off;
    p.hi = tc_plane0(m, L.out_t); p.lo = tc_plane1(m, L.out_t); p.ldo = tout.Cp;
    p.op = S.op; p.op_plane = tout.plane_elems;
    p.rows = rows; p.C = L.Cout; p.act = L.act; p.alpha = m.d.lrelu_alpha;
    p.keep = (L.dropout && training) ? keep_prob : 1.f;
    p.seed = seed; p.stream_id = L.drop_stream;
    if (L.res.size() > 0) {
      const TcTensor& rs = S.tt[L.res[0].src];
      p.res0 = tc_plane0(m, L.res[0].src); p.idx0 = L.res[0].idx; p.ld0 = rs.Cp; p.has0 = 1;
      p.pat0 = L.res[0].pattern == 2 ? 2 : (L.res[0].pattern == 3 && L.res[0].step == 2 ? 3 : 0);
      p.res0_op = reinterpret_cast<const uint16_t*>(tc_plane1(m, L.res[0].src)); p.res0_plane = rs.plane_elems;
    }
    if (L.res.size() > 1) {
      const TcTensor& rs = S.tt[L.res[1].src];
      p.res1 = tc_plane0(m, L.res[1].src); p.idx1 = L.res[1].idx; p.ld1 = rs.Cp; p.has1 = 1;
      p.pat1 = L.res[1].pattern == 2 ? 2 : (L.res[1].pattern == 3 && L.res[1].step == 2 ? 3 : 0);
      p.res1_op = reinterpret_cast<const uint16_t*>(tc_plane1(m, L.res[1].src)); p.res1_plane = rs.plane_elems;
    }
    const double bytes = 4.0 * rows * L.Cout * ((p.hi ? 3 : 2) + L.res.size());
    // (a 2-D row-lane form with the channel statistics held in registers measured 7-15 % slower, twice: the flat float4
    // walk keeps more rows in flight, and the kernel is bound by memory latency, not by its instruction count)
    TC_PROF("tc_bn_apply_kernel", bytes, (tc_bn_apply_kernel<4><<<tc_grid(rows * cdiv(L.Cout, 4)), 256, 0, st>>>(p)));
    if (L.lrn) {
      TC_PROF("tc_lrn_fwd_kernel", 16.0 * rows * L.Cout,
              (tc_lrn_fwd_kernel<<<tc_grid(rows * 32), 256, 8 * L.Cout * sizeof(float), st>>>(
                  p.hi, p.lo, reinterpret_cast<float*>(m.ws + T.u_off), tout.Cp, L.Cout, rows, S.op, tout.plane_elems)));
    }
  }
  return HYP_OK;
}

static int tc_backward(hyp_model& m, const uint8_t* labels, int64_t B, float* loss_out, cudaStream_t st) {
  TcState& S = *m.tc;
  const hyp_model_desc& d = m.d;
  const int nl = (int)m.layers.size();
  std::vector<char> ginit(m.tensors.size(), 0);
  HYP_CUDA(cudaMemsetAsync(m.grads, 0, (size_t)m.n_params * sizeof(float), st));
  HYP_CUDA(cudaMemsetAsync(m.ws + S.mse_off, 0, 256, st));
  if (S.op == OP_F16X3) HYP_CUDA(cudaMemsetAsync(m.ws + S.gzs_off, 0, S.gzs_bytes, st));  // |gz| bounds: atomic maxima
  float* ce = reinterpret_cast<float*>(m.ws + S.ce_off);
  double* mse_acc = reinterpret_cast<double*>(m.ws + S.mse_off);
  const TcTensor& tlog = S.tt[m.logits_t];
  TC_PROF("tc_ce_loss_kernel", 8.0 * B * d.classes,
          (tc_ce_loss_kernel<<<(unsigned)cdiv(B * 32, 256), 256, 0, st>>>(tc_plane0(m, m.logits_t), tlog.Cp, labels, B,
                                                                          d.classes, ce, tc_grad(m, m.logits_t), tlog.Cp,
                                                                          1.f / (float)B)));
  ginit[m.logits_t] = 1;
  int64_t D = 1;
  if (m.recon_t >= 0) {  // reconstruction branch (HYPELCNN training graph)
    const TcTensor& trec = S.tt[m.recon_t];
    const TcTensor& tx = S.tt[0];
    D = (int64_t)tx.PP * tx.C;
    TC_PROF("tc_mse_kernel", 12.0 * B * D,
            (tc_mse_kernel<<<tc_grid(B * D), 256, 0, st>>>(tc_plane0(m, m.recon_t), trec.Cp, tc_plane0(m, 0), tx.Cp, (int)B,
                                                           tx.PP, tx.C, mse_acc, tc_grad(m, m.recon_t), trec.Cp,
                                                           1.f / (float)(B * D))));
    ginit[m.recon_t] = 1;
  }
  loss_finalize_kernel<<<1, 256, 0, st>>>(ce, B, m.recon_t >= 0 ? mse_acc : nullptr, (double)(B * D), loss_out, nullptr);
  HYP_LAUNCHED();

  const float keep_prob = m.keep_prob;
  float* gz0 = reinterpret_cast<float*>(m.ws + S.gz_off);  // 16-bit formats: the start of the operand planes
  float* gz1 = gz0 + S.gz_plane_elems;                     // (3xTF32 only)
  float* bpart = reinterpret_cast<float*>(m.ws + S.bpart_off);
  for (int li = nl - 1; li >= 0; li--) {
    Layer& L = m.layers[li];
    TcLayer& T = S.tl[li];
    const TcTensor& tout = S.tt[L.out_t];
    const int64_t rows = B * tout.PP;
    if (!ginit[L.out_t]) return fail(HYP_E_STATE, "backward: no gradient reached " + L.scope);
    if (L.lrn) {  // the gradient arrives w.r.t. the normalised tensor: rewrite it w.r.t. the LRN input, in place
      TC_PROF("tc_lrn_bwd_kernel", 12.0 * rows * L.Cout,
              (tc_lrn_bwd_kernel<<<tc_grid(rows * 32), 256, 8 * 3 * L.Cout * sizeof(float), st>>>(
                  tc_grad(m, L.out_t), reinterpret_cast<const float*>(m.ws + T.u_off), tout.Cp, tout.Cp, L.Cout, rows)));
    }
    if (L.share != 1) {  // share 1: gz of the second part (processed just before) is still in the scratch planes
    TcBnBwdArgs p{};
    p.gout = tc_grad(m, L.out_t); p.ldg = tout.Cp;
    p.z = reinterpret_cast<float*>(m.ws + T.z_off); p.ldz = tout.Cp;
    p.mean = reinterpret_cast<float*>(m.ws + T.mean_off);
    p.rstd = reinterpret_cast<float*>(m.ws + T.rstd_off);
    p.beta = m.params + L.beta_off;
    p.rows = rows; p.C = L.Cout; p.act = L.act; p.alpha = d.lrelu_alpha;
    p.keep = L.dropout ? keep_prob : 1.f;
    p.seed = m.last_seed; p.stream_id = L.drop_stream;
    p.part = bpart;
    float* s1 = reinterpret_cast<float*>(m.ws + T.s1_off);
    float* s2 = reinterpret_cast<float*>(m.ws + T.s2_off);
    p.s1 = s1; p.s2 = s2; p.gz_hi = gz0; p.gz_lo = gz1; p.ldgz = T.Gp;
    p.op = S.op; p.gz_plane = S.gz_plane_elems;
    float* gzs = reinterpret_cast<float*>(m.ws + T.gzs_off);
    if (S.op == OP_F16X3) { p.gmax_bits = reinterpret_cast<unsigned int*>(gzs); p.gz_scale_out = gzs + 1; }
    p.gcols = T.kind == 1 ? T.Gp : L.Cout;
    p.fpad = T.kind == 1 ? T.fpad : 0; p.f = T.f; p.R = T.R; p.nt = T.nt; p.ft = T.ft;
    // Pass order and sweep directions over this layer's gradient tensor: residual pushes (alternating, the last one
    // descending), statistics ascending, gz descending -- each pass starts on the rows the previous one left in L2, and
    // the dgrad / wgrad GEMMs that follow start at row 0, where gz was written last.
    const int zz = tc_zigzag();
    int n_push = 0;
    for (const Resid& r : L.res) n_push += m.tensors[r.src].needs_grad ? 1 : 0;
    for (const Resid& r : L.res) {
      const Tensor& src = m.tensors[r.src];
      if (!src.needs_grad) continue;
      const int resid_rev = zz ? (n_push--) & 1 : 0;
      const EwGrid gs = ew_grid2(src.C, rows);
      const double bytes = 4.0 * rows * (L.Cout + (ginit[r.src] ? 2.0 : 1.0) * src.C);
      const int mode = r.identity ? 0 : (r.pattern == 2 ? 2 : (r.pattern == 3 ? 3 : 1));
#define TC_RESID_LAUNCH(TXV, MODEV)                                                                              \
  tc_resid_bwd_v4_kernel<TXV, MODEV><<<dim3(gs.gx, gs.rblocks), 256, 0, st>>>(                                  \
      p.gout, tout.Cp, tc_grad(m, r.src), S.tt[r.src].Cp, src.C, r.lo, r.hi, rows, ginit[r.src], resid_rev, r.step)
#define TC_RESID_TX(MODEV)                                                                                       \
  do {                                                                                                           \
    if (gs.TX == 32) TC_RESID_LAUNCH(32, MODEV); else if (gs.TX == 16) TC_RESID_LAUNCH(16, MODEV); else TC_RESID_LAUNCH(8, MODEV); \
  } while (0)
      g_prof.begin(st, "tc_resid_bwd_kernel", 0.0, bytes);
      if (mode == 0) TC_RESID_TX(0); else if (mode == 2) TC_RESID_TX(2); else if (mode == 3) TC_RESID_TX(3); else TC_RESID_TX(1);
      g_prof.end(st);
      HYP_LAUNCHED();
#undef TC_RESID_TX
#undef TC_RESID_LAUNCH
      ginit[r.src] = 1;
    }
    const EwGrid gr = ew_grid2(L.Cout, rows);
    const bool drop = p.keep < 1.f;
#define TC_REDUCE_LAUNCH(TXV)                                                                                         \
  do {                                                                                                                \
    if (drop) tc_bn_bwd_reduce_v4_kernel<TXV, true><<<dim3(gr.gx, gr.rblocks), 256, 0, st>>>(p, 0);                   \
    else tc_bn_bwd_reduce_v4_kernel<TXV, false><<<dim3(gr.gx, gr.rblocks), 256, 0, st>>>(p, 0);                       \
  } while (0)
    g_prof.begin(st, "tc_bn_bwd_reduce_kernel", 0.0, 8.0 * rows * L.Cout);
    if (gr.TX == 32) TC_REDUCE_LAUNCH(32); else if (gr.TX == 16) TC_REDUCE_LAUNCH(16); else TC_REDUCE_LAUNCH(8);
    g_prof.end(st);
    HYP_LAUNCHED();
#undef TC_REDUCE_LAUNCH
    tc_bn_bwd_finalize8_kernel<<<(unsigned)cdiv(L.Cout, 8), 8 * FIN_LANES, 0, st>>>(bpart, gr.rblocks, L.Cout, (double)rows, s1, s2,
                                                                         m.grads + L.beta_off, L.bias_mode ? 1 : 0, p.rstd,
                                                                         S.op == OP_F16X3 ? reinterpret_cast<unsigned int*>(gzs) : nullptr);
    HYP_LAUNCHED();
    {
      const EwGrid ga = ew_grid2(p.gcols, rows);
      const bool shift = p.fpad != 0 && (p.f % 4 != 0 || p.ft % 4 != 0);
#define TC_APPLY_LAUNCH(TXV)                                                                                          \
  do {                                                                                                                \
    if (drop) tc_bn_bwd_apply_v4_kernel<TXV, false, true><<<dim3(ga.gx, ga.rblocks), 256, 0, st>>>(p, zz);            \
    else if (shift) tc_bn_bwd_apply_v4_kernel<TXV, true, false><<<dim3(ga.gx, ga.rblocks), 256, 0, st>>>(p, zz);      \
    else tc_bn_bwd_apply_v4_kernel<TXV, false, false><<<dim3(ga.gx, ga.rblocks), 256, 0, st>>>(p, zz);                \
  } while (0)
      g_prof.begin(st, "tc_bn_bwd_apply_kernel", 0.0, 4.0 * rows * (2.0 * L.Cout + 2.0 * p.gcols));
      if (ga.TX == 32) TC_APPLY_LAUNCH(32); else if (ga.TX == 16) TC_APPLY_LAUNCH(16); else TC_APPLY_LAUNCH(8);
      g_prof.end(st);
      HYP_LAUNCHED();
#undef TC_APPLY_LAUNCH
    }
    }
    // (a two-input FC's first part reuses the gz planes of the second part, processed just before: same scale slot)
    const float* gz_inv = S.op == OP_F16X3 ? reinterpret_cast<const float*>(m.ws + S.tl[L.share == 1 ? li + 1 : li].gzs_off) + 2 : nullptr;
    int rc = tc_run(m, T.wg, m.grads, nullptr, 0, EPI_ATOMIC, "tc_gemm_kernel/wgrad", layer_flops(L, B), st, L.scope.c_str(),
                    1.f, gz_inv);
    if (rc) return rc;
    if (m.tensors[L.in_t].needs_grad) {
      // A gradient tensor that already holds contributions is accumulated into with fire-and-forget vector reductions
      // (red.global.add.v4.f32: every element belongs to exactly one tile of the launch, so the result is the same as
      // load + add + store, without the load's latency inside the epilogue).  HYP_DGRAD_RMW=1: load + add + store.
      static const bool dgrad_rmw = getenv("HYP_DGRAD_RMW") && getenv("HYP_DGRAD_RMW")[0] == '1';
      rc = tc_run(m, T.dg, tc_grad(m, L.in_t), nullptr, 0, ginit[L.in_t] ? (dgrad_rmw ? EPI_ACCUM : EPI_ATOMIC) : EPI_STORE, "tc_gemm_kernel/dgrad",
                  layer_flops(L, B), st, L.scope.c_str(), 1.f / S.w_scale, gz_inv);
      if (rc) return rc;
      ginit[L.in_t] = 1;
    }
    grad_notify(m, li, st);
  }
  return HYP_OK;
}

}  // namespace tc
}  // namespace hyp
