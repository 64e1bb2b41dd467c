// hypelcnn_b200 engine: builds the HYPELCNN layer plan from a descriptor, lays out the
// caller-owned parameter / state / workspace buffers, and drives the CUDA kernels for
// forward, loss, backward, Adam, metrics and the patch gather.  C ABI in
// include/hypelcnn_b200.h.  No CPU fallback: every compute entry point needs a CUDA device.
//
// Reference behaviour restated here (never copied): nnmodel/HYPELCNNModel.py:34-183,
// common/common_nn_ops.py:208-240,546-564.
#include <cmath>
#include <map>
#include <memory>
#include <vector>

#include "hyp_common.cuh"
#include "hyp_gemm_simt.cuh"
#include "hyp_kernels.cuh"
#include "hyp_tc.cuh"

namespace hyp {
thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;
thread_local Profiler g_prof;

struct Variable {
  std::string name;
  int kind;  // 0 weights 1 beta 2 moving_mean 3 moving_variance
  int64_t offset;
  int shape[4];
  int rank;
};

struct Resid {
  int src;          // tensor id
  bool identity;
  int* idx = nullptr;  // device [Cout]   out channel -> src channel
  int* lo = nullptr;   // device [Csrc]   src channel -> [lo, hi) out channels (idx is monotone)
  int* hi = nullptr;
  int pattern = 0;     // 0 generic; 2: tf.repeat by 2 (idx[j] = j / 2); 3: strided gather (idx[j] = step * j)
  int step = 0;
};

struct Tensor {
  std::string name;
  int C;              // channels per row
  int rows_per_sample;
  size_t a_off = 0;   // byte offsets in workspace
  size_t g_off = 0;
  bool external = false;  // a view of the input x: channels [x_c0, x_c0 + C), spatial crop of x_crop pixels per side
  bool needs_grad = true;
  int P = 1;              // spatial edge (rows_per_sample = P * P)
  int x_c0 = 0, x_crop = 0;
};

struct Layer {
  std::string scope;      // level layers: "connector_0" (individual convs are named per kernel)
  bool is_fc;
  int P;                  // spatial edge for shifts (1 for FC)
  int in_t, out_t;
  int Cin, Cout;
  int rows_per_sample;
  std::vector<int> ksizes;     // conv kernels of a level (1x1 conv: {1}); FC: {1}
  int f;                       // channels per kernel
  std::vector<std::string> kscopes;
  int act;
  bool dropout;
  std::vector<Resid> res;
  // parameter offsets (elements)
  std::vector<int64_t> w_off;  // per kernel
  int64_t beta_off, mm_off;    // beta in params; moving mean/var in state (mv = mm + Cout)
  // workspace byte offsets
  size_t z_off, wt_off, stats_off, bstats_off, mean_off, rstd_off, s1_off, s2_off;
  int seg_begin, seg_count;
  std::vector<int> kseg_begin;  // first segment of kernel j
  uint32_t drop_stream;
  bool bias_mode = false;       // slim conv2d / fully_connected without normalizer_fn: z + biases instead of BatchNorm
  // an FC fed by the concatenation of two flattened tensors (DUALCNN fc1) is two layers that share one
  // pre-activation buffer and one weight variable: share = 1 first part (GEMM only), 2 second part (accumulates
  // into the first part's z, then bias / activation / dropout)
  int share = 0;
  bool lrn = false;  // tf.nn.local_response_normalization (TF defaults) after the activation (CONCNNModel.py:37,41)
};

namespace tc {
struct TcState;
}
}  // namespace hyp

using namespace hyp;

struct hyp_model {
  hyp_model_desc d;
  std::vector<Tensor> tensors;
  std::vector<Layer> layers;
  std::vector<Variable> vars;
  std::map<std::string, int> tensor_by_name;
  int64_t n_params = 0, n_state = 0;
  size_t ws_bytes = 0;
  size_t gz_off = 0, ce_off = 0, mse_off = 0, stats_region_off = 0, stats_region_bytes = 0;
  size_t bstats_region_off = 0, bstats_region_bytes = 0;
  int logits_t = -1, recon_t = -1, last_eval_layer = -1;
  float keep_prob = 1.f;  // dropout keep probability (HYPELCNN: 1 - drop_out_ratio; DUALCNN: drop_out_ratio itself)
  // library-owned device metadata
  Seg* segs_dev = nullptr;
  std::vector<Seg> segs_host;
  TransposeJob* tjobs_dev = nullptr;
  int4* tmap_dev = nullptr;
  int n_ttiles = 0;
  std::vector<void*> owned;
  // bound buffers
  float *params = nullptr, *grads = nullptr, *state = nullptr;
  char* ws = nullptr;
  bool segs_bound = false;
  // call state
  int64_t last_B = -1;
  bool last_training = false;
  uint64_t last_seed = 0;
  const float* last_x = nullptr;
  hyp::tc::TcState* tc = nullptr;  // tensor-core engine state (HYP_PRECISION_3XTF32)
  // gradient-ready notification (hyp_model_set_grad_notify)
  struct Notify { cudaEvent_t event; int64_t offset; };
  std::vector<Notify> notify;
};

namespace hyp {

static std::vector<int> scale_index(int cin, int cout) {
  // common/common_nn_ops.py:546-564 — double arithmetic and banker's rounding like Python
  const double ratio = (double)cin / (double)cout;
  const double inv = 1.0 / ratio;
  std::vector<int> idx;
  if (std::floor(inv) == inv) {
    const int rep = (int)inv;
    for (int j = 0; j < cin * rep; j++) idx.push_back(j / rep);
  } else {
    for (int j = 0; j < cout; j++) {
      const int t = (int)std::nearbyint((double)j * ratio);  // FE_TONEAREST = half to even
      idx.push_back(t < cin - 1 ? t : cin - 1);
    }
  }
  return idx;
}

struct Builder {
  hyp_model& m;
  int seg_total = 0;
  explicit Builder(hyp_model& mm) : m(mm) {}

  int add_tensor(const std::string& name, int C, int rps, bool external = false) {
    Tensor t;
    t.name = name;
    t.C = C;
    t.rows_per_sample = rps;
    t.P = (int)std::lround(std::sqrt((double)rps));
    t.external = external;
    t.needs_grad = !external;
    m.tensors.push_back(t);
    m.tensor_by_name[name] = (int)m.tensors.size() - 1;
    return (int)m.tensors.size() - 1;
  }

  int add_resid(Layer& L, int src) {
    const int cs = m.tensors[src].C;
    std::vector<int> idx = scale_index(cs, L.Cout);
    if ((int)idx.size() != L.Cout) return -1;
    Resid r;
    r.src = src;
    r.identity = (cs == L.Cout);
    if (r.identity)
      for (int j = 0; j < L.Cout; j++) r.identity = r.identity && idx[j] == j;
    if (!r.identity) {
      std::vector<int> lo(cs, 0), hi(cs, 0);
      for (int c = 0; c < cs; c++) { lo[c] = L.Cout; hi[c] = 0; }
      for (int j = 0; j < L.Cout; j++) {
        lo[idx[j]] = std::min(lo[idx[j]], j);
        hi[idx[j]] = std::max(hi[idx[j]], j + 1);
      }
      for (int c = 0; c < cs; c++)
        if (hi[c] == 0) lo[c] = 0;
      // monotone index table => each source channel's consumers are contiguous
      for (int c = 0; c < cs; c++)
        for (int j = lo[c]; j < hi[c]; j++)
          if (idx[j] != c) return -1;
      bool rep2 = L.Cout == 2 * cs, strided = cs % L.Cout == 0 && cs > L.Cout;
      for (int j = 0; j < L.Cout; j++) {
        rep2 = rep2 && idx[j] == j / 2;
        strided = strided && idx[j] == (cs / L.Cout) * j;
      }
      if (rep2 && cs % 4 == 0) r.pattern = 2;
      else if (strided) { r.pattern = 3; r.step = cs / L.Cout; }
      r.idx = upload(idx);
      r.lo = upload(lo);
      r.hi = upload(hi);
      if (!r.idx || !r.lo || !r.hi) return -2;
    }
    L.res.push_back(r);
    return 0;
  }

  int* upload(const std::vector<int>& v) {
    int* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    m.owned.push_back(d);
    return d;
  }

  // conv level (one or several square kernels sharing the input, outputs concatenated) or FC
  Layer& add_layer(const std::string& scope, bool fc, int P, int in_t, const std::string& out_name, int f,
                   const std::vector<int>& ks, const std::vector<std::string>& kscopes, int act, bool dropout) {
    Layer L;
    L.scope = scope;
    L.is_fc = fc;
    L.P = fc ? 1 : P;
    L.in_t = in_t;
    L.Cin = fc ? m.tensors[in_t].C * m.tensors[in_t].rows_per_sample : m.tensors[in_t].C;
    L.f = f;
    L.ksizes = ks;
    L.kscopes = kscopes;
    L.Cout = f * (int)ks.size();
    L.rows_per_sample = fc ? 1 : m.tensors[in_t].rows_per_sample;
    L.act = act;
    L.dropout = dropout;
    L.out_t = add_tensor(out_name, L.Cout, L.rows_per_sample);
    L.drop_stream = (uint32_t)m.layers.size() + 1;
    L.seg_begin = seg_total;
    for (int k : ks) {
      L.kseg_begin.push_back(seg_total);
      seg_total += k * k;
    }
    L.seg_count = seg_total - L.seg_begin;
    m.layers.push_back(L);
    return m.layers.back();
  }
};

static int build_hypelcnn(hyp_model& m) {
  const hyp_model_desc& d = m.d;
  Builder b(m);
  const int P = d.patch, PP = P * P;
  const bool res = d.use_residual != 0;
  m.keep_prob = 1.f - d.drop_out_ratio;  // HYPELCNNModel.py:123
  int cur = b.add_tensor("x", d.channels, PP, true);
  const int x_t = cur;
  // spectral encoder / decoder (HYPELCNNModel.py:146-164, :54-64)
  for (int enc = 1; enc >= 0; enc--) {
    const int block_in = cur;
    for (int i = 0; i < d.spectral_levels; i++) {
      const int cout = enc ? d.filter_count >> ((d.spectral_levels - 1) - i) : d.filter_count >> i;
      if (cout <= 0) return fail(HYP_E_INVALID, "filter_count too small for spectral levels");
      const std::string name = std::string(enc ? "conv_enc_" : "conv_dec_") + std::to_string(i);
      Layer& L = b.add_layer(name, false, P, cur, name, cout, {1}, {name}, ACT_LRELU, false);
      if (res) {
        if (b.add_resid(L, cur)) return fail(HYP_E_INVALID, "residual table for " + name);
        if (i == d.spectral_levels - 1 && b.add_resid(L, block_in))
          return fail(HYP_E_INVALID, "block residual table for " + name);
      }
      cur = L.out_t;
    }
  }
  const int net2 = cur;
  // spatial blocks (HYPELCNNModel.py:128-143, :167-183)
  const int lf = m.tensors[net2].C / 2;
  for (int i = 0; i < d.spatial_levels; i++) {
    const int f = lf >> i;
    if (f <= 0) return fail(HYP_E_INVALID, "filter_count too small for spatial levels");
    std::vector<int> ks;
    std::vector<std::string> kn;
    const std::string lvl = "connector_" + std::to_string(i);
    for (int k = 1; k <= P; k += 2) {
      ks.push_back(k);
      kn.push_back(lvl + "_conv" + std::to_string(k) + "x" + std::to_string(k));
    }
    Layer& L = b.add_layer(lvl, false, P, cur, lvl, f, ks, kn, ACT_LRELU, false);
    if (res && b.add_resid(L, cur)) return fail(HYP_E_INVALID, "residual table for " + lvl);
    const int lvl_t = L.out_t;
    const std::string cn = "connector_conv_" + std::to_string(i);
    Layer& Cn = b.add_layer(cn, false, P, lvl_t, cn, m.tensors[lvl_t].C, {1}, {cn}, ACT_LRELU, false);
    if (res) {
      if (b.add_resid(Cn, lvl_t)) return fail(HYP_E_INVALID, "residual table for " + cn);
      if (i == d.spatial_levels - 1 && b.add_resid(Cn, net2)) return fail(HYP_E_INVALID, "net3 residual table");
    }
    cur = Cn.out_t;
  }
  // FC block (HYPELCNNModel.py:115-125): stages = floor(log_deg(flat / classes))
  const int flat = PP * m.tensors[cur].C;
  if (d.degradation < 2) return fail(HYP_E_INVALID, "degradation_coeff must be >= 2");
  const int stages = (int)std::floor(std::log((double)flat / (double)d.classes) / std::log((double)d.degradation));
  int size = flat;
  for (int i = 0; i < stages - 1; i++) {
    size = size / d.degradation;
    const std::string name = "fc_" + std::to_string(i);
    Layer& L = b.add_layer(name, true, 1, cur, name, size, {1}, {name}, ACT_LRELU, true);
    cur = L.out_t;
  }
  {
    Layer& L = b.add_layer("fc_final", true, 1, cur, "fc_final", d.classes, {1}, {"fc_final"}, ACT_NONE, false);
    cur = L.out_t;
    m.logits_t = cur;
    m.last_eval_layer = (int)m.layers.size() - 1;
  }
  // decoder, training graph only (HYPELCNNModel.py:84-94)
  int mult = 3;
  for (int i = 1; i <= 3; i++, mult *= 3) {
    const std::string name = "image_gen_net_" + std::to_string(i);
    Layer& L = b.add_layer(name, true, 1, cur, name, d.classes * mult, {1}, {name}, ACT_LRELU, false);
    cur = L.out_t;
  }
  {
    Layer& L = b.add_layer("image_gen_net_4", true, 1, cur, "image_gen_net_4", PP * d.channels, {1},
                           {"image_gen_net_4"}, ACT_SIGMOID, false);
    m.recon_t = L.out_t;
  }
  (void)x_t;
  return HYP_OK;
}

// DUALCNNModel (nnmodel/DUALCNNModel.py:11-104): the HSI bands (window cropped by hs_lidar_diff) and the LiDAR
// channel run through separate stacks of multi-kernel levels + 1x1 connectors, are flattened, concatenated and
// classified by four FCs.  No BatchNorm: every conv / FC has a bias; LeakyReLU; dropout keep_prob = drop_out_ratio
// (slim dropout's second positional argument, DUALCNNModel.py:49).
static int build_dualcnn(hyp_model& m) {
  const hyp_model_desc& d = m.d;
  Builder b(m);
  const int P0 = d.patch, diff = d.reserved;
  if (d.channels < 2) return fail(HYP_E_INVALID, "DUALCNN needs at least one HSI band and the LiDAR channel");
  const bool crop = P0 > 1;  // DUALCNNModel.py:24-26
  const int Ph = crop ? P0 - 2 * diff : P0;
  if (diff < 0 || Ph < 1) return fail(HYP_E_INVALID, "hs_lidar_diff leaves no HSI window");
  if (d.filter_count < 32) return fail(HYP_E_INVALID, "filter_count must be >= 32 (level8 has filter_count / 32 filters)");
  m.keep_prob = d.drop_out_ratio;
  const int hs = b.add_tensor("x_hs", d.channels - 1, Ph * Ph, true);
  m.tensors[hs].x_c0 = 0; m.tensors[hs].x_crop = crop ? diff : 0;
  const int li_t = b.add_tensor("x_lidar", 1, P0 * P0, true);
  m.tensors[li_t].x_c0 = d.channels - 1; m.tensors[li_t].x_crop = 0;
  auto level_and_connector = [&](int cur, int P, int f, const std::string& lname, const std::string& cname) {
    std::vector<int> ks;
    std::vector<std::string> kn;
    for (int k = 1; k <= P; k += 2) { ks.push_back(k); kn.push_back(lname + "_conv" + std::to_string(k) + "x" + std::to_string(k)); }
    Layer& L = b.add_layer(lname, false, P, cur, lname, f, ks, kn, ACT_LRELU, false);
    L.bias_mode = true;
    const int lvl_t = L.out_t;
    Layer& Cn = b.add_layer(cname, false, P, lvl_t, cname, m.tensors[lvl_t].C, {1}, {cname}, ACT_LRELU, false);
    Cn.bias_mode = true;
    return Cn.out_t;
  };
  int cur = hs;
  const int F = d.filter_count;
  const int hs_f[8] = {F / 4, F / 2, F, F / 2, F / 4, F / 8, F / 16, F / 32};
  for (int i = 0; i < 8; i++)
    cur = level_and_connector(cur, Ph, hs_f[i], "level" + std::to_string(i + 1), "connector_conv" + std::to_string(i + 1));
  const int hs_net = cur;
  cur = li_t;
  const int li_f[3] = {2, 4, 8};
  for (int i = 0; i < 3; i++)
    cur = level_and_connector(cur, P0, li_f[i], "lidar_level" + std::to_string(i + 1),
                              "lidar_connector_conv" + std::to_string(i + 1));
  const int lidar_net = cur;
  // fc1 over concat(flatten(hs_net), flatten(lidar_net)): two layers, one z, one weight variable
  {
    Layer& A = b.add_layer("fc1/hs", true, 1, hs_net, "fc1", d.classes * 9, {1}, {"fc1"}, ACT_LRELU, false);
    A.bias_mode = true; A.share = 1;
    const int out_t = A.out_t;
    Layer& Bp = b.add_layer("fc1", true, 1, lidar_net, "fc1_lidar_tmp", d.classes * 9, {1}, {"fc1"}, ACT_LRELU, true);
    Bp.bias_mode = true; Bp.share = 2;
    m.tensor_by_name.erase("fc1_lidar_tmp");
    m.tensors.pop_back();
    m.layers.back().out_t = out_t;
    cur = out_t;
  }
  const int mult[2] = {6, 3};
  for (int i = 0; i < 2; i++) {
    const std::string name = "fc" + std::to_string(i + 2);
    Layer& L = b.add_layer(name, true, 1, cur, name, d.classes * mult[i], {1}, {name}, ACT_LRELU, true);
    L.bias_mode = true;
    cur = L.out_t;
  }
  {
    Layer& L = b.add_layer("fc4", true, 1, cur, "fc4", d.classes, {1}, {"fc4"}, ACT_NONE, false);
    L.bias_mode = true;
    m.logits_t = L.out_t;
    m.last_eval_layer = (int)m.layers.size() - 1;
  }
  return HYP_OK;
}

// CONCNNModel (nnmodel/CONCNNModel.py:23-64): 1x1 / 3x3 / 5x5 convs concatenated -> LRN -> eight 1x1 convs (LRN after
// the first, identity residuals net13 += net11, net22 += net13, dropout after conv31 / conv32) -> flatten -> FC.
// slim defaults: ReLU (the caller passes lrelu_alpha = 0), biases, no BatchNorm; dropout keep_prob = drop_out_ratio.
static int build_concnn(hyp_model& m) {
  const hyp_model_desc& d = m.d;
  Builder b(m);
  const int P = d.patch, F = d.filter_count;
  if (3 * F > 512) return fail(HYP_E_UNSUPPORTED, "CONCNN: filter_count above 170 (the LRN kernels keep 3 rows of 3*F channels per warp in shared memory)");
  m.keep_prob = d.drop_out_ratio;
  const int x = b.add_tensor("x", d.channels, P * P, true);
  int cur;
  {
    Layer& L = b.add_layer("conv0", false, P, x, "net0_out", F, {1, 3, 5}, {"conv0_1x1", "conv0_3x3", "conv0_5x5"},
                           ACT_LRELU, false);
    L.bias_mode = true; L.lrn = true;
    cur = L.out_t;
  }
  const int C1 = 3 * F;
  bool res_failed = false;
  auto conv = [&](const std::string& name, int in_t, bool lrn, bool dropout, int res_src) -> int {
    Layer& L = b.add_layer(name, false, P, in_t, name, C1, {1}, {name}, ACT_LRELU, dropout);
    L.bias_mode = true; L.lrn = lrn;
    if (res_src >= 0 && b.add_resid(L, res_src)) res_failed = true;
    return L.out_t;
  };
  const int net11 = conv("conv11", cur, true, false, -1);
  const int net12 = conv("conv12", net11, false, false, -1);
  const int net13 = conv("conv13", net12, false, false, net11);   // net13 = net13 + net11   (:45)
  const int net21 = conv("conv21", net13, false, false, -1);
  const int net22 = conv("conv22", net21, false, false, net13);   // net22 = net22 + net13   (:50)
  const int net31 = conv("conv31", net22, false, true, -1);
  const int net32 = conv("conv32", net31, false, true, -1);
  const int net33 = conv("conv33", net32, false, false, -1);
  if (res_failed) return fail(HYP_E_INVALID, "residual table");
  Layer& L = b.add_layer("fc", true, 1, net33, "fc", d.classes, {1}, {"fc"}, ACT_NONE, false);
  L.bias_mode = true;
  m.logits_t = L.out_t;
  m.last_eval_layer = (int)m.layers.size() - 1;
  return HYP_OK;
}

static int layout(hyp_model& m) {
  // ---- parameters: per layer W_0..W_{nk-1} (each 128-byte aligned), then beta[Cout] ----
  int64_t po = 0, so = 0;
  for (size_t li = 0; li < m.layers.size(); li++) {
    Layer& L = m.layers[li];
    for (size_t j = 0; j < L.ksizes.size(); j++) {
      const int k = L.ksizes[j];
      if (L.share != 2) po = (int64_t)align_up((size_t)po, 32);  // share 2: the rows continue the first part's variable
      L.w_off.push_back(po);
      if (L.share != 2) {
        Variable v;
        v.name = "nn_core/" + L.kscopes[j] + "/weights";
        v.kind = 0;
        v.offset = po;
        if (L.is_fc) {
          v.rank = 2;
          v.shape[0] = L.Cin + (L.share == 1 ? m.layers[li + 1].Cin : 0);
          v.shape[1] = L.f; v.shape[2] = v.shape[3] = 0;
        } else {
          v.rank = 4;
          v.shape[0] = k; v.shape[1] = k; v.shape[2] = L.Cin; v.shape[3] = L.f;
        }
        m.vars.push_back(v);
      }
      po += (int64_t)k * k * L.Cin * L.f;
    }
    so = (int64_t)align_up((size_t)so, 32);
    L.mm_off = so;
    so += 2 * (int64_t)L.Cout;
    if (L.share == 1) {  // bias / activation belong to the second part
      L.beta_off = po;
      continue;
    }
    po = (int64_t)align_up((size_t)po, 32);
    L.beta_off = po;
    for (size_t j = 0; j < L.ksizes.size(); j++) {
      const char* names[3] = {"beta", "moving_mean", "moving_variance"};
      for (int q = 0; q < (L.bias_mode ? 1 : 3); q++) {
        Variable v;
        v.name = "nn_core/" + L.kscopes[j] + (L.bias_mode ? "/biases" : std::string("/BatchNorm/") + names[q]);
        v.kind = 1 + q;
        v.offset = (q == 0 ? L.beta_off : (q == 1 ? L.mm_off : L.mm_off + L.Cout)) + (int64_t)j * L.f;
        v.rank = 1;
        v.shape[0] = L.f; v.shape[1] = v.shape[2] = v.shape[3] = 0;
        m.vars.push_back(v);
      }
    }
    po += L.Cout;
  }
  m.n_params = (int64_t)align_up((size_t)po, 32);
  m.n_state = (int64_t)align_up((size_t)so, 32);

  // ---- workspace ----
  const size_t B = (size_t)m.d.max_batch;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    const size_t o = off;
    off += bytes;
    return o;
  };
  for (Tensor& t : m.tensors) {
    if (t.external) continue;
    const size_t bytes = B * t.rows_per_sample * t.C * sizeof(float);
    t.a_off = take(bytes);
    t.g_off = take(bytes);
  }
  size_t gz_max = 0;
  for (Layer& L : m.layers) {
    const size_t zb = B * L.rows_per_sample * L.Cout * sizeof(float);
    L.z_off = take(zb);
    gz_max = std::max(gz_max, zb);
    size_t wn = 0;
    for (int k : L.ksizes) wn += (size_t)k * k * L.Cin * L.f;
    L.wt_off = take(wn * sizeof(float));
    L.mean_off = take(L.Cout * sizeof(float));
    L.rstd_off = take(L.Cout * sizeof(float));
    L.s1_off = take(L.Cout * sizeof(float));
    L.s2_off = take(L.Cout * sizeof(float));
  }
  m.gz_off = take(gz_max);
  m.ce_off = take(B * sizeof(float));
  m.mse_off = take(256);
  // statistics regions are contiguous so one memset clears them
  off = align_up(off, 256);
  m.stats_region_off = off;
  for (Layer& L : m.layers) L.stats_off = take(2 * (size_t)L.Cout * sizeof(double));
  m.stats_region_bytes = off - m.stats_region_off;
  off = align_up(off, 256);
  m.bstats_region_off = off;
  for (Layer& L : m.layers) L.bstats_off = take(2 * (size_t)L.Cout * sizeof(double));
  m.bstats_region_bytes = off - m.bstats_region_off;
  m.ws_bytes = align_up(off, 256);
  return HYP_OK;
}

// segment + transpose tables depend on the bound parameter / workspace pointers
static int bind_tables(hyp_model& m) {
  m.segs_host.clear();
  std::vector<TransposeJob> jobs;
  std::vector<int4> tmap;
  for (Layer& L : m.layers) {
    size_t wt_elems = 0;
    for (size_t j = 0; j < L.ksizes.size(); j++) {
      const int k = L.ksizes[j], h = k / 2;
      const float* w = m.params + L.w_off[j];
      float* gw = m.grads + L.w_off[j];
      float* wt = reinterpret_cast<float*>(m.ws + L.wt_off) + wt_elems;
      for (int ky = 0; ky < k; ky++)
        for (int kx = 0; kx < k; kx++) {
          const size_t tap = (size_t)(ky * k + kx) * L.Cin * L.f;
          Seg s;
          s.w = w + tap;
          s.gw = gw + tap;
          s.wt = wt + tap;
          s.ldw = L.f;
          s.ldwt = L.Cin;
          s.col0 = (int)j * L.f;
          s.width = L.f;
          s.dy = ky - h;
          s.dx = kx - h;
          m.segs_host.push_back(s);
          TransposeJob tj;
          tj.src = s.w;
          tj.dst = wt + tap;
          tj.rows = L.Cin;
          tj.cols = L.f;
          const int job = (int)jobs.size();
          jobs.push_back(tj);
          for (int tr = 0; tr < (int)cdiv(tj.rows, 32); tr++)
            for (int tc = 0; tc < (int)cdiv(tj.cols, 32); tc++) tmap.push_back(make_int4(job, tr, tc, 0));
        }
      wt_elems += (size_t)k * k * L.Cin * L.f;
    }
  }
  if (!m.segs_dev) {
    HYP_CUDA(cudaMalloc(&m.segs_dev, m.segs_host.size() * sizeof(Seg)));
    HYP_CUDA(cudaMalloc(&m.tjobs_dev, jobs.size() * sizeof(TransposeJob)));
    HYP_CUDA(cudaMalloc(&m.tmap_dev, tmap.size() * sizeof(int4)));
  }
  HYP_CUDA(cudaMemcpy(m.segs_dev, m.segs_host.data(), m.segs_host.size() * sizeof(Seg), cudaMemcpyHostToDevice));
  HYP_CUDA(cudaMemcpy(m.tjobs_dev, jobs.data(), jobs.size() * sizeof(TransposeJob), cudaMemcpyHostToDevice));
  HYP_CUDA(cudaMemcpy(m.tmap_dev, tmap.data(), tmap.size() * sizeof(int4), cudaMemcpyHostToDevice));
  m.n_ttiles = (int)tmap.size();
  m.segs_bound = true;
  return HYP_OK;
}

// useful MACs*2 of one conv kernel over B samples: SAME-padding zero taps are not counted
static double conv_flops(int64_t B, int P, int k, int cin, int cout) {
  int64_t v1 = 0;
  for (int d = -(k / 2); d <= k / 2; d++) v1 += P - (d < 0 ? -d : d);
  return 2.0 * (double)B * (double)(v1 * v1) * cin * cout;
}
static double layer_flops(const Layer& L, int64_t B) {
  double f = 0;
  for (int k : L.ksizes) f += conv_flops(B, L.P, k, L.Cin, L.f);
  return f;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int ew_grid(int64_t total) { return (int)std::min<int64_t>(cdiv(total, 256), 148 * 16); }

// records a registered event once every layer whose parameters start at or above its offset is done (backward walks
// the layers last to first, parameters are laid out in layer order)
static void grad_notify(const hyp_model& m, int li, cudaStream_t st) {
  for (const hyp_model::Notify& n : m.notify) {
    const bool here = m.layers[li].w_off[0] >= n.offset && (li == 0 || m.layers[li - 1].w_off[0] < n.offset);
    if (here) cudaEventRecord(n.event, st);
  }
}

static const float* act_ptr(const hyp_model& m, int t, const float* x) {
  return m.tensors[t].external ? x : reinterpret_cast<const float*>(m.ws + m.tensors[t].a_off);
}
static float* grad_ptr(const hyp_model& m, int t) { return reinterpret_cast<float*>(m.ws + m.tensors[t].g_off); }

static int forward_impl(hyp_model& m, const float* x, int64_t B, bool training, bool update_moving, uint64_t seed,
                        cudaStream_t st) {
  const int nl = training ? (int)m.layers.size() : m.last_eval_layer + 1;
  if (training) HYP_CUDA(cudaMemsetAsync(m.ws + m.stats_region_off, 0, m.stats_region_bytes, st));
  const float keep_prob = 1.f - m.d.drop_out_ratio;
  for (int li = 0; li < nl; li++) {
    Layer& L = m.layers[li];
    const int64_t rows = B * L.rows_per_sample;
    const float* A = act_ptr(m, L.in_t, x);
    float* Z = reinterpret_cast<float*>(m.ws + L.z_off);
    double* stats = reinterpret_cast<double*>(m.ws + L.stats_off);
    for (size_t j = 0; j < L.ksizes.size(); j++) {
      RowGemmArgs a;
      a.A = A; a.lda = L.Cin; a.M = (int)rows; a.P = L.P;
      a.segs = m.segs_dev + L.kseg_begin[j];
      a.nseg = L.ksizes[j] * L.ksizes[j];
      a.mode = 0; a.Kfix = L.Cin;
      a.C = Z; a.ldc = L.Cout; a.c_col0 = (int)j * L.f; a.N = L.f;
      a.accumulate = 0;
      a.stats = training ? stats : nullptr;
      a.stats_ld = L.Cout;
      a.a_vec = (L.Cin % 4 == 0) && al16(A);
      a.b_vec = (L.f % 4 == 0) && al16(m.params + L.w_off[j]) && (((size_t)L.Cin * L.f) % 4 == 0);
      int rc = launch_rowgemm(a, st, "fwd", conv_flops(B, L.P, L.ksizes[j], L.Cin, L.f));
      if (rc) return rc;
    }
    float* mean = reinterpret_cast<float*>(m.ws + L.mean_off);
    float* rstd = reinterpret_cast<float*>(m.ws + L.rstd_off);
    PROF("bn_finalize_kernel", 32.0 * L.Cout,
         (bn_finalize_kernel<<<(unsigned)cdiv(L.Cout, 128), 128, 0, st>>>(
             stats, L.Cout, (double)rows, m.d.bn_eps, m.d.bn_decay, m.state + L.mm_off,
             m.state + L.mm_off + L.Cout, mean, rstd, training ? 1 : 0, (training && update_moving) ? 1 : 0)));
    ApplyArgs p;
    p.z = Z; p.mean = mean; p.rstd = rstd; p.beta = m.params + L.beta_off;
    p.out = reinterpret_cast<float*>(m.ws + m.tensors[L.out_t].a_off);
    p.rows = rows; p.C = L.Cout; p.act = L.act; p.alpha = m.d.lrelu_alpha;
    p.keep = (L.dropout && training) ? keep_prob : 1.f;
    p.seed = seed; p.stream_id = L.drop_stream;
    p.res0 = p.res1 = nullptr; p.idx0 = p.idx1 = nullptr; p.C0 = p.C1 = 0;
    if (L.res.size() > 0) {
      p.res0 = act_ptr(m, L.res[0].src, x); p.idx0 = L.res[0].idx; p.C0 = m.tensors[L.res[0].src].C;
    }
    if (L.res.size() > 1) {
      p.res1 = act_ptr(m, L.res[1].src, x); p.idx1 = L.res[1].idx; p.C1 = m.tensors[L.res[1].src].C;
    }
    PROF("bn_apply_fwd_kernel", 4.0 * rows * L.Cout * (2 + L.res.size()),
         (bn_apply_fwd_kernel<<<ew_grid(rows * L.Cout), 256, 0, st>>>(p)));
  }
  return HYP_OK;
}

static int backward_impl(hyp_model& m, const float* x, const uint8_t* labels, int64_t B, float* loss_out,
                         cudaStream_t st) {
  const hyp_model_desc& d = m.d;
  const int nl = (int)m.layers.size();
  std::vector<char> ginit(m.tensors.size(), 0);
  HYP_CUDA(cudaMemsetAsync(m.ws + m.bstats_region_off, 0, m.bstats_region_bytes, st));
  HYP_CUDA(cudaMemsetAsync(m.grads, 0, (size_t)m.n_params * sizeof(float), st));
  HYP_CUDA(cudaMemsetAsync(m.ws + m.mse_off, 0, 256, st));
  // transposed weights for dgrad
  PROF("transpose_kernel", 8.0 * m.n_params, (transpose_kernel<<<m.n_ttiles, 256, 0, st>>>(m.tjobs_dev, m.tmap_dev)));
  // losses: mean_B(CE_i + mse)  (HYPELCNNModel.py:101-112, common_nn_ops.py:214)
  float* ce = reinterpret_cast<float*>(m.ws + m.ce_off);
  double* mse_acc = reinterpret_cast<double*>(m.ws + m.mse_off);
  const float* logits = act_ptr(m, m.logits_t, x);
  PROF("ce_loss_kernel", 8.0 * B * d.classes,
       (ce_loss_kernel<<<(unsigned)cdiv(B * 32, 256), 256, 0, st>>>(logits, labels, B, d.classes, ce,
                                                                    grad_ptr(m, m.logits_t), 1.f / (float)B)));
  ginit[m.logits_t] = 1;
  const int64_t D = (int64_t)d.patch * d.patch * d.channels;
  PROF("mse_kernel", 12.0 * B * D,
       (mse_kernel<<<ew_grid(B * D), 256, 0, st>>>(act_ptr(m, m.recon_t, x), x, B * D, mse_acc,
                                                   grad_ptr(m, m.recon_t), 1.f / (float)(B * D))));
  ginit[m.recon_t] = 1;
  loss_finalize_kernel<<<1, 256, 0, st>>>(ce, B, mse_acc, (double)(B * D), loss_out, nullptr);
  HYP_LAUNCHED();

  const float keep_prob = 1.f - d.drop_out_ratio;
  float* gz = reinterpret_cast<float*>(m.ws + m.gz_off);
  for (int li = nl - 1; li >= 0; li--) {
    Layer& L = m.layers[li];
    const int64_t rows = B * L.rows_per_sample;
    if (!ginit[L.out_t]) return fail(HYP_E_STATE, "backward: no gradient reached " + L.scope);
    BnBwdArgs p;
    p.gout = grad_ptr(m, L.out_t);
    p.z = reinterpret_cast<float*>(m.ws + L.z_off);
    p.mean = reinterpret_cast<float*>(m.ws + L.mean_off);
    p.rstd = reinterpret_cast<float*>(m.ws + L.rstd_off);
    p.beta = m.params + L.beta_off;
    p.rows = rows; p.C = L.Cout; p.act = L.act; p.alpha = d.lrelu_alpha;
    p.keep = L.dropout ? keep_prob : 1.f;
    p.seed = m.last_seed; p.stream_id = L.drop_stream;
    p.sums = reinterpret_cast<double*>(m.ws + L.bstats_off);
    float* s1 = reinterpret_cast<float*>(m.ws + L.s1_off);
    float* s2 = reinterpret_cast<float*>(m.ws + L.s2_off);
    p.s1 = s1; p.s2 = s2; p.gz = gz;
    {
      const int cblocks = (int)cdiv(L.Cout, 32);
      int rblocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(rows, 64), cdiv(148 * 8, cblocks)));
      const int rpb = (int)cdiv(rows, rblocks);
      rblocks = (int)cdiv(rows, rpb);
      PROF("bn_bwd_reduce_kernel", 8.0 * rows * L.Cout,
           (bn_bwd_reduce_kernel<<<dim3(cblocks, rblocks), 256, 0, st>>>(p, rpb)));
    }
    bn_bwd_finalize_kernel<<<(unsigned)cdiv(L.Cout, 128), 128, 0, st>>>(p.sums, L.Cout, (double)rows, s1, s2,
                                                                         m.grads + L.beta_off);
    HYP_LAUNCHED();
    PROF("bn_bwd_apply_kernel", 12.0 * rows * L.Cout,
         (bn_bwd_apply_kernel<<<ew_grid(rows * L.Cout), 256, 0, st>>>(p)));
    // residual pushes
    for (const Resid& r : L.res) {
      const Tensor& src = m.tensors[r.src];
      if (!src.needs_grad) continue;
      PROF("resid_bwd_kernel", 4.0 * rows * (L.Cout + 2.0 * src.C),
           (resid_bwd_kernel<<<ew_grid(rows * src.C), 256, 0, st>>>(p.gout, L.Cout, grad_ptr(m, r.src), src.C, r.lo,
                                                                    r.hi, rows, ginit[r.src])));
      ginit[r.src] = 1;
    }
    const float* A = act_ptr(m, L.in_t, x);
    // wgrad
    {
      WgradArgs a;
      a.A = A; a.lda = L.Cin; a.M = (int)rows; a.P = L.P; a.Cin = L.Cin;
      a.G = gz; a.ldg = L.Cout;
      a.segs = m.segs_dev + L.seg_begin; a.nseg = L.seg_count;
      const int64_t tiles = cdiv(L.Cin, GEMM_BM) * cdiv(L.f, L.f > 64 ? 128 : L.f) * L.seg_count;
      int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(148 * 6, tiles), cdiv(rows, 256)));
      int rps = (int)cdiv(cdiv(rows, ksplit), GEMM_BK) * GEMM_BK;
      ksplit = (int)cdiv(rows, rps);
      a.rows_per_split = rps;
      a.a_vec = (L.Cin % 4 == 0) && al16(A);
      a.g_vec = (L.Cout % 4 == 0) && (L.f % 4 == 0) && al16(gz);
      int rc = launch_wgrad(a, L.f, ksplit, st, layer_flops(L, B));
      if (rc) return rc;
    }
    // dgrad
    if (m.tensors[L.in_t].needs_grad) {
      RowGemmArgs a;
      a.A = gz; a.lda = L.Cout; a.M = (int)rows; a.P = L.P;
      a.segs = m.segs_dev + L.seg_begin; a.nseg = L.seg_count;
      a.mode = 1; a.Kfix = 0;
      a.C = grad_ptr(m, L.in_t); a.ldc = L.Cin; a.c_col0 = 0; a.N = L.Cin;
      a.accumulate = ginit[L.in_t];
      a.stats = nullptr; a.stats_ld = 0;
      a.a_vec = (L.Cout % 4 == 0) && (L.f % 4 == 0) && al16(gz);
      a.b_vec = (L.Cin % 4 == 0) && al16(m.ws + L.wt_off) && (((size_t)L.Cin * L.f) % 4 == 0);
      int rc = launch_rowgemm(a, st, "dgrad", layer_flops(L, B));
      if (rc) return rc;
      ginit[L.in_t] = 1;
    }
    grad_notify(m, li, st);
  }
  return HYP_OK;
}

}  // namespace hyp

#include "hyp_tc_engine.cuh"

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int hyp_version(void) { return HYP_ABI_VERSION; }
const char* hyp_last_error(void) { return g_last_error.c_str(); }
int64_t hyp_launch_count(int reset) {
  const int64_t v = g_launch_count;
  if (reset) g_launch_count = 0;
  return v;
}

int hyp_profile_enable(int on) {
  g_prof.clear();
  g_prof.on = on != 0;
  return HYP_OK;
}

// aggregates all records per kernel tag (synchronises the device); idx enumerates tags
int hyp_profile_get(int idx, char name[64], double* total_ms, int64_t* launches, double* flops, double* bytes) {
  HYP_CHECK_ARG(name && total_ms && launches && flops && bytes, "null argument");
  if (idx < 0 || idx >= (int)g_prof.names.size()) return fail(HYP_E_INVALID, "hyp_profile_get: index out of range");
  HYP_CUDA(cudaDeviceSynchronize());
  double ms = 0, fl = 0, by = 0;
  int64_t n = 0;
  for (const ProfRec& r : g_prof.recs) {
    if (r.tag != idx) continue;
    float t = 0.f;
    HYP_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t; fl += r.flops; by += r.bytes; n++;
  }
  strncpy(name, g_prof.names[idx].c_str(), 63);
  name[63] = 0;
  *total_ms = ms; *launches = n; *flops = fl; *bytes = by;
  return HYP_OK;
}

int hyp_model_create(const hyp_model_desc* desc, hyp_model** out) {
  HYP_CHECK_ARG(desc && out, "null argument");
  HYP_CHECK_ARG(desc->kind == HYP_MODEL_HYPELCNN || desc->kind == HYP_MODEL_DUALCNN || desc->kind == HYP_MODEL_CONCNN,
                "unknown model kind");
  const bool tensor_core = desc->precision_mode == HYP_PRECISION_3XTF32 || desc->precision_mode == HYP_PRECISION_3XF16 ||
                           desc->precision_mode == HYP_PRECISION_BF16;
  if (desc->kind != HYP_MODEL_HYPELCNN && !tensor_core)
    return fail(HYP_E_UNSUPPORTED, "hyp_model_create: DUALCNN / CONCNN are built for the tensor-core engine only");
  HYP_CHECK_ARG(desc->patch >= 1 && desc->patch % 2 == 1 && desc->patch <= 15, "patch must be odd, 1..15");
  HYP_CHECK_ARG(desc->channels >= 1 && desc->classes >= 2 && desc->classes <= 255, "channels/classes out of range");
  HYP_CHECK_ARG(desc->filter_count >= 8, "filter_count out of range");
  HYP_CHECK_ARG(desc->kind != HYP_MODEL_HYPELCNN || (desc->spectral_levels >= 1 && desc->spatial_levels >= 1),
                "levels out of range");
  HYP_CHECK_ARG(desc->max_batch >= 1, "max_batch must be positive");
  HYP_CHECK_ARG(desc->drop_out_ratio >= 0.f && desc->drop_out_ratio < 1.f, "drop_out_ratio in [0,1)");
  if (desc->precision_mode != HYP_PRECISION_FP32 && !tensor_core)
    return fail(HYP_E_UNSUPPORTED, "hyp_model_create: unknown precision mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(HYP_E_CUDA, "hyp_model_create: no CUDA device (this library has no CPU fallback)");
  std::unique_ptr<hyp_model> m(new hyp_model());
  m->d = *desc;
  int rc = desc->kind == HYP_MODEL_DUALCNN ? build_dualcnn(*m)
           : (desc->kind == HYP_MODEL_CONCNN ? build_concnn(*m) : build_hypelcnn(*m));
  if (rc) return rc;
  rc = layout(*m);
  if (rc) return rc;
  if (tensor_core) {
    rc = hyp::tc::tc_layout(*m);
    if (rc) { hyp::tc::tc_destroy(*m); return rc; }
    m->ws_bytes = m->tc->ws_bytes;
  }
  if ((int64_t)m->d.max_batch * m->d.patch * m->d.patch > (int64_t)INT32_MAX / 2)
    return fail(HYP_E_INVALID, "hyp_model_create: max_batch too large");
  *out = m.release();
  return HYP_OK;
}

void hyp_model_destroy(hyp_model* m) {
  if (!m) return;
  hyp::tc::tc_destroy(*m);
  for (void* p : m->owned) cudaFree(p);
  if (m->segs_dev) cudaFree(m->segs_dev);
  if (m->tjobs_dev) cudaFree(m->tjobs_dev);
  if (m->tmap_dev) cudaFree(m->tmap_dev);
  delete m;
}

int hyp_model_sizes(const hyp_model* m, int64_t* n_params, int64_t* n_bn_state, int64_t* workspace_bytes,
                    int32_t* n_variables) {
  HYP_CHECK_ARG(m, "null model");
  if (n_params) *n_params = m->n_params;
  if (n_bn_state) *n_bn_state = m->n_state;
  if (workspace_bytes) *workspace_bytes = (int64_t)m->ws_bytes;
  if (n_variables) *n_variables = (int32_t)m->vars.size();
  return HYP_OK;
}

int hyp_model_variable(const hyp_model* m, int idx, char name[128], int32_t* kind, int64_t* offset, int32_t shape[4],
                       int32_t* rank) {
  HYP_CHECK_ARG(m && name && kind && offset && shape && rank, "null argument");
  HYP_CHECK_ARG(idx >= 0 && idx < (int)m->vars.size(), "variable index out of range");
  const Variable& v = m->vars[idx];
  strncpy(name, v.name.c_str(), 127);
  name[127] = 0;
  *kind = v.kind;
  *offset = v.offset;
  for (int i = 0; i < 4; i++) shape[i] = v.shape[i];
  *rank = v.rank;
  return HYP_OK;
}

int hyp_model_bind(hyp_model* m, float* params, float* grads, float* bn_state, void* workspace, size_t ws_bytes) {
  HYP_CHECK_ARG(m && params && grads && bn_state && workspace, "null argument");
  HYP_CHECK_ARG(ws_bytes >= m->ws_bytes, "workspace too small");
  HYP_CHECK_ARG(al16(params) && al16(grads) && al16(bn_state) && ((uintptr_t)workspace & 255) == 0,
                "buffers must be 16-byte (workspace 256-byte) aligned");
  m->params = params;
  m->grads = grads;
  m->state = bn_state;
  m->ws = static_cast<char*>(workspace);
  m->last_B = -1;
  if (m->tc) {
    m->segs_bound = true;
    return hyp::tc::tc_bind(*m);
  }
  return bind_tables(*m);
}

int hyp_model_forward(hyp_model* m, const float* x, int64_t B, int is_training, int update_moving,
                      uint64_t dropout_seed, float* logits, float* recon, void* stream) {
  HYP_CHECK_ARG(m && x, "null argument");
  if (!m->segs_bound) return fail(HYP_E_STATE, "hyp_model_forward: call hyp_model_bind first");
  HYP_CHECK_ARG(B >= 1 && B <= m->d.max_batch, "B out of range");
  HYP_CHECK_ARG(!is_training || B >= 2, "training-mode BatchNorm needs B >= 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = m->tc ? hyp::tc::tc_forward(*m, x, B, is_training != 0, update_moving != 0, dropout_seed, st)
                 : forward_impl(*m, x, B, is_training != 0, update_moving != 0, dropout_seed, st);
  if (rc) return rc;
  m->last_B = B;
  m->last_training = is_training != 0;
  m->last_seed = dropout_seed;
  m->last_x = x;
  if (m->tc) {  // padded row-major FC outputs -> dense
    if (logits)
      HYP_CUDA(cudaMemcpy2DAsync(logits, (size_t)m->d.classes * sizeof(float), hyp::tc::tc_plane0(*m, m->logits_t),
                                 (size_t)m->tc->tt[m->logits_t].Cp * sizeof(float), (size_t)m->d.classes * sizeof(float),
                                 (size_t)B, cudaMemcpyDeviceToDevice, st));
    if (recon) {
      if (!is_training || m->recon_t < 0)
        return fail(HYP_E_INVALID, "hyp_model_forward: recon only exists in HYPELCNN's training graph");
      const size_t D = (size_t)m->d.patch * m->d.patch * m->d.channels;
      HYP_CUDA(cudaMemcpy2DAsync(recon, D * sizeof(float), hyp::tc::tc_plane0(*m, m->recon_t),
                                 (size_t)m->tc->tt[m->recon_t].Cp * sizeof(float), D * sizeof(float), (size_t)B,
                                 cudaMemcpyDeviceToDevice, st));
    }
    return HYP_OK;
  }
  if (logits)
    HYP_CUDA(cudaMemcpyAsync(logits, act_ptr(*m, m->logits_t, x), (size_t)B * m->d.classes * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  if (recon) {
    if (!is_training) return fail(HYP_E_INVALID, "hyp_model_forward: recon only exists in the training graph");
    const size_t D = (size_t)m->d.patch * m->d.patch * m->d.channels;
    HYP_CUDA(cudaMemcpyAsync(recon, act_ptr(*m, m->recon_t, x), (size_t)B * D * sizeof(float),
                             cudaMemcpyDeviceToDevice, st));
  }
  return HYP_OK;
}

int hyp_model_loss(hyp_model* m, const float* logits, const float* recon, const float* x, const uint8_t* labels,
                   int64_t B, float* per_sample_loss, void* stream) {
  HYP_CHECK_ARG(m && logits && labels && per_sample_loss, "null argument");
  HYP_CHECK_ARG(!recon || x, "x is required with recon");
  if (!m->ws) return fail(HYP_E_STATE, "hyp_model_loss: call hyp_model_bind first");
  HYP_CHECK_ARG(B >= 1 && B <= m->d.max_batch, "B out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ce = reinterpret_cast<float*>(m->ws + (m->tc ? m->tc->ce_off : m->ce_off));
  double* mse_acc = reinterpret_cast<double*>(m->ws + (m->tc ? m->tc->mse_off : m->mse_off));
  HYP_CUDA(cudaMemsetAsync(mse_acc, 0, 256, st));
  ce_loss_kernel<<<(unsigned)cdiv(B * 32, 256), 256, 0, st>>>(logits, labels, B, m->d.classes, ce, nullptr, 0.f);
  HYP_LAUNCHED();
  const int64_t D = (int64_t)m->d.patch * m->d.patch * m->d.channels;
  if (recon) {
    mse_kernel<<<ew_grid(B * D), 256, 0, st>>>(recon, x, B * D, mse_acc, nullptr, 0.f);
    HYP_LAUNCHED();
  }
  loss_finalize_kernel<<<1, 256, 0, st>>>(ce, B, recon ? mse_acc : nullptr, (double)(B * D), nullptr, per_sample_loss);
  HYP_LAUNCHED();
  return HYP_OK;
}

int hyp_model_loss_backward(hyp_model* m, const float* x, const uint8_t* labels, int64_t B, float* loss_out,
                            void* stream) {
  HYP_CHECK_ARG(m && x && labels && loss_out, "null argument");
  if (!m->segs_bound) return fail(HYP_E_STATE, "hyp_model_loss_backward: call hyp_model_bind first");
  if (!m->last_training || m->last_B != B || m->last_x != x)
    return fail(HYP_E_STATE, "hyp_model_loss_backward: needs the preceding hyp_model_forward(is_training=1) on the same x/B");
  if (m->tc) return hyp::tc::tc_backward(*m, labels, B, loss_out, static_cast<cudaStream_t>(stream));
  return backward_impl(*m, x, labels, B, loss_out, static_cast<cudaStream_t>(stream));
}

int hyp_model_set_grad_notify(hyp_model* m, int64_t param_offset, void* event) {
  HYP_CHECK_ARG(m, "null model");
  if (!event) {  // no event: forget every registration
    m->notify.clear();
    return HYP_OK;
  }
  HYP_CHECK_ARG(param_offset >= 0 && param_offset < m->n_params, "param_offset outside the parameter buffer");
  for (hyp_model::Notify& n : m->notify)
    if (n.event == static_cast<cudaEvent_t>(event)) {
      n.offset = param_offset;
      return HYP_OK;
    }
  HYP_CHECK_ARG(m->notify.size() < 8, "at most 8 gradient notifications");
  m->notify.push_back({static_cast<cudaEvent_t>(event), param_offset});
  return HYP_OK;
}

int hyp_adam_step(float* params, const float* grads, float* mbuf, float* vbuf, int64_t n, float lr, float b1,
                  float b2, float eps, int64_t t, float grad_scale, void* stream) {
  HYP_CHECK_ARG(params && grads && mbuf && vbuf, "null argument");
  HYP_CHECK_ARG(n >= 0 && t >= 1, "n >= 0 and t >= 1 required");
  if (n == 0) return HYP_OK;
  const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)b2, (double)t)) / (1.0 - std::pow((double)b1, (double)t));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("adam_kernel", 28.0 * n,
       (adam_kernel<<<ew_grid(n), 256, 0, st>>>(params, grads, mbuf, vbuf, n, (float)lr_t, 1.f - b1, 1.f - b2, eps,
                                                grad_scale)));
  return HYP_OK;
}

int hyp_momentum_step(float* params, const float* grads, float* accum, int64_t n, float lr, float momentum,
                      float grad_scale, void* stream) {
  HYP_CHECK_ARG(params && grads && accum, "null argument");
  HYP_CHECK_ARG(n >= 0, "n >= 0 required");
  if (n == 0) return HYP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("momentum_kernel", 20.0 * n,
       (momentum_kernel<<<ew_grid(n), 256, 0, st>>>(params, grads, accum, n, lr, momentum, grad_scale)));
  return HYP_OK;
}

int hyp_augment_patches(const float* in, float* out, int64_t B, int patch, int channels, int do_rotation,
                        int do_reflection, float spectral, uint64_t seed, uint8_t* choices_out, float* deltas_out,
                        void* stream) {
  HYP_CHECK_ARG(in && out && in != out, "in / out must be distinct non-null buffers");
  HYP_CHECK_ARG(B >= 0 && patch >= 1 && channels >= 1 && spectral >= 0.f, "bad shape");
  HYP_CHECK_ARG(B <= INT32_MAX, "too many patches for one call");
  if (B == 0) return HYP_OK;
  AugmentArgs a;
  a.in = in; a.out = out; a.B = B; a.P = patch; a.C = channels; a.do_rot = do_rotation != 0; a.do_flip = do_reflection != 0;
  a.spectral = spectral; a.seed = seed; a.choices = choices_out; a.deltas = deltas_out;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("augment_kernel", 8.0 * B * patch * patch * channels, (augment_kernel<<<(unsigned)B, 256, 0, st>>>(a)));
  return HYP_OK;
}

int hyp_gan_generator_forward(const float* in, int ld_in, float* out, int ld_out, int64_t rows, int bands,
                              int copy_extra, const float* weights, int encoder_only, int clip_invalid_values,
                              int is_shadow_graph, void* stream) {
  HYP_CHECK_ARG(in && out && weights, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= GAN_MAX_C, "bands out of range (8..512)");
  HYP_CHECK_ARG(rows >= 0 && copy_extra >= 0 && ld_in >= bands + copy_extra && ld_out >= bands + copy_extra, "bad shape");
  if (rows == 0) return HYP_OK;
  GanGenArgs a;
  a.in = in; a.out = out; a.rows = rows; a.C = bands; a.ld_in = ld_in; a.ld_out = ld_out; a.copy_extra = copy_extra;
  a.weights = weights; a.nlayers = encoder_only ? 4 : 7; a.clip = clip_invalid_values != 0; a.is_shadow = is_shadow_graph != 0;
  a.nets = nullptr;
  const int K[7] = {bands, bands / 2, bands / 4, bands / 8, bands / 4, bands / 2, bands};
  int nw = 0;
  for (int l = 0; l < a.nlayers; l++) nw += K[l] + 1;
  const size_t smem = (size_t)(((nw + 3) & ~3) + 4 * 3 * bands) * sizeof(float);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, 4), 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_generator_fwd_kernel", 4.0 * rows * (ld_in + ld_out),
       (gan_generator_fwd_kernel<<<grid, 128, smem, st>>>(a)));
  return HYP_OK;
}

// ---- GAN training kernels (CycleGAN: gan/wrappers/cycle_gan_wrapper.py, gan/shadow_data_models.py) ----
static int gan_gen_weight_count(int bands) {
  const int K[7] = {bands, bands / 2, bands / 4, bands / 8, bands / 4, bands / 2, bands};
  int nw = 0;
  for (int l = 0; l < 7; l++) nw += K[l] + 1;
  return nw;
}
int hyp_gan_generator_train_forward(const float* x, int64_t rows, int bands, const float* weights, float* nets,
                                    void* stream) {
  HYP_CHECK_ARG(x && weights && nets, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= GAN_MAX_C && rows >= 0, "bad shape");
  if (rows == 0) return HYP_OK;
  GanGenArgs a;
  a.in = x; a.rows = rows; a.C = bands; a.ld_in = bands; a.copy_extra = 0; a.weights = weights; a.nlayers = 7;
  a.clip = 0; a.is_shadow = 0; a.nets = nets;
  a.out = nets + 7 * (size_t)bands;  // row r's output lands in nets[r][7][:] (stride 8 * bands)
  a.ld_out = 8 * bands;
  const int nw = gan_gen_weight_count(bands);
  const size_t smem = (size_t)(((nw + 3) & ~3) + 4 * 3 * bands) * sizeof(float);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, 4), 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_generator_fwd_kernel", 4.0 * rows * 9 * bands, (gan_generator_fwd_kernel<<<grid, 128, smem, st>>>(a)));
  return HYP_OK;
}

static int gan_generator_backward_impl(const float* nets, const float* gout, const float* gout_enc, int64_t rows,
                                       int bands, const float* weights, float* gin, float* gweights, void* stream) {
  HYP_CHECK_ARG(nets && (gout || gout_enc) && weights && gweights, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= 256 && rows >= 0, "bands out of range (8..256)");
  if (rows == 0) return HYP_OK;
  GanGenBwdArgs a;
  a.nets = nets; a.gout = gout; a.gout_enc = gout_enc; a.rows = rows; a.C = bands; a.weights = weights; a.gin = gin;
  a.gweights = gweights;
  const int nw = gan_gen_weight_count(bands);
  const size_t smem = (size_t)(2 * ((nw + 3) & ~3) + 4 * 17 * bands) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    HYP_CUDA(cudaFuncSetAttribute(gan_generator_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  HYP_CHECK_ARG(smem <= 96 * 1024, "bands too large for the backward kernel's shared memory");
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, 4), 148 * 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_generator_bwd_kernel", 4.0 * rows * 10 * bands, (gan_generator_bwd_kernel<<<grid, 128, smem, st>>>(a)));
  return HYP_OK;
}

int hyp_gan_generator_backward(const float* nets, const float* gout, int64_t rows, int bands, const float* weights,
                               float* gin, float* gweights, void* stream) {
  HYP_CHECK_ARG(gout, "null argument");
  return gan_generator_backward_impl(nets, gout, nullptr, rows, bands, weights, gin, gweights, stream);
}

int hyp_gan_generator_backward_enc(const float* nets, const float* gout, const float* gout_enc, int64_t rows, int bands,
                                   const float* weights, float* gin, float* gweights, void* stream) {
  return gan_generator_backward_impl(nets, gout, gout_enc, rows, bands, weights, gin, gweights, stream);
}

static int gan_disc_launch(bool backward, const GanDiscArgs& a, cudaStream_t st) {
  const int C = a.C, H = C / 2;
  const int nw = C * C + C + C * C + C + C * H + H, nwp = (nw + 3) & ~3;
  static bool attr = false;
  if (!attr) {
    HYP_CUDA(cudaFuncSetAttribute(gan_discriminator_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HYP_CUDA(cudaFuncSetAttribute(gan_discriminator_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(a.rows, 8), 148 * 2);
  if (!backward) {
    const size_t smem = (size_t)(nwp + 8 * 3 * C) * sizeof(float);
    PROF("gan_discriminator_fwd_kernel", 4.0 * a.rows * 3.5 * C, (gan_discriminator_fwd_kernel<<<grid, 256, smem, st>>>(a)));
  } else {
    const int wpad = ((2 * C * (C + 1) + C * (H + 1)) + 3) & ~3;
    const size_t smem = (size_t)(wpad + (a.gweights ? nwp : 0) + 8 * 6 * C) * sizeof(float);
    HYP_CHECK_ARG(smem <= 227 * 1024, "bands too large for the discriminator backward kernel's shared memory");
    PROF("gan_discriminator_bwd_kernel", 4.0 * a.rows * 4.5 * C, (gan_discriminator_bwd_kernel<<<grid, 256, smem, st>>>(a)));
  }
  return HYP_OK;
}
int hyp_gan_discriminator_forward(const float* x, int64_t rows, int bands, const float* weights, float* hidden,
                                  float* out, void* stream) {
  HYP_CHECK_ARG(x && weights && hidden && out, "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= 96 && bands % 2 == 0 && rows >= 0, "bands out of range (even, 8..96: the weights live in shared memory)");
  if (rows == 0) return HYP_OK;
  GanDiscArgs a{};
  a.x = x; a.rows = rows; a.C = bands; a.weights = weights; a.h = hidden; a.out = out;
  return gan_disc_launch(false, a, static_cast<cudaStream_t>(stream));
}
int hyp_gan_discriminator_backward(const float* x, const float* hidden, const float* gout, int64_t rows, int bands,
                                   const float* weights, float* gin, float* gweights, void* stream) {
  HYP_CHECK_ARG(x && weights && hidden && gout && (gin || gweights), "null argument");
  HYP_CHECK_ARG(bands >= 8 && bands <= 96 && bands % 2 == 0 && rows >= 0, "bands out of range (even, 8..96)");
  if (rows == 0) return HYP_OK;
  GanDiscArgs a{};
  a.x = x; a.rows = rows; a.C = bands; a.weights = weights; a.h = const_cast<float*>(hidden); a.gout = gout; a.gin = gin;
  a.gweights = gweights;
  return gan_disc_launch(true, a, static_cast<cudaStream_t>(stream));
}
int hyp_gan_loss_grad(int mode, const float* a, const float* b, float target, float scale, int64_t numel, float* grad,
                      int accumulate, double* loss_acc, void* stream) {
  HYP_CHECK_ARG(a && (mode == 0 || mode == 2 || (mode == 1 && b)) && numel >= 0, "bad argument");
  if (numel == 0) return HYP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_loss_grad_kernel", 8.0 * numel,
       (gan_loss_grad_kernel<<<ew_grid(numel), 256, 0, st>>>(mode, a, b, target, scale, numel, grad, accumulate, loss_acc)));
  return HYP_OK;
}
int hyp_gan_l2_regularizer(const float* weights, float* grads, int64_t n, float scale, double* loss_acc, void* stream) {
  HYP_CHECK_ARG(weights && n >= 0, "bad argument");
  if (n == 0) return HYP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PROF("gan_l2_reg_kernel", 8.0 * n, (gan_l2_reg_kernel<<<ew_grid(n), 256, 0, st>>>(weights, grads, n, scale, loss_acc)));
  return HYP_OK;
}

static int featdisc_launch(bool backward, GanFeatArgs a, int patch_count, cudaStream_t st) {
  HYP_CHECK_ARG(a.x && a.weights && a.z && a.sumsq, "null argument");
  HYP_CHECK_ARG(patch_count >= 1 && a.C >= patch_count, "bad patch_count");
  a.ps = a.C / patch_count;
  HYP_CHECK_ARG(a.ps >= 4 && a.ps <= 32 && a.E >= 1 && a.E <= NCE_MAX_E, "slice width (4..32) / embedding size (1..8)");
  const int slices = (a.C + a.ps - 1) / a.ps;
  HYP_CHECK_ARG(slices <= NCE_MAX_S, "too many slices");
  if (a.rows == 0) return HYP_OK;
  const int nrows_act = 2 * a.ps + a.ps / 4 + a.ps / 2 + a.E;
  const size_t smem = (size_t)(((featdisc_slice_weights(a.ps, a.E) + 3) & ~3) + (backward ? 2 : 1) * nrows_act * FD_PITCH) *
                      sizeof(float);
  static bool attr = false;
  if (!attr) {
    HYP_CUDA(cudaFuncSetAttribute(gan_featdisc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    HYP_CUDA(cudaFuncSetAttribute(gan_featdisc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr = true;
  }
  dim3 grid((unsigned)cdiv(a.rows, FD_ROWS), (unsigned)slices);
  if (backward)
    PROF("gan_featdisc_bwd_kernel", 4.0 * a.rows * (2 * a.C + 2 * slices * a.E),
         (gan_featdisc_bwd_kernel<<<grid, FD_ROWS, smem, st>>>(a)));
  else
    PROF("gan_featdisc_fwd_kernel", 4.0 * a.rows * (a.C + slices * a.E),
         (gan_featdisc_fwd_kernel<<<grid, FD_ROWS, smem, st>>>(a)));
  return HYP_OK;
}

int64_t hyp_gan_feature_discriminator_weight_count(int bands, int patch_count, int embedded_feature_size) {
  if (patch_count < 1 || bands < patch_count) return -1;
  const int ps = bands / patch_count;
  return (int64_t)((bands + ps - 1) / ps) * featdisc_slice_weights(ps, embedded_feature_size);
}

int hyp_gan_feature_discriminator_forward(const float* x, int64_t rows, int bands, int patch_count,
                                          int embedded_feature_size, const float* weights, float* z, float* sumsq,
                                          void* stream) {
  HYP_CHECK_ARG(rows >= 0 && bands >= 4 && bands <= GAN_MAX_C, "bad shape");
  GanFeatArgs a{};
  a.x = x; a.rows = rows; a.C = bands; a.E = embedded_feature_size; a.weights = weights; a.z = z; a.sumsq = sumsq;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sumsq && patch_count >= 1 && bands >= patch_count) {
    const int ps = bands / patch_count;
    HYP_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * ((bands + ps - 1) / ps), st));
  }
  return featdisc_launch(false, a, patch_count, st);
}

int hyp_gan_feature_discriminator_backward(const float* x, const float* z, const float* sumsq, const float* gf,
                                           const float* dot, int64_t rows, int bands, int patch_count,
                                           int embedded_feature_size, const float* weights, float* gin, float* gweights,
                                           void* stream) {
  HYP_CHECK_ARG(rows >= 0 && bands >= 4 && bands <= GAN_MAX_C, "bad shape");
  HYP_CHECK_ARG(gf && dot && (gin || gweights), "null argument");
  GanFeatArgs a{};
  a.x = x; a.rows = rows; a.C = bands; a.E = embedded_feature_size; a.weights = weights;
  a.z = const_cast<float*>(z); a.sumsq = const_cast<float*>(sumsq); a.gf = gf; a.dot = dot; a.gin = gin;
  a.gweights = gweights;
  return featdisc_launch(true, a, patch_count, static_cast<cudaStream_t>(stream));
}

int hyp_gan_patchnce(const float* z_gen, const float* z_real, const float* sumsq_gen, const float* sumsq_real,
                     int64_t rows, int slices, int embedded_feature_size, float tau, float scale, int fused_grad,
                     float* g_gen, float* g_real, float* dot_gen, float* dot_real, double* loss_acc, void* stream) {
  HYP_CHECK_ARG(z_gen && z_real && sumsq_gen && sumsq_real, "null argument");
  HYP_CHECK_ARG(rows >= 0 && slices >= 1 && slices <= NCE_MAX_S && embedded_feature_size >= 1 &&
                    embedded_feature_size <= NCE_MAX_E && tau > 0.f,
                "bad shape");
  HYP_CHECK_ARG((g_gen != nullptr) == (g_real != nullptr), "g_gen and g_real go together");
  HYP_CHECK_ARG(!g_gen || (dot_gen && dot_real), "dot_gen / dot_real are required with the gradients");
  if (rows == 0) return HYP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_gen) {
    HYP_CUDA(cudaMemsetAsync(dot_gen, 0, sizeof(float) * slices, st));
    HYP_CUDA(cudaMemsetAsync(dot_real, 0, sizeof(float) * slices, st));
  }
  GanNceArgs a;
  a.zg = z_gen; a.zr = z_real; a.ssg = sumsq_gen; a.ssr = sumsq_real; a.rows = rows; a.S = slices;
  a.E = embedded_feature_size; a.inv_tau = 1.f / tau; a.scale = scale; a.fused_grad = fused_grad; a.gfg = g_gen;
  a.gfr = g_real; a.dotg = dot_gen; a.dotr = dot_real; a.loss_acc = loss_acc;
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows, 4), 148 * 16);
  PROF("gan_patchnce_kernel", 16.0 * rows * slices * embedded_feature_size, (gan_patchnce_kernel<<<grid, 128, 0, st>>>(a)));
  return HYP_OK;
}

int hyp_argmax_confusion(const float* logits, const uint8_t* labels, int64_t B, int classes, uint8_t* pred,
                         int32_t* confusion, void* stream) {
  HYP_CHECK_ARG(logits && (pred || confusion), "null argument");
  HYP_CHECK_ARG(classes >= 1 && classes <= 256 && B >= 0, "classes/B out of range");
  HYP_CHECK_ARG(!confusion || labels, "labels are required to accumulate a confusion matrix");
  if (B == 0) return HYP_OK;
  argmax_confusion_kernel<<<(unsigned)cdiv(B, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, B, classes, pred, confusion);
  HYP_LAUNCHED();
  return HYP_OK;
}

int hyp_scatter_class_map(const uint8_t* pred, const int32_t* targets_xy, int64_t N, int H, int W,
                          uint8_t* class_map, void* stream) {
  HYP_CHECK_ARG(pred && targets_xy && class_map && H > 0 && W > 0 && N >= 0, "bad argument");
  if (N == 0) return HYP_OK;
  scatter_class_map_kernel<<<(unsigned)cdiv(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, targets_xy, N, H, W, class_map);
  HYP_LAUNCHED();
  return HYP_OK;
}

// ---- tensor-core building block probe (tests/test_gpu_tc.py) --------------------------------
// mn = 0: A [M,K], B [N,K] row-major, D = A * B^T.   mn = 1: A [K,M], B [K,N], D = A^T * B.
// Splits both operands into TF32 (hi, lo) planes, builds tensor maps and tile tables exactly
// as the engine does, and runs tc_gemm_kernel.  ksplit > 1 (mn = 1 only) splits K over CTAs
// that accumulate with atomics (D must be zeroed by the caller).
// host-only: the static tile schedule on a plain cost vector (tests/test_schedule.py)
int hyp_debug_schedule(const double* costs, int units, int groups, int windowed, int32_t* group_of_unit,
                       int32_t* rank_in_group) {
  HYP_CHECK_ARG(costs && group_of_unit && rank_in_group && units >= 0 && groups >= 1, "bad argument");
  std::vector<std::pair<double, int>> cost((size_t)units);
  for (int u = 0; u < units; u++) cost[u] = {costs[u], u};
  const std::vector<std::vector<int>> lists = tc::assign_units(cost, groups, windowed != 0);
  for (int u = 0; u < units; u++) group_of_unit[u] = rank_in_group[u] = -1;
  for (int g = 0; g < groups; g++)
    for (size_t i = 0; i < lists[g].size(); i++) {
      const int u = lists[g][i];
      HYP_CHECK_ARG(u >= 0 && u < units && group_of_unit[u] < 0, "schedule lost or duplicated a unit");
      group_of_unit[u] = g;
      rank_in_group[u] = (int32_t)i;
    }
  return HYP_OK;
}

// host-only: the tap groups of a level's wgrad launch and the gz position every (group, set, source position) reads
// (tests/test_schedule.py).  Rows of `out`: group, set, dy, dx, source position q, output position (255 = outside the
// patch); one row per set of every (group, q) the launch visits.
int hyp_debug_level_tap_groups(int P, int R, int nt, int fpad, int max_sets, int32_t* out, int cap_rows, int* rows_out) {
  HYP_CHECK_ARG(P >= 1 && R >= 1 && nt >= 1 && fpad >= 16 && fpad % 16 == 0 && max_sets >= 1 && rows_out, "bad argument");
  const int h = std::min(R - 1, P - 1), NS = R * nt, PP = P * P;
  const int spg = std::min(std::min(256 / fpad, tc::TC_MAX_CB), NS), ngroups = (int)cdiv(NS, spg);  // as tc_plan
  const std::vector<tc::TapGroup> groups = tc::level_tap_groups(h, R, nt, NS, fpad, ngroups, PP, std::min(max_sets, tc::TC_MAX_SETS));
  int n = 0;
  for (size_t g = 0; g < groups.size(); g++)
    for (int q = 0; q < PP; q++) {
      bool any = false;
      for (auto& tp : groups[g].t) any = any || tc::tap_output_position(P, q, tp.first, tp.second) != 255;
      if (!any) continue;
      for (size_t j = 0; j < groups[g].t.size(); j++) {
        if (out && n < cap_rows) {
          int32_t* r = out + (size_t)n * 6;
          r[0] = (int32_t)g; r[1] = (int32_t)j; r[2] = groups[g].t[j].first; r[3] = groups[g].t[j].second; r[4] = q;
          r[5] = tc::tap_output_position(P, q, groups[g].t[j].first, groups[g].t[j].second);
        }
        n++;
      }
    }
  *rows_out = n;
  return HYP_OK;
}

int hyp_debug_tc_gemm(int mn_flags, const float* A, const float* B, int M, int N, int K, float* D, float* stats,
                      int raw_hi, int bn, int ksplit, int chunk_kb, void* stream) {
  using namespace hyp::tc;
  const int mn = mn_flags & 1;
  const int cg = (mn_flags & 2) ? 2 : 1;
  const int op = (mn_flags >> 2) & 3;   // OP_TF32X3 / OP_F16X3 / OP_BF16
  HYP_CHECK_ARG(A && B && D && M > 0 && N > 0 && K > 0, "bad argument");
  HYP_CHECK_ARG(op <= OP_BF16, "unknown operand format");
  HYP_CHECK_ARG(N % 4 == 0, "N must be a multiple of 4 (output row alignment)");
  HYP_CHECK_ARG(ksplit >= 1 && (mn || ksplit == 1), "ksplit needs mn = 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kbe = op_kbe(op), npl = op_planes(op), esz = op_esize(op);
  const int lda_src = mn ? M : K, ldb_src = mn ? N : K;
  const int64_t a_rows = mn ? K : M, b_rows_src = mn ? K : N;
  const int ldA = (int)align_up(lda_src, 16 / esz), ldB = (int)align_up(ldb_src, 16 / esz);
  // the 16-bit formats multiply scaled operands (A by 4, B by 64: powers of two, exact) and undo it in the epilogue --
  // the same mechanism the engine uses to keep fp16 remainders in the normal range
  const float sa_ = op == OP_F16X3 ? 4.f : 1.f, sb_ = op == OP_F16X3 ? 64.f : 1.f;
  void *Ap = nullptr, *Bp = nullptr;
  HYP_CUDA(cudaMalloc(&Ap, (size_t)npl * a_rows * ldA * esz));
  HYP_CUDA(cudaMalloc(&Bp, (size_t)npl * b_rows_src * ldB * esz));
  if (op == OP_TF32X3) {
    float *Af = static_cast<float*>(Ap), *Bf = static_cast<float*>(Bp);
    split_planes_kernel<<<256, 256, 0, st>>>(A, a_rows, lda_src, lda_src, Af, Af + a_rows * ldA, ldA, raw_hi);
    HYP_LAUNCHED();
    split_planes_kernel<<<256, 256, 0, st>>>(B, b_rows_src, ldb_src, ldb_src, Bf, Bf + b_rows_src * ldB, ldB, raw_hi);
    HYP_LAUNCHED();
  } else {
    split_planes16_kernel<<<256, 256, 0, st>>>(A, a_rows, lda_src, lda_src, static_cast<uint16_t*>(Ap), (size_t)a_rows * ldA, ldA, op, sa_);
    HYP_LAUNCHED();
    split_planes16_kernel<<<256, 256, 0, st>>>(B, b_rows_src, ldb_src, ldb_src, static_cast<uint16_t*>(Bp), (size_t)b_rows_src * ldB, ldB, op, sb_);
    HYP_LAUNCHED();
  }
  const int ntile_n = 256;
  if (bn <= 0) bn = (int)std::min<int64_t>(256, align_up(N, 16));
  HYP_CHECK_ARG(bn % (8 * cg) == 0 && bn <= 256, "bn must be a multiple of 8 per CTA, <= 256");
  CUtensorMap tmA, tmB;
  int rc;
  {
    const uint64_t da[4] = {(uint64_t)lda_src, (uint64_t)a_rows, 1, (uint64_t)npl};
    const uint64_t sa[3] = {(uint64_t)ldA, (uint64_t)a_rows * ldA, (uint64_t)a_rows * ldA};
    const uint32_t ba[4] = {(uint32_t)kbe, mn ? (uint32_t)kbe : 128u, 1, 1};
    if ((rc = make_map(&tmA, Ap, da, sa, ba, mn != 0, op))) return rc;
    const uint64_t db[4] = {(uint64_t)ldb_src, (uint64_t)b_rows_src, 1, (uint64_t)npl};
    const uint64_t sb[3] = {(uint64_t)ldB, (uint64_t)b_rows_src * ldB, (uint64_t)b_rows_src * ldB};
    const uint32_t bb[4] = {(uint32_t)kbe, mn ? (uint32_t)kbe : (uint32_t)(bn / cg), 1, 1};
    if ((rc = make_map(&tmB, Bp, db, sb, bb, mn != 0, op))) return rc;
  }
  std::vector<TcSeg> segs;
  std::vector<TcTile> tiles;
  const int mt = (int)cdiv(M, 128), nt = (int)cdiv(N, ntile_n);
  const int kblocks = (int)cdiv(K, kbe);
  const int kb_per = (int)cdiv(kblocks, ksplit);
  int max_brows = 0;
  // one segment per (N tile, K split); the M tiles that share it are consecutive (CTA pairs)
  for (int in = 0; in < nt; in++)
    for (int ks = 0; ks < ksplit; ks++) {
      const int kb0 = ks * kb_per, kb1 = std::min(kblocks, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int n0 = in * ntile_n, nw = std::min(ntile_n, N - n0);
      TcSeg s{};
      s.nk = kb1 - kb0;
      s.n_mma = (int)align_up(nw, 16);
      if (!mn) {
        s.b1 = n0;
        s.nb = (int)cdiv(s.n_mma, bn);
        max_brows = std::max(max_brows, s.nb * bn);
      } else {
        s.a1 = kb0 * kbe;
        s.b0 = n0; s.b1 = kb0 * kbe;
        s.nb = (int)cdiv(s.n_mma, kbe);
        max_brows = std::max(max_brows, cg * (int)cdiv(s.n_mma / cg, kbe) * kbe);
      }
      const int sidx = (int)segs.size();
      segs.push_back(s);
      const int mt_pad = (int)align_up(mt, cg);
      for (int im = 0; im < mt_pad; im++) {
        TcTile t{};
        t.seg_begin = sidx;
        t.seg_count = 1;
        t.total_kb = s.nk;
        t.m_valid = im < mt ? std::min(128, M - im * 128) : 0;
        if (mn) t.a0_add = im * 128; else t.a1_add = im * 128;
        t.ncb = 1;
        t.ld_out = N;
        t.stats_row = im;
        t.cb[0].out_off = (int64_t)im * 128 * N + n0;
        t.cb[0].tcol = 0;
        t.cb[0].width = nw;
        t.cb[0].stats_col = n0;
        tiles.push_back(t);
      }
    }
  HYP_CHECK_ARG(mn || cg == 1 || bn * (int)cdiv(max_brows, bn) == max_brows, "bad bn");
  if ((rc = tc_finalize_tiles(tiles, segs))) return rc;
  TcSeg* dsegs = nullptr;
  TcTile* dtiles = nullptr;
  HYP_CUDA(cudaMalloc(&dsegs, segs.size() * sizeof(TcSeg)));
  HYP_CUDA(cudaMalloc(&dtiles, tiles.size() * sizeof(TcTile)));
  HYP_CUDA(cudaMemcpyAsync(dsegs, segs.data(), segs.size() * sizeof(TcSeg), cudaMemcpyHostToDevice, st));
  HYP_CUDA(cudaMemcpyAsync(dtiles, tiles.data(), tiles.size() * sizeof(TcTile), cudaMemcpyHostToDevice, st));
  TcParams p{};
  p.segs = dsegs; p.tiles = dtiles; p.out = D; p.stats = stats; p.stats_ld = N;
  p.op = op; p.out_scale = 1.f / (sa_ * sb_);
  p.epi = ksplit > 1 ? EPI_ATOMIC : EPI_STORE;
  p.b_rows = max_brows; p.bn = bn; p.chunk_kb = chunk_kb; p.stages = 0;
  rc = mn ? (cg == 2 ? launch_tc<true, 2>(tmA, tmB, p, (int)tiles.size(), st)
                     : launch_tc<true, 1>(tmA, tmB, p, (int)tiles.size(), st))
          : (cg == 2 ? launch_tc<false, 2>(tmA, tmB, p, (int)tiles.size(), st)
                     : launch_tc<false, 1>(tmA, tmB, p, (int)tiles.size(), st));
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(Ap); cudaFree(Bp); cudaFree(dsegs); cudaFree(dtiles);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(HYP_E_CUDA, std::string("hyp_debug_tc_gemm: ") + cudaGetErrorString(e));
  return HYP_OK;
}

int hyp_model_debug_tensor(hyp_model* m, const char* name, int what, float** ptr, int64_t* numel) {
  HYP_CHECK_ARG(m && name && ptr && numel, "null argument");
  if (!m->ws || m->last_B < 0) return fail(HYP_E_STATE, "hyp_model_debug_tensor: no forward has run");
  const std::string n(name);
  if (m->tc) {  // position-major padded tensors -> dense [B][PP][C] in the debug scratch
    using namespace hyp::tc;
    TcState& S = *m->tc;
    int t = -1;
    const float* src = nullptr;
    if (what == 3) {
      for (size_t li = 0; li < m->layers.size(); li++)
        if (m->layers[li].scope == n) {
          HYP_CUDA(cudaDeviceSynchronize());
          *ptr = reinterpret_cast<float*>(m->ws + S.tl[li].mean_off);
          *numel = m->layers[li].Cout;
          return HYP_OK;
        }
      return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown layer " + n);
    }
    if (what == 1) {
      for (size_t li = 0; li < m->layers.size(); li++)
        if (m->layers[li].scope == n) { t = m->layers[li].out_t; src = reinterpret_cast<float*>(m->ws + S.tl[li].z_off); }
      if (t < 0) return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown layer " + n);
    } else {
      auto it = m->tensor_by_name.find(n);
      if (it == m->tensor_by_name.end()) return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown tensor " + n);
      t = it->second;
      if (what == 2 && !m->tensors[t].needs_grad) return fail(HYP_E_INVALID, "hyp_model_debug_tensor: tensor has no gradient");
      src = what == 2 ? tc_grad(*m, t) : tc_plane0(*m, t);
    }
    const TcTensor& T = S.tt[t];
    float* dst = reinterpret_cast<float*>(m->ws + S.dbg_off);
    const int64_t total = m->last_B * T.PP * T.C;
    HYP_CUDA(cudaDeviceSynchronize());
    if (src) tc_extract_kernel<<<tc_grid(total), 256>>>(src, T.Cp, (int)m->last_B, T.PP, T.C, dst);
    else  // an activation without a value plane: rebuilt from its operand planes
      tc_extract_planes_kernel<<<tc_grid(total), 256>>>(reinterpret_cast<const uint16_t*>(tc_plane1(*m, t)), T.plane_elems, T.Cp,
                                                        (int)m->last_B, T.PP, T.C, dst);
    HYP_LAUNCHED();
    HYP_CUDA(cudaDeviceSynchronize());
    *ptr = dst;
    *numel = total;
    return HYP_OK;
  }
  if (what == 3) {
    for (Layer& L : m->layers)
      if (L.scope == n) {
        *ptr = reinterpret_cast<float*>(m->ws + L.mean_off);
        *numel = L.Cout;
        return HYP_OK;
      }
    return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown layer " + n);
  }
  if (what == 1) {
    for (Layer& L : m->layers)
      if (L.scope == n) {
        *ptr = reinterpret_cast<float*>(m->ws + L.z_off);
        *numel = m->last_B * L.rows_per_sample * L.Cout;
        return HYP_OK;
      }
    return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown layer " + n);
  }
  auto it = m->tensor_by_name.find(n);
  if (it == m->tensor_by_name.end() || m->tensors[it->second].external)
    return fail(HYP_E_INVALID, "hyp_model_debug_tensor: unknown tensor " + n);
  const Tensor& t = m->tensors[it->second];
  *ptr = reinterpret_cast<float*>(m->ws + (what == 2 ? t.g_off : t.a_off));
  *numel = m->last_B * t.rows_per_sample * t.C;
  return HYP_OK;
}

int hyp_model_dropout_mask(hyp_model* m, const char* layer_scope, uint64_t seed, int64_t B, uint8_t* mask_out,
                           void* stream) {
  HYP_CHECK_ARG(m && layer_scope && mask_out && B >= 1, "bad argument");
  for (Layer& L : m->layers)
    if (L.scope == layer_scope && L.dropout) {
      const int64_t n = B * L.rows_per_sample * L.Cout;
      dropout_mask_kernel<<<(unsigned)cdiv(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          seed, L.drop_stream, n, m->keep_prob, mask_out);
      HYP_LAUNCHED();
      return HYP_OK;
    }
  return fail(HYP_E_INVALID, std::string("hyp_model_dropout_mask: no dropout layer ") + layer_scope);
}

}  // extern "C"
