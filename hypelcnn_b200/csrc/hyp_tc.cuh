// tcgen05 / TMEM / TMA building blocks and the segment-GEMM kernel of the tensor-core path.
//
// One launch = a list of output tiles.  A tile is a 128-row accumulator held in TMEM
// (fp32, up to 256 columns) that sums a list of K-segments:
//     D[128, n] (+)= A_seg[128, K_seg] * B_seg[K_seg, n]
// Every operand arrives through TMA (SWIZZLE_128B boxes of 32 fp32 = 128 B rows) from tensors
// stored as two fp32 planes (hi, lo): plane 0 = the value rounded to TF32, plane 1 = the TF32
// rounding of the remainder.  Each K step of 8 issues three kind::tf32 MMAs
// (hi*hi + lo*hi + hi*lo, "3xTF32"), which reproduces fp32 products to ~2^-22 relative error
// while running on the 5th-generation tensor cores.
//
// Two operand arrangements:
//   MN = false : A [rows, K] and B [n, K] are K-major (forward, dgrad)
//   MN = true  : A [K, rows] and B [K, n] are MN-major (wgrad: K runs over batch rows)
//
// Warp roles (384 threads = three warpgroups): warpgroup 0 = warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer,
// warps 2..3 idle; warpgroups 1..2 = warps 4..11 epilogue.  The first warpgroup hands most of its registers back
// (setmaxnreg.dec 104), the epilogue warpgroups take them (setmaxnreg.inc 200): the epilogue keeps a 128-column fp32
// accumulator per thread and spilled at the 168 registers a uniform 384-thread CTA gets.  Epilogue (TMEM chunks -> fp32 register sums -> global, per-tile BatchNorm
// partial sums).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "hyp_common.cuh"

namespace hyp {
namespace tc {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_EPI_WARP0 = 4;            // first epilogue warp (warpgroup-aligned)
constexpr int TC_THREADS = 32 * (TC_EPI_WARP0 + TC_EPI_WARPS);
constexpr int TC_BM = 128;
constexpr int TC_KB = 32;                  // fp32 elements of K per pipeline stage
constexpr int TC_PLANE_A = TC_BM * 128;    // bytes of one A plane per stage
constexpr int TC_MAX_COLS = 256;
constexpr int TC_MAX_BP = 8;               // distinct MMA widths per tile
constexpr int TC_MAX_SEGS = 128;           // segments per tile (held in registers across a warp's lanes)
constexpr int TC_MAX_CB = 8;               // column blocks per tile           // accumulator columns per tile
constexpr int TC_SMEM_LIMIT = 226 * 1024;
// after the stage ring: epilogue staging, statistics partials [2][4][256] floats, barriers + TMEM slot (256 B)
// [8 warps][32 rows][20 floats]
constexpr int TC_SMEM_FIXED = 256 + 2 * 4 * TC_MAX_COLS * 4 + TC_EPI_WARPS * 32 * 20 * 4;

struct alignas(16) TcSeg {
  int32_t a0, a1, a2;  // A box coordinates (tensor dims 0..2) at K block 0; dim 3 = plane
  int32_t b0, b1, b2;
  int32_t nk;          // K blocks of TC_KB
  int32_t n_mma;       // N of this segment's MMAs: multiple of 16 in [16, 256]
  int32_t nb;          // B boxes per K block and plane
  int32_t ks_last;     // K steps (quarters of a K block) the LAST K block of the segment needs: 1..4, 0 = 4.  The tail of
                       // a K range that is not a multiple of the block is zero padding: its MMAs are skipped
  int32_t nsets;       // MN-major only: 2..4 = the segment multiplies its A tile with that many B sets (one MMA group
                       // each, accumulator columns [s * n_mma, (s + 1) * n_mma)); the sets share b0 / b1 / n_mma and
                       // differ in b2: set 0 reads b2, set s > 0 reads byte s - 1 of b2x.  0 / 1 = one set.
                       // A b2 outside the tensor is a set without a contribution (the TMA unit fills zeros).
  int32_t b2x;
};
constexpr int TC_MAX_SETS = 4;

struct alignas(16) TcColBlock {
  int32_t tcol;       // first accumulator column (multiple of 16; blocks are listed by increasing tcol)
  int32_t width;      // valid columns
  int32_t stats_col;  // first column in the statistics row
  int32_t pad;
  int64_t out_off;    // element offset of (tile row 0, block column 0) inside `out`
  int64_t pad2;
};

struct alignas(16) TcTile {
  int32_t seg_begin, seg_count;
  int32_t m_valid;    // valid accumulator rows
  int32_t ncb;
  int32_t ld_out;     // elements between rows of `out`
  int32_t stats_row;
  int32_t total_kb;   // sum of nk over the tile's segments
  int32_t a1_add;     // added to every segment's a1 (K-major: first row of the tile)
  int32_t b1_add;     // added to every segment's b1 (K-major: first B row of the tile's N range)
  int32_t a0_add;     // added to every segment's a0 (MN-major: first A column of the tile)
  int32_t n_cols;     // accumulator columns the tile uses (max n_mma of its segments)
  int32_t nbp;        // breakpoints below: K block from which the MMA width is bp_n (n_mma never increases)
  int32_t bp_kb[TC_MAX_BP];
  int32_t bp_n[TC_MAX_BP];
  TcColBlock cb[TC_MAX_CB];
};

enum { EPI_STORE = 0, EPI_ACCUM = 1, EPI_ATOMIC = 2 };

// operand formats: how the two GEMM operands are stored and multiplied
//   OP_TF32X3 : two fp32 planes (value, TF32 remainder), 3 kind::tf32 MMAs per K step of 8      (fp32-accurate)
//   OP_F16X3  : two fp16 planes (hi, lo = fp16(v - hi)), 3 kind::f16 MMAs per K step of 16      (fp32-accurate, 2x rate)
//   OP_BF16   : one bf16 plane, 1 kind::f16 MMA per K step of 16                                 (fast mode)
// A K block is always one 128-byte row per operand row: 32 fp32 or 64 16-bit elements.
enum { OP_TF32X3 = 0, OP_F16X3 = 1, OP_BF16 = 2 };
__host__ __device__ inline int op_planes(int op) { return op == OP_BF16 ? 1 : 2; }
__host__ __device__ inline int op_kbe(int op) { return op == OP_TF32X3 ? 32 : 64; }       // K elements per 128-byte block
__host__ __device__ inline int op_esize(int op) { return op == OP_TF32X3 ? 4 : 2; }

struct TcParams {
  const TcSeg* segs;
  const TcTile* tiles;
  float* out;
  int32_t op;             // OP_*
  float out_scale;        // every output value is multiplied by out_scale * (*out_scale_ptr) (operand scaling of the
  const float* out_scale_ptr;  // 16-bit formats; nullable)
  float* stats;       // nullable: [stats rows][2][stats_ld] per-tile column sum / sum of squares
  int32_t stats_ld;
  int32_t epi;
  int32_t b_rows;     // B rows (n) reserved per stage and plane, multiple of 8
  int32_t bn;         // rows of one B box (K-major); MN-major boxes are 32 columns x 32 rows
  int32_t chunk_kb;   // K blocks the tensor core accumulates before the fp32 register merge
  int32_t stages;
  int32_t ntiles;     // tiles of this launch (multiple of the CTA group size)
  unsigned long long* timing;  // nullable: [CTA][8] cycle counters (HYP_TC_TIMING diagnostics)
  int32_t out_cols;   // extent of the 2-D view behind tma_out
  int64_t out_rows;
  int32_t tma_out;    // 1 (EPI_ATOMIC only): tiles with one column block hand their 32 x 16 slabs to the TMA unit as f32
                      // add reductions through the third tensor map (a 2-D fp32 view of `out`)
};

// b_rows = B rows held by ONE CTA per stage and plane
inline size_t tc_smem_bytes(int b_rows, int stages, int planes = 2) {
  return 1024 + (size_t)stages * planes * (TC_PLANE_A + (size_t)b_rows * 128) + TC_SMEM_FIXED;
}
inline int tc_pick_stages(int b_rows, int planes = 2) {
  const size_t fixed = 1024 + TC_SMEM_FIXED;
  const size_t st = (size_t)planes * (TC_PLANE_A + (size_t)b_rows * 128);
  int s = (int)((TC_SMEM_LIMIT - fixed) / st);
  return s > 8 ? 8 : s;
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B, Blackwell version bit set
// layout: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only
// layout kind::tf32 accepts for MN-major operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: D fp32, A/B format fmt (kind::tf32: 2 = tf32; kind::f16: 0 = f16, 1 = bf16), M = 128 per CTA
// of the group
__device__ __forceinline__ uint32_t make_idesc(int n, bool mn_major, int cg, uint32_t fmt) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= fmt << 7;
  d |= fmt << 10;
  if (mn_major) d |= (1u << 15) | (1u << 16);
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)((TC_BM * cg) >> 4) << 24;
  return d;
}

// sum over the 32 lanes of v[j], for every j: afterwards lane l holds the total of column l in v[0]
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// ------------------------------------------------------------------------------------------
// The tensor core adds each MMA result into the TMEM accumulator with round-toward-zero
// (measured: scripts/probe_tc_accum.py, profiles/r01_tc_accum_probe.txt), so the error of a
// long K accumulation grows linearly with the number of MMAs.  The kernel therefore lets the
// tensor core accumulate at most `chunk_kb` K blocks in one TMEM buffer, and the epilogue
// warps add the finished chunks into fp32 registers (round-to-nearest FADD) while the next
// chunk runs in the other buffer.
//
// Persistent: CTA (or CTA pair) c works on tiles c, c + G, c + 2G, ...; the smem stage ring and
// the TMEM chunk ring keep running across tiles, so the producer prefetches the next tile's
// operands and the tensor core starts its first two chunks while the epilogue warps are still
// writing the previous tile.
//
// CG = 2 (cta_group::2): the two CTAs of a cluster own tiles 2i and 2i+1 which share their B
// operand (same segment list).  Each CTA loads its own A rows and HALF of the B rows; the
// leader CTA issues M = 256 MMAs that read both halves, so B crosses L2 -> SM once per pair.
template <int CG>
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// arrive on `bar` of every CTA of the group once all MMAs issued so far have completed
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  } else {  // data lands in this CTA's smem, the transaction bytes are counted on the leader's barrier
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- warp-uniform issue helpers --------------------------------------------------------------
// The producer and MMA roles run with all 32 lanes in uniform control flow and predicate the
// single-thread instructions on `pred` (1 on the elected lane).  With every operand provably
// warp-uniform the compiler keeps descriptors in uniform registers; a `if (lane == 0)` branch
// instead makes it wrap every UTCHMMA / UTMALDG in an ELECT + R2UR waterfall loop (~20 extra
// instructions per MMA, measured: the issue loop, not the tensor pipe, paced small-N tiles).
__device__ __forceinline__ uint32_t elect_one_pred() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(pred));
  return pred;
}
template <typename T>
__device__ __forceinline__ T uni(T v) { return __shfl_sync(0xffffffffu, v, 0); }

template <int CG>
__device__ __forceinline__ void mma_tf32_u(uint32_t pred, uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t desc_hi32,
                                           uint32_t idesc, uint32_t accum) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accum), "r"(pred)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accum), "r"(pred)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void mma_f16_u(uint32_t pred, uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t desc_hi32,
                                          uint32_t idesc, uint32_t accum) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accum), "r"(pred)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accum), "r"(pred)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_commit_u(uint32_t pred, uint32_t bar) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar), "r"(pred) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}"
        ::"r"(bar), "r"(pred), "h"(mask) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_4d_u(uint32_t pred, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                              int c2, int c3) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
        "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(pred)
        : "memory");
  } else {  // data lands in this CTA's smem, the transaction bytes are counted on the leader's barrier
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
        "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(pred)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t pred, uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
      ::"r"(bar), "r"(bytes), "r"(pred) : "memory");
}

constexpr int TC_STAGE_LD = 20;                            // floats per staged row (16 + pad, keeps float4 alignment)
constexpr int TC_STAGE_FLOATS = 32 * TC_STAGE_LD;          // per epilogue warp

template <bool MN, int CG, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t b_plane = (uint32_t)(p.b_rows / CG) * 128u;  // B rows held by THIS CTA
  const uint32_t npl = (uint32_t)op_planes(p.op);
  const int kbe = op_kbe(p.op);                 // K elements per 128-byte block
  const int mnb = kbe;                          // MN-major boxes: mnb columns x kbe rows (32 x 32 fp32, 64 x 64 16-bit)
  const uint32_t mn_box_bytes = (uint32_t)kbe * 128u;
  const uint32_t stage_bytes = npl * (TC_PLANE_A + b_plane);
  // after the ring: epilogue staging [8 warps][2560 B] (512-byte aligned: the TMA slabs are SWIZZLE_64B boxes), the
  // statistics partials [2][4][TC_MAX_COLS], then barriers + TMEM slot (256 B)
  const uint32_t fixed_head = (uint32_t)(TC_EPI_WARPS * TC_STAGE_FLOATS * 4 + 2 * 4 * TC_MAX_COLS * 4);
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes + fixed_head;
  float* s_stage = reinterpret_cast<float*>(gen + (size_t)p.stages * stage_bytes);       // [8 warps][32][TC_STAGE_LD]
  float* s_part = s_stage + TC_EPI_WARPS * TC_STAGE_FLOATS;                              // [2][4][TC_MAX_COLS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (size_t)p.stages * stage_bytes + fixed_head + 192);
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * p.stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * p.stages + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tm_kernel0 = p.timing ? clock64() : 0;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  // tile walk: group g = blockIdx.x / CG handles tile slots g, g + G, ...; slot t = tiles [t*CG, t*CG + CG)
  const int ngroups = (int)gridDim.x / CG;
  const int group = (int)blockIdx.x / CG;
  const int nslots = p.ntiles / CG;
  const int CH = p.chunk_kb;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (p.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    for (int s = 0; s < p.stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), TC_EPI_WARPS * CG);  // the leader's barrier collects both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uni(*reinterpret_cast<volatile uint32_t*>(tmem_slot));

  if (warp < TC_EPI_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 0) {
    // ===================== TMA producer (one warp per CTA, one elected lane issues) =====================
    const uint32_t pred = elect_one_pred();
    int stage = 0;
    uint32_t phase = 0;
    long long tm_wait_empty = 0;
    // tile headers are fetched one tile ahead; a tile's segments are loaded once, spread over the lanes
    // (segment l + 32 i in register set i of lane l) -- the L1 left beside ~220 KB of smem does not keep
    // the tables, and an L2 round trip per segment paced short-K tiles
    int4 hdr0 = make_int4(0, 0, 0, 0), hdrm = hdr0, hdr1 = hdr0;
    if (group < nslots) {
      const int4* hp = reinterpret_cast<const int4*>(p.tiles + (size_t)group * CG + rank);
      hdr0 = __ldg(hp); hdrm = __ldg(hp + 1); hdr1 = __ldg(hp + 2);
    }
    for (int slot = group; slot < nslots; slot += ngroups) {
      // TcTile words: [0] seg_begin seg_count m_valid ncb | [1] ld_out stats_row total_kb a1_add | [2] b1_add a0_add n_cols nbp
      const int seg_begin = uni(hdr0.x), seg_count = uni(hdr0.y);
      const int a1_add = uni(hdrm.w), b1_add = uni(hdr1.x), a0_add = uni(hdr1.y);
      if (slot + ngroups < nslots) {
        const int4* hp = reinterpret_cast<const int4*>(p.tiles + (size_t)(slot + ngroups) * CG + rank);
        hdr0 = __ldg(hp); hdrm = __ldg(hp + 1); hdr1 = __ldg(hp + 2);
      }
      int4 sA[4], sB[4], sC[4];  // TcSeg words: a0 a1 a2 b0 | b1 b2 nk n_mma | nb - - -
#pragma unroll
      for (int i = 0; i < 4; i++) {
        sA[i] = sB[i] = sC[i] = make_int4(0, 0, 0, 0);
        if (lane + 32 * i < seg_count) {
          const int4* sp4 = reinterpret_cast<const int4*>(p.segs + seg_begin + lane + 32 * i);
          sA[i] = __ldg(sp4); sB[i] = __ldg(sp4 + 1); sC[i] = __ldg(sp4 + 2);
        }
      }
      for (int si = 0; si < seg_count; si++) {
        const int rs = si >> 5, src = si & 31;
        const int4 wA = rs == 0 ? sA[0] : (rs == 1 ? sA[1] : (rs == 2 ? sA[2] : sA[3]));
        const int4 wB = rs == 0 ? sB[0] : (rs == 1 ? sB[1] : (rs == 2 ? sB[2] : sB[3]));
        const int4 wC = rs == 0 ? sC[0] : (rs == 1 ? sC[1] : (rs == 2 ? sC[2] : sC[3]));
        const int sa0 = __shfl_sync(0xffffffffu, wA.x, src) + a0_add, sa1 = __shfl_sync(0xffffffffu, wA.y, src) + a1_add;
        const int sa2 = __shfl_sync(0xffffffffu, wA.z, src), sb0 = __shfl_sync(0xffffffffu, wA.w, src);
        const int sb1 = __shfl_sync(0xffffffffu, wB.x, src) + b1_add, sb2 = __shfl_sync(0xffffffffu, wB.y, src);
        const int snk = __shfl_sync(0xffffffffu, wB.z, src), sn = __shfl_sync(0xffffffffu, wB.w, src);
        const int snb = __shfl_sync(0xffffffffu, wC.x, src);
        const int nsets = MN ? max(1, __shfl_sync(0xffffffffu, wC.z, src)) : 1;
        const int sb2x = __shfl_sync(0xffffffffu, wC.w, src);
        // B boxes this CTA loads: K-major boxes of bn/CG rows, MN-major boxes of 32 columns
        const uint32_t b_box_bytes = MN ? mn_box_bytes : (uint32_t)(p.bn / CG) * 128u;
        // MN-major: this CTA's half of the N columns starts at column rank * n / CG and is loaded as mnb-column boxes
        // (the last one may run past the half: harmless extra columns)
        const int nb = MN ? (sn / CG + mnb - 1) / mnb : snb;
        const int b_col0 = MN ? (int)rank * (sn / CG) : 0;
        const int b_row0 = MN ? 0 : (int)rank * (sn / CG);          // first B row
        (void)snb;
        const uint32_t tx_cta = npl * (TC_PLANE_A + (uint32_t)(nsets * nb) * b_box_bytes);
        const int a_boxes = TC_BM / mnb;  // MN-major A: boxes of mnb columns
        for (int kb = 0; kb < snk; kb++) {
          { const long long t0 = p.timing ? clock64() : 0; mbar_wait(empty_bar(stage), phase ^ 1u); if (p.timing) tm_wait_empty += clock64() - t0; }
          const uint32_t fb_local = full_bar(stage);
          const uint32_t fb = CG == 2 ? mapa_shared(fb_local, 0) : fb_local;
          if (leader) mbar_expect_tx_u(pred, fb_local, tx_cta * CG);
          const uint32_t a_s = base + (uint32_t)stage * stage_bytes;
          const uint32_t b_s = a_s + npl * TC_PLANE_A;
          for (int pl = 0; pl < (int)npl; pl++) {
            if (!MN) {
              tma_load_4d_u<CG>(pred, a_s + pl * TC_PLANE_A, &tmA, fb, sa0 + kb * kbe, sa1, sa2, pl);
              for (int j = 0; j < nb; j++)
                tma_load_4d_u<CG>(pred, b_s + pl * b_plane + j * b_box_bytes, &tmB, fb, sb0 + kb * kbe,
                                  sb1 + b_row0 + j * (p.bn / CG), sb2, pl);
            } else {
              for (int i = 0; i < a_boxes; i++)
                tma_load_4d_u<CG>(pred, a_s + pl * TC_PLANE_A + i * mn_box_bytes, &tmA, fb, sa0 + i * mnb, sa1 + kb * kbe, sa2, pl);
              for (int sx = 0; sx < nsets; sx++) {
                const int b2s = sx == 0 ? sb2 : ((sb2x >> (8 * (sx - 1))) & 0xff);
                for (int j = 0; j < nb; j++)
                  tma_load_4d_u<CG>(pred, b_s + pl * b_plane + (sx * nb + j) * mn_box_bytes, &tmB, fb, sb0 + b_col0 + j * mnb,
                                    sb1 + kb * kbe, b2s, pl);
              }
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    if (p.timing && lane == 0) p.timing[blockIdx.x * 8 + 1] = (unsigned long long)tm_wait_empty;
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; warp-uniform, elected lane issues) =====================
    if (leader) {
      const uint32_t pred = elect_one_pred();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t gchunk = 0;  // chunks issued so far, all tiles
      long long tm_wait_full = 0, tm_wait_tempty = 0;
      // smem matrix descriptors: only the 14-bit start address field changes between operands
      //   K-major : LBO 16 B, SBO 1024 B, SWIZZLE_128B;  MN-major: LBO 4096 B, SBO 512 B, SWIZZLE_128B_BASE32B
      //             16-bit MN-major: LBO = one box (8192 B), SBO 1024 B (8 K rows of 128 B), plain SWIZZLE_128B
      const bool op16 = p.op != OP_TF32X3;
      const uint32_t mn_sbo = op16 ? 1024u : 512u, mn_layout = op16 ? 2u : 1u;
      const uint32_t desc_hi32 = (MN ? (mn_sbo >> 4) : (1024u >> 4)) | (1u << 14) | ((MN ? mn_layout : 2u) << 29);
      const uint32_t desc_lo_fixed = (MN ? (mn_box_bytes >> 4) : (16u >> 4)) << 16;
      // descriptor address units per K step (8 fp32 / 16 16-bit elements = 32 bytes of a K-major row, or that many
      // 128-byte K rows of an MN-major box)
      const uint32_t ks_step = MN ? ((op16 ? 2048u : 1024u) >> 4) : (32u >> 4);
      const uint32_t ab_fmt = p.op == OP_TF32X3 ? 2u : (p.op == OP_F16X3 ? 0u : 1u);
      int4 hdr0 = make_int4(0, 0, 0, 0), hdr1 = make_int4(0, 0, 0, 0);
      if (group < nslots) {
        const int4* hp = reinterpret_cast<const int4*>(p.tiles + (size_t)group * CG);
        hdr0 = __ldg(hp); hdr1 = __ldg(hp + 1);
      }
      for (int slot = group; slot < nslots; slot += ngroups) {
        const int seg_begin = uni(hdr0.x), seg_count = uni(hdr0.y), total_kb = uni(hdr1.z);
        if (slot + ngroups < nslots) {
          const int4* hp = reinterpret_cast<const int4*>(p.tiles + (size_t)(slot + ngroups) * CG);
          hdr0 = __ldg(hp); hdr1 = __ldg(hp + 1);
        }
        int2 sN[4];  // (nk, n_mma) of segment lane + 32 i
        int2 sK[4];  // (ks_last, nsets)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          sN[i] = make_int2(0, 0);
          sK[i] = make_int2(0, 0);
          if (lane + 32 * i < seg_count) {
            sN[i] = __ldg(reinterpret_cast<const int2*>(&p.segs[seg_begin + lane + 32 * i].nk));
            const int4 w3 = __ldg(reinterpret_cast<const int4*>(&p.segs[seg_begin + lane + 32 * i].nb));  // nb ks_last nsets b2x
            sK[i] = make_int2(w3.y, w3.z);
          }
        }
        uint32_t accum = 0;
        int kcount = 0;  // K blocks of this tile issued so far
        uint32_t buf = 0;
        for (int si = 0; si < seg_count; si++) {
          const int rs = si >> 5, src = si & 31;
          const int2 wN = rs == 0 ? sN[0] : (rs == 1 ? sN[1] : (rs == 2 ? sN[2] : sN[3]));
          const int snk = __shfl_sync(0xffffffffu, wN.x, src);
          const int2 wK = rs == 0 ? sK[0] : (rs == 1 ? sK[1] : (rs == 2 ? sK[2] : sK[3]));
          const int sks = __shfl_sync(0xffffffffu, wK.x, src);
          const int nsets = MN ? max(1, __shfl_sync(0xffffffffu, wK.y, src)) : 1;
          const int sn = __shfl_sync(0xffffffffu, wN.y, src);
          const uint32_t idesc = make_idesc(sn, MN, CG, ab_fmt);
          // B sets of the segment (MN-major): boxes of set s start nb boxes behind those of set s - 1
          const uint32_t set_step = MN ? (uint32_t)((sn / CG + mnb - 1) / mnb) * (mn_box_bytes >> 4) : 0u;
          for (int kb = 0; kb < snk; kb++, kcount++) {
            if (kcount % CH == 0) {  // new chunk: its TMEM buffer must have been drained
              buf = gchunk & 1u;
              { const long long t0 = p.timing ? clock64() : 0; mbar_wait(tempty_bar(buf), ((gchunk >> 1) & 1u) ^ 1u); if (p.timing) tm_wait_tempty += clock64() - t0; }
              tc_fence_after();
              accum = 0;
            }
            const uint32_t tmem_d = tmem_base + buf * TC_MAX_COLS;
            { const long long t0 = p.timing ? clock64() : 0; mbar_wait(full_bar(stage), phase); if (p.timing) tm_wait_full += clock64() - t0; }
            tc_fence_after();
            const uint32_t a_s = base + (uint32_t)stage * stage_bytes;
            const uint32_t a_hi = desc_lo_fixed | ((a_s & 0x3FFFFu) >> 4);
            const uint32_t a_lo = a_hi + (TC_PLANE_A >> 4);
            const uint32_t b_hi = a_hi + (npl * TC_PLANE_A >> 4);
            const uint32_t b_lo = b_hi + (b_plane >> 4);
            const int nks = (kb == snk - 1 && sks != 0) ? sks : 4;
            for (int sx = 0; sx < nsets; sx++) {  // one MMA group per B set; they share the A tile of the stage
              const uint32_t td = tmem_d + (uint32_t)(sx * sn);
              const uint32_t bh = b_hi + (uint32_t)sx * set_step, bl = b_lo + (uint32_t)sx * set_step;
              uint32_t acc = accum;
              if (p.op == OP_TF32X3) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                  if (ks >= nks) break;
                  const uint32_t o = (uint32_t)ks * ks_step;
                  mma_tf32_u<CG>(pred, td, a_lo + o, bh + o, desc_hi32, idesc, acc);
                  mma_tf32_u<CG>(pred, td, a_hi + o, bl + o, desc_hi32, idesc, 1);
                  mma_tf32_u<CG>(pred, td, a_hi + o, bh + o, desc_hi32, idesc, 1);
                  acc = 1;
                }
              } else if (p.op == OP_F16X3) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                  if (ks >= nks) break;
                  const uint32_t o = (uint32_t)ks * ks_step;
                  mma_f16_u<CG>(pred, td, a_lo + o, bh + o, desc_hi32, idesc, acc);
                  mma_f16_u<CG>(pred, td, a_hi + o, bl + o, desc_hi32, idesc, 1);
                  mma_f16_u<CG>(pred, td, a_hi + o, bh + o, desc_hi32, idesc, 1);
                  acc = 1;
                }
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                  if (ks >= nks) break;
                  mma_f16_u<CG>(pred, td, a_hi + (uint32_t)ks * ks_step, bh + (uint32_t)ks * ks_step, desc_hi32, idesc, acc);
                  acc = 1;
                }
              }
            }
            accum = 1;
            tc_commit_u<CG>(pred, empty_bar(stage));  // frees the stage (in both CTAs) once these MMAs have read it
            if (kcount % CH == CH - 1 || kcount == total_kb - 1) {
              tc_commit_u<CG>(pred, tfull_bar(buf));
              gchunk++;
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (p.timing && lane == 0) {
        p.timing[blockIdx.x * 8 + 2] = (unsigned long long)tm_wait_full;
        p.timing[blockIdx.x * 8 + 3] = (unsigned long long)tm_wait_tempty;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
    const int q = warp & 3;
    const int half = (warp - TC_EPI_WARP0) >> 2;
    const int row = q * 32 + lane;
    float* st = s_stage + (warp - TC_EPI_WARP0) * TC_STAGE_FLOATS;
    uint32_t gchunk = 0;
    long long tm_wait_tfull = 0, tm_store = 0, tm_drain = 0, tm_tiles = 0;
    const float oscale = p.out_scale * (p.out_scale_ptr ? __ldg(p.out_scale_ptr) : 1.f);
    const uint32_t tempty_remote0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    const uint32_t tempty_remote1 = CG == 2 ? mapa_shared(tempty_bar(1), 0) : tempty_bar(1);
    for (int slot = group; slot < nslots; slot += ngroups) {
      const TcTile* T = p.tiles + (size_t)slot * CG + rank;
      // everything the tile needs from its descriptor goes to registers up front: header words, the MMA width
      // breakpoints (lane i < nbp) and the column blocks (lane b < ncb); later lookups are ballots + shuffles
      const int4 h0 = __ldg(reinterpret_cast<const int4*>(T));
      const int4 h1 = __ldg(reinterpret_cast<const int4*>(T) + 1);
      const int4 h2 = __ldg(reinterpret_cast<const int4*>(T) + 2);
      const int m_valid = h0.z, ncb = h0.w, ld_out = h1.x, stats_row = h1.y, total_kb = h1.z, n_cols = h2.z, nbp = h2.w;
      int my_bp_kb = 0x7fffffff, my_bp_n = 0;
      if (lane < nbp) { my_bp_kb = __ldg(&T->bp_kb[lane]); my_bp_n = __ldg(&T->bp_n[lane]); }
      const int my_bp_n_first = __shfl_sync(0xffffffffu, my_bp_n, 0);  // MMA width at K block 0
      int my_tcol = 0x7fffffff, my_width = 0, my_scol = 0;
      int64_t my_off = 0;
      if (lane < ncb) {
        my_off = (int64_t)__ldg(reinterpret_cast<const long long*>(&T->cb[lane].out_off));
        const int4 w = __ldg(reinterpret_cast<const int4*>(&T->cb[lane].tcol));
        my_tcol = w.x; my_width = w.y; my_scol = w.z;
      }
      // the two column halves split the tile's accumulator columns (multiples of 32 each)
      const int hw = ((n_cols + 63) >> 6) << 5;
      const int cbase = half * hw;
      // ---- one 16-column slab of the tile: a16 = this lane's row, accumulator columns [c0, c0 + 16).  Each warp stages
      //      32 rows x 16 columns in smem and writes them back as row segments (4 lanes x float4 = 64 contiguous bytes
      //      per row, 8 rows per instruction), so global stores / read-modify-writes are sector-complete instead of
      //      one row per lane.  Column blocks map accumulator columns to output columns; they start at multiples of 16.
      const int r_lane = lane >> 2, c4 = (lane & 3) * 4;
      const bool all_rows = m_valid == TC_BM;
      int one_tcol = 0, one_width = 0;
      int64_t one_off = 0;
      if (ncb == 1) {  // the common case (1x1 convolutions, FC layers, every dgrad): no per-slab lookup
        one_tcol = __shfl_sync(0xffffffffu, my_tcol, 0);
        one_width = __shfl_sync(0xffffffffu, my_width, 0);
        one_off = __shfl_sync(0xffffffffu, my_off, 0);
      }
      // TMA slabs: (column, row) of the tile's first element in the 2-D view of `out`
      //      (rows past m_valid and columns past the block's width must fall outside the view, where the unit clips
      //      them; phantom tiles and odd tiles in the middle of a tensor keep the plain path)
      bool use_tma = false;
      int tma_row0 = 0, tma_col0 = 0;
      if (p.tma_out != 0 && ncb == 1 && EPI == EPI_ATOMIC && m_valid > 0) {
        tma_row0 = (int)(one_off / ld_out);
        tma_col0 = (int)(one_off - (int64_t)tma_row0 * ld_out);
        use_tma = (m_valid == TC_BM || tma_row0 + m_valid == p.out_rows) &&
                  ((one_width & 15) == 0 || tma_col0 + one_width == p.out_cols) && one_tcol == 0;
      }
      const uint32_t st_u32 = smem_u32(st);
      auto store_slab = [&](const float* a16, const int c0) {
          // column block holding the slab: blocks are listed by increasing tcol (one block: uniform registers)
          int cb_tcol = one_tcol, cb_width = one_width;
          int64_t cb_off = one_off;
          if (ncb != 1) {
            const uint32_t cbm = __ballot_sync(0xffffffffu, my_tcol <= c0);
            if (cbm == 0) return;
            const int cb_i = 31 - __clz((int)cbm);
            cb_tcol = __shfl_sync(0xffffffffu, my_tcol, cb_i);
            cb_width = __shfl_sync(0xffffffffu, my_width, cb_i);
            cb_off = __shfl_sync(0xffffffffu, my_off, cb_i);
          } else if (c0 < cb_tcol) {
            return;
          }
          if (c0 >= cb_tcol + cb_width) return;  // padding columns of the block
          if (p.tma_out) {  // the TMA unit may still be reading the staging rows of this warp's previous slab
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          const int ocol = c0 + c4 - cb_tcol;             // output column of this lane's 4 values
          float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
          if (use_tma) {
            // ---- the slab leaves through the TMA unit: dense 32 x 64-byte rows in the SWIZZLE_64B pattern (16-byte
            //      chunk index ^ (row >> 1) & 3: conflict-free for the row-per-lane writes), one f32 add reduction
            //      (cp.reduce.async.bulk.tensor) per slab.  The warp never waits on a global atomic; rows and columns
            //      outside the tensor are clipped by the unit.
            const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
            for (int k = 0; k < 4; k++)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_u32 + (uint32_t)lane * 64u + (((uint32_t)k ^ sw) << 4)),
                           "f"(a16[4 * k] * oscale), "f"(a16[4 * k + 1] * oscale), "f"(a16[4 * k + 2] * oscale),
                           "f"(a16[4 * k + 3] * oscale) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              const int tc = tma_col0 + (c0 - cb_tcol), tr = tma_row0 + q * 32;
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(tc), "r"(tr), "r"(st_u32) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else {
#pragma unroll
          for (int k = 0; k < 4; k++)
            *reinterpret_cast<float4*>(st + lane * TC_STAGE_LD + 4 * k) =
                make_float4(a16[4 * k] * oscale, a16[4 * k + 1] * oscale,
                            a16[4 * k + 2] * oscale, a16[4 * k + 3] * oscale);
          __syncwarp();
          float* const obase = p.out + cb_off + (int64_t)(q * 32 + r_lane) * ld_out + ocol;
          const int64_t ostep = (int64_t)8 * ld_out;     // rows advance by 8 per pass
          // warp-uniform fast path: whole slab valid, every lane's 4 values 16-byte aligned in the output
          const bool fast = all_rows && (c0 + 16 <= cb_tcol + cb_width) && ((ld_out & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.out + cb_off + (c0 - cb_tcol)) & 15) == 0);
          if (fast) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) v[i] = *reinterpret_cast<const float4*>(st + (i * 8 + r_lane) * TC_STAGE_LD + c4);
            if (EPI == EPI_ACCUM) {
              float4 o[4];
#pragma unroll
              for (int i = 0; i < 4; i++) o[i] = *reinterpret_cast<const float4*>(obase + i * ostep);
#pragma unroll
              for (int i = 0; i < 4; i++) { v[i].x += o[i].x; v[i].y += o[i].y; v[i].z += o[i].z; v[i].w += o[i].w; }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
              if (EPI == EPI_ATOMIC) atomicAdd(reinterpret_cast<float4*>(obase + i * ostep), v[i]);
              else *reinterpret_cast<float4*>(obase + i * ostep) = v[i];
              if (EPI == EPI_STORE) {
                s1[0] += v[i].x; s1[1] += v[i].y; s1[2] += v[i].z; s1[3] += v[i].w;
                s2[0] += v[i].x * v[i].x; s2[1] += v[i].y * v[i].y; s2[2] += v[i].z * v[i].z; s2[3] += v[i].w * v[i].w;
              }
            }
          } else {
            const int nval = min(4, cb_width - ocol);  // valid values of this lane (<= 0: none)
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float4 v4 = *reinterpret_cast<const float4*>(st + (i * 8 + r_lane) * TC_STAGE_LD + c4);
              const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
              if (q * 32 + i * 8 + r_lane < m_valid) {
                float* optr = obase + i * ostep;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                  if (t < nval) {
                    float x = vv[t];
                    if (EPI == EPI_ATOMIC) atomicAdd(optr + t, x);
                    else {
                      if (EPI == EPI_ACCUM) x += optr[t];
                      optr[t] = x;
                    }
                    s1[t] += x;
                    s2[t] += x * x;
                  }
                }
              }
            }
          }
          }
          if (EPI == EPI_STORE && p.stats) {
            // the 8 lanes that share this lane's 4 columns (lane >> 2 = row lane) hold 8 partial sums each
            // (s1[0..3], s2[0..3]); a halving butterfly leaves each of them with ONE fully reduced value
            const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
            float w4[4], w2[2], w1;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const float keep = u16 ? s2[k] : s1[k], send = u16 ? s1[k] : s2[k];
              w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
              const float keep = u8 ? w4[2 + k] : w4[k], send = u8 ? w4[k] : w4[2 + k];
              w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
              const float keep = u4 ? w2[1] : w2[0], send = u4 ? w2[0] : w2[1];
              w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            const int which = (u8 ? 2 : 0) + (u4 ? 1 : 0);  // column inside the lane group's 4; u16 selects s1 / s2
            s_part[((u16 ? 1 : 0) * 4 + q) * TC_MAX_COLS + c0 + c4 + which] = w1;
          }
          __syncwarp();  // the staging rows are rewritten by the next slab
      };
      // ---- 32 accumulator columns of the lane's row at once, for tiles with one column block: the warp stages 16 rows x
      //      128 bytes per pass and writes FOUR ROWS x 128 BYTES per instruction -- whole lines.  The 16-column slabs
      //      above touch eight half lines per instruction, and the store path's cost follows the number of lines an
      //      instruction touches, not its bytes (a lane-per-row 32-byte store form measured slower still).
      //      Returns false when the block is not whole (edge tiles, odd widths): the caller falls back to two slabs.
      auto store_block32 = [&](const float* a32, const int c0) -> bool {
        if (ncb != 1 || use_tma || !all_rows || c0 < one_tcol || c0 + 32 > one_tcol + one_width || (ld_out & 3) != 0 ||
            (reinterpret_cast<uintptr_t>(p.out + one_off + (c0 - one_tcol)) & 15) != 0)
          return false;
        if (p.tma_out) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        constexpr int LD32 = 36;  // floats per staged row (32 + pad: conflict-free 16-byte accesses both ways)
        const int r4 = lane >> 3, c8 = (lane & 7) * 4;
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if ((lane >> 4) == h) {
#pragma unroll
            for (int k = 0; k < 8; k++)
              *reinterpret_cast<float4*>(st + (lane & 15) * LD32 + 4 * k) =
                  make_float4(a32[4 * k] * oscale, a32[4 * k + 1] * oscale, a32[4 * k + 2] * oscale, a32[4 * k + 3] * oscale);
          }
          __syncwarp();
          float* const obase = p.out + one_off + (int64_t)(q * 32 + h * 16 + r4) * ld_out + (c0 - one_tcol) + c8;
          const int64_t ostep = (int64_t)4 * ld_out;
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; i++) v[i] = *reinterpret_cast<const float4*>(st + (i * 4 + r4) * LD32 + c8);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            if (EPI == EPI_ACCUM) {
              const float4 o = *reinterpret_cast<const float4*>(obase + i * ostep);
              v[i].x += o.x; v[i].y += o.y; v[i].z += o.z; v[i].w += o.w;
            }
            if (EPI == EPI_ATOMIC) atomicAdd(reinterpret_cast<float4*>(obase + i * ostep), v[i]);
            else *reinterpret_cast<float4*>(obase + i * ostep) = v[i];
            if (EPI == EPI_STORE) {
              s1[0] += v[i].x; s1[1] += v[i].y; s1[2] += v[i].z; s1[3] += v[i].w;
              s2[0] += v[i].x * v[i].x; s2[1] += v[i].y * v[i].y; s2[2] += v[i].z * v[i].z; s2[3] += v[i].w * v[i].w;
            }
          }
          __syncwarp();  // the staging rows are rewritten by the next pass
        }
        if (EPI == EPI_STORE && p.stats) {
          // the 4 lanes that share this lane's 4 columns (lane >> 3 = row lane) hold 8 partial sums each; a halving
          // butterfly leaves each of them with TWO fully reduced values
          const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0;
          float w4[4], w2[2];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float keep = u16 ? s2[k] : s1[k], send = u16 ? s1[k] : s2[k];
            w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const float keep = u8 ? w4[2 + k] : w4[k], send = u8 ? w4[k] : w4[2 + k];
            w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          float* sp = s_part + ((u16 ? 1 : 0) * 4 + q) * TC_MAX_COLS + c0 + c8 + (u8 ? 2 : 0);
          sp[0] = w2[0];
          sp[1] = w2[1];
        }
        return true;
      };
      const int nchunks = (total_kb + CH - 1) / CH;
      if (nchunks == 1) {
        // ---- single-chunk tiles (K <= chunk_kb blocks: the 1x1 convolutions): straight from TMEM to the slabs, 32
        //      columns at a time, no register accumulator; the TMEM buffer is handed back after the last read
        const uint32_t buf = gchunk & 1u;
        const int n_c = my_bp_n_first;
        long long t0 = p.timing ? clock64() : 0;
        mbar_wait(tfull_bar(buf), (gchunk >> 1) & 1u);
        if (p.timing) { const long long t1 = clock64(); tm_wait_tfull += t1 - t0; t0 = t1; }
        tc_fence_after();
        gchunk++;
        int last_g = -1;
#pragma unroll
        for (int g = 0; g < 4; g++)
          if (g * 32 < hw && cbase + g * 32 < n_c) last_g = g;
        if (last_g < 0) {  // this warp's column half is empty: only the hand-back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t tb = buf ? tempty_remote1 : tempty_remote0;
            if (CG == 2) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tb) : "memory");
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb) : "memory");
          }
        }
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const int tc0 = cbase + g * 32;
          if (g * 32 < hw && tc0 < n_c) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_MAX_COLS + tc0), v);
            if (g == last_g) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                const uint32_t tb = buf ? tempty_remote1 : tempty_remote0;
                if (CG == 2)  // TMEM reads are complete (tcgen05.wait::ld): no memory ordering needed, skip the release fence
                  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tb) : "memory");
                else
                  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb) : "memory");
              }
            }
            if (tc0 + 32 > n_c) {  // columns the MMAs never wrote
#pragma unroll
              for (int j = 0; j < 32; j++) v[j] = (tc0 + j < n_c) ? v[j] : 0.f;
            }
            if (!store_block32(v, tc0)) {
              if (tc0 < n_cols) store_slab(v, tc0);
              if (tc0 + 16 < n_cols) store_slab(v + 16, tc0 + 16);
            }
          }
        }
        if (p.timing) tm_store += clock64() - t0;
      } else {
      float acc[TC_MAX_COLS / 2];
#pragma unroll
      for (int i = 0; i < TC_MAX_COLS / 2; i++) acc[i] = 0.f;
      {
        for (int c = 0; c < nchunks; c++, gchunk++) {
          const uint32_t buf = gchunk & 1u;
          // columns the chunk accumulated = MMA width at its first K block (widths never increase inside a tile)
          const uint32_t bpm = __ballot_sync(0xffffffffu, my_bp_kb <= c * CH);
          const int n_c = __shfl_sync(0xffffffffu, my_bp_n, 31 - __clz((int)(bpm | 1u)));
          long long t0 = p.timing ? clock64() : 0;
          mbar_wait(tfull_bar(buf), (gchunk >> 1) & 1u);
          if (p.timing) { const long long t1 = clock64(); tm_wait_tfull += t1 - t0; t0 = t1; }
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int tc0 = cbase + g * 32;
            if (g * 32 < hw && tc0 < n_c) {
              float v[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_MAX_COLS + tc0), v);
              if (tc0 + 32 <= n_c) {
#pragma unroll
                for (int j = 0; j < 32; j++) acc[g * 32 + j] += v[j];
              } else {
#pragma unroll
                for (int j = 0; j < 32; j++) acc[g * 32 + j] += (tc0 + j < n_c) ? v[j] : 0.f;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t tb = buf ? tempty_remote1 : tempty_remote0;
            if (CG == 2)  // TMEM reads are complete (tcgen05.wait::ld): no memory ordering needed, skip the release fence
              asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(tb) : "memory");
            else
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb) : "memory");
          }
          if (p.timing) tm_drain += clock64() - t0;
        }
      }
      const long long t_store0 = p.timing ? clock64() : 0;
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int tc0 = cbase + g * 32;
        if (g * 32 >= hw || tc0 >= n_cols) continue;  // warp-uniform
        if (!store_block32(acc + g * 32, tc0)) {
          store_slab(acc + g * 32, tc0);
          if (tc0 + 16 < n_cols) store_slab(acc + g * 32 + 16, tc0 + 16);
        }
      }
      if (p.timing) tm_store += clock64() - t_store0;
      }
      if (EPI == EPI_STORE && p.stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 32 * TC_EPI_WARP0;
        float* srow = p.stats + (size_t)stats_row * 2 * p.stats_ld;
        for (int b = 0; b < (m_valid > 0 ? ncb : 0); b++) {  // phantom tiles (m_valid == 0) own no statistics row
          const int b_tcol = __shfl_sync(0xffffffffu, my_tcol, b), b_width = __shfl_sync(0xffffffffu, my_width, b);
          const int b_scol = __shfl_sync(0xffffffffu, my_scol, b);
          const int c = t - b_tcol;
          if (c >= 0 && c < b_width) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; w++) {
              a1 += s_part[(0 * 4 + w) * TC_MAX_COLS + t];
              a2 += s_part[(1 * 4 + w) * TC_MAX_COLS + t];
            }
            srow[b_scol + c] = a1;
            srow[p.stats_ld + b_scol + c] = a2;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // s_part is reused by the next tile
      }
      if (p.timing) tm_tiles++;
    }
    // the staging rows must outlive the TMA unit's reads (global completion is the kernel boundary's business)
    if (p.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (p.timing && warp == TC_EPI_WARP0 && lane == 0) {
      p.timing[blockIdx.x * 8 + 4] = (unsigned long long)tm_wait_tfull;
      p.timing[blockIdx.x * 8 + 5] = (unsigned long long)tm_store;
      p.timing[blockIdx.x * 8 + 6] = (unsigned long long)tm_drain;
      p.timing[blockIdx.x * 8 + 7] = (unsigned long long)tm_tiles;
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (p.timing && threadIdx.x == 0) p.timing[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - tm_kernel0);
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps and launch
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// fp32 tensor of rank 4 (dim 0 innermost, contiguous).  strides_elems: element strides of dims 1..3.
// mn_major: boxes feed MN-major tf32 operands -> SWIZZLE_128B_ATOM_32B, else SWIZZLE_128B.
// op = OP_*: fp32 elements (MN-major boxes feed tf32 operands -> SWIZZLE_128B_ATOM_32B) or 16-bit elements (plain
// SWIZZLE_128B for both majors)
inline int make_map(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides_elems[3],
                    const uint32_t box[4], bool mn_major = false, int op = OP_TF32X3) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; i++) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 3; i++) {
    gs[i] = strides_elems[i] * (uint64_t)op_esize(op);
    if (gs[i] % 16) return fail(HYP_E_INVALID, "tensor map stride is not a multiple of 16 bytes");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(HYP_E_INVALID, "tensor map base is not 16-byte aligned");
  const CUresult r = fn(map, op == OP_TF32X3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                        const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        (mn_major && op == OP_TF32X3) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return HYP_OK;
}

// 2-D fp32 view of a GEMM output for the epilogue's TMA slabs: `cols` valid columns (the unit clips what lies beyond),
// rows `ld` floats apart, boxes of 32 rows x 16 columns in the SWIZZLE_64B smem pattern
inline int make_out_map(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  if ((ld * 4) % 16 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(HYP_E_INVALID, "output tensor map: rows are not 16-byte aligned");
  cuuint64_t gd[2] = {cols, rows}, gs[1] = {ld * 4};
  cuuint32_t bx[2] = {16, 32}, es[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled (output map) failed with CUresult " + std::to_string((int)r));
  return HYP_OK;
}

constexpr int TC_DEFAULT_CHUNK_KB = 8;  // 96 MMAs per TMEM chunk: error ~2e-6 of max|D|; 12 already fails the 3e-6 building-block test

static thread_local const char* g_tc_timing_tag = nullptr;  // label of the next launch in HYP_TC_TIMING output

// fills n_cols and the MMA-width breakpoints of every tile; checks the limits the kernel relies on
inline int tc_finalize_tiles(std::vector<TcTile>& tiles, const std::vector<TcSeg>& segs) {
  static_assert(sizeof(TcSeg) == 48 && sizeof(TcColBlock) == 32 && offsetof(TcTile, bp_kb) == 48 &&
                    offsetof(TcTile, cb) == 48 + 8 * TC_MAX_BP && offsetof(TcSeg, nk) == 24 && offsetof(TcSeg, nb) == 32 &&
                    offsetof(TcSeg, nsets) == 40,
                "descriptor layouts are read as raw words by the kernel");
  for (TcTile& t : tiles) {
    if (t.seg_count > TC_MAX_SEGS) return fail(HYP_E_UNSUPPORTED, "tc gemm: more than 128 segments in one tile");
    t.n_cols = 0;
    t.nbp = 0;
    int kb = 0, last = -1;
    for (int i = 0; i < t.seg_count; i++) {
      const TcSeg& sg = segs[t.seg_begin + i];
      const int width = sg.n_mma * std::max(1, sg.nsets);  // accumulator columns the segment's MMA groups cover
      if (sg.nsets > TC_MAX_SETS || width > TC_MAX_COLS) return fail(HYP_E_INVALID, "tc gemm: too many B sets in one segment");
      t.n_cols = std::max(t.n_cols, width);
      if (width != last) {
        if (last >= 0 && width > last) return fail(HYP_E_INVALID, "tc gemm: segment widths must not increase");
        if (t.nbp == TC_MAX_BP) return fail(HYP_E_UNSUPPORTED, "tc gemm: too many distinct MMA widths in one tile");
        t.bp_kb[t.nbp] = kb;
        t.bp_n[t.nbp] = width;
        t.nbp++;
        last = width;
      }
      kb += sg.nk;
    }
    for (int b = 1; b < t.ncb; b++)
      if (t.cb[b].tcol <= t.cb[b - 1].tcol) return fail(HYP_E_INVALID, "tc gemm: column blocks must be listed by increasing tcol");
  }
  return HYP_OK;
}

inline int tc_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    // HYP_TC_SMS: SMs the persistent GEMMs (and the grids sized after them) may occupy.  A GEMM CTA takes a whole SM
    // (64 K registers), so a collective running beside the backward pass cannot share one: data-parallel runs can
    // leave its CTAs a few SMs instead of letting them displace GEMM CTAs into a second wave.
    if (const char* e = getenv("HYP_TC_SMS")) {
      const int v = atoi(e);
      if (v >= 2 && v <= sms) sms = v & ~1;
    }
  }
  return sms;
}

// CG = 2: tiles 2i, 2i+1 form a pair sharing its segment list (ntiles even); tmB boxes hold bn/2 rows
template <bool MN, int CG, int EPI>
inline int launch_tc_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, int ntiles, cudaStream_t st,
                         const CUtensorMap* tmO);

// tmO: optional 2-D view of p.out (make_out_map); with it, EPI_ATOMIC tiles of one column block accumulate through the
// TMA unit (measured: plain stores are better off as 128-byte row blocks from the warps, profiles/r02_store_paths.md)
template <bool MN, int CG>
inline int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, int ntiles, cudaStream_t st,
                     const CUtensorMap* tmO = nullptr) {
  if (p.epi == EPI_STORE) return launch_tc_epi<MN, CG, EPI_STORE>(tmA, tmB, p, ntiles, st, nullptr);
  if (p.epi == EPI_ACCUM) return launch_tc_epi<MN, CG, EPI_ACCUM>(tmA, tmB, p, ntiles, st, nullptr);
  return launch_tc_epi<MN, CG, EPI_ATOMIC>(tmA, tmB, p, ntiles, st, tmO);
}

template <bool MN, int CG, int EPI>
inline int launch_tc_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, int ntiles, cudaStream_t st,
                         const CUtensorMap* tmO) {
  p.tma_out = tmO ? 1 : 0;
  if (ntiles <= 0) return HYP_OK;
  if (ntiles % CG) return fail(HYP_E_INVALID, "tc gemm: tile count is not a multiple of the CTA group size");
  if (p.b_rows % (8 * CG)) return fail(HYP_E_INVALID, "tc gemm: B rows must be a multiple of 8 per CTA");
  if (p.stages <= 0) p.stages = tc_pick_stages(p.b_rows / CG, op_planes(p.op));
  if (p.out_scale == 0.f) p.out_scale = 1.f;
  if (p.chunk_kb <= 0) {
    static const int env_chunk = getenv("HYP_TC_CHUNK_KB") ? atoi(getenv("HYP_TC_CHUNK_KB")) : 0;  // experiments only
    p.chunk_kb = env_chunk > 0 ? env_chunk : TC_DEFAULT_CHUNK_KB;
  }
  if (p.stages < 2) return fail(HYP_E_INVALID, "tc gemm: B tile too large for two pipeline stages");
  p.ntiles = ntiles;
  const size_t smem = tc_smem_bytes(p.b_rows / CG, p.stages, op_planes(p.op));
  static bool attr_set = false;
  if (!attr_set) {
    HYP_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MN, CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_set = true;
  }
  const int groups = std::min(ntiles / CG, tc_sm_count() / CG);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG > 1 ? 1 : 0;
  static const bool timing_on = getenv("HYP_TC_TIMING") != nullptr;
  unsigned long long* tbuf = nullptr;
  if (timing_on) {
    HYP_CUDA(cudaMalloc(&tbuf, (size_t)groups * CG * 8 * sizeof(unsigned long long)));
    HYP_CUDA(cudaMemsetAsync(tbuf, 0, (size_t)groups * CG * 8 * sizeof(unsigned long long), st));
    p.timing = tbuf;
  }
  HYP_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<MN, CG, EPI>, tmA, tmB, tmO ? *tmO : tmA, p));
  HYP_LAUNCHED();
  if (timing_on) {
    std::vector<unsigned long long> h((size_t)groups * CG * 8);
    HYP_CUDA(cudaStreamSynchronize(st));
    HYP_CUDA(cudaMemcpy(h.data(), tbuf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(tbuf);
    double s[8] = {0}, mx0 = 0, mn0 = 1e30;
    const int n = groups * CG;
    for (int c = 0; c < n; c++) {
      for (int k = 0; k < 8; k++) s[k] += (double)h[(size_t)c * 8 + k];
      mx0 = std::max(mx0, (double)h[(size_t)c * 8]);
      mn0 = std::min(mn0, (double)h[(size_t)c * 8]);
    }
    const int nl = n / CG;  // MMA issuers
    fprintf(stderr,
            "[tc_timing] %-28s mn=%d cg=%d ctas=%d tiles=%d stages=%d brows=%d | kcyc total avg %.0f min %.0f max %.0f | "
            "prod_wait_empty %.0f | mma_wait_full %.0f mma_wait_tempty %.0f | epi_wait_tfull %.0f epi_drain %.0f epi_store %.0f | "
            "tiles/cta %.1f\n",
            g_tc_timing_tag ? g_tc_timing_tag : "?", (int)MN, CG, n, ntiles, p.stages, p.b_rows, s[0] / n / 1e3, mn0 / 1e3,
            mx0 / 1e3, s[1] / n / 1e3, s[2] / nl / 1e3, s[3] / nl / 1e3, s[4] / n / 1e3, s[6] / n / 1e3, s[5] / n / 1e3,
            s[7] / n);
  }
  return HYP_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) TF32 planes.  raw_hi = 1 keeps the unrounded value in plane 0 and relies on
// the tensor core ignoring the 13 low mantissa bits (probe only).
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - hi);
}
// 16-bit operand formats.  OP_F16X3: hi = fp16(x), lo = fp16(x - hi) — 22 significand bits as long as lo stays a
// normal fp16 number (|x| >= 2^-3); below that the absolute error is bounded by half the subnormal spacing, 2^-25, so
// tensors are scaled to O(1)..O(2^10) magnitudes before the split (see hyp_tc_engine.cuh).  OP_BF16: one bf16 plane.
__device__ __forceinline__ void split_f16(float x, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(x);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
}
__device__ __forceinline__ uint16_t to_bf16(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
// the planes of 4 consecutive elements, 8 bytes per plane; `planes` points at element 0 of plane 0
__device__ __forceinline__ void store_op16x4(uint16_t* planes, size_t plane_stride, int op, float v0, float v1, float v2, float v3) {
  if (op == OP_F16X3) {
    uint16_t h[4], l[4];
    split_f16(v0, h[0], l[0]); split_f16(v1, h[1], l[1]); split_f16(v2, h[2], l[2]); split_f16(v3, h[3], l[3]);
    *reinterpret_cast<uint2*>(planes) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(planes + plane_stride) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
  } else {
    *reinterpret_cast<uint2*>(planes) = make_uint2(to_bf16(v0) | ((uint32_t)to_bf16(v1) << 16),
                                                   to_bf16(v2) | ((uint32_t)to_bf16(v3) << 16));
  }
}
__device__ __forceinline__ void store_op16(uint16_t* planes, size_t plane_stride, int op, float v) {
  if (op == OP_F16X3) {
    uint16_t h, l;
    split_f16(v, h, l);
    planes[0] = h;
    planes[plane_stride] = l;
  } else {
    planes[0] = to_bf16(v);
  }
}
__global__ void split_planes16_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld_src, uint16_t* planes,
                                      size_t plane_stride, int ld_dst, int op, float scale) {
  const int64_t total = rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    store_op16(planes + i, plane_stride, op, c < cols ? src[r * ld_src + c] * scale : 0.f);
  }
}
__global__ void split_planes_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld_src, float* hi,
                                    float* lo, int ld_dst, int raw_hi) {
  const int64_t total = rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    float x = c < cols ? src[r * ld_src + c] : 0.f;
    float h, l;
    if (raw_hi) {
      h = x;
      l = tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
    } else {
      split_tf32(x, h, l);
    }
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace tc
}  // namespace hyp
