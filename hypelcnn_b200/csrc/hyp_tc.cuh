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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM chunks -> fp32 register sums -> global, per-tile BatchNorm
// partial sums).
#pragma once
#include <cuda.h>

#include <algorithm>

#include "hyp_common.cuh"

namespace hyp {
namespace tc {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_BM = 128;
constexpr int TC_KB = 32;                  // fp32 elements of K per pipeline stage
constexpr int TC_PLANE_A = TC_BM * 128;    // bytes of one A plane per stage
constexpr int TC_MAX_COLS = 256;
constexpr int TC_MAX_CB = 8;               // column blocks per tile           // accumulator columns per tile
constexpr int TC_SMEM_LIMIT = 226 * 1024;

struct alignas(16) TcSeg {
  int32_t a0, a1, a2;  // A box coordinates (tensor dims 0..2) at K block 0; dim 3 = plane
  int32_t b0, b1, b2;
  int32_t nk;          // K blocks of TC_KB
  int32_t n_mma;       // N of this segment's MMAs: multiple of 16 in [16, 256]
  int32_t nb;          // B boxes per K block and plane
  int32_t pad[3];
};

struct alignas(16) TcColBlock {
  int64_t out_off;    // element offset of (tile row 0, block column 0) inside `out`
  int32_t tcol;       // first accumulator column
  int32_t width;      // valid columns
  int32_t stats_col;  // first column in the statistics row
  int32_t pad;
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
  int32_t pad[2];
  TcColBlock cb[TC_MAX_CB];
};

enum { EPI_STORE = 0, EPI_ACCUM = 1, EPI_ATOMIC = 2 };

struct TcParams {
  const TcSeg* segs;
  const TcTile* tiles;
  float* out;
  float* stats;       // nullable: [stats rows][2][stats_ld] per-tile column sum / sum of squares
  int32_t stats_ld;
  int32_t epi;
  int32_t b_rows;     // B rows (n) reserved per stage and plane, multiple of 8
  int32_t bn;         // rows of one B box (K-major); MN-major boxes are 32 columns x 32 rows
  int32_t chunk_kb;   // K blocks the tensor core accumulates before the fp32 register merge
  int32_t stages;
  int32_t ntiles;     // tiles of this launch (multiple of the CTA group size)
};

// b_rows = B rows held by ONE CTA per stage and plane
inline size_t tc_smem_bytes(int b_rows, int stages) {
  return 1024 + (size_t)stages * (2 * TC_PLANE_A + 2 * (size_t)b_rows * 128) + 256 + 2 * 4 * TC_MAX_COLS * sizeof(float);
}
inline int tc_pick_stages(int b_rows) {
  const size_t fixed = 1024 + 256 + 2 * 4 * TC_MAX_COLS * sizeof(float);
  const size_t st = 2 * TC_PLANE_A + 2 * (size_t)b_rows * 128;
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
// instruction descriptor: D fp32, A/B tf32, M = 128 per CTA of the group
__device__ __forceinline__ uint32_t make_idesc_tf32(int n, bool mn_major, int cg = 1) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
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

template <bool MN, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t b_plane = (uint32_t)(p.b_rows / CG) * 128u;  // B rows held by THIS CTA
  const uint32_t stage_bytes = 2u * TC_PLANE_A + 2u * b_plane;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (size_t)p.stages * stage_bytes + 192);
  float* s_part = reinterpret_cast<float*>(gen + (size_t)p.stages * stage_bytes + 256);  // [2][4][TC_MAX_COLS]
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * p.stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (uint32_t)(2 * p.stages + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer (one per CTA) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int slot = group; slot < nslots; slot += ngroups) {
        const TcTile* T = p.tiles + (size_t)slot * CG + rank;
        const int seg_begin = T->seg_begin, seg_count = T->seg_count;
        const int a0_add = T->a0_add, a1_add = T->a1_add, b1_add = T->b1_add;
        for (int si = 0; si < seg_count; si++) {
          TcSeg sg = p.segs[seg_begin + si];
          sg.a0 += a0_add;
          sg.a1 += a1_add;
          sg.b1 += b1_add;
          // B boxes this CTA loads: K-major boxes of bn/CG rows, MN-major boxes of 32 columns
          const uint32_t b_box_bytes = MN ? 4096u : (uint32_t)(p.bn / CG) * 128u;
          const int nb = MN ? sg.nb / CG : sg.nb;
          const int b_first = MN ? (int)rank * nb : 0;                          // first 32-column box
          const int b_row0 = MN ? 0 : (int)rank * (sg.n_mma / CG);              // first B row
          const uint32_t tx_cta = 2u * TC_PLANE_A + 2u * (uint32_t)nb * b_box_bytes;
          for (int kb = 0; kb < sg.nk; kb++) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t fb_local = full_bar(stage);
            const uint32_t fb = CG == 2 ? mapa_shared(fb_local, 0) : fb_local;
            if (leader) mbar_expect_tx(fb_local, tx_cta * CG);
            const uint32_t a_s = base + (uint32_t)stage * stage_bytes;
            const uint32_t b_s = a_s + 2u * TC_PLANE_A;
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
              if (!MN) {
                tma_load_4d<CG>(a_s + pl * TC_PLANE_A, &tmA, fb, sg.a0 + kb * TC_KB, sg.a1, sg.a2, pl);
                for (int j = 0; j < nb; j++)
                  tma_load_4d<CG>(b_s + pl * b_plane + j * b_box_bytes, &tmB, fb, sg.b0 + kb * TC_KB,
                                  sg.b1 + b_row0 + j * (p.bn / CG), sg.b2, pl);
              } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                  tma_load_4d<CG>(a_s + pl * TC_PLANE_A + i * 4096, &tmA, fb, sg.a0 + i * 32, sg.a1 + kb * TC_KB, sg.a2, pl);
                for (int j = 0; j < nb; j++)
                  tma_load_4d<CG>(b_s + pl * b_plane + j * 4096, &tmB, fb, sg.b0 + (b_first + j) * 32, sg.b1 + kb * TC_KB,
                                  sg.b2, pl);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t gchunk = 0;  // chunks issued so far, all tiles
      for (int slot = group; slot < nslots; slot += ngroups) {
        const TcTile* T = p.tiles + (size_t)slot * CG;
        const int seg_begin = T->seg_begin, seg_count = T->seg_count, total_kb = T->total_kb;
        uint32_t accum = 0;
        int kcount = 0;  // K blocks of this tile issued so far
        uint32_t buf = 0;
        for (int si = 0; si < seg_count; si++) {
          const TcSeg sg = p.segs[seg_begin + si];
          const uint32_t idesc = make_idesc_tf32(sg.n_mma, MN, CG);
          for (int kb = 0; kb < sg.nk; kb++, kcount++) {
            if (kcount % CH == 0) {  // new chunk: its TMEM buffer must have been drained
              buf = gchunk & 1u;
              mbar_wait(tempty_bar(buf), ((gchunk >> 1) & 1u) ^ 1u);
              tc_fence_after();
              accum = 0;
            }
            const uint32_t tmem_d = tmem_base + buf * TC_MAX_COLS;
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_s = base + (uint32_t)stage * stage_bytes;
            const uint32_t b_s = a_s + 2u * TC_PLANE_A;
#pragma unroll
            for (int ks = 0; ks < TC_KB / 8; ks++) {
              uint64_t a_hi, a_lo, b_hi, b_lo;
              if (!MN) {  // rows of 128 B, 8-row groups 1024 B apart; K advances 32 B inside the swizzled row
                a_hi = make_smem_desc(a_s + ks * 32, 16, 1024);
                a_lo = make_smem_desc(a_s + TC_PLANE_A + ks * 32, 16, 1024);
                b_hi = make_smem_desc(b_s + ks * 32, 16, 1024);
                b_lo = make_smem_desc(b_s + b_plane + ks * 32, 16, 1024);
              } else {    // 32-column atoms 4096 B apart (LBO), 4-row K groups 512 B apart (SBO), 32-byte swizzle
                a_hi = make_smem_desc(a_s + ks * 1024, 4096, 512, 1);
                a_lo = make_smem_desc(a_s + TC_PLANE_A + ks * 1024, 4096, 512, 1);
                b_hi = make_smem_desc(b_s + ks * 1024, 4096, 512, 1);
                b_lo = make_smem_desc(b_s + b_plane + ks * 1024, 4096, 512, 1);
              }
              mma_tf32<CG>(tmem_d, a_lo, b_hi, idesc, accum);
              mma_tf32<CG>(tmem_d, a_hi, b_lo, idesc, 1);
              mma_tf32<CG>(tmem_d, a_hi, b_hi, idesc, 1);
              accum = 1;
            }
            tc_commit<CG>(empty_bar(stage));  // frees the stage (in both CTAs) once these MMAs have read it
            if (kcount % CH == CH - 1 || kcount == total_kb - 1) {
              tc_commit<CG>(tfull_bar(buf));
              gchunk++;
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    uint32_t gchunk = 0;
    const uint32_t tempty_remote0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    const uint32_t tempty_remote1 = CG == 2 ? mapa_shared(tempty_bar(1), 0) : tempty_bar(1);
    for (int slot = group; slot < nslots; slot += ngroups) {
      const TcTile* T = p.tiles + (size_t)slot * CG + rank;
      const int seg_begin = T->seg_begin, seg_count = T->seg_count, total_kb = T->total_kb;
      const int m_valid = T->m_valid, ld_out = T->ld_out, ncb = T->ncb;
      const bool rvalid = row < m_valid;
      float acc[TC_MAX_COLS / 2];
#pragma unroll
      for (int i = 0; i < TC_MAX_COLS / 2; i++) acc[i] = 0.f;
      {
        const int nchunks = (total_kb + CH - 1) / CH;
        int si = 0, left = 0;  // segment holding the chunk's first K block; K blocks of it still ahead
        int n_first = 0;
        if (seg_count > 0) { left = p.segs[seg_begin].nk; n_first = p.segs[seg_begin].n_mma; }
        for (int c = 0; c < nchunks; c++, gchunk++) {
          const uint32_t buf = gchunk & 1u;
          const int n_c = n_first;  // segments are ordered by non-increasing n_mma
          // advance the walker by CH K blocks
          int adv = CH;
          while (adv > 0 && si < seg_count) {
            if (left > adv) { left -= adv; adv = 0; }
            else {
              adv -= left;
              si++;
              if (si < seg_count) { left = p.segs[seg_begin + si].nk; n_first = p.segs[seg_begin + si].n_mma; }
            }
          }
          mbar_wait(tfull_bar(buf), (gchunk >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int tc0 = half * (TC_MAX_COLS / 2) + g * 32;
            if (tc0 < n_c) {
              float v[32];
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_MAX_COLS + tc0), v);
#pragma unroll
              for (int j = 0; j < 32; j++) acc[g * 32 + j] += (tc0 + j < n_c) ? v[j] : 0.f;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t tb = buf ? tempty_remote1 : tempty_remote0;
            if (CG == 2)
              asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tb) : "memory");
            else
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tb) : "memory");
          }
        }
      }
      // ---- write the tile: column blocks map accumulator columns to output columns ----
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int tc0 = half * (TC_MAX_COLS / 2) + g * 32;
        float s1 = 0.f, s2 = 0.f;  // lane l: column tc0 + l
        bool any = false;
        for (int b = 0; b < ncb; b++) {
          const TcColBlock cb = T->cb[b];
          const int lo = max(cb.tcol, tc0) - tc0, hi = min(cb.tcol + cb.width, tc0 + 32) - tc0;
          if (lo >= hi) continue;
          any = true;
          if (rvalid) {
            // output column of accumulator column tc0 + j is (tc0 + j - cb.tcol)
            float* orow = p.out + cb.out_off + (int64_t)row * ld_out + (tc0 - cb.tcol);
            const bool vec = ((cb.out_off | (int64_t)ld_out | (int64_t)(tc0 - cb.tcol)) & 3) == 0;
            if (p.epi == EPI_ATOMIC) {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (j >= lo && j < hi) atomicAdd(orow + j, acc[g * 32 + j]);
            } else {
              if (p.epi == EPI_ACCUM) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                  if (j >= lo && j < hi) acc[g * 32 + j] += orow[j];
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (vec && j >= lo && j + 3 < hi) {
                  *reinterpret_cast<float4*>(orow + j) =
                      make_float4(acc[g * 32 + j], acc[g * 32 + j + 1], acc[g * 32 + j + 2], acc[g * 32 + j + 3]);
                } else {
#pragma unroll
                  for (int t = 0; t < 4; t++)
                    if (j + t >= lo && j + t < hi) orow[j + t] = acc[g * 32 + j + t];
                }
              }
            }
          }
        }
        if (p.stats && __any_sync(0xffffffffu, any)) {
          float v[32], sq[32];
#pragma unroll
          for (int j = 0; j < 32; j++) {
            v[j] = rvalid ? acc[g * 32 + j] : 0.f;
            sq[j] = v[j] * v[j];
          }
          s1 = warp_transpose_sum(v, lane);
          s2 = warp_transpose_sum(sq, lane);
          s_part[(0 * 4 + q) * TC_MAX_COLS + tc0 + lane] = s1;
          s_part[(1 * 4 + q) * TC_MAX_COLS + tc0 + lane] = s2;
        }
      }
      if (p.stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 64;
        float* srow = p.stats + (size_t)T->stats_row * 2 * p.stats_ld;
        for (int b = 0; b < (m_valid > 0 ? ncb : 0); b++) {  // phantom tiles (m_valid == 0) own no statistics row
          const int c = t - T->cb[b].tcol;
          if (c >= 0 && c < T->cb[b].width) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; w++) {
              a1 += s_part[(0 * 4 + w) * TC_MAX_COLS + t];
              a2 += s_part[(1 * 4 + w) * TC_MAX_COLS + t];
            }
            srow[T->cb[b].stats_col + c] = a1;
            srow[p.stats_ld + T->cb[b].stats_col + c] = a2;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // s_part is reused by the next tile
      }
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
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
inline int make_map(CUtensorMap* map, const float* base, const uint64_t dims[4], const uint64_t strides_elems[3],
                    const uint32_t box[4], bool mn_major = false) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; i++) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < 3; i++) {
    gs[i] = strides_elems[i] * sizeof(float);
    if (gs[i] % 16) return fail(HYP_E_INVALID, "tensor map stride is not a multiple of 16 bytes");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(HYP_E_INVALID, "tensor map base is not 16-byte aligned");
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HYP_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return HYP_OK;
}

constexpr int TC_DEFAULT_CHUNK_KB = 4;

inline int tc_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// CG = 2: tiles 2i, 2i+1 form a pair sharing its segment list (ntiles even); tmB boxes hold bn/2 rows
template <bool MN, int CG>
inline int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams p, int ntiles, cudaStream_t st) {
  if (ntiles <= 0) return HYP_OK;
  if (ntiles % CG) return fail(HYP_E_INVALID, "tc gemm: tile count is not a multiple of the CTA group size");
  if (p.b_rows % (8 * CG)) return fail(HYP_E_INVALID, "tc gemm: B rows must be a multiple of 8 per CTA");
  if (p.stages <= 0) p.stages = tc_pick_stages(p.b_rows / CG);
  if (p.chunk_kb <= 0) p.chunk_kb = TC_DEFAULT_CHUNK_KB;
  if (p.stages < 2) return fail(HYP_E_INVALID, "tc gemm: B tile too large for two pipeline stages");
  p.ntiles = ntiles;
  const size_t smem = tc_smem_bytes(p.b_rows / CG, p.stages);
  static bool attr_set = false;
  if (!attr_set) {
    HYP_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<MN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
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
  HYP_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<MN, CG>, tmA, tmB, p));
  HYP_LAUNCHED();
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
