// Shared helpers: error convention, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/hypelcnn_b200.h"

namespace hyp {

extern thread_local std::string g_last_error;
extern thread_local int64_t g_launch_count;

inline int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define HYP_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) return ::hyp::fail(HYP_E_INVALID, std::string(__func__) + ": " + (msg)); \
  } while (0)

#define HYP_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::hyp::fail(HYP_E_CUDA, std::string(__func__) + ": " #expr ": " + cudaGetErrorString(_e)); \
  } while (0)

// every kernel launch goes through this so gpu_launches can be counted and launch errors
// surface at the call that caused them
#define HYP_LAUNCHED()                                                                      \
  do {                                                                                      \
    ::hyp::g_launch_count++;                                                                \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return ::hyp::fail(HYP_E_CUDA, std::string(__func__) + ": launch: " + cudaGetErrorString(_e)); \
  } while (0)

// ---- optional per-kernel-class profiler: CUDA events around tagged launches ----------------
// (bench.py enables it to report the dominant kernel's live duration; off by default)
struct ProfRec {
  int tag;
  cudaEvent_t a, b;
  double flops, bytes;
};
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<std::string> names;
  int tag_of(const char* name) {
    for (size_t i = 0; i < names.size(); i++)
      if (names[i] == name) return (int)i;
    names.push_back(name);
    return (int)names.size() - 1;
  }
  void begin(cudaStream_t st, const char* name, double flops, double bytes) {
    if (!on) return;
    ProfRec r;
    r.tag = tag_of(name);
    r.flops = flops;
    r.bytes = bytes;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    recs.push_back(r);
  }
  void end(cudaStream_t st) {
    if (!on || recs.empty()) return;
    cudaEventRecord(recs.back().b, st);
  }
  void clear() {
    for (ProfRec& r : recs) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    recs.clear();
  }
};
extern thread_local Profiler g_prof;

#define PROF(name, bytes, launch)             \
  do {                                        \
    ::hyp::g_prof.begin(st, name, 0.0, (double)(bytes)); \
    launch;                                   \
    ::hyp::g_prof.end(st);                           \
    HYP_LAUNCHED();                           \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_SIGMOID = 2 };

// ---- Philox4x32-10, counter-based RNG for dropout (mask is recomputed in backward) ----
__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}

// uniform in [0,1) for (seed, stream id, element index)
__host__ __device__ inline float philox_uniform(uint64_t seed, uint32_t stream_id, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream_id, 0u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; i++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace hyp
