// exchange_kernel.cu -- the cross-GPU best-pick (lib.rs:397-413 over restart ranges sharded across GPUs) as direct
// peer-to-peer stores over NVLink instead of a collective library call.
//
// Every rank owns one SYMMETRIC buffer (same layout on every GPU, mapped into every peer's address space by the host:
// torch symmetric memory / cudaIpc -- plumbing):   records[NSLOT][W][len] doubles | flags[NSLOT][W] u64
//   push    rank r stores its candidate record (OPTIK_RECORD_HEAD + n doubles, see optik_b200.h) into slot seq % NSLOT,
//           row r of EVERY peer's buffer, fences, then stores flags[slot][r] = seq there (release, system scope)
//   select  one warp waits (acquire, system scope) until flags[slot][w] >= seq for all w, then applies the reference's
//           selection rule to the W rows: converged first, lowest score, lowest restart index
// The payload is 15 doubles per rank: the cost of the exchange is launch + one NVLink round trip (~ a few us) where
// ncclAllGather costs a kernel launch plus its protocol (~40 us at 8 GPUs, profiles/r01b).  A rank can be at most
// `pipeline depth` sequence numbers ahead of a peer (its own select for seq needs every peer's push of seq, and a peer
// submits seq + depth only after its own select of seq), so NSLOT = 32 slots are never overwritten before they are read
// for depth <= 16.  The wait is bounded (2 s): a dead peer yields
// found = -1 in the output record instead of a hung GPU.
#include <cuda_runtime.h>

#include <cstdint>

#include "solver_params.h"

namespace optik {

constexpr int EX_NSLOT = OPTIK_EXCHANGE_NSLOT;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one warp per peer
__global__ void __launch_bounds__(256) exchange_push_kernel(const double* __restrict__ rec, int len, const uint64_t* __restrict__ peers,
                                                            int rank, int W, unsigned long long seq) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned slot = (unsigned)(seq % EX_NSLOT);
  for (int p = warp; p < W; p += nwarps) {
    double* base = (double*)peers[p];
    double* dst = base + ((size_t)slot * W + rank) * len;
    for (int j = lane; j < len; j += 32) dst[j] = rec[j];
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
      unsigned long long* flags = (unsigned long long*)(base + (size_t)EX_NSLOT * W * len);
      st_release_sys(flags + (size_t)slot * W + rank, seq);
    }
  }
}

__global__ void __launch_bounds__(32) exchange_select_kernel(const double* __restrict__ buf, int W, int n, unsigned long long seq,
                                                             double* __restrict__ out) {
  const int len = 8 + n;
  const unsigned lane = threadIdx.x;
  const unsigned slot = (unsigned)(seq % EX_NSLOT);
  const unsigned long long* flags = (const unsigned long long*)(buf + (size_t)EX_NSLOT * W * len) + (size_t)slot * W;
  bool ok = true;
  if ((int)lane < W) {
    const unsigned long long t0 = gtimer();
    while (ld_acquire_sys(flags + lane) < seq) {
      if (gtimer() - t0 > 2000000000ull) { ok = false; break; }
      __nanosleep(200);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  const double* rec = buf + (size_t)slot * W * len;
  double has = -1.0, score = 0.0, restart = 0.0;
  unsigned idx = 0;
  for (unsigned c = lane; c < (unsigned)W; c += 32) {
    const double h = rec[c * len + 0], s = rec[c * len + 1], r = rec[c * len + 2];
    const bool better = (h > has) || (h == has && (s < score || (s == score && r < restart)));
    if (better) { has = h; score = s; restart = r; idx = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double h2 = __shfl_xor_sync(0xffffffffu, has, o), s2 = __shfl_xor_sync(0xffffffffu, score, o),
                 r2 = __shfl_xor_sync(0xffffffffu, restart, o);
    const unsigned i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    const bool better = (h2 > has) || (h2 == has && (s2 < score || (s2 == score && (r2 < restart || (r2 == restart && i2 < idx)))));
    if (better) { has = h2; score = s2; restart = r2; idx = i2; }
  }
  for (int j = lane; j < len; j += 32) out[j] = rec[idx * len + j];
  if (!ok && lane == 0) out[0] = -1.0;  // a peer never delivered
}

}  // namespace optik

extern "C" int optik_exchange_nslot(void) { return optik::EX_NSLOT; }
extern "C" int optik_launch_exchange_push(const double* rec, int len, const uint64_t* peers_dev, int rank, int W,
                                          unsigned long long seq, void* stream) {
  const int threads = W * 32 > 256 ? 256 : W * 32;
  optik::exchange_push_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(rec, len, peers_dev, rank, W, seq);
  return (int)cudaGetLastError();
}
extern "C" int optik_launch_exchange_select(const double* buf, int W, int n, unsigned long long seq, double* out, void* stream) {
  optik::exchange_select_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(buf, W, n, seq, out);
  return (int)cudaGetLastError();
}
