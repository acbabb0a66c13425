// solve_t1_kernel.cu -- the restart solve with ONE THREAD PER RESTART SEED (throughput layout; any chain of up to 32
// joints: 128-thread blocks, 32-thread blocks for long chains so that more of their per-thread rows fit an SM).
//
// Same path, objective, stop rules, seeds and LM step as solve_kernel.cu (reference: crates/optik/src/lib.rs:297-413,
// objective.rs:40-110, kinematics.rs:123-196, math.rs:40-203), but the evaluation order is sequential per seed:
//   * one BACKWARD recursion on the inverse pose C_j = B_j^-1 (B_{j-1} = L_j B_j, B_n = tip) in the BASE frame yields
//     joint j's body-Jacobian column [t_C x (R_C a_j); R_C a_j] with one rotation and one cross product; the pose error
//     is X = (C_0 T_tgt)^-1 (objective.rs:48-49), one isometry product after the loop
//   * on accept the body columns become task columns by two explicit 3x3 matrices (Jlog6 = [[J, C J],[0, J]], 27 fma
//     per column)
//   * the current point's task columns live in a per-thread shared-memory row (128-bit accesses, odd 16-byte row stride
//     => conflict-free); the trial point's body columns in a second shared row (ROWS = 2, two blocks per SM) or in
//     thread-local memory (ROWS = 1: L1/L2-resident, written once per evaluation, read back only on accept; three
//     blocks per SM)
//   * the 6x6 Gram matrix is accumulated joint by joint with fma; the LDL^T solve is thread-private.
// Restart seeds (lib.rs:360-370) come from a table in HBM written by seed_table_kernel (one ChaCha8 block per restart,
// all lanes busy) -- a lone lane drawing a block inside this kernel stalls the 31 other lanes of its warp for ~800
// instructions; the in-kernel draw remains as the fallback for restart indices beyond the table.
// Scheduling (lib.rs:297-301, 381-387, 393-413): a flattened state machine, every lane runs ONE uniform loop body
// (evaluate -> stop tests -> step) and takes its next attempt when one ends.  The transition code is entered by the
// whole warp (converged) whenever some lane needs work, so that queue fetches are warp-aggregated:
//   sched 2  per-attempt records: job = one restart of the one target (optik_gpu_ik_attempts, BASELINE config 2)
//   sched 0  static jobs (target, chunk): chunk c runs restarts r_begin+c, +C, ... in order (Quality batches, small
//            Speed batches with the `found` early exit)
//   sched 1  dynamic Speed chains: a lane that takes a target claims its restarts one by one until one converges; once
//            every target has been taken, lanes left without work join the chains that still run in THEIR warp and claim
//            restarts of the same targets in parallel.  A claimed index always runs unless a LOWER index has already
//            converged (found[t], lib.rs:308, 382-384), so the per-target answer is the lowest-index converged restart
//            (lib.rs:409-412 with one thread) regardless of timing; no host round trip, no tail of unlucky targets.  A
//            warp leaves when every target has been taken and none of its lanes runs a chain.
#include <cuda_runtime.h>

#include "ik_math.cuh"
#include "ldl6.cuh"
#include "solver_params.h"

namespace optik {

#ifndef OPTIK_T1_THREADS
#define OPTIK_T1_THREADS 128
#endif
constexpr int T1_THREADS_STD = OPTIK_T1_THREADS;  // threads per block
// long chains (one-row layout, n > 12): 32-thread blocks.  A 128-thread block of the 20-joint snake needs 166 KB, so one
// block = 4 warps per SM; five 32-thread blocks (44.8 KB each) fit.
constexpr int T1_THREADS_LONG = 32;
constexpr int T1_LONG_N = 12;
constexpr unsigned FULLMASK = 0xffffffffu;
constexpr unsigned DYN_NONE = 0xffffffffu;

// hot joint loops: fully unrolled when the joint count is a template constant (NS), rolled otherwise
#define T1_HOT_UNROLL _Pragma("unroll (NS ? NS : 1)")

DEV int t1_row_units(int n) { return (3 * n) | 1; }  // 16-byte units per column row, forced odd

// ChaCha8 block `counter` of stream `stream` -> 8 u64 (joint j draws u64 number j: block j / 8, word pair j % 8).  Rolled and not inlined: it is the
// cold path of the solve kernel (restart indices beyond the seed table) and the whole of seed_table_kernel.
__device__ __noinline__ void chacha8_block(const uint32_t* key, uint64_t counter, uint64_t stream, uint64_t* out8) {
  uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                    key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                    (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
    OPTIK_QR(x[0], x[4], x[8], x[12]) OPTIK_QR(x[1], x[5], x[9], x[13])
    OPTIK_QR(x[2], x[6], x[10], x[14]) OPTIK_QR(x[3], x[7], x[11], x[15])
    OPTIK_QR(x[0], x[5], x[10], x[15]) OPTIK_QR(x[1], x[6], x[11], x[12])
    OPTIK_QR(x[2], x[7], x[8], x[13]) OPTIK_QR(x[3], x[4], x[9], x[14])
  }
#pragma unroll
  for (int k = 0; k < 8; k++)
    out8[k] = (uint64_t)(x[2 * k] + s[2 * k]) | ((uint64_t)(x[2 * k + 1] + s[2 * k + 1]) << 32);
}

// Restart seeds for indices [r_begin, r_begin + count): seed[i][j] = clamp(uniform(ChaCha8(42).stream(r).u64[j]))
// (lib.rs:86-91, 360-370).  Index 0 is the caller's seed and is left untouched.  `chain` is the device blob.
__global__ void __launch_bounds__(128) seed_table_kernel(const double* __restrict__ chain, int n, const uint32_t* __restrict__ key_g,
                                                         unsigned long long r_begin, unsigned long long count,
                                                         double* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const unsigned long long r = r_begin + i;
  if (r == 0) return;
  uint32_t key[8];
#pragma unroll
  for (int k = 0; k < 8; k++) key[k] = key_g[k];
  for (int j0 = 0; j0 < n; j0 += 8) {
    uint64_t u[8];
    chacha8_block(key, (uint64_t)(j0 >> 3), r, u);
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (j0 + j < n) {
        const double* jc = chain + OPTIK_CHAIN_STRIDE * (j0 + j);
        out[i * n + j0 + j] = fmin(fmax(uniform_f64(u[j], jc[14], jc[15]), jc[12]), jc[13]);
      }
  }
}

// Known-answer hook: the raw 16-word ChaCha8 block (counter 0) of `stream` under `key`, as the kernels compute it
__global__ void chacha8_kat_kernel(const uint32_t* __restrict__ key_g, unsigned long long stream, uint32_t* __restrict__ out16) {
  uint32_t key[8];
  for (int k = 0; k < 8; k++) key[k] = key_g[k];
  uint64_t u[8];
  chacha8_block(key, 0, stream, u);
  for (int k = 0; k < 8; k++) { out16[2 * k] = (uint32_t)u[k]; out16[2 * k + 1] = (uint32_t)(u[k] >> 32); }
}

// ---- per-target word of the dynamic scheduler: high half = lowest converged relative restart index so far (DYN_NONE:
// none), low half = 1 while a writer holds the target's record (anything else: free).  Initialised to all ones.
DEV unsigned dyn_found_of(const SolveParams& P, unsigned long long t) { return ((volatile unsigned*)(P.dyn_word + t))[1]; }

// ---- warp-level job pool.  `mask` = lanes that want a job (warp-uniform).  The pool [pool_next, pool_end) is a range
// of the global queue claimed `chunk` jobs at a time by ONE atomic, so that most job starts cost no memory round trip;
// on a refill `nb` is the new range's first job (the caller prefetches its inputs) else ~0.
DEV unsigned long long warp_take(unsigned long long* queue, int lane, unsigned mask, unsigned chunk,
                                 unsigned long long& pool_next, unsigned long long& pool_end, unsigned long long& nb) {
  const unsigned k = __popc(mask), rank = __popc(mask & ((1u << lane) - 1u));
  const unsigned long long avail = pool_end - pool_next;
  nb = ~0ull;
  if (avail >= k) {
    const unsigned long long job = pool_next + rank;
    pool_next += k;
    return job;
  }
  const unsigned take = chunk > k ? chunk : k;
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(queue, (unsigned long long)take);
  base = __shfl_sync(FULLMASK, base, leader);
  const unsigned long long job = rank < avail ? pool_next + rank : base + (rank - avail);
  pool_next = base + (k - avail);
  pool_end = base + take;
  nb = base;
  return job;
}
// L1 prefetch of the byte range [p, p + bytes) by the whole warp (at most 64 lines)
DEV void warp_prefetch(const void* p, unsigned long long bytes, int lane) {
  const char* b = (const char*)((unsigned long long)p & ~127ull);
  const char* e = (const char*)p + bytes;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const char* a = b + 128ull * (unsigned)(lane + 32 * i);
    if (a < e) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
  }
}

template <bool GENERAL, int ROWS, int NS, int TB>
__global__ void __launch_bounds__(TB, (ROWS == 1 ? 3 : 2) * (128 / TB)) solve_t1_kernel(const __grid_constant__ SolveParams P) {
  constexpr int T1_THREADS = TB;
  extern __shared__ __align__(128) double smem[];
  const int n = NS ? NS : P.n;  // NS: the joint count as a compile-time constant (joint loops fully unrolled), 0 = any
  // layout: chain blob | tip^-1 pose8 | mbarrier (16 B) | rows[ROWS][T1_THREADS][units] (16 B units) | qc[n][T] | qt[n][T] |
  //         per-joint constants origin_q (x) (axis, 0) [n][4]
  double* s_chain = smem;
  double* s_tip = smem + OPTIK_CHAIN_STRIDE * n + 8;
  uint64_t* s_bar = (uint64_t*)(s_tip + 8);
  const int units = t1_row_units(n);
  double2* s_rows = (double2*)(s_tip + 10);
  double* s_qc = (double*)(s_rows + ROWS * T1_THREADS * units);
  double* s_qt = s_qc + n * T1_THREADS;
  double* s_oa = s_qt + n * T1_THREADS;
  stage_chain_tma(s_chain, s_bar, P.chain, P.chain_bytes);
  if (threadIdx.x == 0) {  // (fixed tip joint * ee_offset)^-1, once per block
    const se3 tip = se3mul(load_pose8(s_chain + OPTIK_CHAIN_STRIDE * n), load_pose8(P.ee_offset));
    const qt iq = qconj(tip.q);
    const v3 it = neg3(qrot(iq, tip.t));
    s_tip[0] = iq.x; s_tip[1] = iq.y; s_tip[2] = iq.z; s_tip[3] = iq.w;
    s_tip[4] = it.x; s_tip[5] = it.y; s_tip[6] = it.z; s_tip[7] = 0.0;
  }
  if (threadIdx.x < n) {  // origin_q (x) (axis, 0): L_j.q = cos * origin_q + sin * this
    const double* jc = s_chain + OPTIK_CHAIN_STRIDE * threadIdx.x;
    qt oq, qa;
    oq.x = jc[4]; oq.y = jc[5]; oq.z = jc[6]; oq.w = jc[7];
    qa.x = jc[8]; qa.y = jc[9]; qa.z = jc[10]; qa.w = 0.0;
    const qt oa = qmul(oq, qa);
    double* d = s_oa + 4 * threadIdx.x;
    d[0] = oa.x; d[1] = oa.y; d[2] = oa.z; d[3] = oa.w;
  }
  __syncthreads();

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  double* qc = s_qc + tid;  // qc[j * T1_THREADS]
  double* qt_ = s_qt + tid;
  double2* const row0 = s_rows + (size_t)tid * units;  // rows[k] = row0 + k * row_stride
  const size_t row_stride = (size_t)T1_THREADS * units;
  double2 ltrial[ROWS == 1 ? 3 * OPTIK_MAX_DOF : 1];  // ROWS == 1: the trial point's body columns (thread-local memory)
  const int sched = P.sched;
  const unsigned long long njobs = sched == 1 ? P.T : P.T * (unsigned long long)P.C;
  const bool speed = (P.mode == 2);
  const unsigned long long t_start = P.max_ns ? globaltimer_ns() : 0ull;
  const unsigned long long nrest = P.r_end - P.r_begin;
  // dynamic chains: jobs of the queue and the offset of the shared restart counters (see the transition code)
  const unsigned long long dyn_jobs = P.T * (unsigned long long)(P.dyn_k0 + 1u), dyn_base = P.dyn_k0 ? P.dyn_k0 + 1u : 0u;

  // ---- per-thread state.  sched 0: a "job" is one (target, chunk); sched 1: a "chain" on one target; sched 2: one attempt
  unsigned long long job = 0, tgt_id = 0, r_idx = 0, r_next = 0;
  bool running = false, job_open = false, best_has = false, rec_any = false, done = false;
  bool seed_clamped = false;  // this attempt started from a caller seed that had to be clamped into the limits
  bool pre_ok = false;  // dynamic, shared target: pre_rel is a claimed restart index that may run
  unsigned long long pre_rel = 0;
  bool no_help = false; // dynamic: my chain cannot use helpers any more (no restart left / a restart of the target converged)
  bool excl = false;    // dynamic: no other lane works on my target (nobody joined it yet) -> no atomics, no record word
  unsigned my_next = 0; // dynamic, exclusive chain: the next relative restart index of my target
  bool wfd = false;     // warp-uniform: every target of a dynamic launch has been taken
  unsigned long long pool_next = 0, pool_end = 0;  // warp-uniform: this warp's claimed range of the job queue
  double fc = 0.0, rc[6] = {0, 0, 0, 0, 0, 0}, lambda = P.lambda0, best_score = 0.0;
  int have_cur = 0, slow = 0, evals = 0, job_evals = 0, cur = 0;
  unsigned n_attempts = 0, n_evals = 0, n_conv = 0;

  int sel_has = -1;  // sched 2 with a fused selection: this lane's best attempt (converged?, score, restart index)
  double sel_score = 0.0;
  unsigned long long sel_restart = ~0ull;
  unsigned pass = 0;  // warp-uniform loop counter
  for (;;) {
    pass++;
    // ---------------- transitions: lanes without an attempt take the next one.  Entered by the WHOLE warp (converged)
    // whenever some lane needs work, so that queue fetches are one atomic per warp.
    const unsigned need = __ballot_sync(FULLMASK, !running && !done);
    if (need) {
      const bool idle = !running && !done;
      // lib.rs:260-264; warp-uniform (the timer is read per lane): the branches below hold warp-level primitives
      const bool late = P.max_ns && __any_sync(FULLMASK, (globaltimer_ns() - t_start) > P.max_ns);
      bool got = false;
      if (sched == 2) {
        unsigned long long nb;
        const unsigned long long mine = warp_take(P.queue, lane, need, P.pool_chunk, pool_next, pool_end, nb);
        if (nb < njobs && P.seed_tab && P.r_begin + nb >= P.seed_begin)  // the new range's seeds -> L1
          warp_prefetch(P.seed_tab + (P.r_begin + nb - P.seed_begin) * n, (pool_end - nb) * n * sizeof(double), lane);
        if (idle) {
          job = mine;
          if (job >= njobs) done = true;
          else {
            r_idx = P.r_begin + job;
            if (!late) got = true;
            else {  // past the deadline: the attempt is recorded as not run
#pragma unroll 1
              for (int j = 0; j < n; j++) P.cand_q[job * n + j] = P.x0[j];
              P.cand_f[job] = 0.0; P.cand_score[job] = 0.0; P.cand_restart[job] = r_idx;
              P.cand_status[job] = OPTIK_ST_SKIPPED; P.cand_evals[job] = 0;
            }
          }
        }
      } else if (sched == 0) {
        if (idle) {
          for (;;) {
            if (!job_open) {
              job = atomicAdd(P.queue, 1ull);  // dynamic job queue
              if (job >= njobs) { done = true; break; }
              const unsigned long long c = (P.C == 1) ? 0ull : (P.T == 1 ? job : job % P.C);
              tgt_id = (P.C == 1) ? job : (P.T == 1 ? 0ull : job / P.C);
              r_next = P.r_begin + c;
              best_has = false; rec_any = false; best_score = 0.0; job_evals = 0;
              job_open = true;
            }
            bool go = r_next < P.r_end && !late;
            if (go && speed && P.found) go = !(*((volatile unsigned long long*)(P.found + tgt_id)) < r_next);
            if (go) { r_idx = r_next; r_next += P.C; got = true; break; }
            if (!rec_any) {  // the chunk ran no attempt: its record says so
#pragma unroll 1
              for (int j = 0; j < n; j++) P.cand_q[job * n + j] = P.x0[tgt_id * n + j];
              P.cand_f[job] = 0.0; P.cand_score[job] = 0.0; P.cand_restart[job] = r_next;
              P.cand_status[job] = OPTIK_ST_SKIPPED;
            }
            P.cand_evals[job] = job_evals;  // the chunk is finished
            job_open = false;
          }
        }
      } else {
        // (A) continue my chain: claim the next restart unless a restart of the target already converged
        if (idle && job_open) {
          if (excl) {  // nobody else knows this target: its restart counter lives in a register
            if (!late && my_next < nrest) { r_idx = P.r_begin + my_next; my_next++; got = true; }
          } else if (!late && pre_ok) {  // shared target: the index was claimed when the last attempt failed (one round trip)
            r_idx = P.r_begin + pre_rel; got = true;
          }
          pre_ok = false;
          if (!got) job_open = false;  // chain over
        }
        // (B) fresh jobs, one queue fetch per warp.  Job f = first restart of target f; with fewer targets than
        // resident lanes (dyn_k0 > 0) the queue holds (dyn_k0 + 1) jobs per target -- job f = restart f / T of target
        // f % T -- so that every lane starts a restart at once (the counter next[t] then counts from dyn_k0 + 1)
        unsigned want = __ballot_sync(FULLMASK, idle && !got);
        if (want && !wfd) {
          unsigned long long nb;
          const unsigned long long f = warp_take(P.queue, lane, want, P.pool_chunk, pool_next, pool_end, nb);
          if (P.dyn_k0 == 0 && nb < P.T) {  // the new range's seeds and targets -> L1
            warp_prefetch(P.x0 + nb * n, (pool_end - nb) * n * sizeof(double), lane);
            warp_prefetch(P.targets + nb * 8, (pool_end - nb) * 64, lane);
          }
          if (idle && !got && f < dyn_jobs) {
            const unsigned long long t = P.dyn_k0 ? f % P.T : f, rel0 = P.dyn_k0 ? f / P.T : 0ull;
            if (late) {  // past the deadline: the target's record says that nothing ran
              if (rel0 == 0) {
                for (int j = 0; j < n; j++) P.cand_q[t * n + j] = P.x0[t * n + j];
                P.cand_f[t] = 0.0; P.cand_status[t] = OPTIK_ST_SKIPPED;
                if (P.cand_restart) P.cand_restart[t] = P.r_begin;
              }
            } else if (rel0 < nrest) {
              tgt_id = t; r_idx = P.r_begin + rel0; got = true; job_open = true; no_help = false;
              excl = P.dyn_k0 == 0;
              my_next = 1;
            }
          }
          if (pool_next >= dyn_jobs) wfd = true;
          want = __ballot_sync(FULLMASK, idle && !got);
        }
        // (B2) speculation inside the warp: once every target has been taken, lanes left without work join the chains
        // that still run in THIS warp and claim the next restarts of their targets in parallel (idle lane number i
        // helps chain number i mod #chains).  Nothing waits for a failure: the geometric tail of unlucky targets
        // (one attempt after the other) becomes a few rounds of parallel restarts on lanes that had nothing to do.
        // A helper's restart index is above its owner's, so it is dropped at the next poll if the owner converges.
        // Looking for help work costs global round trips that the warp's RUNNING lanes wait for.  A chain whose target
        // has no restart left or has a converged restart cannot use helpers any more: it is marked (no_help) and left
        // alone, so every failed claim below happens once.
        if (want && wfd && !late) {
          const unsigned chains = __ballot_sync(FULLMASK, running && job_open && !no_help);
          if (chains) {
            const unsigned nch = __popc(chains), k = __popc(want);
            const unsigned below = (1u << lane) - 1u;
            const bool owner = (chains >> lane) & 1u;
            if (owner && __popc(chains & below) < k && excl) {  // my target gets a helper: it is shared from here on
              atomicExch(P.dyn_next + tgt_id, my_next);
              __threadfence();
              excl = false;
            }
            __syncwarp();
            const bool helper = idle && !got;
            const int src = helper ? (int)__fns(chains, 0, (__popc(want & below) % nch) + 1) : lane;
            const unsigned long long ht = __shfl_sync(FULLMASK, tgt_id, src);
            bool refused = false;
            if (helper) {
              refused = true;
              if (dyn_found_of(P, ht) == DYN_NONE) {
                const unsigned long long rel = dyn_base + atomicAdd(P.dyn_next + ht, 1u);
                if (rel < nrest) { tgt_id = ht; r_idx = P.r_begin + rel; got = true; job_open = true; excl = false; no_help = false; refused = false; }
              }
            }
            if ((__reduce_or_sync(FULLMASK, refused ? (1u << src) : 0u) >> lane) & 1u) no_help = true;
            want = __ballot_sync(FULLMASK, idle && !got);
          }
        }
        // leave when nothing is left to do here: every target has been taken and no lane of this warp runs a chain.
        // (Help across warps -- tickets that failing chains pushed and other warps popped -- was measured to add nothing
        // once the speculation above existed, and was removed: the warps that could have helped are gone by then.)
        if (wfd && !__any_sync(FULLMASK, running || got)) { if (idle) done = true; }
      }
      if (got) {  // restart 0 = the caller's seed, i >= 1 = ChaCha8 stream i (lib.rs:360-370)
        if (r_idx == 0) {  // (loads issued eight at a time: one memory latency per group, not per joint)
          for (int j0 = 0; j0 < n; j0 += 8) {
            double sv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) if (j0 + j < n) sv[j] = P.x0[tgt_id * n + j0 + j];
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (j0 + j < n) {
                const double* jc = s_chain + OPTIK_CHAIN_STRIDE * (j0 + j);
                const double cl = fmin(fmax(sv[j], jc[12]), jc[13]);
                seed_clamped |= !(cl == sv[j]);  // the reference panics here (lib.rs:251-254); see OPTIK_STATUS_FLAG_SEED_CLAMPED
                qt_[(j0 + j) * T1_THREADS] = cl;
              }
          }
        } else if (r_idx - P.seed_begin < P.seed_count) {
          const double* sd = P.seed_tab + (r_idx - P.seed_begin) * n;
          for (int j0 = 0; j0 < n; j0 += 8) {
            double sv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) if (j0 + j < n) sv[j] = __ldg(sd + j0 + j);
#pragma unroll
            for (int j = 0; j < 8; j++) if (j0 + j < n) qt_[(j0 + j) * T1_THREADS] = sv[j];
          }
        } else {
          for (int j0 = 0; j0 < n; j0 += 8) {
            uint64_t u[8];
            chacha8_block(P.key, (uint64_t)(j0 >> 3), r_idx, u);
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (j0 + j < n) {
                const double* jc = s_chain + OPTIK_CHAIN_STRIDE * (j0 + j);
                qt_[(j0 + j) * T1_THREADS] = fmin(fmax(uniform_f64(u[j], jc[14], jc[15]), jc[12]), jc[13]);
              }
          }
        }
        if (r_idx != 0) seed_clamped = false;
        have_cur = 0; slow = 0; evals = 0; lambda = P.lambda0;
        running = true;
      }
      if (__all_sync(FULLMASK, done)) break;
    }
    if (!running) continue;  // meets the rest of the warp again at the ballot

    // ---------------- evaluate the trial point: backward recursion on the inverse pose, base frame
    double2* trow = (ROWS == 1) ? ltrial : row0 + (size_t)(cur ^ 1) * row_stride;  // trial body columns
    se3 Ci = load_pose8(s_tip);  // C_n = tip^-1
    const se3 tgt = load_pose8(P.targets + 8 * tgt_id);  // issued here: the recursion hides its latency
    // unrolled instances: sin/cos of joint j - 1 is computed during joint j's quaternion chain (independent work under
    // a dependent chain; measured slower in the rolled loop)
    double s_nx = 0.0, c_nx = 0.0;
    if (NS) dsincos(0.5 * qt_[(n - 1) * T1_THREADS], s_nx, c_nx);
T1_HOT_UNROLL
    for (int j = n - 1; j >= 0; j--) {
      const double2* jc = (const double2*)(s_chain + OPTIK_CHAIN_STRIDE * j);
      const double2 c0 = jc[0], c1 = jc[1], c2 = jc[2], c3 = jc[3], c4 = jc[4], c5 = jc[5];
      const double2 a0 = ((const double2*)s_oa)[2 * j], a1 = ((const double2*)s_oa)[2 * j + 1];
      const bool pris = GENERAL && ((int)c1.y != 0);
      const v3 ax = mk3(c4.x, c4.y, c5.x);
      v3 lin, ang = qrot(Ci.q, ax);
      if (!pris) lin = cross3(Ci.t, ang);
      else { lin = ang; ang = mk3(0, 0, 0); }
      trow[3 * j + 0] = make_double2(lin.x, lin.y);
      trow[3 * j + 1] = make_double2(lin.z, ang.x);
      trow[3 * j + 2] = make_double2(ang.y, ang.z);
      const double qj = qt_[j * T1_THREADS];
      qt lq;  // conj(L_j.q)
      v3 lt = mk3(c0.x, c0.y, c1.x);
      if (!pris) {
        double s, c;
        if (NS) {
          s = s_nx; c = c_nx;
          if (j > 0) dsincos(0.5 * qt_[(j - 1) * T1_THREADS], s_nx, c_nx);
        } else dsincos(0.5 * qj, s, c);
        lq.x = -fma(c, c2.x, s * a0.x); lq.y = -fma(c, c2.y, s * a0.y); lq.z = -fma(c, c3.x, s * a1.x);
        lq.w = fma(c, c3.y, s * a1.y);
      } else {
        qt oq;
        oq.x = c2.x; oq.y = c2.y; oq.z = c3.x; oq.w = c3.y;
        lq = qconj(oq);
        lt = add3(lt, qrot(oq, scale3(ax, qj)));
      }
      Ci.q = qmul(Ci.q, lq);
      Ci.t = sub3(Ci.t, qrot(Ci.q, lt));
    }
    se3 B;  // X = (C_0 T_tgt)^-1
    {
      const qt xiq = qmul(Ci.q, tgt.q);
      const v3 xit = add3(Ci.t, qrot(Ci.q, tgt.t));
      B.q = qconj(xiq);
      B.t = neg3(qrot(B.q, xit));
    }
    ErrCoef ec;
    v3 elin;
    error_terms(B.q, B.t, ec, elin);
    v3 rl = elin, ra = ec.w;
    const bool weighted = GENERAL && P.weighted;
    if (weighted) { rl = weight3(tgt.q, P.wl, elin); ra = weight3(tgt.q, P.wa, ec.w); }
    double rt[6] = {rl.x, rl.y, rl.z, ra.x, ra.y, ra.z};
    const double ft = dot6(rt, rt);

    // ---------------- bookkeeping (mirrors NLopt's stop tests as the reference configures them, lib.rs:345-347)
    int status = OPTIK_ST_NONE;
    bool accept = false;
    evals++;
    if (ft != ft) status = OPTIK_ST_NAN;
    else if (ft < P.tol_f) status = OPTIK_ST_STOPVAL;
    else if (!have_cur) accept = true;
    else if (ft < fc) {
      accept = true;
      const double df = fc - ft;
      if (df < P.tol_df_eff) status = OPTIK_ST_FTOL;
      else if (P.tol_dx > 0.0) {
        double dx = 0.0;
#pragma unroll 1
        for (int j = 0; j < n; j++) dx = fmax(dx, fabs(qt_[j * T1_THREADS] - qc[j * T1_THREADS]));
        if (dx < P.tol_dx) status = OPTIK_ST_XTOL;
      }
      slow = (df < P.stall_rel * fc) ? slow + 1 : 0;
      if (status == OPTIK_ST_NONE && slow >= P.stall_count) status = OPTIK_ST_STUCK;
      lambda = fmax(lambda * P.lambda_dec, P.lambda_min);
    } else {
      lambda = lambda * P.lambda_inc;
      if (lambda > P.lambda_max) status = OPTIK_ST_STUCK;
    }
    if (status == OPTIK_ST_NONE && evals >= P.max_evals) status = OPTIK_ST_ITERCAP;
    if (status == OPTIK_ST_NONE && P.max_ns && (globaltimer_ns() - t_start) > P.max_ns) status = OPTIK_ST_SKIPPED;
    // should_exit (lib.rs:308): a LOWER restart index of this target converged.  Polled every 4th pass of the WARP (a
    // warp-uniform condition: a per-lane one splits the warp in two for the rest of the loop body).
    if (speed && (pass & 3u) == 0u) {
      if (sched == 1) {
        if (status == OPTIK_ST_NONE && !excl &&
            dyn_found_of(P, tgt_id) < (unsigned)(r_idx - P.r_begin)) status = OPTIK_ST_SKIPPED;
      } else if (P.found) {
        if (status == OPTIK_ST_NONE && *((volatile unsigned long long*)(P.found + tgt_id)) < r_idx) status = OPTIK_ST_SKIPPED;
      }
    }

    if (status != OPTIK_ST_NONE) {  // attempt over
      const bool success = (P.tol_f >= 0.0 && status == OPTIK_ST_STOPVAL) ||
                           (P.tol_df_user >= 0.0 && status == OPTIK_ST_FTOL) ||
                           (P.tol_dx >= 0.0 && status == OPTIK_ST_XTOL);  // lib.rs:376-379
      n_attempts++; n_evals += evals;
      if (success) n_conv++;
      if (sched != 1) {
        double score = 0.0;
        if (success && !speed) {  // Quality score ||q - x0||^2 (lib.rs:402-407)
#pragma unroll 1
          for (int j = 0; j < n; j++) {
            const double d = qt_[j * T1_THREADS] - P.x0[tgt_id * n + j];
            score = fma(d, d, score);
          }
        }
        job_evals += evals;
        const bool record = sched == 2 || (success ? (!best_has || score < best_score) : !rec_any);  // failures: keep the first
        if (record) {
          rec_any = true;
#pragma unroll 1
          for (int j = 0; j < n; j++) P.cand_q[job * n + j] = qt_[j * T1_THREADS];
          P.cand_f[job] = ft; P.cand_score[job] = score; P.cand_restart[job] = r_idx;
          // per-attempt records: the caller's seed (restart 0) carries the clamped-seed flag itself (no extra launch)
          P.cand_status[job] = (sched == 2 && seed_clamped) ? (status | 0x100) : status;
          if (sched == 2) P.cand_evals[job] = evals;
        }
        if (sched == 2 && P.fused_record) {  // this lane's best attempt so far, for the in-kernel selection pass
          const int has = success ? 1 : 0;
          const double sc = speed ? (double)r_idx : score;  // Speed: lowest converged index (lib.rs:409-412)
          if (has > sel_has || (has == sel_has && (sc < sel_score || (sc == sel_score && r_idx < sel_restart)))) {
            sel_has = has; sel_score = sc; sel_restart = r_idx;
          }
        }
        if (success) {
          if (record) { best_has = true; best_score = score; }
          if (speed && sched == 0) {  // first success ends the chunk (lib.rs:381-387, 411)
            if (P.found) atomicMin(P.found + tgt_id, r_idx);
            r_next = P.r_end;
          }
        }
      } else {
        // dynamic chains: the target's record is the lowest-index converged attempt, or the first attempt's failure
        const unsigned rel = (unsigned)(r_idx - P.r_begin);
        if (P.cand_evals) atomicAdd(P.cand_evals + tgt_id, evals);  // (a reduction: no result, no wait)
        if (excl) {  // the only chain on this target: plain stores
          if (success || rel == 0u) {
#pragma unroll 1
            for (int j = 0; j < n; j++) P.cand_q[tgt_id * n + j] = qt_[j * T1_THREADS];
            P.cand_f[tgt_id] = ft; P.cand_status[tgt_id] = status;
            if (P.cand_restart) P.cand_restart[tgt_id] = r_idx;
          }
          if (success) job_open = false;  // first success ends the chain (lib.rs:381-387, 411)
          // (a failing chain stays exclusive: its counter is published when an idle lane of its warp joins it)
        } else {
          // shared target.  A failed attempt claims the chain's next restart and reads found[t] right here -- two
          // independent requests in flight together, consumed by the next transition pass.  A record
          // is written under the target's word: one CAS takes the record and publishes the index (only a new lowest
          // index gets it; a higher one sees that and leaves), the row is written, fenced and the word released.
          if (!success && status != OPTIK_ST_SKIPPED) {
            const unsigned fnow = dyn_found_of(P, tgt_id);
            pre_rel = dyn_base + atomicAdd(P.dyn_next + tgt_id, 1u);
            pre_ok = fnow == DYN_NONE && pre_rel < nrest;
          }
          if (success || (rel == 0u && status != OPTIK_ST_SKIPPED)) {
            const unsigned pub = success ? rel : DYN_NONE;  // a failure record leaves found[t] empty
            unsigned long long w = ~0ull;                   // optimistic: nothing recorded, nobody writing
            for (;;) {
              const unsigned f = (unsigned)(w >> 32);
              if (success ? f <= rel : f != DYN_NONE) break;            // a lower index (any index) is already recorded
              if ((unsigned)w == 1u) {                                   // a writer holds the record
                if (!success) break;                                     // ... and it can only be a converged attempt
                w = *((volatile unsigned long long*)(P.dyn_word + tgt_id));
                continue;
              }
              const unsigned long long old = atomicCAS(P.dyn_word + tgt_id, w, ((unsigned long long)pub << 32) | 1ull);
              if (old == w) {
#pragma unroll 1
                for (int j = 0; j < n; j++) P.cand_q[tgt_id * n + j] = qt_[j * T1_THREADS];
                P.cand_f[tgt_id] = ft; P.cand_status[tgt_id] = status;
                if (P.cand_restart) P.cand_restart[tgt_id] = r_idx;
                // release store: the row is visible before the word says "free" (a release fence, not the sequentially
                // consistent one of __threadfence())
                asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(P.dyn_word + tgt_id), "l"((unsigned long long)pub << 32) : "memory");
                break;
              }
              w = old;
            }
          }
        }
      }
      running = false;
      continue;
    }

    if (accept) {  // current point <- trial point; body columns -> task columns
T1_HOT_UNROLL
      for (int j = 0; j < n; j++) qc[j * T1_THREADS] = qt_[j * T1_THREADS];
      fc = ft; have_cur = 1;
#pragma unroll
      for (int i = 0; i < 6; i++) rc[i] = rt[i];
      double Jm[9], CJ[9];
      task_mats(ec, Jm, CJ);
      double2* wrow = (ROWS == 1) ? row0 : trow;  // ROWS == 2: converted in place, then the rows swap
      double2 n0 = trow[0], n1 = trow[1], n2 = trow[2];  // software-pipelined: column j + 1 is in flight during column j
T1_HOT_UNROLL
      for (int j = 0; j < n; j++) {
        const double2 a0 = n0, a1 = n1, a2 = n2;
        if (j + 1 < n) { n0 = trow[3 * j + 3]; n1 = trow[3 * j + 4]; n2 = trow[3 * j + 5]; }
        v3 top, bot;
        task_col_m(Jm, CJ, mk3(a0.x, a0.y, a1.x), mk3(a1.y, a2.x, a2.y), top, bot);
        if (weighted) { top = weight3(tgt.q, P.wl, top); bot = weight3(tgt.q, P.wa, bot); }
        wrow[3 * j + 0] = make_double2(top.x, top.y);
        wrow[3 * j + 1] = make_double2(top.z, bot.x);
        wrow[3 * j + 2] = make_double2(bot.y, bot.z);
      }
      if (ROWS == 2) cur ^= 1;
    }

    // ---------------- LM step from the current point: y = (Jm Jm^T + lambda I)^-1 r ; dq = -Jm^T y ; project on bounds
    const double2* crow = (ROWS == 1) ? row0 : row0 + (size_t)cur * row_stride;
    double Ap[21];
#pragma unroll
    for (int e = 0; e < 21; e++) Ap[e] = 0.0;
    unsigned free_mask = 0;
T1_HOT_UNROLL
    for (int j = 0; j < n; j++) {
      const double2 a0 = crow[3 * j + 0], a1 = crow[3 * j + 1], a2 = crow[3 * j + 2];
      const double c[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
      const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
      const double g = dot6(rc, c);
      const double qcj = qc[j * T1_THREADS];
      const bool pinned = (qcj <= jc[12] && g > 0.0) || (qcj >= jc[13] && g < 0.0);
      const double m = pinned ? 0.0 : 1.0;
      free_mask |= pinned ? 0u : (1u << j);
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const double jm = m * c[a];
#pragma unroll
        for (int b = 0; b <= a; b++) Ap[a * (a + 1) / 2 + b] = fma(jm, c[b], Ap[a * (a + 1) / 2 + b]);
      }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) Ap[a * (a + 1) / 2 + a] = Ap[a * (a + 1) / 2 + a] + lambda;
    double y[6];
    ldl6_solve(Ap, rc, y);
T1_HOT_UNROLL
    for (int j = 0; j < n; j++) {
      const double2 a0 = crow[3 * j + 0], a1 = crow[3 * j + 1], a2 = crow[3 * j + 2];
      const double m = ((free_mask >> j) & 1u) ? 1.0 : 0.0;
      const double c[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
      const double* jc = s_chain + OPTIK_CHAIN_STRIDE * j;
      // pinned joints stay put (m = 0); the projection on [lb, ub] as two compares (fmin / fmax carry NaN-quieting
      // code this loop does not need: a NaN step stays NaN and ends the attempt as OPTIK_ST_NAN at the next evaluation)
      double x = qc[j * T1_THREADS] - m * dot6(c, y);
      x = x < jc[12] ? jc[12] : x;
      x = x > jc[13] ? jc[13] : x;
      qt_[j * T1_THREADS] = x;
    }
  }

  // ---------------- fused selection (lib.rs:397-413) for per-attempt launches: lane -> warp -> block -> the last block
  // to finish reduces the per-block winners and writes the ONE candidate record (and, across GPUs, stores it straight
  // into every peer's exchange buffer over NVLink: csrc/exchange_kernel.cu's push without its launch)
  if (sched == 2 && P.fused_record) {
    __shared__ int s_has[T1_THREADS / 32 + 1];
    __shared__ double s_score[T1_THREADS / 32 + 1];
    __shared__ unsigned long long s_restart[T1_THREADS / 32 + 1];
    __shared__ bool s_last;
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int h2 = __shfl_xor_sync(FULLMASK, sel_has, o);
      const double c2 = __shfl_xor_sync(FULLMASK, sel_score, o);
      const unsigned long long r2 = __shfl_xor_sync(FULLMASK, sel_restart, o);
      if (h2 > sel_has || (h2 == sel_has && (c2 < sel_score || (c2 == sel_score && r2 < sel_restart)))) {
        sel_has = h2; sel_score = c2; sel_restart = r2;
      }
    }
    if (lane == 0) { s_has[tid >> 5] = sel_has; s_score[tid >> 5] = sel_score; s_restart[tid >> 5] = sel_restart; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < T1_THREADS / 32; w++)
        if (s_has[w] > sel_has || (s_has[w] == sel_has && (s_score[w] < sel_score || (s_score[w] == sel_score && s_restart[w] < sel_restart)))) {
          sel_has = s_has[w]; sel_score = s_score[w]; sel_restart = s_restart[w];
        }
      P.fused_part_has[blockIdx.x] = sel_has; P.fused_part_score[blockIdx.x] = sel_score; P.fused_part_restart[blockIdx.x] = sel_restart;
      __threadfence();  // this block's records and its partial are visible before it counts as finished
      s_last = atomicAdd(P.fused_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      int h = -1;
      double sc = 0.0;
      unsigned long long rs = ~0ull;
      for (unsigned b = tid; b < gridDim.x; b += T1_THREADS) {
        const int h2 = ((volatile int*)P.fused_part_has)[b];
        const double c2 = ((volatile double*)P.fused_part_score)[b];
        const unsigned long long r2 = ((volatile unsigned long long*)P.fused_part_restart)[b];
        if (h2 > h || (h2 == h && (c2 < sc || (c2 == sc && r2 < rs)))) { h = h2; sc = c2; rs = r2; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int h2 = __shfl_xor_sync(FULLMASK, h, o);
        const double c2 = __shfl_xor_sync(FULLMASK, sc, o);
        const unsigned long long r2 = __shfl_xor_sync(FULLMASK, rs, o);
        if (h2 > h || (h2 == h && (c2 < sc || (c2 == sc && r2 < rs)))) { h = h2; sc = c2; rs = r2; }
      }
      __syncthreads();
      if (lane == 0) { s_has[tid >> 5] = h; s_score[tid >> 5] = sc; s_restart[tid >> 5] = rs; }
      __syncthreads();
      if (tid < 32) {  // warp 0 finishes and writes the record [found, score, restart, cost, status, 0,0,0, q...]
        h = s_has[0]; sc = s_score[0]; rs = s_restart[0];
        for (int w = 1; w < T1_THREADS / 32; w++)
          if (s_has[w] > h || (s_has[w] == h && (s_score[w] < sc || (s_score[w] == sc && s_restart[w] < rs)))) {
            h = s_has[w]; sc = s_score[w]; rs = s_restart[w];
          }
        const unsigned long long win = (h >= 0) ? rs - P.r_begin : 0ull;  // job index of the winning attempt
        const int len = 8 + n;
        double v = 0.0;
        if (tid == 0) v = h > 0 ? 1.0 : 0.0;
        else if (tid == 1) v = speed ? (double)rs : ((volatile double*)P.cand_score)[win];
        else if (tid == 2) v = (double)rs;
        else if (tid == 3) v = ((volatile double*)P.cand_f)[win];
        else if (tid == 4) v = (double)(((volatile int*)P.cand_status)[win] & 0xff);
        else if (tid >= 8 && tid < len) v = ((volatile double*)P.cand_q)[win * n + (tid - 8)];
        if (tid < len) P.fused_record[tid] = v;
        if (P.push_peers) {  // cross-GPU: row `push_rank` of slot push_seq % NSLOT in every peer's buffer, then its flag
          const unsigned slot = (unsigned)(P.push_seq % (unsigned long long)OPTIK_EXCHANGE_NSLOT);
          for (int p = 0; p < P.push_world; p++) {
            double* base = (double*)P.push_peers[p];
            if (tid < len) base[((size_t)slot * P.push_world + P.push_rank) * len + tid] = v;
          }
          __threadfence_system();
          __syncwarp();
          if (tid < P.push_world) {
            unsigned long long* flags = (unsigned long long*)((double*)P.push_peers[tid] + (size_t)OPTIK_EXCHANGE_NSLOT * P.push_world * len);
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags + (size_t)slot * P.push_world + P.push_rank), "l"(P.push_seq) : "memory");
          }
        }
      }
    }
  }

  if (P.counters) {  // one atomic triple per warp: every thread of the warp reaches this point exactly once
    __syncwarp();
    unsigned a = n_attempts, e = n_evals, c = n_conv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(FULLMASK, a, o); e += __shfl_xor_sync(FULLMASK, e, o); c += __shfl_xor_sync(FULLMASK, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(P.counters + 0, (unsigned long long)a);
      atomicAdd(P.counters + 1, (unsigned long long)e);
      atomicAdd(P.counters + 2, (unsigned long long)c);
    }
  }
}

static int t1_threads(int n, int rows) { return (rows == 1 && n > T1_LONG_N) ? T1_THREADS_LONG : T1_THREADS_STD; }
static int t1_smem_bytes(int n, int rows) {
  const size_t units = (size_t)((3 * n) | 1), tb = (size_t)t1_threads(n, rows);
  return (int)(sizeof(double) * (OPTIK_CHAIN_STRIDE * n + 8 + 8 + 2) + 16 * rows * tb * units +
               sizeof(double) * 2 * n * tb + sizeof(double) * 4 * n);
}
template <bool GENERAL, int ROWS, int NS, int TB>
static int t1_launch(const SolveParams* p, int blocks, cudaStream_t s) {
  const int smem = t1_smem_bytes(p->n, ROWS);
  cudaError_t e = cudaFuncSetAttribute(solve_t1_kernel<GENERAL, ROWS, NS, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  solve_t1_kernel<GENERAL, ROWS, NS, TB><<<blocks, TB, smem, s>>>(*p);
  return (int)cudaGetLastError();
}
// revolute chains with unit weights and 6 or 7 joints (the common arms) run instances with the joint count baked in
template <int ROWS>
static int t1_dispatch(const SolveParams* p, int general, int blocks, cudaStream_t s) {
  if (t1_threads(p->n, ROWS) == T1_THREADS_LONG) {
    if (ROWS != 1) return (int)cudaErrorInvalidValue;
    return general ? t1_launch<true, 1, 0, T1_THREADS_LONG>(p, blocks, s) : t1_launch<false, 1, 0, T1_THREADS_LONG>(p, blocks, s);
  }
  if (general) return t1_launch<true, ROWS, 0, T1_THREADS_STD>(p, blocks, s);
  if (p->n == 7) return t1_launch<false, ROWS, 7, T1_THREADS_STD>(p, blocks, s);
  if (p->n == 6) return t1_launch<false, ROWS, 6, T1_THREADS_STD>(p, blocks, s);
  return t1_launch<false, ROWS, 0, T1_THREADS_STD>(p, blocks, s);
}
template <int ROWS>
static int t1_occupancy(int n, int* blocks_per_sm) {  // same resources for both GENERAL variants' launch bounds
  const int smem = t1_smem_bytes(n, ROWS);
  int dev = 0, optin = 0;  // a layout that cannot fit is "0 blocks", not an API error
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
    return (int)cudaGetLastError();
  if (smem > optin) { *blocks_per_sm = 0; return 0; }
  const void* fn = (ROWS == 1 && t1_threads(n, ROWS) == T1_THREADS_LONG) ? (const void*)solve_t1_kernel<false, 1, 0, T1_THREADS_LONG>
                                                                          : (const void*)solve_t1_kernel<false, ROWS, 0, T1_THREADS_STD>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, fn, t1_threads(n, ROWS), smem);
}

}  // namespace optik

// general = the chain has a prismatic joint or the config has non-unit weights; rows = 1 or 2 (see the file header)
extern "C" int optik_launch_solve_t1(const SolveParams* p, int general, int rows, int blocks, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  return rows == 1 ? optik::t1_dispatch<1>(p, general, blocks, s) : optik::t1_dispatch<2>(p, general, blocks, s);
}
extern "C" int optik_solve_t1_threads(int n, int rows) { return optik::t1_threads(n, rows); }
extern "C" int optik_solve_t1_occupancy(int n, int rows, int* blocks_per_sm) {
  return rows == 1 ? optik::t1_occupancy<1>(n, blocks_per_sm) : optik::t1_occupancy<2>(n, blocks_per_sm);
}
extern "C" int optik_launch_seed_table(const double* chain, int n, const uint32_t* key_dev, unsigned long long r_begin,
                                       unsigned long long count, double* out, void* stream) {
  if (count == 0) return 0;
  const unsigned blocks = (unsigned)((count + 127) / 128);
  optik::seed_table_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(chain, n, key_dev, r_begin, count, out);
  return (int)cudaGetLastError();
}
extern "C" int optik_launch_chacha8_kat(const uint32_t* key_dev, unsigned long long stream_id, uint32_t* out16_dev, void* stream) {
  optik::chacha8_kat_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(key_dev, stream_id, out16_dev);
  return (int)cudaGetLastError();
}
