// robot.cpp -- host side of liboptik_b200.so: the Robot object, launch plumbing and the C ABI.
//
// Mirrors, for the ik() path only, the reference's Robot / SolverConfig surface
//   crates/optik/src/lib.rs:36-99, 241-415      Robot, joint_limits, random_configuration, fk, joint_jacobian, ik
//   crates/optik/src/config.rs:22-65            SolverConfig
//   crates/optik-cpp/src/lib.rs:10-183          the extern "C" layer whose symbols are re-exported here
// The rayon restart loop and the NLopt solve live in solve_kernel.cu; this file only validates, stages buffers,
// launches, and applies the reference's budget semantics (max_time / max_restarts, lib.rs:260-277) as waves.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <atomic>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/optik_b200.h"
#include "solver_params.h"
#include "urdf.hpp"

extern "C" int optik_launch_solve(const SolveParams* p, int tile, int blocks, void* stream);
extern "C" int optik_launch_select(const SelectParams* p, unsigned slices, void* partial_scratch, void* stream);
extern "C" int optik_select_partial_bytes(void);
extern "C" int optik_solve_occupancy(int tile, int* blocks_per_sm);
extern "C" int optik_launch_eval(const EvalParams* p, int blocks, void* stream);
extern "C" int optik_eval_smem_bytes(int n, int cols);
extern "C" int optik_eval_threads(int n);
extern "C" int optik_launch_diffik(const DiffIkParams* p, int blocks, void* stream);
extern "C" int optik_eval_occupancy(int n, int cols, int* blocks_per_sm);
extern "C" int optik_launch_solve_t1(const SolveParams* p, int general, int rows, int blocks, void* stream);
extern "C" int optik_solve_t1_occupancy(int n, int rows, int* blocks_per_sm);
extern "C" int optik_solve_t1_threads(int n, int rows);
extern "C" int optik_launch_seed_table(const double* chain, int n, const uint32_t* key_dev, unsigned long long r_begin,
                                       unsigned long long count, double* out, void* stream);
extern "C" int optik_launch_chacha8_kat(const uint32_t* key_dev, unsigned long long stream_id, uint32_t* out16_dev, void* stream);
extern "C" int optik_exchange_nslot(void);
extern "C" int optik_launch_exchange_push(const double* rec, int len, const uint64_t* peers_dev, int rank, int W,
                                          unsigned long long seq, void* stream);
extern "C" int optik_launch_exchange_select(const double* buf, int W, int n, unsigned long long seq, double* out, void* stream);
extern "C" int optik_launch_flag_clamped(const double* chain, int n, const double* x0, unsigned long long T, int* status,
                                         unsigned long long status_stride, void* stream);
extern "C" int optik_launch_select_records(const double* rec, unsigned count, int n, double* out, void* stream);

namespace {

thread_local std::string g_last_error;
std::atomic<bool> g_urdf_correct_fold{false};

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
[[noreturn]] void panic(const std::string& msg) {  // a Rust panic inside extern "C" aborts the process
  fprintf(stderr, "optik_b200: panicked: %s\n", msg.c_str());
  fflush(stderr);
  abort();
}
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (cudaError_t)(expr);                                                      \
    if (_e != cudaSuccess)                                                                     \
      return fail(OPTIK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

// rand_core's SeedableRng::seed_from_u64 (PCG32 expansion) -> the ChaCha8 key used at lib.rs:359-361
void seed_key_from_u64(uint64_t state, uint32_t key[8]) {
  for (int i = 0; i < 8; i++) {
    state = state * 6364136223846793005ULL + 11634580027462260723ULL;
    const uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
    const uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
}
const uint64_t RNG_SEED = 42;  // lib.rs:359
constexpr unsigned long long SEED_CACHE = 4096;           // restart indices whose seeds every robot keeps resident
constexpr unsigned long long SEED_TABLE_MAX = 1ull << 21;  // largest per-call seed table (restarts)

void pose8_identity(double* p) {
  for (int i = 0; i < 8; i++) p[i] = 0;
  p[3] = 1;
}

// 3x3 rotation (column-major 4x4 input) -> unit quaternion (Shepperd); optik-cpp/src/lib.rs:141-144 uses
// UnitQuaternion::from_matrix, which also returns the nearest rotation for an already-orthonormal block.
void pose8_from_colmajor4x4(const double* m, double* p) {
  const double r00 = m[0], r10 = m[1], r20 = m[2], r01 = m[4], r11 = m[5], r21 = m[6], r02 = m[8], r12 = m[9], r22 = m[10];
  double q[4];
  const double tr = r00 + r11 + r22;
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q[3] = 0.25 * s; q[0] = (r21 - r12) / s; q[1] = (r02 - r20) / s; q[2] = (r10 - r01) / s;
  } else if (r00 > r11 && r00 > r22) {
    const double s = sqrt(1.0 + r00 - r11 - r22) * 2;
    q[3] = (r21 - r12) / s; q[0] = 0.25 * s; q[1] = (r01 + r10) / s; q[2] = (r02 + r20) / s;
  } else if (r11 > r22) {
    const double s = sqrt(1.0 + r11 - r00 - r22) * 2;
    q[3] = (r02 - r20) / s; q[0] = (r01 + r10) / s; q[1] = 0.25 * s; q[2] = (r12 + r21) / s;
  } else {
    const double s = sqrt(1.0 + r22 - r00 - r11) * 2;
    q[3] = (r10 - r01) / s; q[0] = (r02 + r20) / s; q[1] = (r12 + r21) / s; q[2] = 0.25 * s;
  }
  const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) p[i] = q[i] / nrm;
  p[4] = m[12]; p[5] = m[13]; p[6] = m[14]; p[7] = 0;
}
void colmajor4x4_from_pose8(const double* p, double* m) {
  const double x = p[0], y = p[1], z = p[2], w = p[3];
  m[0] = 1 - 2 * (y * y + z * z); m[4] = 2 * (x * y - z * w);     m[8] = 2 * (x * z + y * w);      m[12] = p[4];
  m[1] = 2 * (x * y + z * w);     m[5] = 1 - 2 * (x * x + z * z); m[9] = 2 * (y * z - x * w);      m[13] = p[5];
  m[2] = 2 * (x * z - y * w);     m[6] = 2 * (y * z + x * w);     m[10] = 1 - 2 * (x * x + y * y); m[14] = p[6];
  m[3] = 0; m[7] = 0; m[11] = 0; m[15] = 1;
}

struct DevBuf {  // grow-only device scratch
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    cap = bytes;
    return 0;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

// One in-flight Robot::ik() call: its own stream, persistent device scratch (control words + one wave of candidate
// records) and a mapped pinned host block the kernel reads its inputs from and writes its result + completion flag to.
constexpr int IK_SLOTS = 8;
#ifndef OPTIK_IK_WAVE
#define OPTIK_IK_WAVE 256
#endif
constexpr uint64_t IK_WAVE = OPTIK_IK_WAVE;  // restarts per launch of a single-target call
struct IkSlot {
  std::atomic<bool> busy{false};
  bool ready = false;
  cudaStream_t stream = nullptr;
  char* dev = nullptr;       // [queue 8 | fused_done 8 | found 8 | pad] then the candidate arrays
  double* host = nullptr;    // mapped: [0..7] target, [8..40) x0, [64..64+8+n) record, [128] completion flag
  double* host_dev = nullptr;
  unsigned long long seq = 0;
};

struct optik_robot {
  std::vector<optik::Joint> joints;  // articulated joints, then at most one fixed tip joint
  int n = 0;                         // num_positions
  bool kernel_ok = true;             // chain shape supported by the kernels
  std::string why_not;
  unsigned parallelism = 0;
  int device = 0;
  // lazily created GPU state
  mutable std::mutex mu;
  mutable bool gpu_ready = false;
  mutable DevBuf chain_dev;
  mutable uint32_t chain_bytes = 0;
  mutable int sm_count = 0;
  mutable int occ[3] = {0, 0, 0};  // resident blocks/SM for TILE 8,16,32
  mutable int occ_t1[2] = {0, 0};  // resident blocks/SM of the thread-per-seed kernel: rows = 1, rows = 2
  mutable DevBuf scratch;          // single-call scratch (ik / fk / jacobian), guarded by mu
  mutable cudaStream_t stream = nullptr;
  mutable void* pinned = nullptr;
  mutable size_t pinned_cap = 0;
  mutable cudaMemPool_t pool = nullptr;  // library-owned stream-ordered pool (the process's default pool is left alone)
  mutable DevBuf key_dev;                // ChaCha8 key = seed_from_u64(42), 8 words
  mutable DevBuf seed_cache;             // restart seeds for indices [0, SEED_CACHE) (lib.rs:360-370)
  // the first long restart range a caller asks for (e.g. a rank's 65 536 restarts) keeps its seed table too
  mutable std::mutex seed_mu;
  mutable DevBuf seed_big;
  mutable uint64_t seed_big_begin = 0, seed_big_count = 0;
  mutable cudaEvent_t seed_big_ready = nullptr;
  mutable IkSlot ik_slots[IK_SLOTS];

  ~optik_robot() {
    if (stream) cudaStreamDestroy(stream);
    if (seed_big_ready) cudaEventDestroy(seed_big_ready);
    for (IkSlot& sl : ik_slots) {
      if (sl.stream) cudaStreamDestroy(sl.stream);
      if (sl.dev) cudaFree(sl.dev);
      if (sl.host) cudaFreeHost(sl.host);
    }
    if (pinned) cudaFreeHost(pinned);
    if (pool) cudaMemPoolDestroy(pool);
  }
  void finish_init() {
    n = 0;
    for (size_t i = 0; i < joints.size(); i++) {
      if (joints[i].type != optik::FIXED) n++;
      else if (i + 1 != joints.size()) { kernel_ok = false; why_not = "fixed joint in the middle of the chain"; }
    }
    if (n > OPTIK_MAX_DOF) { kernel_ok = false; why_not = "more than 32 articulated joints"; }
  }
  std::vector<double> flat_chain() const {
    std::vector<double> c(joints.size() * OPTIK_CHAIN_STRIDE, 0.0);
    for (size_t i = 0; i < joints.size(); i++) {
      double* j = &c[i * OPTIK_CHAIN_STRIDE];
      const optik::Joint& J = joints[i];
      j[0] = J.origin.t[0]; j[1] = J.origin.t[1]; j[2] = J.origin.t[2]; j[3] = (double)J.type;
      j[4] = J.origin.q[0]; j[5] = J.origin.q[1]; j[6] = J.origin.q[2]; j[7] = J.origin.q[3];
      j[8] = J.axis[0]; j[9] = J.axis[1]; j[10] = J.axis[2];
      j[12] = J.lower; j[13] = J.upper;
    }
    return c;
  }
  // device blob: n joint records (128 B each, with seed-sampling limits) + fixed tip pose8
  int ensure_gpu() const {
    if (gpu_ready) return 0;
    if (!kernel_ok) return fail(OPTIK_ERR_UNSUPPORTED, "chain not supported by the GPU kernels: " + why_not);
    CUDA_TRY(cudaSetDevice(device));
    std::vector<double> blob((size_t)n * OPTIK_CHAIN_STRIDE + 8, 0.0);
    const double PI = 3.14159265358979311600e+00;
    int k = 0;
    pose8_identity(&blob[(size_t)n * OPTIK_CHAIN_STRIDE]);
    for (const optik::Joint& J : joints) {
      if (J.type == optik::FIXED) {
        double* t = &blob[(size_t)n * OPTIK_CHAIN_STRIDE];
        t[0] = J.origin.q[0]; t[1] = J.origin.q[1]; t[2] = J.origin.q[2]; t[3] = J.origin.q[3];
        t[4] = J.origin.t[0]; t[5] = J.origin.t[1]; t[6] = J.origin.t[2]; t[7] = 0;
        continue;
      }
      double* j = &blob[(size_t)k * OPTIK_CHAIN_STRIDE];
      j[0] = J.origin.t[0]; j[1] = J.origin.t[1]; j[2] = J.origin.t[2]; j[3] = (double)J.type;
      j[4] = J.origin.q[0]; j[5] = J.origin.q[1]; j[6] = J.origin.q[2]; j[7] = J.origin.q[3];
      j[8] = J.axis[0]; j[9] = J.axis[1]; j[10] = J.axis[2];
      j[12] = J.lower; j[13] = J.upper;
      const bool finite = std::isfinite(J.lower) && std::isfinite(J.upper);
      j[14] = finite ? J.lower : -PI;  // infinite limits: the reference's uniform draw is undefined; we use [-pi, pi]
      j[15] = finite ? J.upper : PI;
      k++;
    }
    chain_bytes = (uint32_t)(blob.size() * sizeof(double));
    if (chain_dev.reserve(chain_bytes)) return fail(OPTIK_ERR_CUDA, "cudaMalloc(chain) failed");
    CUDA_TRY(cudaMemcpy(chain_dev.p, blob.data(), chain_bytes, cudaMemcpyHostToDevice));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    sm_count = prop.multiProcessorCount;
    const int tiles[3] = {8, 16, 32};
    for (int i = 0; i < 3; i++) CUDA_TRY(optik_solve_occupancy(tiles[i], &occ[i]));
    for (int rows = 1; rows <= 2; rows++)  // 0 blocks/SM (or an error) = this chain's rows do not fit
      if (optik_solve_t1_occupancy(n, rows, &occ_t1[rows - 1]) != 0) { occ_t1[rows - 1] = 0; cudaGetLastError(); }
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    // stream-ordered scratch comes from a pool this robot owns; its release threshold keeps blocks cached across
    // synchronisations (with the default of 0 every host-path call would turn into a fresh cudaMalloc)
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CUDA_TRY(cudaMemPoolCreate(&pool, &props));
    unsigned long long keep = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    // ChaCha8 key and the restart seeds every call shares (restart i >= 1 is the same draw for every target)
    uint32_t key[8];
    seed_key_from_u64(RNG_SEED, key);
    if (key_dev.reserve(sizeof(key))) return fail(OPTIK_ERR_CUDA, "cudaMalloc(key) failed");
    CUDA_TRY(cudaMemcpy(key_dev.p, key, sizeof(key), cudaMemcpyHostToDevice));
    if (occ_t1[0] > 0) {
      if (seed_cache.reserve((size_t)SEED_CACHE * n * sizeof(double))) return fail(OPTIK_ERR_CUDA, "cudaMalloc(seeds) failed");
      CUDA_TRY(optik_launch_seed_table((const double*)chain_dev.p, n, (const uint32_t*)key_dev.p, 0, SEED_CACHE,
                                       (double*)seed_cache.p, stream));
      CUDA_TRY(cudaStreamSynchronize(stream));
    }
    gpu_ready = true;
    return 0;
  }
};

namespace {

optik_robot* robot_from_chain(std::vector<optik::Joint>&& joints) {
  auto* r = new optik_robot();
  r->joints = std::move(joints);
  r->finish_init();
  return r;
}

optik_robot* try_from_urdf_str(const char* urdf, const char* base, const char* ee) {
  try {
    if (!urdf || !base || !ee) throw std::runtime_error("null argument");
    return robot_from_chain(optik::chain_from_urdf(urdf, base, ee, g_urdf_correct_fold.load()));
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return nullptr;
  }
}

struct Plan {
  int tile, blocks, tiles_per_block, resident_tiles, rows;
};
// lanes per restart seed: 1 = thread-per-seed kernel, 8/16/32 = tile kernel.  `batch`: auto picks the
// throughput layout (1) when it exists, otherwise / for single-target latency the smallest tile that fits.
int choose_tile(const optik_robot* r, uint32_t want, bool batch) {
  const int n = r->n;
  const bool t1_ok = r->occ_t1[0] > 0;  // the per-thread rows of this chain fit an SM's shared memory
  if (want == 1) return t1_ok ? 1 : 0;
  if (want == 8 || want == 16 || want == 32) return (int)want >= n ? (int)want : 0;
  if (want != 0) return 0;
  if (batch && t1_ok) return 1;
  return n <= 8 ? 8 : (n <= 16 ? 16 : 32);
}
// thread-per-seed kernel: rows = 2 keeps the trial columns in shared memory (2 blocks/SM), rows = 1 in local memory
// (3 blocks/SM).  opts.variant forces one; the default is the measured better one.
int t1_rows(const optik_robot* r, const optik_gpu_batch_opts* o, double attempts, bool speed_chains) {
  if (r->occ_t1[1] <= 0) return 1;  // long chains: two column rows per thread do not fit
  if (o && (o->variant == 1 || o->variant == 2)) return (int)o->variant;
  // Speed batches of 7-joint arms: the one-row instance keeps its trial columns (6n doubles) in registers and spills
  // from n = 7 on; with the divergence of Speed chains the two-row layout is then 3-6 % faster at every batch size
  // (Panda 1 Mi targets 6.04 -> 5.83 ms; UR5, n = 6, is 4 % faster with one row; Quality batches 5 % faster with one row)
  if (speed_chains && r->n >= 7) return 2;
  // three blocks per SM win once every lane runs several attempts (throughput); with about one attempt per lane the
  // launch is a latency chain and two blocks per SM with the trial columns in shared memory are faster
  const double lanes1 = (double)r->sm_count * (r->occ_t1[0] > 0 ? r->occ_t1[0] : 3) * (double)optik_solve_t1_threads(r->n, 1);
  return attempts >= 4.0 * lanes1 ? 1 : 2;
}
Plan make_plan(const optik_robot* r, int tile, int rows, uint32_t blocks_req, unsigned long long njobs) {
  Plan p;
  p.tile = tile;
  p.rows = rows;
  p.tiles_per_block = tile == 1 ? optik_solve_t1_threads(r->n, rows) : 128 / tile;
  const int oi = tile == 8 ? 0 : (tile == 16 ? 1 : 2);
  int per_sm = r->occ[oi] > 0 ? r->occ[oi] : 1;
  if (tile == 1) per_sm = r->occ_t1[rows - 1] > 0 ? r->occ_t1[rows - 1] : 1;
  const long long resident_blocks = (long long)r->sm_count * per_sm;  // one full wave: a multiple of the SM count
  p.resident_tiles = (int)(resident_blocks * p.tiles_per_block);
  long long need = (long long)((njobs + p.tiles_per_block - 1) / p.tiles_per_block);
  long long b = blocks_req ? blocks_req : (need < resident_blocks ? need : resident_blocks);
  if (b < 1) b = 1;
  p.blocks = (int)b;
  return p;
}

void fill_common(const optik_robot* r, const optik_solver_config* cfg, const double* ee_offset, uint32_t max_evals,
                 SolveParams& P) {
  P.chain = (const double*)r->chain_dev.p;
  P.n = r->n;
  P.chain_bytes = r->chain_bytes;
  P.mode = cfg->solution_mode;
  P.tol_f = cfg->tol_f;
  P.tol_df_eff = cfg->tol_df > 0.0 ? cfg->tol_df : 1e-3 * cfg->tol_f;  // lib.rs:283-293
  P.tol_df_user = cfg->tol_df;
  P.tol_dx = cfg->tol_dx;
  P.weighted = 0;
  for (int i = 0; i < 3; i++) {
    P.wl[i] = cfg->linear_weight[i];
    P.wa[i] = cfg->angular_weight[i];
    if (P.wl[i] != 1.0 || P.wa[i] != 1.0) P.weighted = 1;
  }
  P.has_prismatic = 0;
  for (const optik::Joint& J : r->joints) if (J.type == optik::PRISMATIC) P.has_prismatic = 1;
  P.max_evals = max_evals ? (int)max_evals : OPTIK_LM_MAX_EVALS;
  P.lambda0 = OPTIK_LM_LAMBDA0; P.lambda_dec = OPTIK_LM_LAMBDA_DEC; P.lambda_inc = OPTIK_LM_LAMBDA_INC;
  P.lambda_min = OPTIK_LM_LAMBDA_MIN; P.lambda_max = OPTIK_LM_LAMBDA_MAX;
  P.stall_rel = OPTIK_LM_STALL_REL; P.stall_count = OPTIK_LM_STALL_COUNT;
  seed_key_from_u64(RNG_SEED, P.key);
  if (ee_offset) for (int i = 0; i < 8; i++) P.ee_offset[i] = ee_offset[i];
  else pose8_identity(P.ee_offset);
}

bool config_valid(const optik_solver_config* c) {
  return c && (c->solution_mode == OPTIK_MODE_QUALITY || c->solution_mode == OPTIK_MODE_SPEED);
}

// stream-ordered scratch from the robot's pool, released (stream-ordered) on every exit path
struct StreamBuf {
  char* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(const optik_robot* r, size_t bytes, cudaStream_t stream) {
    s = stream;
    return cudaMallocFromPoolAsync((void**)&p, bytes ? bytes : 1, r->pool, stream);
  }
  ~StreamBuf() { if (p) cudaFreeAsync(p, s); }
};

#ifndef OPTIK_POOL_DIV
#define OPTIK_POOL_DIV 32  /* a warp claims 1/(32 * warps) of the jobs per queue atomic (at most 64): measured against 8, 16
                              and 128 -- smaller private pools balance the tail of 1 Mi-target batches 2-3 % better */
#endif
constexpr uint64_t UNBOUNDED_RESTARTS = 0xfffffffeull;  // max_restarts == 0 with a max_time: restarts until the deadline

// Device-side batched solve on `stream`; every pointer is device memory.  Stream-ordered scratch, no host sync.
int solve_device(const optik_robot* r, const optik_solver_config* cfg, const optik_gpu_batch_opts* o,
                 const double* d_targets, const double* d_x0, uint64_t T, uint64_t r_begin, uint64_t R, double* d_q,
                 double* d_f, int32_t* d_status, uint64_t* d_restart, int32_t* d_evals, uint64_t* d_counters,
                 unsigned long long max_ns, bool per_attempt_records, cudaStream_t s, double* d_best_record = nullptr) {
  const int tile = choose_tile(r, o ? o->tile : 0, true);
  if (!tile) return fail(OPTIK_ERR_INVALID, "opts.tile must be 1 (chains whose rows fit shared memory), 8, 16 or 32 and >= num_positions");
  const bool speed = cfg->solution_mode == OPTIK_MODE_SPEED;
  const int rows = t1_rows(r, o, per_attempt_records ? (double)R : (speed ? 2.0 * (double)T : (double)T * (double)(R < 4096 ? R : 4096)),
                           speed && !per_attempt_records);
  SolveParams P{};
  fill_common(r, cfg, o ? o->ee_offset : nullptr, o ? o->max_evals : 0, P);
  P.targets = d_targets; P.x0 = d_x0; P.T = T; P.r_begin = r_begin; P.r_end = r_begin + R;
  P.max_ns = max_ns;
  P.counters = (unsigned long long*)d_counters;
  const int n = r->n;
  Plan plan0 = make_plan(r, tile, rows, o ? o->blocks : 0, ~0ull);
  // ---- scheduling.  Speed batches of the thread-per-seed kernel with enough targets (a quarter of the resident lanes:
  // below that the static schedule's private candidate records beat the shared-record protocol, measured 5 000-8 192
  // Panda targets) run as dynamic chains (one launch, restarts claimed on the device); everything else as static
  // (target, chunk) jobs.
  const bool dyn = tile == 1 && speed && !per_attempt_records && !(o && o->chunks) && !(o && (o->flags & OPTIK_BATCH_STATIC)) &&
                   T * 4 >= (uint64_t)plan0.resident_tiles && T < 0xfffffffeull && R <= UNBOUNDED_RESTARTS;
  uint64_t C = o ? o->chunks : 0;
  if (per_attempt_records) C = R;
  if (dyn) C = 1;
  if (C == 0) {
    const uint64_t want = 2ull * (uint64_t)plan0.resident_tiles;
    C = T >= want ? 1 : (want + T - 1) / T;
  }
  if (C > R) C = R;
  if (C < 1) C = 1;
  if (C > 0xffffffffull) return fail(OPTIK_ERR_INVALID, "too many chunks");
  P.C = (uint32_t)C;
  const unsigned long long njobs = T * C;
  if (C > 1 && !per_attempt_records && T > 0x7fffffffull)
    return fail(OPTIK_ERR_INVALID, "T too large for a selection pass; use chunks = 1");
  Plan plan = make_plan(r, tile, rows, o ? o->blocks : 0, njobs);
  const bool direct = (C == 1) || per_attempt_records;  // candidate records ARE the outputs
  // ---- scratch: optional outputs the caller did not ask for + candidate arrays when a selection pass follows
  size_t bytes = 0;
  auto carve = [&](size_t b) { size_t off = bytes; bytes += (b + 255) & ~size_t(255); return off; };
  const size_t off_q = direct ? 0 : carve(njobs * n * sizeof(double));
  const size_t off_f = direct ? 0 : carve(njobs * sizeof(double));
  const size_t off_st = direct ? 0 : carve(njobs * sizeof(int));
  const size_t off_score = dyn ? 0 : carve(njobs * sizeof(double));
  const bool restart_direct = direct && d_restart && !per_attempt_records;  // per-attempt: restart_out is the winner's
  const size_t off_rs = (restart_direct || dyn) ? 0 : carve(njobs * sizeof(unsigned long long));
  const size_t off_ev = ((direct && d_evals) || dyn) ? 0 : carve(njobs * sizeof(int));
  const bool use_found = speed && C > 1 && !per_attempt_records;
  const size_t off_found = use_found ? carve(T * sizeof(unsigned long long)) : 0;
  // dynamic chains: [queue (one 128-byte line) | next[T]] zeroed together, the per-target found / record words [T] set to ~0
  const size_t dyn_zero = 128 + T * sizeof(unsigned);
  const size_t off_zero = carve(dyn ? dyn_zero : 16);  // static: [queue (8 B) | fused_done (4 B)]
  const bool fused = tile == 1 && per_attempt_records && T == 1 && d_best_record != nullptr;  // selection inside the solve launch
  const size_t off_fpart = fused ? carve((size_t)plan.blocks * 24) : 0;
  const size_t off_dfound = dyn ? carve(T * sizeof(unsigned long long)) : 0;
  // selection: slice the candidate range when one target has very many chunks
  unsigned slices = 1;
  if (C >= 4096) { slices = (unsigned)((C + 1023) / 1024); if (slices > 256) slices = 256; }
  const size_t off_part = slices > 1 ? carve((size_t)T * slices * optik_select_partial_bytes()) : 0;
  // restart seeds: the robot's resident table, a per-call table for long restart ranges, in-kernel draws beyond
  size_t off_seed = 0;
  bool own_table = false;
  if (tile == 1) {
    if (r_begin + R <= SEED_CACHE || R > SEED_TABLE_MAX) {
      P.seed_tab = (const double*)r->seed_cache.p; P.seed_begin = 0; P.seed_count = SEED_CACHE;
    } else {
      std::lock_guard<std::mutex> lk(r->seed_mu);
      if (r->seed_big_count == 0) {  // first long range of this robot: build its table once, on this stream
        if (r->seed_big.reserve((size_t)R * n * sizeof(double)) == 0 &&
            cudaEventCreateWithFlags(&r->seed_big_ready, cudaEventDisableTiming) == cudaSuccess) {
          CUDA_TRY(optik_launch_seed_table(P.chain, n, (const uint32_t*)r->key_dev.p, r_begin, R, (double*)r->seed_big.p, s));
          CUDA_TRY(cudaEventRecord(r->seed_big_ready, s));
          r->seed_big_begin = r_begin; r->seed_big_count = R;
        }
      }
      if (r->seed_big_count && r_begin >= r->seed_big_begin && r_begin + R <= r->seed_big_begin + r->seed_big_count) {
        CUDA_TRY(cudaStreamWaitEvent(s, r->seed_big_ready, 0));
        P.seed_tab = (const double*)r->seed_big.p; P.seed_begin = r->seed_big_begin; P.seed_count = r->seed_big_count;
      } else {
        own_table = true;
        off_seed = carve((size_t)R * n * sizeof(double));
      }
    }
  }
  StreamBuf scratch;
  CUDA_TRY(scratch.alloc(r, bytes, s));
  char* sc = scratch.p;
  if (own_table) {
    P.seed_tab = (const double*)(sc + off_seed); P.seed_begin = r_begin; P.seed_count = R;
    CUDA_TRY(optik_launch_seed_table(P.chain, n, (const uint32_t*)r->key_dev.p, r_begin, R, (double*)(sc + off_seed), s));
  }
  P.cand_q = direct ? d_q : (double*)(sc + off_q);
  P.cand_f = direct ? d_f : (double*)(sc + off_f);
  P.cand_status = direct ? d_status : (int*)(sc + off_st);
  P.cand_score = dyn ? nullptr : (double*)(sc + off_score);
  P.cand_restart = (restart_direct || dyn) ? (unsigned long long*)d_restart : (unsigned long long*)(sc + off_rs);
  P.cand_evals = ((direct && d_evals) || dyn) ? d_evals : (int*)(sc + off_ev);
  P.queue = (unsigned long long*)(sc + off_zero);
  P.found = nullptr;
  if (dyn) {
    P.sched = 1;
    P.dyn_next = (unsigned*)(sc + off_zero + 128);
    P.dyn_word = (unsigned long long*)(sc + off_dfound);
    const uint64_t lanes = (uint64_t)plan0.resident_tiles;
    P.dyn_k0 = T >= lanes ? 0u : (unsigned)((lanes + T - 1) / T - 1);
    if (P.dyn_k0 > 7) P.dyn_k0 = 7;
    CUDA_TRY(cudaMemsetAsync(sc + off_zero, 0, dyn_zero, s));
    CUDA_TRY(cudaMemsetAsync(P.dyn_word, 0xff, T * sizeof(unsigned long long), s));
    if (d_evals) CUDA_TRY(cudaMemsetAsync(d_evals, 0, T * sizeof(int32_t), s));
  } else {
    P.sched = (per_attempt_records && tile == 1 && T == 1) ? 2 : 0;
    CUDA_TRY(cudaMemsetAsync(P.queue, 0, 16, s));
    if (fused) {
      P.fused_record = d_best_record;
      P.fused_done = (unsigned*)(sc + off_zero + 8);
      P.fused_part_score = (double*)(sc + off_fpart);
      P.fused_part_restart = (unsigned long long*)(sc + off_fpart + (size_t)plan.blocks * 8);
      P.fused_part_has = (int*)(sc + off_fpart + (size_t)plan.blocks * 16);
      if (o && o->push_peers) {
        if (o->push_world == 0 || o->push_world > 32 || o->push_rank >= o->push_world || o->push_seq == 0)
          return fail(OPTIK_ERR_INVALID, "bad exchange arguments in opts (1 <= push_world <= 32, push_rank < push_world, push_seq >= 1)");
        P.push_peers = o->push_peers; P.push_rank = (int)o->push_rank; P.push_world = (int)o->push_world; P.push_seq = o->push_seq;
      }
    }
    if (use_found) {
      P.found = (unsigned long long*)(sc + off_found);
      CUDA_TRY(cudaMemsetAsync(P.found, 0xff, T * sizeof(unsigned long long), s));
    }
  }
  {  // warp-level job pools: large enough to amortise the queue atomic, small enough to keep the tail balanced
    const unsigned long long warps = (unsigned long long)plan.blocks * (unsigned long long)(plan.tile == 1 ? optik_solve_t1_threads(n, rows) / 32 : 4);
    unsigned long long chunk = (dyn ? T * (unsigned long long)(P.dyn_k0 + 1u) : njobs) / (warps * (unsigned long long)OPTIK_POOL_DIV);
    P.pool_chunk = (unsigned)(chunk < 1 ? 1 : (chunk > 64 ? 64 : chunk));
  }
  if (plan.tile == 1) CUDA_TRY(optik_launch_solve_t1(&P, (P.has_prismatic || P.weighted) ? 1 : 0, rows, plan.blocks, s));
  else CUDA_TRY(optik_launch_solve(&P, plan.tile, plan.blocks, s));
  if (per_attempt_records && d_best_record && !fused) {  // selection pass over the per-attempt records -> one packed record
    SelectParams S{};
    S.T = 1; S.C = P.C; S.n = n; S.mode = cfg->solution_mode;
    S.cand_q = P.cand_q; S.cand_f = P.cand_f; S.cand_score = P.cand_score; S.cand_restart = P.cand_restart;
    S.cand_status = P.cand_status; S.cand_evals = P.cand_evals;
    S.tol_f = cfg->tol_f; S.tol_df_user = cfg->tol_df; S.tol_dx = cfg->tol_dx;
    S.record_out = d_best_record;
    CUDA_TRY(optik_launch_select(&S, slices, slices > 1 ? sc + off_part : nullptr, s));
  }
  if (!direct) {
    SelectParams S{};
    S.T = T; S.C = P.C; S.n = n; S.mode = cfg->solution_mode;
    S.cand_q = P.cand_q; S.cand_f = P.cand_f; S.cand_score = P.cand_score; S.cand_restart = P.cand_restart;
    S.cand_status = P.cand_status; S.cand_evals = P.cand_evals;
    S.tol_f = cfg->tol_f; S.tol_df_user = cfg->tol_df; S.tol_dx = cfg->tol_dx;
    S.q_out = d_q; S.f_out = d_f; S.status_out = d_status;
    S.restart_out = (unsigned long long*)d_restart;  // optional
    S.evals_out = d_evals;                            // optional
    CUDA_TRY(optik_launch_select(&S, slices, slices > 1 ? sc + off_part : nullptr, s));
  }
  return OPTIK_OK;
}

int check_seeds_host(const optik_robot* r, const double* x0, uint64_t T) {  // lib.rs:251-254
  const int n = r->n;
  for (uint64_t t = 0; t < T; t++) {
    int k = 0;
    for (const optik::Joint& J : r->joints) {
      if (J.type == optik::FIXED) continue;
      const double q = x0[t * n + k];
      if (q < J.lower || q > J.upper) return fail(OPTIK_ERR_SEED_LIMITS, "seed joint position outside of joint limits");
      k++;
    }
  }
  return OPTIK_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* optik_last_error(void) { return g_last_error.c_str(); }
void optik_set_urdf_correct_fold(int on) { g_urdf_correct_fold.store(on != 0); }

optik_robot* optik_robot_try_from_urdf_str(const char* urdf, const char* base_link, const char* ee_link) {
  return try_from_urdf_str(urdf, base_link, ee_link);
}
optik_robot* optik_robot_from_urdf_str(const char* urdf, const char* base_link, const char* ee_link) {
  optik_robot* r = try_from_urdf_str(urdf, base_link, ee_link);
  if (!r) panic(g_last_error);
  return r;
}
optik_robot* optik_robot_from_urdf_file(const char* path, const char* base_link, const char* ee_link) {
  if (!path) panic("null path");
  std::ifstream f(path, std::ios::binary);
  if (!f) panic("error parsing URDF file!: cannot open '" + std::string(path) + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  return optik_robot_from_urdf_str(ss.str().c_str(), base_link, ee_link);
}
optik_robot* optik_robot_from_chain(const double* chain, unsigned int njoints) {
  if (!chain || njoints == 0) { g_last_error = "kinematic chain is empty"; return nullptr; }
  std::vector<optik::Joint> joints(njoints);
  int nq = 0;
  for (unsigned i = 0; i < njoints; i++) {
    const double* j = chain + (size_t)i * OPTIK_CHAIN_STRIDE;
    optik::Joint& J = joints[i];
    J.type = (int)j[3];
    if (J.type < 0 || J.type > 2) { g_last_error = "joint type not supported"; return nullptr; }
    J.origin.t[0] = j[0]; J.origin.t[1] = j[1]; J.origin.t[2] = j[2];
    J.origin.q[0] = j[4]; J.origin.q[1] = j[5]; J.origin.q[2] = j[6]; J.origin.q[3] = j[7];
    J.axis[0] = j[8]; J.axis[1] = j[9]; J.axis[2] = j[10];
    J.lower = j[12]; J.upper = j[13];
    nq += (J.type != optik::FIXED);
  }
  if (nq == 0) { g_last_error = "kinematic chain is empty"; return nullptr; }
  return robot_from_chain(std::move(joints));
}
void optik_robot_free(optik_robot* robot) { delete robot; }

void optik_robot_set_parallelism(optik_robot* robot, unsigned int n) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  robot->parallelism = n;
}
unsigned int optik_robot_num_positions(const optik_robot* robot) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  return (unsigned)robot->n;
}
unsigned int optik_robot_num_joints(const optik_robot* robot) { return robot ? (unsigned)robot->joints.size() : 0; }
int optik_robot_chain(const optik_robot* robot, double* out) {
  if (!robot || !out) return fail(OPTIK_ERR_INVALID, "null argument");
  std::vector<double> c = robot->flat_chain();
  memcpy(out, c.data(), c.size() * sizeof(double));
  return OPTIK_OK;
}
int optik_robot_set_device(optik_robot* robot, int device) {
  if (!robot) return fail(OPTIK_ERR_INVALID, "null robot");
  std::lock_guard<std::mutex> lk(robot->mu);
  if (robot->gpu_ready && robot->device != device) return fail(OPTIK_ERR_INVALID, "device already initialised");
  robot->device = device;
  return OPTIK_OK;
}
int optik_status_is_success(const optik_solver_config* c, int st) {
  if (!c) return 0;
  st &= OPTIK_STATUS_CODE_MASK;
  return (c->tol_f >= 0.0 && st == OPTIK_STATUS_STOPVAL) || (c->tol_df >= 0.0 && st == OPTIK_STATUS_FTOL) ||
         (c->tol_dx >= 0.0 && st == OPTIK_STATUS_XTOL);
}

double* optik_robot_joint_limits(const optik_robot* robot) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  double* out = (double*)malloc(sizeof(double) * 2 * robot->n);
  int k = 0;
  for (const optik::Joint& J : robot->joints) {
    if (J.type == optik::FIXED) continue;
    out[k] = J.lower;
    out[robot->n + k] = J.upper;
    k++;
  }
  return out;
}
double* optik_robot_random_configuration(const optik_robot* robot) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  static thread_local std::mt19937_64 rng{std::random_device{}()};  // rand::rng(): thread-local, OS-seeded
  double* out = (double*)malloc(sizeof(double) * robot->n);
  int k = 0;
  for (const optik::Joint& J : robot->joints) {
    if (J.type == optik::FIXED) continue;
    if (!(std::isfinite(J.lower) && std::isfinite(J.upper))) panic("cannot sample a joint with infinite limits");
    out[k++] = std::uniform_real_distribution<double>(J.lower, std::nextafter(J.upper, INFINITY))(rng);
  }
  return out;
}

int optik_gpu_select_records(const optik_robot* robot, const double* records, uint32_t count, double* best_record_out,
                             void* stream) {
  if (!robot || !records || !best_record_out || count == 0) return fail(OPTIK_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(robot->device));
  CUDA_TRY(optik_launch_select_records(records, count, robot->n, best_record_out, stream));
  return OPTIK_OK;
}

int optik_gpu_restart_seeds(const optik_robot* robot, uint64_t restart_begin, uint64_t count, int memory, double* seeds_out,
                            void* stream) {
  if (!robot || !seeds_out) return fail(OPTIK_ERR_INVALID, "null argument");
  if (restart_begin == 0 && count) return fail(OPTIK_ERR_INVALID, "restart 0 is the caller's seed; restart_begin must be >= 1");
  if (count == 0) return OPTIK_OK;
  {
    std::lock_guard<std::mutex> lk(robot->mu);
    if (int rc = robot->ensure_gpu()) return rc;
  }
  CUDA_TRY(cudaSetDevice(robot->device));
  const int n = robot->n;
  if (memory == 1) {
    CUDA_TRY(optik_launch_seed_table((const double*)robot->chain_dev.p, n, (const uint32_t*)robot->key_dev.p, restart_begin,
                                     count, seeds_out, stream));
    return OPTIK_OK;
  }
  std::unique_lock<std::mutex> lk(robot->mu, std::defer_lock);
  cudaStream_t s = (cudaStream_t)stream;
  if (!s) { lk.lock(); s = robot->stream; }
  StreamBuf buf;
  CUDA_TRY(buf.alloc(robot, count * n * sizeof(double), s));
  CUDA_TRY(optik_launch_seed_table((const double*)robot->chain_dev.p, n, (const uint32_t*)robot->key_dev.p, restart_begin,
                                   count, (double*)buf.p, s));
  CUDA_TRY(cudaMemcpyAsync(seeds_out, buf.p, count * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return OPTIK_OK;
}

int optik_gpu_chacha8_block(const optik_robot* robot, const uint32_t* key8, uint64_t stream_id, uint32_t* words16_out) {
  if (!robot || !key8 || !words16_out) return fail(OPTIK_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(robot->mu);
  if (int rc = robot->ensure_gpu()) return rc;
  CUDA_TRY(cudaSetDevice(robot->device));
  StreamBuf buf;
  CUDA_TRY(buf.alloc(robot, 256, robot->stream));
  CUDA_TRY(cudaMemcpyAsync(buf.p, key8, 32, cudaMemcpyHostToDevice, robot->stream));
  CUDA_TRY(optik_launch_chacha8_kat((const uint32_t*)buf.p, stream_id, (uint32_t*)(buf.p + 128), robot->stream));
  CUDA_TRY(cudaMemcpyAsync(words16_out, buf.p + 128, 64, cudaMemcpyDeviceToHost, robot->stream));
  CUDA_TRY(cudaStreamSynchronize(robot->stream));
  return OPTIK_OK;
}

uint64_t optik_gpu_exchange_bytes(const optik_robot* robot, uint32_t world) {
  if (!robot || world == 0) return 0;
  const uint64_t len = OPTIK_RECORD_HEAD + (uint64_t)robot->n;
  return (uint64_t)optik_exchange_nslot() * world * (len * sizeof(double) + sizeof(uint64_t));
}
int optik_gpu_exchange_push(const optik_robot* robot, const double* record, const uint64_t* peer_buffers_dev, uint32_t rank,
                            uint32_t world, uint64_t seq, void* stream) {
  if (!robot || !record || !peer_buffers_dev || world == 0 || world > 32 || rank >= world || seq == 0)
    return fail(OPTIK_ERR_INVALID, "bad exchange arguments (1 <= world <= 32, rank < world, seq >= 1)");
  CUDA_TRY(cudaSetDevice(robot->device));
  CUDA_TRY(optik_launch_exchange_push(record, OPTIK_RECORD_HEAD + robot->n, peer_buffers_dev, (int)rank, (int)world, seq, stream));
  return OPTIK_OK;
}
int optik_gpu_exchange_select(const optik_robot* robot, const double* local_buffer, uint32_t world, uint64_t seq,
                              double* best_record_out, void* stream) {
  if (!robot || !local_buffer || !best_record_out || world == 0 || world > 32 || seq == 0)
    return fail(OPTIK_ERR_INVALID, "bad exchange arguments (1 <= world <= 32, seq >= 1)");
  CUDA_TRY(cudaSetDevice(robot->device));
  CUDA_TRY(optik_launch_exchange_select(local_buffer, (int)world, robot->n, seq, best_record_out, stream));
  return OPTIK_OK;
}

void* optik_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { g_last_error = "cudaMallocHost failed"; return nullptr; }
  return p;
}
void optik_host_free(void* p) { if (p) cudaFreeHost(p); }

int optik_gpu_stream_create(const optik_robot* robot, void** stream_out) {
  if (!robot || !stream_out) return fail(OPTIK_ERR_INVALID, "null argument");
  {
    std::lock_guard<std::mutex> lk(robot->mu);
    if (int rc = robot->ensure_gpu()) return rc;
  }
  CUDA_TRY(cudaSetDevice(robot->device));
  cudaStream_t s = nullptr;
  CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream_out = (void*)s;
  return OPTIK_OK;
}
int optik_gpu_stream_sync(void* stream) {
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return OPTIK_OK;
}
void optik_gpu_stream_destroy(void* stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }

// ---------------------------------------------------------------- batched evaluator
int optik_gpu_eval_batch(const optik_robot* robot, const double* q, const double* targets, int shared_target,
                         uint64_t B, const double* linear_weight, const double* angular_weight,
                         const double* ee_offset, int memory, double* ee_out, double* jac_out, double* f_out,
                         double* grad_out, void* stream) {
  if (!robot || !q) return fail(OPTIK_ERR_INVALID, "null argument");
  if ((f_out || grad_out) && !targets) return fail(OPTIK_ERR_INVALID, "objective outputs need targets");
  if (memory == 1 && (((uintptr_t)jac_out | (uintptr_t)ee_out) & 15)) return fail(OPTIK_ERR_INVALID, "ee_out / jac_out must be 16-byte aligned");
  if (B == 0) return OPTIK_OK;
  std::unique_lock<std::mutex> lk(robot->mu);
  if (int rc = robot->ensure_gpu()) return rc;
  CUDA_TRY(cudaSetDevice(robot->device));
  const int n = robot->n;
  const int cols = (jac_out || grad_out) ? 1 : 0;
  if (optik_eval_smem_bytes(n, cols) > 227 * 1024) return fail(OPTIK_ERR_UNSUPPORTED, "chain too long for the evaluator tile");
  // device-pointer calls run on the caller's stream (NULL = the legacy default stream); host-pointer calls on ours
  cudaStream_t s = (memory == 1) ? (cudaStream_t)stream : (stream ? (cudaStream_t)stream : robot->stream);
  if (memory == 1) lk.unlock();  // device-pointer calls only touch immutable robot state
  EvalParams P{};
  P.chain = (const double*)robot->chain_dev.p; P.n = n; P.chain_bytes = robot->chain_bytes;
  P.B = B; P.target_stride = shared_target ? 0 : 8;
  P.weighted = 0;
  for (int i = 0; i < 3; i++) {
    P.wl[i] = linear_weight ? linear_weight[i] : 1.0;
    P.wa[i] = angular_weight ? angular_weight[i] : 1.0;
    if (P.wl[i] != 1.0 || P.wa[i] != 1.0) P.weighted = 1;
  }
  if (ee_offset) for (int i = 0; i < 8; i++) P.ee_offset[i] = ee_offset[i];
  else pose8_identity(P.ee_offset);
  // persistent grid: one wave of resident blocks, each striding over tiles of 128 configurations
  static std::atomic<int> occ_cache[2][OPTIK_MAX_DOF + 1];  // blocks/SM per (layout, n) (all GPUs of a box are alike)
  int per_sm = occ_cache[cols][n].load(std::memory_order_relaxed);
  if (per_sm == 0) {
    CUDA_TRY(optik_eval_occupancy(n, cols, &per_sm));
    occ_cache[cols][n].store(per_sm, std::memory_order_relaxed);
  }
  if (per_sm < 1) return fail(OPTIK_ERR_UNSUPPORTED, "evaluator tile does not fit one SM");
  const uint64_t tb = (uint64_t)optik_eval_threads(n);
  const uint64_t nblk = (B + tb - 1) / tb, resident = (uint64_t)robot->sm_count * (uint64_t)per_sm;
  const int blocks = (int)(nblk < resident ? nblk : resident);
  if (memory == 1) {
    P.q = q; P.targets = targets; P.ee_out = ee_out; P.jac_out = jac_out; P.f_out = f_out; P.grad_out = grad_out;
    CUDA_TRY(optik_launch_eval(&P, blocks, s));
    return OPTIK_OK;
  }
  // host pointers: stage through stream-ordered device buffers
  const uint64_t nt = shared_target ? 1 : B;
  size_t bytes = 0;
  auto carve = [&](size_t b) { size_t off = bytes; bytes += (b + 255) & ~size_t(255); return off; };
  const size_t o_q = carve(B * n * 8), o_t = targets ? carve(nt * 64) : 0, o_ee = ee_out ? carve(B * 64) : 0,
               o_j = jac_out ? carve(B * 6 * n * 8) : 0, o_f = f_out ? carve(B * 8) : 0, o_g = grad_out ? carve(B * n * 8) : 0;
  StreamBuf buf;
  CUDA_TRY(buf.alloc(robot, bytes, s));
  char* d = buf.p;
  CUDA_TRY(cudaMemcpyAsync(d + o_q, q, B * n * 8, cudaMemcpyHostToDevice, s));
  if (targets) CUDA_TRY(cudaMemcpyAsync(d + o_t, targets, nt * 64, cudaMemcpyHostToDevice, s));
  P.q = (double*)(d + o_q); P.targets = targets ? (double*)(d + o_t) : nullptr;
  P.ee_out = ee_out ? (double*)(d + o_ee) : nullptr; P.jac_out = jac_out ? (double*)(d + o_j) : nullptr;
  P.f_out = f_out ? (double*)(d + o_f) : nullptr; P.grad_out = grad_out ? (double*)(d + o_g) : nullptr;
  CUDA_TRY(optik_launch_eval(&P, blocks, s));
  if (ee_out) CUDA_TRY(cudaMemcpyAsync(ee_out, d + o_ee, B * 64, cudaMemcpyDeviceToHost, s));
  if (jac_out) CUDA_TRY(cudaMemcpyAsync(jac_out, d + o_j, B * 6 * n * 8, cudaMemcpyDeviceToHost, s));
  if (f_out) CUDA_TRY(cudaMemcpyAsync(f_out, d + o_f, B * 8, cudaMemcpyDeviceToHost, s));
  if (grad_out) CUDA_TRY(cudaMemcpyAsync(grad_out, d + o_g, B * n * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return OPTIK_OK;
}

double* optik_robot_fk(const optik_robot* robot, const double* x) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  double ee[8];
  if (optik_gpu_eval_batch(robot, x, nullptr, 0, 1, nullptr, nullptr, nullptr, 0, ee, nullptr, nullptr, nullptr, nullptr))
    panic("fk: " + g_last_error);
  double* m = (double*)malloc(sizeof(double) * 16);
  colmajor4x4_from_pose8(ee, m);
  return m;
}
double* optik_robot_joint_jacobian(const optik_robot* robot, const double* x) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  for (const optik::Joint& J : robot->joints)
    if (J.type == optik::PRISMATIC) panic("not yet implemented: prismatic joints not yet supported");  // kinematics.rs:185
  double* jac = (double*)malloc(sizeof(double) * 6 * robot->n);
  if (optik_gpu_eval_batch(robot, x, nullptr, 0, 1, nullptr, nullptr, nullptr, 0, nullptr, jac, nullptr, nullptr, nullptr))
    panic("joint_jacobian: " + g_last_error);
  return jac;
}

// ---------------------------------------------------------------- batched ik
static int batch_common(const optik_robot* robot, const optik_solver_config* config, const optik_gpu_batch_opts* opts,
                        const double* targets, const double* x0, uint64_t T, double* q_out, double* cost_out,
                        int32_t* status_out, int32_t* evals_all, bool per_attempt, void* stream) {
  if (!robot || !targets || !x0 || !q_out || !cost_out || !status_out) return fail(OPTIK_ERR_INVALID, "null argument");
  if (!config_valid(config)) return fail(OPTIK_ERR_INVALID, "invalid solver config / solution_mode");
  if (opts && opts->struct_size != sizeof(optik_gpu_batch_opts)) return fail(OPTIK_ERR_INVALID, "opts.struct_size mismatch");
  if (T == 0) return OPTIK_OK;
  uint64_t R = opts ? opts->restarts : 0;
  if (R == 0) {
    // the reference's default budget (config.rs:52-65): no restart limit, max_time bounds the call (lib.rs:260-277)
    if (config->max_restarts == 0 || config->max_restarts > UNBOUNDED_RESTARTS) {
      if (!(config->max_time > 0.0) || per_attempt)
        return fail(OPTIK_ERR_INVALID, "no time or restart limit applied -- solver would run forever");
      R = UNBOUNDED_RESTARTS;
    } else R = config->max_restarts;
  }
  const uint64_t r_begin = opts ? opts->restart_begin : 0;
  const int memory = opts ? opts->memory : 0;
  const bool async = opts && (opts->flags & OPTIK_BATCH_ASYNC);
  if (async && (memory == 1 || !stream))
    return fail(OPTIK_ERR_INVALID, "OPTIK_BATCH_ASYNC is for host-memory calls on a caller-provided stream");
  {
    std::lock_guard<std::mutex> lk(robot->mu);
    if (int rc = robot->ensure_gpu()) return rc;
  }
  // from here on only immutable robot state is read: calls from several threads (or on several streams) overlap
  CUDA_TRY(cudaSetDevice(robot->device));
  const int n = robot->n;
  const unsigned long long max_ns = config->max_time > 0.0 ? (unsigned long long)(config->max_time * 1e9) : 0ull;
  const uint64_t NO = per_attempt ? R : T;  // number of output records
  if (memory == 1) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = solve_device(robot, config, opts, targets, x0, T, r_begin, R, q_out, cost_out, status_out,
                          opts ? opts->restart_out : nullptr, per_attempt ? evals_all : (opts ? opts->evals_out : nullptr),
                          opts ? opts->counters : nullptr, max_ns, per_attempt, s, opts ? opts->best_record_out : nullptr);
    if (rc) return rc;
    // the reference panics on a seed outside the limits (lib.rs:251-254); device seeds cannot be checked without a
    // sync, so they are clamped and the target's status carries OPTIK_STATUS_FLAG_SEED_CLAMPED
    // (per-attempt launches of the thread-per-seed kernel set the flag on restart 0's record themselves)
    if (r_begin == 0 && !(per_attempt && T == 1 && choose_tile(robot, opts ? opts->tile : 0, true) == 1))
      CUDA_TRY(optik_launch_flag_clamped((const double*)robot->chain_dev.p, n, x0, T, status_out, 1, s));
    return OPTIK_OK;
  }
  // blocking host calls reject a seed outside the limits like the reference (lib.rs:251-254); enqueue-only calls leave
  // the check to the device (clamp + OPTIK_STATUS_FLAG_SEED_CLAMPED): walking 1 Mi seeds on one host core costs more
  // than their H2D copy
  if (!async)
    if (int rc = check_seeds_host(robot, x0, T)) return rc;
  // host-memory calls without a stream of their own serialise on the robot's stream
  std::unique_lock<std::mutex> lk(robot->mu, std::defer_lock);
  cudaStream_t s = (cudaStream_t)stream;
  if (!s) { lk.lock(); s = robot->stream; }
  size_t bytes = 0;
  auto carve = [&](size_t b) { size_t off = bytes; bytes += (b + 255) & ~size_t(255); return off; };
  const bool want_rs = opts && opts->restart_out, want_cnt = opts && opts->counters;
  const bool want_best = per_attempt && opts && opts->best_record_out;
  const uint64_t NR = per_attempt ? 1 : T;  // restart_out entries
  int32_t* evals_host = per_attempt ? evals_all : (opts ? opts->evals_out : nullptr);
  const size_t o_t = carve(T * 64), o_x = carve(T * n * 8), o_q = carve(NO * n * 8), o_f = carve(NO * 8),
               o_s = carve(NO * 4), o_r = want_rs ? carve(NR * 8) : 0, o_e = evals_host ? carve(NO * 4) : 0,
               o_c = want_cnt ? carve(24) : 0, o_br = want_best ? carve((OPTIK_RECORD_HEAD + n) * 8) : 0;
  StreamBuf buf;
  CUDA_TRY(buf.alloc(robot, bytes, s));
  char* d = buf.p;
  CUDA_TRY(cudaMemcpyAsync(d + o_t, targets, T * 64, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d + o_x, x0, T * n * 8, cudaMemcpyHostToDevice, s));
  if (want_cnt) CUDA_TRY(cudaMemsetAsync(d + o_c, 0, 24, s));
  if (int rc = solve_device(robot, config, opts, (double*)(d + o_t), (double*)(d + o_x), T, r_begin, R, (double*)(d + o_q),
                            (double*)(d + o_f), (int32_t*)(d + o_s), want_rs ? (uint64_t*)(d + o_r) : nullptr,
                            evals_host ? (int32_t*)(d + o_e) : nullptr, want_cnt ? (uint64_t*)(d + o_c) : nullptr, max_ns,
                            per_attempt, s, want_best ? (double*)(d + o_br) : nullptr))
    return rc;
  if (async && r_begin == 0 && !(per_attempt && T == 1 && choose_tile(robot, opts ? opts->tile : 0, true) == 1))
    CUDA_TRY(optik_launch_flag_clamped((const double*)robot->chain_dev.p, n, (const double*)(d + o_x), T, (int*)(d + o_s), 1, s));
  CUDA_TRY(cudaMemcpyAsync(q_out, d + o_q, NO * n * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(cost_out, d + o_f, NO * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(status_out, d + o_s, NO * 4, cudaMemcpyDeviceToHost, s));
  if (want_rs) CUDA_TRY(cudaMemcpyAsync(opts->restart_out, d + o_r, NR * 8, cudaMemcpyDeviceToHost, s));
  if (want_best)
    CUDA_TRY(cudaMemcpyAsync(opts->best_record_out, d + o_br, (OPTIK_RECORD_HEAD + n) * 8, cudaMemcpyDeviceToHost, s));
  if (evals_host) CUDA_TRY(cudaMemcpyAsync(evals_host, d + o_e, NO * 4, cudaMemcpyDeviceToHost, s));
  uint64_t cnt[3] = {0, 0, 0};
  if (want_cnt) CUDA_TRY(cudaMemcpyAsync(async ? opts->counters : cnt, d + o_c, 24, cudaMemcpyDeviceToHost, s));
  if (async) return OPTIK_OK;  // the caller waits with optik_gpu_stream_sync(stream)
  CUDA_TRY(cudaStreamSynchronize(s));
  if (want_cnt) for (int i = 0; i < 3; i++) opts->counters[i] += cnt[i];
  return OPTIK_OK;
}

int optik_gpu_ik_batch(const optik_robot* robot, const optik_solver_config* config, const optik_gpu_batch_opts* opts,
                       const double* targets, const double* x0, uint64_t T, double* q_out, double* cost_out,
                       int32_t* status_out, void* stream) {
  return batch_common(robot, config, opts, targets, x0, T, q_out, cost_out, status_out, nullptr, false, stream);
}
int optik_gpu_ik_attempts(const optik_robot* robot, const optik_solver_config* config,
                          const optik_gpu_batch_opts* opts, const double* target, const double* x0, double* q_all,
                          double* f_all, int32_t* status_all, int32_t* evals_all, void* stream) {
  if (!evals_all) return fail(OPTIK_ERR_INVALID, "null argument");
  return batch_common(robot, config, opts, target, x0, 1, q_all, f_all, status_all, evals_all, true, stream);
}

}  // extern "C"

// ---------------------------------------------------------------- Robot::ik (single target), lib.rs:241-415
// One call = waves of IK_WAVE restarts, ONE kernel launch per wave and nothing else on the critical path: the kernel reads
// target and seed from mapped pinned host memory, its last block selects the winner (lib.rs:397-413), writes the record
// back into mapped memory and raises a flag the host polls.  Up to IK_SLOTS calls run concurrently on one Robot (the
// reference's ik() takes &self and may be called from many threads): each takes a slot with its own stream and scratch.
// Returns 1 and fills q_out/cost_out when a restart converged, 0 for "no solution" (== None), <0 = -(error code).
static int ik_slot_init(const optik_robot* robot, IkSlot& sl) {
  const int n = robot->n;
  const size_t W = IK_WAVE;
  const size_t bytes = 256 + W * (n * 8 + 8 + 8 + 8 + 4 + 4) + 256;
  CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc((void**)&sl.dev, bytes));
  CUDA_TRY(cudaMemset(sl.dev, 0, bytes));
  CUDA_TRY(cudaMemset(sl.dev + 16, 0xff, 8));  // found[0] = ~0
  CUDA_TRY(cudaHostAlloc((void**)&sl.host, 2048, cudaHostAllocMapped));
  memset(sl.host, 0, 2048);
  CUDA_TRY(cudaHostGetDevicePointer((void**)&sl.host_dev, sl.host, 0));
  CUDA_TRY(cudaDeviceSynchronize());  // the memsets ran on the default stream; the slot's stream does not wait for it
  sl.ready = true;
  return OPTIK_OK;
}

static int ik_single(const optik_robot* robot, const optik_solver_config* config, const double* tgt8, const double* x0,
                     const double* ee_offset, double* q_out, double* cost_out) {
  const int n = robot->n;
  if (int rc = check_seeds_host(robot, x0, 1)) return -rc;
  {
    std::lock_guard<std::mutex> lk(robot->mu);
    if (int rc = robot->ensure_gpu()) return -rc;
  }
  if (cudaSetDevice(robot->device) != cudaSuccess) return -fail(OPTIK_ERR_CUDA, "cudaSetDevice failed");
  // max_time (lib.rs:260-264) runs from here: the one-time device initialisation of a fresh Robot is not solve time
  const auto t_begin = std::chrono::steady_clock::now();
  // ---- take a slot (calls beyond IK_SLOTS wait for one)
  IkSlot* sl = nullptr;
  for (;;) {
    for (IkSlot& c : robot->ik_slots)
      if (!c.busy.exchange(true, std::memory_order_acquire)) { sl = &c; break; }
    if (sl) break;
    std::this_thread::yield();
  }
  struct Release { IkSlot* s; ~Release() { s->busy.store(false, std::memory_order_release); } } release{sl};
  if (!sl->ready)
    if (int rc = ik_slot_init(robot, *sl)) return -rc;
  const int tile = choose_tile(robot, 0, false);  // single target: the low-latency tile layout, one restart per tile
  const uint64_t max_restarts = config->max_restarts > 0 ? config->max_restarts : ~0ull;  // lib.rs:273-277
  const bool speed = config->solution_mode == OPTIK_MODE_SPEED;
  SolveParams P{};
  fill_common(robot, config, ee_offset, 0, P);
  volatile double* h = sl->host;
  for (int i = 0; i < 8; i++) h[i] = tgt8[i];
  for (int i = 0; i < n; i++) h[8 + i] = x0[i];
  P.targets = sl->host_dev; P.x0 = sl->host_dev + 8; P.T = 1;
  char* d = sl->dev;
  P.queue = (unsigned long long*)d;
  P.fused_done = (unsigned*)(d + 8);
  P.found = speed ? (unsigned long long*)(d + 16) : nullptr;
  const size_t W = IK_WAVE;
  P.cand_q = (double*)(d + 256);
  P.cand_f = P.cand_q + W * n;
  P.cand_score = P.cand_f + W;
  P.cand_restart = (unsigned long long*)(P.cand_score + W);
  P.cand_status = (int*)(P.cand_restart + W);
  P.cand_evals = P.cand_status + W;
  P.fused_record = sl->host_dev + 64;
  P.fused_flag = (unsigned long long*)(sl->host_dev + 128);
  P.fused_reset = 1;
  volatile unsigned long long* flag = (volatile unsigned long long*)(sl->host + 128);
  bool have_best = false;
  double best_score = 0;
  uint64_t done = 0;
  while (done < max_restarts) {
    // budget check between waves (lib.rs:260-264, 393-394); inside a wave the kernel checks the same deadline
    double remaining = 0;
    if (config->max_time > 0.0) {
      const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
      remaining = config->max_time - el;
      if (remaining <= 0) {
        if (done > 0) break;
        remaining = 1e-6;  // at least one wave runs (restart 0 is evaluated once before its deadline check)
      }
    }
    const uint64_t R = (max_restarts - done) < W ? (max_restarts - done) : W;
    P.r_begin = done; P.r_end = done + R; P.C = (uint32_t)R;
    P.max_ns = config->max_time > 0.0 ? (unsigned long long)(remaining * 1e9) + 1 : 0ull;
    P.fused_seq = ++sl->seq;
    const int tiles_per_block = 128 / tile;
    const int blocks = (int)((R + tiles_per_block - 1) / tiles_per_block);
    if (int e = optik_launch_solve(&P, tile, blocks, sl->stream))
      return -fail(OPTIK_ERR_CUDA, std::string("ik launch: ") + cudaGetErrorString((cudaError_t)e));
    // poll the completion flag in mapped memory (a stream synchronisation costs several microseconds more)
    const auto t_wait = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*flag != P.fused_seq) {
      if ((++spins & 0x3ff) == 0) {
        if (cudaStreamQuery(sl->stream) != cudaErrorNotReady && *flag != P.fused_seq) {  // finished without a flag: an error
          const cudaError_t e = cudaStreamSynchronize(sl->stream);
          if (*flag == P.fused_seq) break;
          return -fail(OPTIK_ERR_CUDA, std::string("ik: ") + cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e));
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wait).count() > 30.0)
          return -fail(OPTIK_ERR_CUDA, "ik: kernel did not complete within 30 s");
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    done += R;
    if (h[64] > 0.0) {  // a restart of this wave converged: record = [found, score, restart, cost, status, 0,0,0, q...]
      double score = 0;
      for (int i = 0; i < n; i++) score += (h[72 + i] - x0[i]) * (h[72 + i] - x0[i]);
      if (!have_best || score < best_score) {
        have_best = true; best_score = score;
        for (int i = 0; i < n; i++) q_out[i] = h[72 + i];
        *cost_out = h[67];
      }
      if (speed) break;  // the first wave with a success holds the lowest index
    }
  }
  return have_best ? 1 : 0;
}

extern "C" int optik_robot_ik_ex(const optik_robot* robot, const optik_solver_config* config, const double* target_pose8,
                                 const double* x0, const double* ee_offset_pose8, double* q_out, double* cost_out) {
  if (!robot || !config || !target_pose8 || !x0 || !q_out || !cost_out) return -fail(OPTIK_ERR_INVALID, "null argument");
  if (!config_valid(config)) return -fail(OPTIK_ERR_INVALID, "invalid solution mode");
  return ik_single(robot, config, target_pose8, x0, ee_offset_pose8, q_out, cost_out);
}

extern "C" double* optik_robot_ik(const optik_robot* robot, const optik_solver_config* config, const double* target,
                                  const double* x0) {
  if (!robot || !config || !target || !x0) panic("called `Option::unwrap()` on a `None` value (null argument)");
  if (!config_valid(config)) panic("invalid solution mode");
  double tgt8[8], cost = 0;
  pose8_from_colmajor4x4(target, tgt8);
  double* out = (double*)malloc(sizeof(double) * robot->n);
  const int rc = ik_single(robot, config, tgt8, x0, nullptr, out, &cost);
  if (rc == 1) return out;
  free(out);
  if (rc < 0) panic(g_last_error);  // e.g. "seed joint position outside of joint limits" (lib.rs:251-254)
  return nullptr;
}

// ---------------------------------------------------------------- diff_ik (lib.rs:101-239)
extern "C" int optik_gpu_diff_ik_batch(const optik_robot* robot, const double* x0, const double* V_WE, int shared_V,
                                       const double* v_max, int shared_vmax, uint64_t B, const double* ee_offset,
                                       int memory, double* alpha_out, double* v_out, int32_t* status_out, void* stream) {
  if (!robot || !x0 || !V_WE || !v_max || !alpha_out || !v_out || !status_out) return fail(OPTIK_ERR_INVALID, "null argument");
  const int n = robot->n;
  if (n != 6 && n != 7) return fail(OPTIK_ERR_UNSUPPORTED, "diff_ik supports num_positions 6 (the reference's case) and 7");
  if (B == 0) return OPTIK_OK;
  if (memory == 0) {
    const uint64_t nv = shared_vmax ? (uint64_t)n : B * n;
    for (uint64_t k = 0; k < nv; k++)
      if (!(v_max[k] > 0.0)) return fail(OPTIK_ERR_INVALID, "v_max entries must be > 0");
  }
  std::unique_lock<std::mutex> lk(robot->mu);
  if (int rc = robot->ensure_gpu()) return rc;
  CUDA_TRY(cudaSetDevice(robot->device));
  cudaStream_t s = (memory == 1) ? (cudaStream_t)stream : (stream ? (cudaStream_t)stream : robot->stream);
  if (memory == 1) lk.unlock();
  DiffIkParams P{};
  P.chain = (const double*)robot->chain_dev.p; P.n = n; P.chain_bytes = robot->chain_bytes;
  P.B = B; P.shared_V = shared_V ? 1 : 0; P.shared_vmax = shared_vmax ? 1 : 0;
  if (ee_offset) for (int i = 0; i < 8; i++) P.ee_offset[i] = ee_offset[i];
  else pose8_identity(P.ee_offset);
  const uint64_t nblk = (B + 127) / 128, cap = (uint64_t)robot->sm_count * 4;
  const int blocks = (int)(nblk < cap ? nblk : cap);
  if (memory == 1) {
    P.x0 = x0; P.V = V_WE; P.vmax = v_max; P.alpha_out = alpha_out; P.v_out = v_out; P.status_out = status_out;
    CUDA_TRY(optik_launch_diffik(&P, blocks, s));
    return OPTIK_OK;
  }
  size_t bytes = 0;
  auto carve = [&](size_t b) { size_t off = bytes; bytes += (b + 255) & ~size_t(255); return off; };
  const uint64_t nV = shared_V ? 6 : 6 * B, nm = shared_vmax ? (uint64_t)n : B * n;
  const size_t o_x = carve(B * n * 8), o_V = carve(nV * 8), o_m = carve(nm * 8), o_a = carve(B * 8), o_v = carve(B * n * 8),
               o_s = carve(B * 4);
  StreamBuf buf;
  CUDA_TRY(buf.alloc(robot, bytes, s));
  char* d = buf.p;
  CUDA_TRY(cudaMemcpyAsync(d + o_x, x0, B * n * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d + o_V, V_WE, nV * 8, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d + o_m, v_max, nm * 8, cudaMemcpyHostToDevice, s));
  P.x0 = (double*)(d + o_x); P.V = (double*)(d + o_V); P.vmax = (double*)(d + o_m);
  P.alpha_out = (double*)(d + o_a); P.v_out = (double*)(d + o_v); P.status_out = (int*)(d + o_s);
  CUDA_TRY(optik_launch_diffik(&P, blocks, s));
  CUDA_TRY(cudaMemcpyAsync(alpha_out, d + o_a, B * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(v_out, d + o_v, B * n * 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(status_out, d + o_s, B * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return OPTIK_OK;
}

extern "C" int optik_robot_diff_ik_ex(const optik_robot* robot, const double* x0, const double* V_WE, const double* v_max,
                                      const double* ee_offset_pose8, double* alpha_out, double* v_out) {
  if (!robot) return -fail(OPTIK_ERR_INVALID, "null robot");
  int32_t st = 0;
  if (int rc = optik_gpu_diff_ik_batch(robot, x0, V_WE, 1, v_max, 1, 1, ee_offset_pose8, 0, alpha_out, v_out, &st, nullptr))
    return -rc;
  return st == 1 ? 1 : 0;
}

// crates/optik-cpp/src/lib.rs:164-183: n doubles (caller frees) or NULL when there is no solution; alpha is dropped
extern "C" double* optik_robot_diff_ik(const optik_robot* robot, const double* x0, const double* V_WE, const double* v_max) {
  if (!robot) panic("called `Option::unwrap()` on a `None` value (null robot)");
  double alpha = 0.0;
  double* v = (double*)malloc(sizeof(double) * robot->n);
  const int rc = optik_robot_diff_ik_ex(robot, x0, V_WE, v_max, nullptr, &alpha, v);
  if (rc < 0) { free(v); panic("diff_ik: " + g_last_error); }
  if (rc == 0) { free(v); return nullptr; }
  return v;
}
