/*
 * optik_b200.h -- C ABI of liboptik_b200.so, the B200-native drop-in for OptIK's
 * Robot::ik() hot path.
 *
 * Part 1 are EXACTLY the symbols the reference's own FFI layer exports and its
 * C++ wrapper binds (kylc/optik @ 355e463):
 *     Rust side   crates/optik-cpp/src/lib.rs:26-183   (#[no_mangle] extern "C")
 *     C++ side    crates/optik-cpp/src/lib.cpp:5-31    (extern "C" declarations)
 * with the reference's ownership contract: every returned double* is a
 * malloc()-compatible buffer the CALLER releases with free()
 * (lib.cpp:72,87,100,115,129); optik_robot* is opaque and released only by
 * optik_robot_free().  Where the reference panics across the FFI boundary
 * (null robot, bad URDF, seed outside the joint limits, lib.rs:251-254) these
 * functions print the same message to stderr and abort(), which is what a Rust
 * panic in an extern "C" fn does.
 *
 * Part 2 are additive, batched entry points (plain pointers and sizes, no C++
 * or torch types) through which a host -- the reference's Rust crate via an
 * `extern "C"` block, see INTEGRATION.md -- drives the sm_100a kernels
 * directly.  They return 0 on success or a nonzero code and never abort;
 * optik_last_error() describes the last failure on the calling thread.
 *
 * There is no CPU fallback: every numerical entry point runs on the GPU and
 * fails loudly (error code / NULL + message) when no CUDA device is usable.
 */
#ifndef OPTIK_B200_H
#define OPTIK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* == optik::detail::robot (crates/optik-cpp/include/optik.hpp:9-10) */
typedef struct optik_robot optik_robot;

/* == SolutionMode, crates/optik/src/config.rs:3-8 (repr(C)) */
enum { OPTIK_MODE_QUALITY = 1, OPTIK_MODE_SPEED = 2 };

/* == CSolverConfig, crates/optik-cpp/src/lib.rs:10-20 == optik::SolverConfig, include/optik.hpp:18-27.
 * 96 bytes on LP64.  max_restarts == 0 means "no restart limit" (lib.rs:273-277). */
typedef struct optik_solver_config {
  int32_t solution_mode;
  double max_time;
  unsigned long max_restarts;
  double tol_f;
  double tol_df;
  double tol_dx;
  double linear_weight[3];
  double angular_weight[3];
} optik_solver_config;

/* ------------------------------------------------------------------------
 * Part 1 -- the reference's FFI surface
 * ------------------------------------------------------------------------ */

/* crates/optik-cpp/src/lib.rs:26-37.  Aborts on unreadable/invalid URDF or unknown links. */
optik_robot* optik_robot_from_urdf_file(const char* path, const char* base_link, const char* ee_link);
/* crates/optik-cpp/src/lib.rs:39-50 */
optik_robot* optik_robot_from_urdf_str(const char* urdf, const char* base_link, const char* ee_link);
/* crates/optik-cpp/src/lib.rs:52-57 */
void optik_robot_free(optik_robot* robot);
/* crates/optik-cpp/src/lib.rs:59-65.  rayon thread count in the reference.  Here the restart fan-out is the CUDA grid:
 * the value is stored and otherwise has no effect (results never depend on it -- the per-target answer is always the
 * one the reference gives with set_parallelism(1), see optik_gpu_ik_batch). */
void optik_robot_set_parallelism(optik_robot* robot, unsigned int n);
/* crates/optik-cpp/src/lib.rs:67-73 */
unsigned int optik_robot_num_positions(const optik_robot* robot);
/* crates/optik-cpp/src/lib.rs:75-88.  2n doubles: [lb_0..lb_{n-1}, ub_0..ub_{n-1}] */
double* optik_robot_joint_limits(const optik_robot* robot);
/* crates/optik-cpp/src/lib.rs:118-125.  n doubles, uniform in the limits, non-deterministic RNG */
double* optik_robot_random_configuration(const optik_robot* robot);
/* crates/optik-cpp/src/lib.rs:90-104.  6*n doubles, column-major 6 x n body-frame Jacobian, identity ee_offset */
double* optik_robot_joint_jacobian(const optik_robot* robot, const double* x);
/* crates/optik-cpp/src/lib.rs:106-116.  16 doubles, column-major 4x4 homogeneous end-effector pose */
double* optik_robot_fk(const optik_robot* robot, const double* x);
/* crates/optik-cpp/src/lib.rs:127-162.  target: 16 doubles column-major 4x4; x0: n doubles.
 * Returns n doubles, or NULL when no restart converged within max_time / max_restarts.
 * Aborts with "seed joint position outside of joint limits" like lib.rs:251-254. */
double* optik_robot_ik(const optik_robot* robot, const optik_solver_config* config, const double* target,
                       const double* x0);
/* crates/optik-cpp/src/lib.rs:164-183.  n doubles (joint velocities; alpha is dropped like the reference's wrapper
 * does) or NULL when there is no solution.  Runs the batched diff_ik kernel on one configuration: the reference's LP
 * (lib.rs:123-239) solved exactly instead of by Clarabel's interior-point iteration; num_positions 6 or 7. */
double* optik_robot_diff_ik(const optik_robot* robot, const double* x0, const double* V_WE, const double* v_max);

/* ------------------------------------------------------------------------
 * Part 2 -- additive batched / introspection entry points
 * ------------------------------------------------------------------------ */

/* Poses in the batched API are "pose8": {qx,qy,qz,qw, tx,ty,tz, 0} = nalgebra's Isometry3<f64> coordinates
 * (rotation i,j,k,w then translation) padded to 64 B so that a pose is four 16-byte vector loads. */

enum {
  OPTIK_OK = 0,
  OPTIK_ERR_INVALID = 1,     /* bad argument */
  OPTIK_ERR_CUDA = 2,        /* CUDA runtime error (no device, launch failure ...) */
  OPTIK_ERR_SEED_LIMITS = 3, /* an x0 lies outside the joint limits (lib.rs:251-254) */
  OPTIK_ERR_UNSUPPORTED = 4  /* chain not supported by the kernel (n > 32, fixed joint mid-chain) */
};

/* per-target / per-attempt status; 1..3 = NLopt SuccessState as classified at lib.rs:376-379 */
enum {
  OPTIK_STATUS_NONE = 0,
  OPTIK_STATUS_STOPVAL = 1,
  OPTIK_STATUS_FTOL = 2,
  OPTIK_STATUS_XTOL = 3,
  OPTIK_STATUS_ITERCAP = 4,
  OPTIK_STATUS_STUCK = 5,
  OPTIK_STATUS_NAN = 6,
  OPTIK_STATUS_SKIPPED = 7
};
/* status words are `code | flags`.  Device-memory batched calls cannot reject a seed outside the joint limits without
 * a host sync (the reference panics, lib.rs:251-254; host-memory calls return OPTIK_ERR_SEED_LIMITS): the seed is
 * clamped into the limits and the target's status carries this flag. */
#define OPTIK_STATUS_CODE_MASK 0xff
#define OPTIK_STATUS_FLAG_SEED_CLAMPED 0x100

const char* optik_last_error(void);
/* Non-aborting constructors (NULL + optik_last_error() on failure). */
optik_robot* optik_robot_try_from_urdf_str(const char* urdf, const char* base_link, const char* ee_link);
/* Robot::new(KinematicChain) analogue (crates/optik/src/lib.rs:42-47): flat chain, 16 doubles per joint:
 * [0..2] origin xyz [3] type (0 revolute,1 prismatic,2 fixed) [4..7] origin quat xyzw [8..10] axis [12] lower [13] upper */
optik_robot* optik_robot_from_chain(const double* chain, unsigned int njoints);
/* Option: fold fixed joints in URDF order (earlier*later) instead of the reference's order
 * (kinematics.rs:70,77).  Must be called before loading; default 0 = reference order. */
void optik_set_urdf_correct_fold(int on);
unsigned int optik_robot_num_joints(const optik_robot* robot); /* chain entries incl. a fixed tip joint */
int optik_robot_chain(const optik_robot* robot, double* out /* num_joints*16 */);
/* CUDA device the robot's kernels run on (default: device 0). */
int optik_robot_set_device(optik_robot* robot, int device);
/* 1 if a satisfying status under `config` (lib.rs:376-379) */
int optik_status_is_success(const optik_solver_config* config, int status);

typedef struct optik_gpu_batch_opts {
  uint32_t struct_size;    /* sizeof(optik_gpu_batch_opts) */
  uint32_t restarts;       /* restart attempts per target (restart 0 = x0, i>=1 = ChaCha8 stream i).
                              0 = config->max_restarts, which must then be finite and > 0 */
  uint64_t restart_begin;  /* first restart index; [restart_begin, restart_begin+restarts) is run */
  uint32_t chunks;         /* parallel chunks per target (each runs its restarts in index order); 0 = auto */
  uint32_t tile;           /* lanes per restart seed: 1 = thread-per-seed kernel (batch throughput layout),
                              8, 16, 32 (= one warp per seed) = tile kernel; 0 = auto (1 for batched calls) */
  uint32_t max_evals;      /* objective evaluations per attempt; 0 = default (24) */
  uint32_t blocks;         /* grid size; 0 = auto (one resident wave: a multiple of the SM count).  Asynchronous callers that keep
                              several calls in flight give each a fraction of the wave (e.g. one block per SM), so that the
                              calls are co-resident and one call's straggler tail runs under the others' bulk */
  int32_t memory;          /* 0: all data pointers are host memory; 1: device memory (async on `stream`) */
  const double* ee_offset; /* pose8 (host memory) or NULL = identity (lib.rs:245) */
  uint64_t* restart_out;   /* [T] winning restart index, optional (ignored by optik_gpu_ik_attempts: see best_record_out[2]) */
  int32_t* evals_out;      /* [T] objective evaluations spent on the target, optional */
  uint64_t* counters;      /* [3] += {attempts run, evaluations, converged attempts}, optional */
  /* optik_gpu_ik_attempts only: also run the selection pass (lib.rs:397-413) over the records and write ONE
   * candidate record of OPTIK_RECORD_HEAD + n doubles:
   *   [0] found (1.0 / 0.0)  [1] score (Quality: ||q-x0||^2, Speed: restart index)  [2] restart index
   *   [3] cost f(q)  [4] status  [5..7] 0  [8..8+n) q
   * This is the record ranks exchange for a cross-GPU best-pick (optik_gpu_select_records). */
  double* best_record_out;
  uint32_t flags;          /* OPTIK_BATCH_* bits */
  uint32_t variant;        /* thread-per-seed kernel: 0 = default, 1 = trial columns in local memory (3 blocks/SM),
                              2 = in shared memory (2 blocks/SM); same results bit for bit */
  /* optik_gpu_ik_attempts with best_record_out, device memory, thread-per-seed kernel: the launch that writes the
   * candidate record also stores it into every peer's exchange buffer (see optik_gpu_exchange_push: same arguments,
   * no extra launch).  push_peers = NULL: off. */
  const uint64_t* push_peers;
  uint32_t push_rank, push_world;
  uint64_t push_seq;
} optik_gpu_batch_opts;

/* opts->flags.  OPTIK_BATCH_ASYNC (host-memory calls only, `stream` must be non-NULL): enqueue the H2D copies, the
 * kernels and the D2H copies on `stream` and return WITHOUT waiting; outputs are valid after
 * optik_gpu_stream_sync(stream).  Inputs and outputs must stay alive until then and should be pinned
 * (optik_host_alloc), otherwise the copies are staged synchronously.  Two streams with two sets of buffers overlap
 * one call's transfers with the next call's kernels.  opts->counters is then SET (not incremented).  Seeds are
 * validated on the device like device-memory calls (clamped + OPTIK_STATUS_FLAG_SEED_CLAMPED), not on the host. */
#define OPTIK_BATCH_ASYNC 1u
/* Speed-mode batches of the thread-per-seed kernel normally run as DYNAMIC CHAINS in one launch: a lane that takes a
 * target claims its restarts one by one on the device; once every target has been taken, idle lanes claim further
 * restarts of the still-unsolved targets in parallel.  A claimed restart always runs unless a lower index of the same
 * target has converged, so the per-target result is the lowest-index converged restart (lib.rs:409-412) whatever the
 * timing.  OPTIK_BATCH_STATIC forces the static (target, chunk) schedule instead (same results; evals_out is then
 * reproducible, with dynamic chains it includes speculative attempts). */
#define OPTIK_BATCH_STATIC 8u

#define OPTIK_RECORD_HEAD 8

/* Robot::ik() over T independent (target, x0) pairs in one launch.
 *   targets [T][8] pose8, x0 [T][n]  ->  q_out [T][n], cost_out [T] (objective value), status_out [T]
 * Per target the SELECTION follows the reference with set_parallelism(1): Speed = the lowest-index converged restart
 * (lib.rs:409-412), Quality = arg-min ||q - x0|| over converged restarts (lib.rs:398-407).  The numerical solution
 * of a restart is the in-kernel Levenberg-Marquardt iterate, not NLopt SLSQP's (replaced by design, DESIGN.md
 * section 3): it satisfies the reference's success predicate f(q) < tol_f inside the limits, it is not bit-equal to
 * what the reference would return.  status_out[t] is the winning attempt's status, or a failure status when none
 * converged (== None).
 * Budgets as in the reference (lib.rs:260-277): config->max_time > 0 bounds the whole launch (device-side deadline,
 * checked per evaluation like lib.rs:308); restarts = opts->restarts, else config->max_restarts, and with
 * max_restarts == 0 (the reference's default, "no limit") restarts are drawn until max_time expires -- max_time == 0
 * together with no restart limit is refused like crates/optik-py/src/lib.rs:45-47. */
int optik_gpu_ik_batch(const optik_robot* robot, const optik_solver_config* config, const optik_gpu_batch_opts* opts,
                       const double* targets, const double* x0, uint64_t T, double* q_out, double* cost_out,
                       int32_t* status_out, void* stream);

/* Per-restart records for ONE target: restarts [restart_begin, restart_begin+restarts) each run to completion.
 *   q_all [R][n], f_all [R], status_all [R], evals_all [R]  (memory space per opts->memory) */
int optik_gpu_ik_attempts(const optik_robot* robot, const optik_solver_config* config,
                          const optik_gpu_batch_opts* opts, const double* target, const double* x0, double* q_all,
                          double* f_all, int32_t* status_all, int32_t* evals_all, void* stream);

/* Batched evaluator = Robot::fk + Robot::joint_jacobian + objective + objective_grad
 * (lib.rs:93-99, objective.rs:40-110) over B configurations.
 *   q [B][n]; targets [B][8] (or one pose8 when shared_target != 0; may be NULL if f_out and grad_out are NULL)
 *   ee_out [B][8] pose8, jac_out [B][6n] column-major 6 x n, f_out [B], grad_out [B][n]; any output may be NULL */
int optik_gpu_eval_batch(const optik_robot* robot, const double* q, const double* targets, int shared_target,
                         uint64_t B, const double* linear_weight, const double* angular_weight,
                         const double* ee_offset, int memory, double* ee_out, double* jac_out, double* f_out,
                         double* grad_out, void* stream);

/* Batched Robot::diff_ik (crates/optik/src/lib.rs:101-239): for each configuration the LP
 *   max alpha  s.t.  J_W(x0) v = alpha V_WE,  |v_i| <= v_max_i,  0 <= alpha <= 1
 * solved exactly (closed form instead of the reference's interior-point solve; see csrc/diffik_kernel.cu).
 *   x0 [B][n]; V_WE [B][6] world-frame twist [linear; angular] (or one twist when shared_V != 0);
 *   v_max [B][n] (or one vector when shared_vmax != 0), every entry > 0
 *   alpha_out [B], v_out [B][n], status_out [B]: 1 = solved.  The LP is always feasible and bounded (the reference
 *   returns Some((alpha, v)), lib.rs:231-238): at a singular configuration a twist outside the Jacobian's range gives
 *   alpha = 0, v = 0, one inside it the basic solution scaled into the velocity box; 0 is kept for "no solution"
 *   (not produced by the current kernel)
 * num_positions must be 6 (the reference's case) or 7; otherwise OPTIK_ERR_UNSUPPORTED.  memory: 0 host, 1 device. */
int optik_gpu_diff_ik_batch(const optik_robot* robot, const double* x0, const double* V_WE, int shared_V,
                            const double* v_max, int shared_vmax, uint64_t B, const double* ee_offset, int memory,
                            double* alpha_out, double* v_out, int32_t* status_out, void* stream);
/* Robot::diff_ik with the arguments the C wrapper drops (alpha, ee_offset).  1 = solved, 0 = none, <0 = -(error). */
int optik_robot_diff_ik_ex(const optik_robot* robot, const double* x0, const double* V_WE, const double* v_max,
                           const double* ee_offset_pose8, double* alpha_out, double* v_out);

/* Robot::ik with the arguments the C wrapper drops (crates/optik/src/lib.rs:241-247): ee_offset and the returned cost.
 * target_pose8 / ee_offset_pose8 are pose8 (ee_offset may be NULL = identity).  Same max_time / max_restarts
 * semantics as optik_robot_ik.  Returns 1 = solution written to q_out[n], *cost_out; 0 = no solution; <0 = -(error). */
int optik_robot_ik_ex(const optik_robot* robot, const optik_solver_config* config, const double* target_pose8,
                      const double* x0, const double* ee_offset_pose8, double* q_out, double* cost_out);

/* The restart seeds the solver uses (lib.rs:86-91, 360-370): seeds_out[i][j] for restart indices
 * [restart_begin, restart_begin + count), restart_begin >= 1 (restart 0 is the caller's x0), written by the same
 * kernel that fills the solve kernels' seed table.  memory: 0 host, 1 device (asynchronous on `stream`). */
int optik_gpu_restart_seeds(const optik_robot* robot, uint64_t restart_begin, uint64_t count, int memory, double* seeds_out,
                            void* stream);
/* Known-answer hook for the seed generator: the 16 output words of the ChaCha8 block (block counter 0) of stream
 * `stream_id` under `key8`, computed by the device function the kernels use.  Host pointers, blocking. */
int optik_gpu_chacha8_block(const optik_robot* robot, const uint32_t* key8, uint64_t stream_id, uint32_t* words16_out);

/* Cross-GPU best-pick: the reference's selection rule (lib.rs:397-413) over `count` candidate records of
 * OPTIK_RECORD_HEAD + n doubles (e.g. the output of an all-gather): converged first, then lowest score, then lowest
 * restart index.  Device pointers, asynchronous on `stream`. */
int optik_gpu_select_records(const optik_robot* robot, const double* records, uint32_t count, double* best_record_out,
                             void* stream);

/* The same best-pick as DIRECT peer-to-peer stores over NVLink (no collective library call; csrc/exchange_kernel.cu).
 * Every rank owns one zero-initialised SYMMETRIC buffer of optik_gpu_exchange_bytes(robot, world) bytes that the host
 * has mapped into every peer (cudaIpc / torch symmetric memory); `peer_buffers_dev` is a device array of the `world`
 * base addresses (this rank's own included).  Call number seq = 1, 2, ... (the same on every rank):
 *   push    stores `record` into row `rank` of every peer's buffer and raises that row's flag to seq
 *   select  waits (bounded, 2 s) for all `world` rows of call seq in the local buffer and applies the selection rule;
 *           best_record_out[0] = -1 if a peer never delivered
 * Both are asynchronous on `stream`; at most 16 calls may be in flight per rank (32 slots). */
uint64_t optik_gpu_exchange_bytes(const optik_robot* robot, uint32_t world);
int optik_gpu_exchange_push(const optik_robot* robot, const double* record, const uint64_t* peer_buffers_dev, uint32_t rank,
                            uint32_t world, uint64_t seq, void* stream);
int optik_gpu_exchange_select(const optik_robot* robot, const double* local_buffer, uint32_t world, uint64_t seq,
                              double* best_record_out, void* stream);

/* Streams for hosts without CUDA bindings of their own (Rust, ctypes): create on the robot's device, wait, destroy.
 * Any cudaStream_t the caller already owns may be passed to the batched calls instead. */
int optik_gpu_stream_create(const optik_robot* robot, void** stream_out);
int optik_gpu_stream_sync(void* stream);
void optik_gpu_stream_destroy(void* stream);

/* Measurement helper (bench.py `roofline_solve`): fp64 FMA throughput of `device` in TFLOP/s, sustained over about
 * `seconds` of launches of a register-only DFMA kernel (8 independent chains per thread); < 0 on error. */
double optik_measure_fp64_peak(int device, double seconds);
/* Measurement only: GB/s of a pure streaming kernel that reads rd_units and writes wr_units 16-byte units per item, fully
 * coalesced, over `items` items (best of `reps` launches): what HBM delivers at the evaluator's read/write mix. */
double optik_measure_hbm_mix(int device, unsigned long long items, int rd_units, int wr_units, int reps);

/* Pinned host allocations for fast, truly asynchronous host<->device copies. */
void* optik_host_alloc(uint64_t bytes);
void optik_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* OPTIK_B200_H */
