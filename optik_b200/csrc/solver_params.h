// solver_params.h -- POD structs shared by the host launcher and the kernels.
#pragma once
#include <stdint.h>

#define OPTIK_MAX_DOF 32
#define OPTIK_CHAIN_STRIDE 16 /* doubles per joint in the flat chain */
/* slots of the cross-GPU exchange buffer (csrc/exchange_kernel.cu): a rank may keep NSLOT / 2 calls in flight */
#define OPTIK_EXCHANGE_NSLOT 32

// Attempt status (one restart).  1..3 map to NLopt's SuccessState as the
// reference classifies them (crates/optik/src/lib.rs:376-379).
enum {
  OPTIK_ST_NONE = 0,
  OPTIK_ST_STOPVAL = 1,  // f < tol_f
  OPTIK_ST_FTOL = 2,     // accepted step with |df| < tol_df
  OPTIK_ST_XTOL = 3,     // accepted step with max|dx| < tol_dx
  OPTIK_ST_ITERCAP = 4,  // evaluation cap reached
  OPTIK_ST_STUCK = 5,    // no descent step / stalled far from tol_f
  OPTIK_ST_NAN = 6,
  OPTIK_ST_SKIPPED = 7   // not run: timeout or Speed-mode early exit
};

// Device-side joint record: 16 doubles = 128 B (one TMA-friendly line per joint).
//  [0..2] origin xyz  [3] type (0 revolute, 1 prismatic, 2 fixed/padding)
//  [4..7] origin quaternion xyzw  [8..10] axis  [11] pad
//  [12] lower [13] upper [14] seed-sampling lower [15] seed-sampling upper
// The blob is n joints followed by the fixed tip pose8 {qx,qy,qz,qw,tx,ty,tz,0}.

// Levenberg-Marquardt constants of the in-warp solver (tuned on Panda/UR5/UR3e/snake).
#define OPTIK_LM_MAX_EVALS 24
#define OPTIK_LM_LAMBDA0 1e-1
#define OPTIK_LM_LAMBDA_DEC 0.3
#define OPTIK_LM_LAMBDA_INC 10.0
#define OPTIK_LM_LAMBDA_MIN 1e-9
#define OPTIK_LM_LAMBDA_MAX 1e6
#define OPTIK_LM_STALL_REL 1e-1
#define OPTIK_LM_STALL_COUNT 2

struct SolveParams {
  // problem
  const double* chain;  // device blob (see above)
  int n;                // articulated joints (<= 32)
  uint32_t chain_bytes; // n*128 + 64
  const double* targets;  // [T][8] pose8
  const double* x0;       // [T][n]
  unsigned long long T;
  unsigned long long r_begin, r_end;  // restart index range [r_begin, r_end); restart 0 = x0
  uint32_t C;                         // parallel chunks per target; chunk c runs r_begin+c, +C, ...
  int mode;                           // 1 = Quality, 2 = Speed (config.rs:5-8)
  // SolverConfig-derived (lib.rs:283-293, 345-347)
  double tol_f, tol_df_eff, tol_df_user, tol_dx;
  double wl[3], wa[3];
  int weighted;
  int has_prismatic;  // any prismatic joint in the chain (grid-uniform: lets the revolute-only path skip dead math)
  // LM constants
  int max_evals;
  double lambda0, lambda_dec, lambda_inc, lambda_min, lambda_max, stall_rel;
  int stall_count;
  // seeds
  uint32_t key[8];  // ChaCha8 key = seed_from_u64(42)
  double ee_offset[8];
  // Speed-mode early exit: per-target lowest converged restart index so far (init ~0ull), or null
  unsigned long long* found;
  unsigned long long max_ns;  // 0 = no deadline; else nanoseconds from kernel start
  unsigned long long* queue;  // job queue head (zeroed before launch): idle tiles/threads pull the next (target, chunk) job
  // candidate records, one per (target, chunk): [T*C]
  double* cand_q;                     // [T*C][n]
  double* cand_f;                     // objective value of the recorded attempt
  double* cand_score;                 // Quality: ||q - x0||^2 ; Speed: 0
  unsigned long long* cand_restart;   // restart index of the recorded attempt
  int* cand_status;                   // status of the recorded attempt
  int* cand_evals;                    // evaluations spent by this chunk (all its attempts)
  unsigned long long* counters;       // [0] attempts run, [1] evaluations, [2] converged attempts (optional)
  // restart seeds for indices [seed_begin, seed_begin + seed_count) written by seed_table_kernel: [count][n], already
  // clamped to the limits; indices outside the table are drawn in-kernel (thread-per-seed kernel only)
  const double* seed_tab;
  unsigned long long seed_begin, seed_count;
  // thread-per-seed kernel, sched = 1: dynamic Speed chains (see solve_t1_kernel.cu).  The per-target record goes to
  // cand_q/f/status/restart[t] directly (cand_evals[t] += evaluations, zeroed by the host; cand_score unused).
  int sched;
  unsigned pool_chunk;                // jobs a warp claims from the queue per atomic (sched 1, 2)
  unsigned dyn_k0;                    // T < resident lanes: restarts 0..dyn_k0 of every target are jobs of the queue (start at once)
  unsigned* dyn_next;                 // [T] next relative restart index to claim            (zeroed)
  unsigned long long* dyn_word;       // [T] high half: lowest converged relative restart index so far, low half: 1 while
                                      //     a writer holds the target's record                  (all ones)
  // thread-per-seed kernel, sched = 2: selection fused into the solve launch (last block done).  fused_record = the
  // packed candidate record [8 + n]; partials are per block; fused_done is zeroed by the host.
  double* fused_record;
  int* fused_part_has;
  double* fused_part_score;
  unsigned long long* fused_part_restart;
  unsigned* fused_done;
  // tile kernel (single-target launches): completion flag the host polls (mapped host memory) and re-arming of the
  // persistent control words (queue, fused_done, found[0]) by the last block
  unsigned long long* fused_flag;
  unsigned long long fused_seq;
  int fused_reset;
  // optional: the last block also stores the record into every peer's exchange buffer (csrc/exchange_kernel.cu layout)
  const uint64_t* push_peers;         // device array of push_world base addresses, or null
  int push_rank, push_world;
  unsigned long long push_seq;
};

namespace optik { struct SelKey; }
struct SelectParams {
  unsigned long long T;
  uint32_t C;
  int n;
  const double* cand_q;
  const double* cand_f;
  const double* cand_score;
  const unsigned long long* cand_restart;
  const int* cand_status;
  const int* cand_evals;
  double tol_f, tol_df_user, tol_dx;
  double* q_out;                     // [T][n]
  double* f_out;                     // [T]
  unsigned long long* restart_out;   // [T]
  int* status_out;                   // [T]
  int* evals_out;                    // [T] total evaluations spent on the target
  double* record_out;                // [T][8+n] packed candidate record (see optik_b200.h), may be null
  int mode;                          // 1 Quality, 2 Speed (record score)
  // two-level reduction for targets with very many chunks (set by optik_launch_select)
  int final_pass;
  unsigned partials;
  const optik::SelKey* partial;
  optik::SelKey* partial_out;
};

struct EvalParams {
  const double* chain;
  int n;
  uint32_t chain_bytes;
  const double* q;        // [B][n]
  const double* targets;  // [B][8] or one shared pose8 (target_stride = 0)
  unsigned long long B;
  uint32_t target_stride;  // 8 or 0
  double wl[3], wa[3];
  int weighted;
  double ee_offset[8];
  double* ee_out;    // [B][8] pose8
  double* jac_out;   // [B][6n] column-major 6 x n body Jacobian (kinematics.rs:166-196), may be null
  double* f_out;     // [B] objective (objective.rs:40-57), may be null
  double* grad_out;  // [B][n] gradient (objective.rs:60-110), may be null
};

struct DiffIkParams {
  const double* chain;
  int n;                 // 6 or 7
  uint32_t chain_bytes;
  const double* x0;      // [B][n]
  const double* V;       // [B][6] world-frame twist [linear; angular], or one shared twist
  const double* vmax;    // [B][n] or one shared vector
  int shared_V, shared_vmax;
  unsigned long long B;
  double ee_offset[8];
  double* alpha_out;     // [B]
  double* v_out;         // [B][n]
  int* status_out;       // [B] 1 = solved, 0 = no solution (rank-deficient Jacobian)
};
