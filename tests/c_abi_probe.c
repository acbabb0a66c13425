/* Compiled by tests/test_host_cpu.py with `gcc -std=c99 -Wall -Werror`: proves include/optik_b200.h is valid C and that
 * a C host can drive the non-compute part of the ABI exactly as crates/optik-cpp/src/lib.cpp:5-31 declares it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "optik_b200.h"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  if (sizeof(optik_solver_config) != 96) return 3; /* CSolverConfig layout, crates/optik-cpp/src/lib.rs:10-20 */
  printf("sizeof(optik_gpu_batch_opts)=%u\n", (unsigned)sizeof(optik_gpu_batch_opts));
  optik_robot* r = optik_robot_from_urdf_file(argv[1], argv[2], argv[3]);
  if (!r) return 4;
  unsigned n = optik_robot_num_positions(r);
  double* lim = optik_robot_joint_limits(r); /* 2n doubles, caller frees (lib.cpp:72) */
  double* q = optik_robot_random_configuration(r);
  int inside = 1;
  for (unsigned i = 0; i < n; i++) inside &= (q[i] >= lim[i] && q[i] <= lim[n + i]);
  printf("n=%u joints=%u inside=%d lb0=%.4f ub0=%.4f\n", n, optik_robot_num_joints(r), inside, lim[0], lim[n]);
  optik_robot_set_parallelism(r, 4);
  optik_solver_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.solution_mode = OPTIK_MODE_SPEED;
  cfg.tol_f = 1e-6; cfg.tol_df = -1; cfg.tol_dx = -1;
  int ok = optik_status_is_success(&cfg, OPTIK_STATUS_STOPVAL) && !optik_status_is_success(&cfg, OPTIK_STATUS_FTOL);
  optik_robot* bad = optik_robot_try_from_urdf_str("<robot><link name='a'/></robot>", "a", "nope");
  printf("try_from_urdf_str(bad) = %s : %s\n", bad ? "non-null" : "NULL", optik_last_error());
  free(lim);
  free(q);
  optik_robot_free(r);
  return (inside && ok && !bad) ? 0 : 5;
}
