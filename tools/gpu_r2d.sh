set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -s 2 -c 1 -f -o gpurun_out/r02d_prof_t1_speed python tools/profile_t1.py speed > gpurun_out/prof_t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_t1 -s 2 -c 1 -f -o gpurun_out/r02d_prof_t1_quality python tools/profile_t1.py quality >> gpurun_out/prof_t1.log 2>&1
tail -3 gpurun_out/prof_t1.log
python -c "
import optik_b200 as ob
print('fp64 peak TFLOP/s', ob.load_library().optik_measure_fp64_peak(0, 2.0))
"
