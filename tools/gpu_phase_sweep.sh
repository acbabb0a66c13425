for cfg in "2 3 2 0" "2 3 2 1" "3 3 2 1" "1 2 2 1" "2 2 1 1" "3 4 2 0" "4 4 1 1"; do
set -- $cfg
echo "== K0=$1 K=$2 FILL=$3 ONESHOT=$4"
OPTIK_EXP=phased OPTIK_PHASE_K0=$1 OPTIK_PHASE_K=$2 OPTIK_PHASE_FILL=$3 OPTIK_PHASE_ONESHOT=$4 python tools/exp_speed_batch.py 2>&1 | tail -4 | cut -c1-110
done
