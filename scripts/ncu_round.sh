#!/bin/bash
# ncu captures of one round: scripts/ncu_round.sh r02   (run on the GPU box through gpurun; outputs in gpurun_out/)
# One `--set full` capture (one launch of the kernel under test) per case, each after the same command has run
# once WITHOUT ncu; plus the launch list of the default bench.
RND=${1:-r02}
W=/tmp/ncu_${RND}          # the .ncu-rep files (~40 MB each) stay on the box; only the summaries go to gpurun_out/
mkdir -p gpurun_out $W
# CASES="gemv fpe8_loguniform" restricts the captures to the cases whose name contains one of the words; NOLAUNCH=1 skips
# the launch list
cap() {  # name, kernel regex, args of scripts/prof_case.py
  local name=$1 regex=$2; shift 2
  if [ -n "$CASES" ]; then local hit=0; for c in $CASES; do case "$name" in *$c*) hit=1;; esac; done; [ $hit = 1 ] || return; fi
  timeout 300 python scripts/prof_case.py "$@" > $W/prof_${RND}_${name}.plain.log 2>&1 || { echo "$name: plain run failed"; return; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" --launch-skip 2 --launch-count 1 -f \
      -o $W/prof_${RND}_${name} python scripts/prof_case.py "$@" > $W/prof_${RND}_${name}.ncu.log 2>&1
  echo "$name: $(tail -1 $W/prof_${RND}_${name}.plain.log)"
}
cap exsum_fpe3_loguniform_2p30        exblas_reduce_kernel  exsum loguniform 3 0 30
cap exsum_fpe8_loguniform_2p30        exblas_reduce_kernel  exsum loguniform 8 0 30
cap exsum_fpe3_loguniform_signed_2p30 exblas_reduce_kernel  exsum loguniform_signed 3 0 30
cap exsum_fpe4_naive_2p30             exblas_reduce_kernel  exsum naive 4 0 30
cap exsum_fpe8_naive_2p30             exblas_reduce_kernel  exsum naive 8 0 30
cap exdot_fpe0_cancel_2p30            exblas_reduce0_kernel exdot cancel 0 0 30
cap exdot_fpe3_cancel_2p30            exblas_reduce_kernel  exdot cancel 3 0 30
cap exdot_fpe8ee_cancel_2p30          exblas_reduce_kernel  exdot cancel 8 1 30
cap exgemv_n_narrow_32768             exgemv_n_win_kernel   gemvN narrow 0 0 30
cap exgemv_t_narrow_32768             exgemv_t_win_kernel   gemvT narrow 0 0 30
# launch list of the default bench (every kernel of the process with its device time)
if [ -z "$NOLAUNCH" ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/launch_bench_${RND}.json 2> gpurun_out/launch_bench_${RND}.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"exblas|exgemv|mb_" -c 400 --csv --log-file $W/launches_${RND}.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/launch_ncu_${RND}.log 2>&1
fi
python scripts/make_profiles.py ${RND} $W gpurun_out/profiles_${RND}
cat $W/*.plain.log > gpurun_out/profiles_${RND}/plain_runs_${RND}.txt
ls -la $W/*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
