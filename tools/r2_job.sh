# Round-2 measurement job (run under gpurun from the repo root): tests, bench, parity statistics,
# launch list, ncu captures.  usage: bash tools/r2_job.sh <tag> [steps...]; steps default to all.
TAG=${1:-r2}; shift
STEPS=${@:-tests bench c3 stats launches caps}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $OUT/tests.log; fi
if has bench; then timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; fi
if has c3; then timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --cpu-sample 256 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; fi
if has c4; then timeout 300 python bench.py --workload c4 --steps 2 --warmup 3 --cpu-sample 128 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; fi
if has c5; then timeout 200 python bench.py --workload c5 --steps 3 --warmup 3 --cpu-sample 256 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; fi
if has stats; then timeout 400 python tools/gpu_parity_stats.py c2:16384:2048 c3:8192:1024 c4:2048:128 > $OUT/parity_stats.log 2>&1; fi
if has mpc; then timeout 200 python tools/mpc_latency.py 1024 16 > $OUT/mpc_1024.json 2>&1; timeout 200 python tools/mpc_latency.py 1024 16 10 > $OUT/mpc_1024_cap10.json 2>&1; timeout 200 python tools/mpc_latency.py 1 16 > $OUT/mpc_1.json 2>&1; fi
if has launches; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv python tools/gpu_one_solve.py phased 16384 > $OUT/one.log 2>&1
  python tools/launch_summary.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
fi
cap() { # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/$1 -f python tools/gpu_one_solve.py phased 16384 > $OUT/$1.log 2>&1
  python tools/ncu_summarize.py $OUT/$1.ncu-rep $OUT/ncu_$1_summary.txt > /dev/null 2>&1
  rm -f $OUT/$1.ncu-rep
}
if has caps; then
  cap k_backward_mat_phased k_backward_mat 24
  cap k_ls_deep k_ls_deep 60
  cap k_ls_wide k_ls_wide 24
  cap k_update_expansions k_update_expansions 24
  cap k_outer_rollout k_outer_rollout 9
fi
if has bpalone; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_backward_mat -s 4 -c 1 -o $OUT/bp_insolve -f python tools/gpu_debug.py ncu_bp_insolve > $OUT/bp_insolve.log 2>&1
  python tools/ncu_summarize.py $OUT/bp_insolve.ncu-rep $OUT/ncu_k_backward_mat_insolve_fullbatch_summary.txt > /dev/null 2>&1
  rm -f $OUT/bp_insolve.ncu-rep
fi
if has c5cap; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_solve_large_mma -s 1 -c 1 -o $OUT/c5_mma -f python tools/gpu_c5_one.py 4096 > $OUT/c5_mma.log 2>&1
  python tools/ncu_summarize.py $OUT/c5_mma.ncu-rep $OUT/ncu_k_solve_large_mma_summary.txt > /dev/null 2>&1
  rm -f $OUT/c5_mma.ncu-rep
fi
if has c3cap; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_backward_coop -s 1 -c 1 -o $OUT/c3_coop -f python tools/gpu_one_solve.py phased 8192 0 c3 > $OUT/c3_coop.log 2>&1
  python tools/ncu_summarize.py $OUT/c3_coop.ncu-rep $OUT/ncu_k_backward_coop_c3_summary.txt > /dev/null 2>&1
  rm -f $OUT/c3_coop.ncu-rep
  timeout 200 python bench.py --workload c3 --batch 65536 --steps 3 --warmup 3 --no-cpu --no-extras > $OUT/bench_c3_b65536.json 2> $OUT/bench_c3_b65536.err
fi
ls -la $OUT
cat $OUT/tests.log 2>/dev/null | tail -8
head -c 1500 $OUT/bench_c2.json 2>/dev/null
cat $OUT/launches_summary.txt 2>/dev/null | head -30
