# Round-end measurement job (run under gpurun from the repo root): tests, bench, launch list, ncu.
set -x
mkdir -p gpurun_out/final
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/final/tests.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/final/bench_c2.json 2> gpurun_out/final/bench_c2.err
for p in 8; do ALTRO_B200_OUTER_PERIOD=$p python tools/gpu_variants.py default | tail -1; done > gpurun_out/final/period8.log 2>&1
timeout 120 python bench.py --workload c3 --steps 3 --warmup 3 --cpu-sample 256 > gpurun_out/final/bench_c3.json 2> gpurun_out/final/bench_c3.err
timeout 150 python bench.py --workload c4 --steps 2 --warmup 3 --cpu-sample 128 > gpurun_out/final/bench_c4.json 2> gpurun_out/final/bench_c4.err
timeout 100 python bench.py --workload c5 --steps 3 --warmup 3 --cpu-sample 256 > gpurun_out/final/bench_c5.json 2> gpurun_out/final/bench_c5.err
timeout 100 python tools/mpc_latency.py 1024 16 > gpurun_out/final/mpc_1024.json 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/final/launches.csv python tools/gpu_one_solve.py phased 16384 > gpurun_out/final/one.log 2>&1
cap() { # name regex skip
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/final/$1 -f python tools/gpu_one_solve.py phased 16384 > gpurun_out/final/$1.log 2>&1
  python tools/ncu_summarize.py gpurun_out/final/$1.ncu-rep gpurun_out/final/ncu_$1_summary.txt > /dev/null 2>&1
  rm -f gpurun_out/final/$1.ncu-rep
}
cap k_backward_mat_phased k_backward_mat 24
cap k_ls_deep k_ls_deep 60
cap k_ls_wide k_ls_wide 24
cap k_update_expansions k_update_expansions 24
cap k_solve_outer k_solve 6
timeout 120 python tools/gpu_debug.py ncu_bp > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_backward_mat -s 1 -c 1 -o gpurun_out/final/bp_alone -f python tools/gpu_debug.py ncu_bp > gpurun_out/final/bp_alone.log 2>&1
python tools/ncu_summarize.py gpurun_out/final/bp_alone.ncu-rep gpurun_out/final/ncu_k_backward_mat_summary.txt > /dev/null 2>&1
rm -f gpurun_out/final/bp_alone.ncu-rep
ls -la gpurun_out/final
cat gpurun_out/final/tests.log gpurun_out/final/period8.log
head -c 600 gpurun_out/final/bench_c2.json
