set -x
mkdir -p gpurun_out/ncu
timeout 200 python tools/gpu_engines.py timing 2>&1 | grep phased
cap() { # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/ncu/$1 -f python tools/gpu_one_solve.py phased 16384 > gpurun_out/ncu/$1.log 2>&1
  python tools/ncu_summarize.py gpurun_out/ncu/$1.ncu-rep gpurun_out/ncu/$1_summary.txt > /dev/null 2>&1
  rm -f gpurun_out/ncu/$1.ncu-rep
}
cap ls_deep_phaseB k_ls_deep 90
cap ls_wide_phaseA k_ls_wide 24
cap expansions_phaseA k_update_expansions 24
cap backward_phaseA k_backward_mat 24
ls -la gpurun_out/ncu
