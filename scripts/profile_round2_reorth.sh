set -x
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > $O/r2_pytest12.log 2>&1; tail -2 $O/r2_pytest12.log
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:'reorth_(dots|update)_kernel' -s 599 -c 4 -o $O/r2_prof_reorth_fp64 $B > $O/r2_prof_reorth_fp64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'reorth_(dots|update)_kernel' -s 599 -c 4 -o $O/r2_prof_reorth_fp32 $B --basis fp32 > $O/r2_prof_reorth_fp32.log 2>&1
timeout 200 python bench.py --no-cpu-baseline > $O/r2_bench12.json 2>/dev/null
