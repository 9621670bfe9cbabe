set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'tfim_sweep' -s 8 -c 6 -o $O/r2_prof_sweeps_L24 python scripts/bench_matvec.py --spins 24 --reps 1 --variants staged > $O/r2_prof_sweeps_L24.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tfim_sweep' -s 8 -c 6 -o $O/r2_prof_sweeps_L26 python scripts/bench_matvec.py --spins 26 --reps 1 --variants staged > $O/r2_prof_sweeps_L26.log 2>&1
timeout 300 python scripts/bench_matvec.py --spins 20 22 24 25 26 27 --variants r1_plan sweeps3 direct staged generic --out $O/r2_matvec_micro_final.json > $O/r2_matvec_micro_final.log 2>&1
timeout 300 python bench.py > $O/r2_bench_final_P1.json 2> $O/r2_bench_final_P1.err
