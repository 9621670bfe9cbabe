#!/usr/bin/env bash
# Final-build refresh of the launch list (one solve) and of the adjoint sweep captures.
set -x
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tfim_sweep_pipe_kernel -s 1 -c 2 -o $O/r2_prof_sweeps_adjoint_L24 python scripts/bench_matvec.py --spins 24 --reps 1 --variants staged > $O/r2_prof_sweeps_adjoint_L24.log 2>&1
