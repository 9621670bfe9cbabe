#!/usr/bin/env bash
# Round-2 profiling pass (run on ONE B200 under gpurun; numbers printed under ncu are never bench values).
#   1. launch list of one bench step: every kernel with its device time (cold-cache, serialised: compare SHARES);
#   2. `ncu --set full` captures of every shipped hot kernel: both re-orthogonalisation passes at m ~ 100 (fp64 basis
#      and the opt-in fp32 shadow basis), the TFIM sweeps of the default plan at L = 24 (pipelined first sweep with TMA
#      staging, pipelined strided sweep with direct bits, pipelined adjoint) and at L = 26 (generic last sweep).
# Summaries are produced afterwards on the CPU box by scripts/ncu_summarise.py into profiles/.
set -x
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r2_launches.csv $B > $O/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'reorth_(dots|update)_kernel' -s 599 -c 4 -o $O/r2_prof_reorth_fp64 $B > $O/r2_prof_reorth_fp64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'reorth_(dots|update)_kernel' -s 599 -c 4 -o $O/r2_prof_reorth_fp32 $B --basis fp32 > $O/r2_prof_reorth_fp32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tfim_sweep' -s 8 -c 8 -o $O/r2_prof_sweeps_L24 python scripts/bench_matvec.py --spins 24 --reps 1 --variants direct > $O/r2_prof_sweeps_L24.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tfim_sweep' -s 8 -c 8 -o $O/r2_prof_sweeps_L26 python scripts/bench_matvec.py --spins 26 --reps 1 --variants direct > $O/r2_prof_sweeps_L26.log 2>&1
ls -la $O/*.ncu-rep
