#!/bin/bash
# 8-GPU measurement session (round 1): sharded parity tests + weak-scaling bench + BASELINE configs 4 and 5.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5) | tee gpurun_out/pytest_mgpu8.log
run() { # name nproc args...
  name=$1; np=$2; shift 2
  (timeout 400 $TR --nproc-per-node $np --master-port $((29600+np)) bench.py --gpus $np "$@" 2>&1 | tail -1) > gpurun_out/$name.json
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(name, d["config"]["workload"], "value", round(d["value"],4), "e2e", round(d["e2e"]["value"],4), d["analytic_check"], d["clocks"],
          {k:(round(v["ms_per_step"],1), round(v.get("frac",0),3)) for k,v in d["kernels"].items()}, "cg", d["cg_iterations_per_solve"][:3])
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/{name}.json").read()[-1500:])
PY
}
run bench_P8_N27 8 --steps 3 --warmup 3
run bench_P4_N26 4 --steps 3 --warmup 3
run bench_P8_N28_k200 8 --spins 28 --k 200 --steps 3 --warmup 3
# (N=28 on 4 GPUs measured earlier in the round: profiles/r1_bench_P4_N28_k200.json)
run bench_P8_N30_kfit 8 --spins 30 --k 0 --steps 3 --warmup 3
