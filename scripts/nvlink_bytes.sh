#!/usr/bin/env bash
# NVLink traffic of the fused exchange, from the driver's per-link data counters (ncu must not wrap a multi-rank job):
#   nvidia-smi nvlink -gt d   before and after   torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 2 --warmup 1 --no-extras
# 6 solves run in total (1 warm-up + 2 timed, then 1 + 2 through host buffers).  Per solve and GPU the model is
# (k + CG matvecs + a few unfused pushes) shards of 8 * 2^24 bytes sent to the one partner.
O=gpurun_out
nvidia-smi nvlink -gt d -i 0 > $O/r2_nvlink_before.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 2 --warmup 1 --no-extras > $O/r2_nvlink_bench.json 2> /dev/null
nvidia-smi nvlink -gt d -i 0 > $O/r2_nvlink_after.txt 2>&1
python - <<'PY'
import json, re
def total(path):
    tx = rx = 0
    for line in open(path):
        m = re.search(r"Data (Tx|Rx):\s*(\d+)\s*KiB", line)
        if m:
            if m.group(1) == "Tx": tx += int(m.group(2))
            else: rx += int(m.group(2))
    return tx * 1024, rx * 1024
b, a = total("gpurun_out/r2_nvlink_before.txt"), total("gpurun_out/r2_nvlink_after.txt")
d = json.loads([l for l in open("gpurun_out/r2_nvlink_bench.json") if l.startswith("{")][-1])
cg = d["cg_iterations_per_solve"]
solves = 6
shard = 8 * 2 ** 24
model = solves * (200 + 1 + sum(cg) / len(cg) + 2) * shard          # K2b pushes + q0 + CG direction pushes + x0 / first d
out = {"gpu0_tx_bytes": a[0] - b[0], "gpu0_rx_bytes": a[1] - b[1], "solves": solves, "cg_iterations_per_solve": cg,
       "model_bytes_per_direction": model, "tx_over_model": (a[0] - b[0]) / model, "rx_over_model": (a[1] - b[1]) / model,
       "s_per_solve": d["value"]}
print(json.dumps(out))
json.dump(out, open("gpurun_out/r2_nvlink_bytes.json", "w"), indent=1)
PY
