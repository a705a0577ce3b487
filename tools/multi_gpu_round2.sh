#!/bin/bash
# One 8-GPU box: (1) host-to-device ceiling of the box at N = 1/2/4/8 (tools/h2d_probe.py: pinned 11.2 GB per rank,
# plain cudaMemcpyAsync, no kernels), with per-rank core affinity and with write-combined memory at N = 8;
# (2) BASELINE config 5 (512-window degradation sweep) sharded over 8 GPUs; (3) bench.py at N = 8 with and without
# per-rank core affinity (device-timed value + end-to-end).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/h2d_probe.jsonl
for n in 1 2 4 8; do
  $TR --master-port 2951$n --nproc-per-node $n tools/h2d_probe.py --reps 2 2>gpurun_out/h2d_err_$n.log | grep '^{' >> gpurun_out/h2d_probe.jsonl
done
$TR --master-port 29521 --nproc-per-node 8 tools/h2d_probe.py --reps 2 --affinity 2>>gpurun_out/h2d_err_8.log | grep '^{' >> gpurun_out/h2d_probe.jsonl
$TR --master-port 29522 --nproc-per-node 8 tools/h2d_probe.py --reps 2 --wc 2>>gpurun_out/h2d_err_8.log | grep '^{' >> gpurun_out/h2d_probe.jsonl
cat gpurun_out/h2d_probe.jsonl
$TR --master-port 29523 --nproc-per-node 8 tools/run_configs.py --config c5 2>gpurun_out/c5_n8_err.log | grep '^{' > gpurun_out/c5_n8.json
cut -c1-400 gpurun_out/c5_n8.json
$TR --master-port 29524 --nproc-per-node 8 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --affinity off 2>gpurun_out/bench_n8_err.log | grep '^{' > gpurun_out/bench_n8_off.json
$TR --master-port 29525 --nproc-per-node 8 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --affinity on 2>>gpurun_out/bench_n8_err.log | grep '^{' > gpurun_out/bench_n8_on.json
python - <<'PY'
import json
for f in ("off", "on"):
    try:
        l = json.loads(open(f"gpurun_out/bench_n8_{f}.json").read())
        print(f, "value", round(l["value"]), "e2e", round(l["e2e"]["value"]), "ms/step e2e", round(l["e2e"]["ms_per_step"], 1), l["bpm_ok"])
    except Exception as e:
        print(f, "ERR", e)
PY
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/lscpu.txt 2>&1; free -g >> gpurun_out/lscpu.txt
