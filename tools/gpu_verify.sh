#!/bin/bash
# One gpurun call that re-establishes the evidence for the tree as it stands: GPU parity tests, the driver's bench
# line, one `ncu --set full` capture of the dominant kernel and the launch list.  Ordered by importance — the call may
# be cut by the GPU budget.      gpurun --timeout 900 -- 'bash tools/gpu_verify.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
date +%s > gpurun_out/t0
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))s" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))s"; cat gpurun_out/bench_n1.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_rec -s 20 -c 1 -f -o gpurun_out/rec256 \
    python bench.py --steps 1 --warmup 1 --inner 20 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --inner 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))s"
timeout 150 python tools/configs_bench.py > gpurun_out/small_configs.jsonl 2> gpurun_out/small_configs.err
echo "small configs rc=$? t=$(( $(date +%s) - $(cat gpurun_out/t0) ))s"; cut -c1-220 gpurun_out/small_configs.jsonl
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
