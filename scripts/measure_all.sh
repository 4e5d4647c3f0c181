#!/bin/bash
# Developer tool: the one-GPU measurement suite of a round, run on the GPU box through gpurun (from the repository root):
# GPU tests, smoke, the four bench workloads, the reference arm, three ncu captures summarised into gpurun_out/ (the
# .ncu-rep files are deleted: gpurun_out/ is limited to 64 MiB), the launch list and the latency-mode crossover.
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2f_bench_push.json 2> gpurun_out/r2f_bench_push.err
timeout 400 python bench.py --workload walk > gpurun_out/r2f_bench_walk.json 2> gpurun_out/r2f_bench_walk.err
timeout 400 python bench.py --workload sweep > gpurun_out/r2f_bench_sweep.json 2> gpurun_out/r2f_bench_sweep.err
timeout 400 python bench.py --workload qp > gpurun_out/r2f_bench_qp.json 2> gpurun_out/r2f_bench_qp.err
timeout 400 python bench.py --impl reference > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
for m in settled:8 push:2 walk:6; do
  mode=${m%%:*}; skip=${m##*:}
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel_multi --launch-skip $skip --launch-count 1 -o gpurun_out/r2f_$mode -f python scripts/prof_target.py $mode > gpurun_out/ncu_$mode.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2f_$mode.ncu-rep gpurun_out/r2f_rollout_$mode.txt
  rm -f gpurun_out/r2f_$mode.ncu-rep
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
PYTHONPATH=. timeout 300 python scripts/latency_mode_crossover.py > gpurun_out/r2f_latency_crossover.log 2>&1
for f in gpurun_out/r2f_bench_*.json; do echo $f; cut -c1-400 $f; done
