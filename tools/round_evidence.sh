#!/bin/bash
# Everything the round's numbers come from, in one gpurun call:  bash tools/round_evidence.sh r02
R=${1:-r02}
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 400 gpurun_out/${R}_bench.json
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${R}_bench_ref_arm.json 2> gpurun_out/${R}_bench_ref_arm.err
for w in mcra ref gsc ph c2hi; do
  timeout 400 python bench.py --workload $w --no-extra --steps 10 --warmup 3 > gpurun_out/${R}_bench_$w.json 2> gpurun_out/${R}_bench_$w.err
done
bash tools/profile_round.sh $R
timeout 600 python tools/hop_latency.py > gpurun_out/${R}_hop_latency.json 2> gpurun_out/${R}_hop_latency.err
ls -la gpurun_out | tail -40
