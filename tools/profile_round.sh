#!/bin/bash
# Round profile artefacts (run on the GPU box under gpurun; outputs in gpurun_out/, copied to profiles/ afterwards):
#   launch lists + DRAM traffic of every workload (bench shapes, ncu --metrics, --clock-control none)
#   ncu --set full captures of the C2, C5, mcra and phase kernels
R=${1:-r02}
K='regex:sel_|das_|frames_|phase_n|mcra_|srp_|save_prev|gss_|ref_kernel|gsc_|zero_hops'
for w in c1 c2 c3l c3g c4 c5 mcra ph; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --csv \
    --log-file gpurun_out/${R}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/${R}_launches_$w.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sel_stream -s 1 -c 1 -o gpurun_out/${R}_ncu_c2 \
  python bench.py --workload c2 --steps 1 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/${R}_ncu_c2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:srp_power_tc -s 1 -c 1 -o gpurun_out/${R}_ncu_c5 \
  python bench.py --workload c5 --steps 1 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/${R}_ncu_c5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mcra_pairs -s 1 -c 1 -o gpurun_out/${R}_ncu_mcra \
  python bench.py --workload mcra --steps 1 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/${R}_ncu_mcra.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:phase_n_kernel -s 1 -c 1 -o gpurun_out/${R}_ncu_ph \
  python bench.py --workload ph --steps 1 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/${R}_ncu_ph.log 2>&1
