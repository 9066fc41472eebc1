cd /root/repo; mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sel_pairs -s 1 -c 1 -o gpurun_out/r02_ncu_c3g \
  python bench.py --workload c3g --steps 1 --warmup 1 --passes 1 --no-e2e --no-cpu --no-extra > gpurun_out/r02_ncu_c3g.log 2>&1
