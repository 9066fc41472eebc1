cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in c4 ph; do
timeout 300 python bench.py --workload $w --no-e2e --no-cpu --no-extra --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'])"
done
