cd /root/repo; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase or mcra or frame or c4 or golden or drop_in" 2>&1 | tail -3
for w in c4 ph; do
timeout 300 python bench.py --workload $w --no-e2e --no-cpu --no-extra --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'])"
done
BF_MCRA_OLD=1 timeout 300 python bench.py --workload mcra --no-e2e --no-cpu --no-extra --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mcra old', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'])"
