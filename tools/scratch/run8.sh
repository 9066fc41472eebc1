cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mcra or golden or drop_in" 2>&1 | tail -5
for e in 0 1; do
if [ $e = 1 ]; then export BF_MCRA_OLD=1; fi
timeout 300 python bench.py --workload mcra --no-e2e --no-cpu --no-extra --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mcra old=$e', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'])"
done
