for m in 0 1 2; do
  for b in 1 32; do
    echo "PDL=$m B=$b: $(MLIC_PDL=$m timeout 100 python bench.py --batch $b --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])')"
  done
done
MLIC_PDL=1 timeout 200 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
