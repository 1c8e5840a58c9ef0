set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_gputests_s3.log
python bench.py > gpurun_out/r02_bench_s3.json 2> gpurun_out/r02_bench_s3.err
timeout 300 python tools/trace.py 32 gpurun_out/trace_s3g.tsv > gpurun_out/r02_trace_s3g.txt 2>&1
for c in sd_decode vbr_sweep 4k_bands; do timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02_bench_s3_$c.json 2>gpurun_out/r02_bench_s3_$c.err; done
timeout 300 python tools/graph_probe.py > gpurun_out/r02_graph_replay_s3.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_s3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_s3.log 2>&1
cat gpurun_out/r02_gputests_s3.log; head -c 600 gpurun_out/r02_bench_s3.json
