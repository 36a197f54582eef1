mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r1d_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --no-cpu > gpurun_out/r1d_bench_s1.json 2> gpurun_out/r1d_bench_s1.err; echo "bench rc=$?"
tail -n 1 gpurun_out/r1d_bench_s1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], {k:round(v['ms_per_step'],3) for k,v in d['breakdown_rank0'].items() if isinstance(v,dict)})"
