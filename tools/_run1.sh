set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_l.log 2>&1; tail -3 gpurun_out/r02_pytest_l.log
python tools/k3_tune.py --planner-only --sizes 448x2560,448x65536,3584x65536 --poly=-2,0,4,8 > gpurun_out/r02_k3_tune_l.jsonl 2>&1; cat gpurun_out/r02_k3_tune_l.jsonl
for ov in 0 1; do python bench.py --steps 20 --warmup 5 --ema-overlap $ov --no-extras --no-cpu-baseline > gpurun_out/r02_bench_l_ov$ov.json 2> gpurun_out/r02_bench_l_ov$ov.err; done
for g in 296 1184 100000; do B200SSL_EMA_GRID=$g python bench.py --steps 20 --warmup 5 --ema-overlap 1 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_l_ov1_g$g.json 2> gpurun_out/r02_bench_l_ov1_g$g.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_l_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step']*1e3,1), 'e2e', round(d['e2e']['ms_per_step']*1e3,1), 'ema', round(d['roofline']['avg_launch_ms']*1e3,1))
    except Exception as e: print(f, 'ERR', e)
PY
