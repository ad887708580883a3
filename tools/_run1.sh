python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_n.log 2>&1; tail -3 gpurun_out/r02_pytest_n.log
python tools/k3_tune.py --planner-only --sizes 448x2560,448x20480,448x65536,1792x16384,3584x65536,14336x65536 --poly=0,8,16,100 > gpurun_out/r02_k3_tune_q.jsonl 2>&1; cat gpurun_out/r02_k3_tune_q.jsonl
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; tail -c 300 gpurun_out/r02_bench_q.err
