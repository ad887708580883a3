#!/bin/bash
# Regenerates the measured evidence of a round on ONE B200 (run through gpurun from the repo root):
#     gpurun --timeout 2400 -- 'bash tools/profile_round.sh r02'
# Every number printed under ncu is a profile, never a bench value: the bench lines come from the plain runs below.
# Output: gpurun_out/${R}_*; the files worth keeping are copied to profiles/ by hand (profiles/README.md lists them).
R=${1:-r02}
O=gpurun_out
mkdir -p $O
set -x
# ---- bench lines (plain runs) --------------------------------------------------------------------------------------
python bench.py --steps 1000 --warmup 20 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
python bench.py --workload cfg1 --steps 200 --warmup 20 --no-extras > $O/${R}_bench_cfg1.json 2> $O/${R}_bench_cfg1.err
python bench.py --workload cfg3 --steps 200 --warmup 20 --no-extras > $O/${R}_bench_cfg3.json 2> $O/${R}_bench_cfg3.err
python bench.py --steps 200 --warmup 20 --ema-overlap 0 --no-extras --no-cpu-baseline > $O/${R}_bench_n1_serial_ema.json 2> $O/${R}_bench_n1_serial_ema.err
# ---- per-kernel times, head sweep (SURVEY 8d cfg 5), fused optimizer + EMA -------------------------------------------
python tools/microbench.py > $O/${R}_microbench_cfg2.txt 2>&1
python tools/microbench.py --rows 3584 --batch 512 --bank 65536 > $O/${R}_microbench_rows3584_bank65536.txt 2>&1
python tools/k3_tune.py --planner-only --sizes 448x2560,448x20480,448x65536,1792x16384,3584x65536,14336x65536 --poly=0,8,100 > $O/${R}_k3_tune.jsonl 2>&1
python tools/k3_f32.py > $O/${R}_k3_fp32_storage_tc.jsonl 2>&1
B200SSL_K3_F32_SIMT=1 python tools/k3_f32.py --sizes 448x2560,448x65536,3584x32768 > $O/${R}_k3_fp32_storage_ffma.jsonl 2>&1
python tools/opt_bench.py > $O/${R}_opt_bench.json 2>&1
python tools/sweep.py --cpu-budget 1.0 > $O/${R}_sweep_cfg5.jsonl 2> $O/${R}_sweep_cfg5.err
for b in tmem_bench mma_bench; do          # the two micro-benchmarks are plain nvcc programs (binaries are not tracked)
  [ -x tools/micro/$b ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iendoscopy-image-classification_b200/csrc -Iinclude \
      -o tools/micro/$b tools/micro/$b.cu
done
tools/micro/tmem_bench > $O/${R}_micro_tmem_mufu_sts.jsonl 2>&1
tools/micro/mma_bench > $O/${R}_micro_mma_issue.jsonl 2>&1
# ---- ncu: launch list of the bench command, then one full capture per hot kernel -------------------------------------
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --graph-profiling node -c 400 --csv \
    --log-file $O/${R}_launches_bench_cfg2.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-extras --no-cpu-baseline > $O/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ema_multi_tensor -s 2 -c 1 -f -o $O/${R}_ncu_ema \
    python tools/microbench.py --reps 2 > $O/${R}_ncu_ema.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bank_smooth_tc -s 2 -c 1 -f -o $O/${R}_ncu_k3_big \
    python tools/k3_only.py 3584 65536 > $O/${R}_ncu_k3_big.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bank_smooth_tc -s 2 -c 1 -f -o $O/${R}_ncu_k3_cfg2 \
    python tools/k3_only.py 448 2560 > $O/${R}_ncu_k3_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:contrast_tc_fwd -s 2 -c 1 -f -o $O/${R}_ncu_contrast_fwd_big \
    python tools/contrast_only.py 3584 > $O/${R}_ncu_contrast_fwd_big.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:contrast_tc_bwd -s 2 -c 1 -f -o $O/${R}_ncu_contrast_bwd_big \
    python tools/contrast_only.py 3584 > $O/${R}_ncu_contrast_bwd_big.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:comatch_rows_fused -s 2 -c 1 -f -o $O/${R}_ncu_rows_fused \
    python tools/microbench.py --reps 2 > $O/${R}_ncu_rows_fused.log 2>&1
# ---- compute-sanitizer (racecheck / synccheck of tools/sanitize_target.py) is closed on this pool ("runs under it have left
# GPUs needing a reset"): the attempt of round 2 is kept as profiles/r02_sanitizer_closed.log.  The intra-kernel protocols are
# covered by the randomised interleaving model (tests/test_peer_protocol_model.py), bounded waits with sticky abort / timeout
# counters in every kernel, and bit-exact replay-vs-eager tests.
python tools/microbench.py --dtype f32 > $O/${R}_microbench_cfg2_f32.txt 2>&1
ls -la $O | grep ${R}_ | tail -40
