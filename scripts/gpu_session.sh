#!/bin/bash
# One gpurun call = several independent stages, each under its own timeout, logs into gpurun_out/.
# usage: scripts/gpu_session.sh stage1 stage2 ...   (stages: gemm pool encoder head all bench micro ncu)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for s in "$@"; do
  case $s in
    gemm)    timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" ;;
    gemmall) timeout 400 python -m pytest tests/test_gpu_gemm.py -q -m gpu > gpurun_out/t_gemm.log 2>&1; echo "gemm rc=$?" ;;
    pool)    timeout 300 python -m pytest tests/test_gpu_voxel_pool.py -q -m gpu -s > gpurun_out/t_pool.log 2>&1; echo "pool rc=$?" ;;
    encoder) timeout 600 python -m pytest tests/test_gpu_encoder.py -q -m gpu > gpurun_out/t_encoder.log 2>&1; echo "encoder rc=$?" ;;
    wattn)   timeout 240 python -m pytest tests/test_gpu_window_attn.py -q -m gpu > gpurun_out/t_wattn.log 2>&1; echo "wattn rc=$?" ;;
    head)    timeout 600 python -m pytest tests/test_gpu_head.py -q -m gpu > gpurun_out/t_head.log 2>&1; echo "head rc=$?" ;;
    full)    timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s > gpurun_out/t_full.log 2>&1; echo "full rc=$?" ;;
    neck)    timeout 600 python -m pytest tests/test_gpu_neck.py tests/test_gpu_eval.py -q -m gpu > gpurun_out/t_neck.log 2>&1; echo "neck rc=$?" ;;
    all)     timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?" ;;
    micro)   timeout 600 python scripts/microbench.py > gpurun_out/micro.log 2>&1; echo "micro rc=$?" ;;
    microelem) timeout 600 python scripts/microbench.py elem > gpurun_out/micro_elem.log 2>&1; echo "microelem rc=$?"; tail -n 20 gpurun_out/micro_elem.log ;;
    bench)   timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.log ;;
    benchq)  timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/benchq.log 2> gpurun_out/benchq.err; echo "benchq rc=$?"; tail -c 3000 gpurun_out/benchq.log; tail -n 5 gpurun_out/benchq.err ;;
    benchref) timeout 900 python bench.py --impl reference --steps 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"; tail -n 2 gpurun_out/bench_ref.log ;;
    ktimes)  timeout 600 python scripts/kernel_times.py gpurun_out/kernel_times.md > gpurun_out/kernel_times.log 2>&1; echo "ktimes rc=$?"; head -n 40 gpurun_out/kernel_times.md ;;
    ktimes4) OCC_BATCH=4 timeout 600 python scripts/kernel_times.py gpurun_out/kernel_times_b4.md > gpurun_out/kernel_times_b4.log 2>&1; echo "ktimes4 rc=$?"; head -n 30 gpurun_out/kernel_times_b4.md ;;
    smoke)   timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-1500} -c ${NCU_COUNT:-900} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "launches rc=$?" ;;
    ncupool) timeout 900 ncu --set full --clock-control none --import-source on -k regex:vp_ -s 12 -c 6 -f -o gpurun_out/prof_vp_pool python scripts/microbench.py pool > gpurun_out/ncu_vp_pool.log 2>&1; echo "ncupool rc=$?" ;;
    ncuhead) timeout 900 ncu --set full --clock-control none --import-source on -k regex:cross_attn_tc -s 11 -c 1 -f -o gpurun_out/prof_cross_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_cross_attn.log 2>&1; echo "ncuhead rc=$?" ;;
    ncuelem) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fuse_kernel|gn_upsample_add_x2" -s 2 -c 2 -f -o gpurun_out/prof_elem python scripts/microbench.py elem > gpurun_out/ncu_elem.log 2>&1; echo "ncuelem rc=$?" ;;
    ncuneckgemm) timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 4 -c 2 -f -o gpurun_out/prof_neck_gemm python scripts/microbench.py neck > gpurun_out/ncu_neck_gemm.log 2>&1; echo "ncuneckgemm rc=$?" ;;
    ncu_*)   k=${s#ncu_}; case $k in gemm_bf16x3*) w=conv ;; swin_qkv*|window_attn*) w=attn ;; swin_mlp*) w=tail ;; vp_pool*|lift_front*) w=lift ;; ms_deform*|token_prep*) w=neck ;; *) w=${NCU_WHICH:-conv} ;; esac
             timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-2} -c 1 -f -o gpurun_out/prof_$k python scripts/microbench.py $w > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?" ;;
    *) echo "unknown stage $s" ;;
  esac
done
tail -n 25 gpurun_out/t_*.log gpurun_out/micro.log 2>/dev/null | tail -n 120
