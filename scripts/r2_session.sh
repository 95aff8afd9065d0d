#!/bin/bash
# One GPU session: tests, bench, A/B legs, light ncu metrics.  Usage: gpurun -- bash scripts/r2_session.sh <tag> [what...]
tag=$1; shift
what=${@:-tests bench ab ncu}
mkdir -p gpurun_out
for w in $what; do
case $w in
tests) timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests_$tag.log ;;
newtests) timeout 1200 python -m pytest tests/test_gpu_config_sizes.py tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests_$tag.log ;;
bench) timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$tag.json ;;
ab) for t in ${AB_TUNINGS:-0 64 32 16}; do GS_TUNING=$t timeout 300 python scripts/stage_times.py > gpurun_out/stages_${tag}_t$t.log 2>&1; echo "tuning $t:"; tail -2 gpurun_out/stages_${tag}_t$t.log; done ;;
ncu) timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_composite|k_preprocess|k_stratum|k_tile_sort" -s 25 -c 10 --csv --log-file gpurun_out/ncu_light_$tag.csv env GS_STEPS=8 python scripts/profile_step.py > gpurun_out/ncu_light_$tag.log 2>&1; echo "ncu rc=$?" ;;
abc4) for t in ${AB_TUNINGS:-0 32 64}; do GS_TUNING=$t GS_P=2000000 GS_V=32 GS_HW=512 GS_STEPS=6 timeout 400 python scripts/stage_times.py > gpurun_out/stages_c4_${tag}_t$t.log 2>&1; echo "C4 tuning $t:"; tail -2 gpurun_out/stages_c4_${tag}_t$t.log; done ;;
abc5) for t in ${AB_TUNINGS:-0 32 64}; do GS_TUNING=$t GS_SCENE=aligned GS_V=3 GS_HW=256 GS_STEPS=20 timeout 300 python scripts/stage_times.py > gpurun_out/stages_c5_${tag}_t$t.log 2>&1; echo "C5-shape tuning $t:"; tail -2 gpurun_out/stages_c5_${tag}_t$t.log; done ;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 --no-moving > gpurun_out/launches_$tag.log 2>&1; echo "launches rc=$?" ;;
abe2e) AB_TUNING=${AB_E2E:-0,131072} timeout 400 python scripts/ab_e2e.py > gpurun_out/ab_e2e_$tag.log 2>&1; echo "ab_e2e rc=$?"; tail -6 gpurun_out/ab_e2e_$tag.log ;;
hosttests) timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -q -s -k "host_buffer or sh_coefficient or variants" > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" ; tail -5 gpurun_out/tests_$tag.log ;;
probe) timeout 120 scripts/probes/pcie_pull_probe > gpurun_out/pcie_probe_$tag.log 2>&1; cat gpurun_out/pcie_probe_$tag.log ;;
esac
done
