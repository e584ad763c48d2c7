#!/bin/bash
# Round 2, GPU call A (1 GPU): where does the random-gather ceiling live (footprint x concurrency sweep, plain and under
# ncu), a full ncu capture of the default (id-space) fold kernel at RMAT-26 and one on the C5 Zipf graph.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
S=gpurun_out/r2a_summary.txt
: > $S
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv >> $S 2>&1
t0=$(date +%s)
stage() { echo "== $1: rc=$2 at +$(( $(date +%s) - t0 ))s" >> $S; }
M=gpu__time_duration.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_stalled.sum,l1tex__m_l1tex2xbar_req_cycles_active.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__sectors_read.sum,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__d_sectors_fill_sysmem.sum,sm__cycles_elapsed.max
timeout 300 profiles/probes/gather_sweep 64 > gpurun_out/r2_gather_sweep.jsonl 2> gpurun_out/r2_gather_sweep.err; stage sweep $?
cat gpurun_out/r2_gather_sweep.jsonl >> $S
timeout 900 ncu --metrics $M --clock-control none -k regex:gather_ --csv --log-file gpurun_out/r2_gather_sweep_ncu.csv \
    profiles/probes/gather_sweep 64 0.25 > gpurun_out/r2_gather_sweep_under_ncu.log 2>&1; stage sweep_ncu $?
timeout 700 ncu --set full --clock-control none --import-source on -k regex:walk_fold_conv -s 1 -c 1 -o gpurun_out/r2_prof_fold_ids -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-exact > gpurun_out/r2_bench_under_ncu_full.log 2>&1; stage ncu_full $?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk_fold_conv -s 1 -c 1 -o gpurun_out/r2_prof_c5 -f \
    python profiles/run_c5.py > gpurun_out/r2_c5_under_ncu.log 2>&1; stage ncu_c5 $?
timeout 300 python profiles/run_c5.py > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; stage c5 $?
cat gpurun_out/r2_c5.json >> $S
cat $S
