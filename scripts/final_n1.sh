#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): full GPU test suite, bench (ours + reference arm), the other configs,
# the launch list of a bench run and ncu --set full summaries.  Only small text files are left in gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; tail -14 gpurun_out/r02_pytest_gpu.log   # LUXB_SKIP_HEAVY=1 skips C2 / C3 (run r2a: 47 passed incl. both)
python bench.py > gpurun_out/r02_bench_n1.out 2> gpurun_out/r02_bench_n1.err; grep "^{" gpurun_out/r02_bench_n1.out > gpurun_out/r02_bench_n1_rmat27.json; cut -c1-1500 gpurun_out/r02_bench_n1_rmat27.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.out 2> gpurun_out/r02_bench_ref.err; grep "^{" gpurun_out/r02_bench_ref.out > gpurun_out/r02_bench_reference_arm.json; cut -c1-600 gpurun_out/r02_bench_reference_arm.json
python scripts/bench_configs.py C1,C3,C4,C5 > gpurun_out/r02_configs_1gpu.txt 2>&1; mv gpurun_out/configs_1gpu.json gpurun_out/r02_configs_1gpu.json; cut -c1-400 gpurun_out/r02_configs_1gpu.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-parity --no-cpu-baseline --no-ref-gpu --no-e2e > /dev/null 2>&1
python scripts/summarize_ncu.py list gpurun_out/launches_r2.csv gpurun_out/r02_launches_bench_rmat27.md "Round 2 — every launch of bench.py --steps 1 --warmup 1 (RMAT-27, 1 B200), ncu gpu__time_duration.sum" > /dev/null 2>&1; rm -f gpurun_out/launches_r2.csv; head -30 gpurun_out/r02_launches_bench_rmat27.md
ncu --set full --clock-control none --import-source on -k regex:seg_tile -s 12 -c 2 -o gpurun_out/prof_seg python scripts/sweep_panel.py 27 "1:48:64" > gpurun_out/r02_ncu_seg.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/prof_seg.ncu-rep gpurun_out/r02_seg_sweep_rmat27_full.md "Round 2 (final shapes) — seg_tile_kernel panel (24 warps x 16 edges/lane, 48 blocks x 32768 values) and main (8 warps x 2 stages x 1 round, 3 CTAs/SM) on RMAT-27, 1 B200" > /dev/null 2>&1; rm -f gpurun_out/prof_seg.ncu-rep
ncu --set full --clock-control none -k regex:"cf_chunk|cf_update|push_relax|push_big|frontier_|seg_tile|combine_hub|pull_fixup" -c 36 -o gpurun_out/prof_other python scripts/prof_other_kernels.py > gpurun_out/r02_ncu_other.log 2>&1
python scripts/summarize_ncu.py full gpurun_out/prof_other.ncu-rep gpurun_out/r02_other_kernels_full.md "Round 2 — the non-PageRank kernels at C4 (SSSP RMAT-24) and C5 (col_filter NetFlix-scale) scale, 1 B200" > /dev/null 2>&1; rm -f gpurun_out/prof_other.ncu-rep
ls -la gpurun_out | tail -20
