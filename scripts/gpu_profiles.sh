#!/bin/bash
# final evidence pass: bench line, one `ncu --set full` capture of the dominant kernel, the launch list of the same command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 12 -o gpurun_out/prof_conv_tc_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err
ls -la gpurun_out/*.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 2760 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
python scripts/summarize_launches.py gpurun_out/launches_final.csv > gpurun_out/launches_final_summary.md 2>&1; head -12 gpurun_out/launches_final_summary.md
