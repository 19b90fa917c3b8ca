#!/bin/bash
# Evidence captures of one round (run on the GPU box through gpurun; everything lands in gpurun_out/):
#   1. ncu launch list of the bench step (time + DRAM bytes per launch)        -> r2_launches_bench_step.csv
#   2. ncu --set full of the FusionNet forward (conv_ss / conv_tc / conv_chain) -> r2_conv_full.ncu-rep
#   3. ncu --set full of the first conv_wt launches of an AdapNet++ forward      -> r2_wt_full.ncu-rep
#   4. ncu --set full of extract / integrate                                     -> r2_memkernels_full.ncu-rep
# profiles/summarize_full.py turns the reports into the text summaries kept under profiles/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches_bench_step.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline --skip-e2e --ncu-range \
    > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_ss_kernel|conv_tc_kernel|conv_chain_kernel' \
    -o gpurun_out/r2_conv_full -f python tools/fusionnet_bench.py --reps 1 --profile > gpurun_out/r2_conv_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'conv_wt_kernel' -s 66 -c 14 \
    -o gpurun_out/r2_wt_full -f python tools/adapnet_once.py > gpurun_out/r2_wt_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'extract_kernel|count_kernel|offsets_kernel|scatter_kernel|rank_kernel|apply_kernel' -c 8 \
    -o gpurun_out/r2_memkernels_full -f python tools/kernel_bench.py --frames 4 --reps 2 --profile > gpurun_out/r2_mem_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
