#!/bin/bash
# Round profile pass: plain bench, ncu launch list of the same command, ncu --set full of every hot kernel.
# Only text summaries are kept (gpurun_out/ is capped at 64 MiB): the .ncu-rep is converted to CSV on the box and removed.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/bench_prof_plain.log 2>&1; echo "plain rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc $?"
timeout 300 python tools/prof_kernels.py 256 3 gpurun_out/r02_kernel_table.md > gpurun_out/prof_kernels.log 2>&1; echo "kernels rc $?"
timeout 1200 ncu --set full --clock-control none -k regex:"gemm_tc2|wgrad_tc|dwconv7|ln_fwd_bf16|ln_bwd_bf16|attn_fwd_tc2|attn_bwd_tc2|rope_qk|mlp_fused" -o /tmp/r02_kernels -f python tools/prof_kernels.py 256 1 > gpurun_out/ncu_kernels.log 2>&1; echo "ncu kernels rc $?"
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r02_kernels_raw.csv
du -sh gpurun_out
