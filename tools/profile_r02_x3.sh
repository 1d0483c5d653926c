HEAD="python bench.py --workload infer_drcnn --no-extras --no-cpu-baseline --steps 1 --warmup 1 --preload 0"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 8 -c 1 -o gpurun_out/r02_conv_tc_x3_block5 \
    $HEAD --precision fp16x3 > gpurun_out/ncu_full_x3b.log 2>&1
echo "ncu full x3 rc=$?"
