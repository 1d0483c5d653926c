#!/bin/bash
# Round-2 profiling recipe (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are copied to profiles/ by hand).
# Every ncu command runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
HEAD="python bench.py --workload infer_drcnn --no-extras --no-cpu-baseline --steps 1 --warmup 1 --preload 0"
for prec in fp16 fp16x3; do
  $HEAD --precision $prec > gpurun_out/plain_$prec.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_infer_drcnn_$prec.csv \
      $HEAD --precision $prec > gpurun_out/ncu_launch_$prec.log 2>&1
  echo "launch list $prec rc=$?"
done
# the split-precision 40->40 fused block kernel (3 MMA passes per product)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 10 -c 2 -o gpurun_out/r02_conv_tc_x3 \
    $HEAD --precision fp16x3 > gpurun_out/ncu_full_x3.log 2>&1
echo "ncu full x3 rc=$?"
# the phase-split conv2 of Unet:M (384 -> 104, 3x1; row-merged operand rows, chunk-group stages): 19th conv launch of a forward
python tools/layer_times.py unet_m 646 fp16 > gpurun_out/plain_unet_layers.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 78 -c 1 -o gpurun_out/r02_conv2_rowmerged \
    python tools/layer_times.py unet_m 646 fp16 > gpurun_out/ncu_full_conv2.log 2>&1
echo "ncu full conv2 rc=$?"
