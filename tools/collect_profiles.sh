#!/bin/bash
# Runs on the GPU box (under gpurun): collects the evidence summarised under profiles/ into gpurun_out/prof/.
#   launches.csv   ncu launch list (gpu__time_duration) of bench.py --steps 1 --warmup 3
#   layers.txt     per-launch CUDA-event times of one forward, grouped by shape
#   ops_<name>.csv ncu --set full raw page of representative heavy shapes run in isolation (tools/ncu_ops.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/prof
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/prof/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/prof/launches_bench.log 2>&1
gzip -f gpurun_out/prof/launches.csv
timeout 300 python tools/layer_profile.py 16 512 > gpurun_out/prof/layers.txt 2>&1
if [ -z "${PGT_PROFILE_SKIP_TRAFFIC:-}" ]; then
# DRAM traffic of every launch of one forward (roofline.traffic of bench.py: profiles/<tag>_traffic.json)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/prof/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/prof/traffic_bench.log 2>&1
gzip -f gpurun_out/prof/traffic.csv
fi
for op in ${PGT_PROFILE_OPS:-halo64 halo128 conv256 linear256 linear512 linear512_f32 up128 swin_mlp rgb gn window_tc mha_tc argmax l2_argmin ln_linear conv_out}; do
  case $op in
    halo64|halo128|up128) rx='regex:conv_halo' ;;
    conv256|linear256|linear512|linear512_f32) rx='regex:gemm_tc_kernel' ;;
    swin_mlp) rx='regex:swin_mlp' ;;
    rgb) rx='regex:rgb_conv' ;;
    gn) rx='regex:gn_apply' ;;
    window_tc) rx='regex:window_attn_tc' ;;
    mha_tc) rx='regex:mha_tc_kernel' ;;
    argmax) rx='regex:argmax_gather' ;;
    l2_argmin) rx='regex:l2_argmin_pair' ;;
    ln_linear) rx='regex:ln_linear' ;;
    conv_out) rx='regex:conv_out_gn' ;;
  esac
  timeout 200 ncu --set full --clock-control none -k "$rx" -s 2 -c 1 -o gpurun_out/prof/ops_$op -f python tools/ncu_ops.py $op > gpurun_out/prof/ops_$op.log 2>&1
  ncu -i gpurun_out/prof/ops_$op.ncu-rep --page raw --csv > gpurun_out/prof/ops_$op.csv 2>/dev/null
  rm -f gpurun_out/prof/ops_$op.ncu-rep
done
ls -la gpurun_out/prof
