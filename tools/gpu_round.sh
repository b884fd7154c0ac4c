#!/bin/bash
# One GPU session (run under gpurun): tests, per-op profiles, ncu launch metrics of a forward, ncu --set full captures
# of the hot kernels, a short bench.  usage: tools/gpu_round.sh TAG [tests] [ops] [fwd] [ncu] [bench] [launches]
# Everything lands in gpurun_out/TAG_*; summarise here with tools/ncu_summary.py and copy what should be judged to profiles/.
TAG=${1:-rX}; shift
WHAT="${*:-tests ops fwd ncu bench}"
O=gpurun_out
mkdir -p $O
has() { [[ " $WHAT " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${TAG}_smi.txt 2>&1

if has tests; then
  timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
  tail -5 $O/${TAG}_tests.log
fi
if has ops; then
  LD_PROFILE_OPS=60 timeout 300 python tools/gpu_profile_ops.py 32 256 mri 3 2> $O/${TAG}_ops_mri_32x256.txt
  LD_PROFILE_OPS=60 timeout 300 python tools/gpu_profile_ops.py 8 512 mri 3 2> $O/${TAG}_ops_mri_8x512.txt
  LD_PROFILE_OPS=60 timeout 300 python tools/gpu_profile_ops.py 32 128 mri_attn8 3 2> $O/${TAG}_ops_attn8_32x128.txt
  LD_PROFILE_OPS=60 timeout 300 python tools/gpu_profile_ops.py 32 256 mri_attn8 3 2> $O/${TAG}_ops_attn8_32x256.txt
  grep "LDPROF total" $O/${TAG}_ops_*.txt | tail -12
fi
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
M="$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"
M="$M,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active"
if has fwd; then
  # every launch of two UNet forwards (N=32, 256x256) with in-situ utilisation metrics; the second call is the warm one
  timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/${TAG}_fwd_metrics.csv python tools/gpu_profile_ops.py 32 256 mri 2 > $O/${TAG}_fwd_metrics.log 2>&1
  tail -2 $O/${TAG}_fwd_metrics.log
fi
cap() {  # cap NAME REGEX SKIP [source] -- cmd...
  local name=$1 rx=$2 skip=$3 src=$4; shift 4
  local rep=$O/${TAG}_ncu_${name}
  timeout 600 ncu --set full --clock-control none $( [ "$src" = src ] && echo --import-source on ) -k regex:$rx -s $skip -c 1 -f -o $rep "$@" > $rep.log 2>&1
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page raw --csv > $rep.raw.csv 2>/dev/null
    [ "$src" = src ] && ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > $rep.source.csv.gz
    # keep the report itself only when it is small (gpurun_out is capped at 64 MiB)
    [ $(stat -c %s $rep.ncu-rep) -gt 6000000 ] && rm -f $rep.ncu-rep
  else
    tail -3 $rep.log
  fi
}
if has ncuxf; then   # only the normalise-on-load and plain variants, with source
  cap conv32_xf conv_tc_kernel 2 src python tools/gpu_conv_pro_one.py
  cap conv32_plain conv_tc_kernel 3 src python tools/gpu_conv_one.py 32 0 256 32 3 0 32
fi
if has ncu64; then   # 64 -> 64 at 128 x 128
  cap conv64_plain conv_tc_kernel 3 src python tools/gpu_variant_shape.py 0 64 0 32 128 64
fi
if has ncu7; then   # only the init conv
  cap conv7 conv7_ 1 nosrc python tools/gpu_profile_ops.py 32 256 mri 2
fi
if has ncu; then
  cap conv32_plain conv_tc_kernel 3 src python tools/gpu_conv_one.py 32 0 256 32 3 0 32
  cap conv32_xf conv_tc_kernel 2 src python tools/gpu_conv_pro_one.py
  cap conv32_dual conv_tc_kernel 2 src python tools/gpu_conv_dual_one.py
  cap conv256 conv_tc_kernel 3 src python tools/gpu_conv_one.py 256 0 32 256 3 0 32
  cap conv32_stats conv_tc_kernel 5 nosrc python tools/gpu_conv_variant_one.py 1
  cap conv_up2 conv_tc_kernel 3 nosrc python tools/gpu_conv_one.py 64 0 256 32 3 1 32 3
  cap la_ctx la_ctx_kernel 1 src python tools/gpu_la_dbg.py 32 32 65536 2
  cap la_out la_out_kernel 1 src python tools/gpu_la_dbg.py 32 32 65536 2
  cap la_out8 la_out_kernel 1 nosrc python tools/gpu_la_dbg.py 32 32 16384 2 8
  cap attn attn_tc_kernel 2 src python tools/gpu_attn_one.py 32 1024 4
  cap step step_kernel 3 nosrc python tools/gpu_step_one.py 16 256 0
  cap knn knn_tc_kernel 1 nosrc python tools/gpu_knn_one.py
  cap attn4096 attn_tc_kernel 1 nosrc python tools/gpu_attn_one.py 8 4096 8
  cap gn_apply gn_apply_bf16_fast 20 nosrc python tools/gpu_profile_ops.py 32 256 mri 2
  cap conv7 conv7_ 1 nosrc python tools/gpu_profile_ops.py 32 256 mri 2
  ls -la $O/${TAG}_ncu_* | head -40
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --timesteps 12 --no-cpu-baseline --no-e2e --no-roofline > $O/${TAG}_launches.log 2>&1
  python profiles/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launch_shares.md 2>&1
fi
if has bench; then
  timeout 900 python bench.py --steps 2 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
  tail -c 1500 $O/${TAG}_bench.json
fi
