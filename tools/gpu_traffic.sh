# fused vs unfused DRAM traffic (ncu), both packages
mkdir -p gpurun_out
for impl in ours reference; do
  python tools/traffic_probe.py --impl $impl > gpurun_out/traffic_plain_$impl.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/traffic_$impl.csv python tools/traffic_probe.py --impl $impl > gpurun_out/traffic_ncu_$impl.log 2>&1
  tail -3 gpurun_out/traffic_plain_$impl.log gpurun_out/traffic_ncu_$impl.log
done
python tools/traffic_probe.py --summarise gpurun_out/traffic_ours.csv gpurun_out/traffic_reference.csv | head -12
