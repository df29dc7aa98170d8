mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_viterbi -s 2 -c 1 -o gpurun_out/vit_r16_v4 python tools/perf_probe.py --targets 30000 HHG_GROUP_JOBS=64 > gpurun_out/ncu7.log 2>&1
tail -n 2 gpurun_out/ncu7.log
