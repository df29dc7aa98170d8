mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_viterbi -s 3 -c 1 -o gpurun_out/vit_bench_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/ncu_bench.log | cut -c1-300
