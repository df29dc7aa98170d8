python -m pytest tests/test_viterbi_gpu.py -x -q -m gpu 2>&1 | tail -3
V=hh-suite_b200/variants
for t in c0q0 c1q0 c0q1 c1q1; do HHG_LIB=$V/libhhg_$t.so python tools/perf_probe.py HHG_GROUP_JOBS=64 HHG_GROUP_JOBS=64,HHG_STRIP_ROWS=8 2>&1 | sed "s/^/$t /"; done
HHG_LIB=$V/libhhg_c0q0.so python tools/perf_probe.py HHG_GROUP_JOBS=1 HHG_GROUP_JOBS=8 HHG_GROUP_JOBS=32 HHG_GROUP_JOBS=296 2>&1 | sed "s/^/c0q0 /"
