python -m pytest tests/test_viterbi_gpu.py -x -q -m gpu 2>&1 | tail -2
python tools/perf_probe.py HHG_GROUP_JOBS=64 HHG_GROUP_JOBS=64,HHG_STRIP_ROWS=8
