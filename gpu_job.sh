V=hh-suite_b200/variants
HHG_LIB=$V/libhhg_paramq.so python tools/perf_probe.py HHG_GROUP_JOBS=1 HHG_GROUP_JOBS=4 HHG_GROUP_JOBS=16 HHG_GROUP_JOBS=296 2>&1 | sed "s/^/paramq /"
python tools/perf_probe.py HHG_GROUP_JOBS=1 HHG_GROUP_JOBS=4 HHG_GROUP_JOBS=16 2>&1 | sed "s/^/base /"
