# oracle/ref_build.mk -- TEST INFRASTRUCTURE, not product code.
#
# Compiles the UNMODIFIED reference (soedinglab/hh-suite, mounted read-only at
# $(REF), default /root/reference) straight from the sources where they lie,
# with the flags of the official release binaries (AVX2, no FMA contraction:
# src/CMakeLists.txt:16-22 "-mavx2", azure-pipelines.yml:58-65), WITHOUT running
# the reference's CMake build system.  Outputs only into oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun).  No reference source is copied into the repo.
#
# The only non-source inputs the reference build needs are
#   * hhsuite_config.h  (three version macros, src/hhsuite_config.h.in)
#   * context_data.crf.h / cs219.lib.h (binary resources embedded with xxd -i,
#     cmake/ResourceCompiler.cmake:23-35).  We embed the same bytes with `ld -r -b binary`
#     and a 4-line shim header that maps the xxd symbol names onto the linker symbols.
#
# Usage:  make -f oracle/ref_build.mk -j8        (from the repo root)
REF    ?= /root/reference
OUT    ?= oracle/_ref
CXX     = /usr/bin/g++
CC      = /usr/bin/gcc
# -ffp-contract=off is the GCC default for ISO mode only; with -mavx2 (no -mfma) no
# FMA can be emitted at all, which is what pins the arithmetic.
ARCH    = -mavx2
CXXFLAGS = -O3 -std=c++11 -fsigned-char -fno-strict-aliasing -fopenmp -DOPENMP -fPIC $(ARCH) -w
CFLAGS   = -O3 -std=c99 -fPIC -w -D_GNU_SOURCE
INC     = -I$(REF)/src -I$(REF)/src/cs -I$(REF)/lib/ffindex/src -I$(REF)/lib/simde -I$(REF)/lib/simd -I$(OUT)/gen

HH_SRC  = hhblits hhdecl hhhit hhmatrices hhsearch hhalign hhhitlist hhposteriordecoder hhutil util \
          hhalignment hhforwardalgorithm hhhmm hhposteriordecoderrunner hhviterbialgorithm hhfullalignment \
          hhhmmsimd hhposteriormatrix hhviterbi hhbacktracemac hhmacalgorithm hhprefilter hhviterbimatrix \
          hhbackwardalgorithm ffindexdatabase hhdatabase hhhalfalignment hhviterbirunner hhfunc
CS_SRC  = aa as assert_helpers blosum_matrix getopt_pp log application
FF_SRC  = ffindex ffutil

OBJS  = $(addprefix $(OUT)/obj/,$(addsuffix .o,$(HH_SRC))) \
        $(OUT)/obj/vit_celloff.o $(OUT)/obj/vit_ss.o $(OUT)/obj/vit_celloff_ss.o \
        $(addprefix $(OUT)/obj/cs_,$(addsuffix .o,$(CS_SRC))) \
        $(addprefix $(OUT)/obj/ff_,$(addsuffix .o,$(FF_SRC))) \
        $(OUT)/obj/res_crf.o $(OUT)/obj/res_cs219.o

all: $(OUT)/libhhref.a $(OUT)/libhhref_shim.so $(OUT)/hh_dropin_check

$(OUT)/gen/.stamp:
	mkdir -p $(OUT)/gen $(OUT)/obj
	printf '#define HHSUITE_VERSION_MAJOR 3\n#define HHSUITE_VERSION_MINOR 3\n#define HHSUITE_VERSION_PATCH 0\n' > $(OUT)/gen/hhsuite_config.h
	printf '#pragma once\nextern "C" const unsigned char _binary_context_data_crf_start[];\nextern "C" const unsigned char _binary_context_data_crf_end[];\n#define context_data_crf _binary_context_data_crf_start\n#define context_data_crf_len ((unsigned int)(_binary_context_data_crf_end-_binary_context_data_crf_start))\n' > $(OUT)/gen/context_data.crf.h
	printf '#pragma once\nextern "C" const unsigned char _binary_cs219_lib_start[];\nextern "C" const unsigned char _binary_cs219_lib_end[];\n#define cs219_lib _binary_cs219_lib_start\n#define cs219_lib_len ((unsigned int)(_binary_cs219_lib_end-_binary_cs219_lib_start))\n' > $(OUT)/gen/cs219.lib.h
	touch $@

$(OUT)/obj/res_crf.o: $(OUT)/gen/.stamp
	cd $(REF)/data && ld -r -b binary -o $(abspath $@) context_data.crf
$(OUT)/obj/res_cs219.o: $(OUT)/gen/.stamp
	cd $(REF)/data && ld -r -b binary -o $(abspath $@) cs219.lib

$(OUT)/obj/%.o: $(REF)/src/%.cpp $(OUT)/gen/.stamp
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@
$(OUT)/obj/vit_celloff.o: $(REF)/src/hhviterbialgorithm.cpp $(OUT)/gen/.stamp
	$(CXX) $(CXXFLAGS) $(INC) -DVITERBI_CELLOFF=1 -c $< -o $@
$(OUT)/obj/vit_ss.o: $(REF)/src/hhviterbialgorithm.cpp $(OUT)/gen/.stamp
	$(CXX) $(CXXFLAGS) $(INC) -DVITERBI_SS_SCORE=1 -c $< -o $@
$(OUT)/obj/vit_celloff_ss.o: $(REF)/src/hhviterbialgorithm.cpp $(OUT)/gen/.stamp
	$(CXX) $(CXXFLAGS) $(INC) -DVITERBI_CELLOFF=1 -DVITERBI_SS_SCORE=1 -c $< -o $@
$(OUT)/obj/cs_%.o: $(REF)/src/cs/%.cc $(OUT)/gen/.stamp
	$(CXX) $(CXXFLAGS) $(INC) -c $< -o $@
$(OUT)/obj/ff_%.o: $(REF)/lib/ffindex/src/%.c $(OUT)/gen/.stamp
	$(CC) $(CFLAGS) $(INC) -c $< -o $@

$(OUT)/libhhref.a: $(OBJS)
	rm -f $@ && ar rcs $@ $(OBJS)

# The C-ABI driver (our own code, oracle/ref_shim.cpp) linked against the reference objects.
shim: $(OUT)/libhhref_shim.so
$(OUT)/libhhref_shim.so: oracle/ref_shim.cpp $(OUT)/libhhref.a
	$(CXX) $(CXXFLAGS) $(INC) -shared -o $@ oracle/ref_shim.cpp -Wl,--whole-archive $(OUT)/libhhref.a -Wl,--no-whole-archive -lgomp

# The drop-in check: reference front half + reference ViterbiRunner vs the C-ABI adapter (needs libhhg.so).
dropin: $(OUT)/hh_dropin_check
$(OUT)/hh_dropin_check: oracle/ref_gpu_adapter.cpp $(OUT)/libhhref.a hh-suite_b200/libhhg.so include/hhg.h
	mkdir -p $(OUT)/data && cp -f $(REF)/data/query.hhm $(REF)/data/query.a3m $(OUT)/data/ && chmod u+w $(OUT)/data/query.hhm $(OUT)/data/query.a3m
	$(CXX) $(CXXFLAGS) $(INC) -o $@ oracle/ref_gpu_adapter.cpp -Wl,--whole-archive $(OUT)/libhhref.a -Wl,--no-whole-archive \
	    -Lhh-suite_b200 -lhhg -Wl,-rpath,'$$ORIGIN/../../hh-suite_b200' -lgomp
