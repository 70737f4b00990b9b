# Top-level build.  Target names follow the reference's Makefile (all / sharedlib / jni / clean);
# `cuda` and `oracle` are ours.
#
#   make cuda       ec504_imageencoder_b200/libm1cu.so      CUDA kernels + C ABI (nvcc, sm_100a)
#   make sharedlib  ec504_imageencoder_b200/libencoder.so   host C: reference API + driver (gcc), links libm1cu
#                   ./libencoder.so                         symlink, where the reference's target puts it
#   make all        ./encoder                               main.c + libencoder
#   make jni        ./libencoder_jni.so                     needs JAVA_HOME with <jni.h>
#   make oracle     oracle/libm1oracle.so (+ oracle/_ref when the reference checkout is present)

CC      ?= gcc
NVCC    ?= nvcc
PKG     := ec504_imageencoder_b200
HOST    := $(PKG)/csrc/host
OBJDIR  := $(HOST)/_build
CFLAGS  := -O2 -g -fPIC -Iinclude -ffp-contract=off -Wall -Wno-unused-result -Wno-stringop-overflow
# make sharedlib SAN=1 -> $(PKG)/libencoder_san.so: the same host C with AddressSanitizer + UBSan
# (tools/run_host_sanitizers.sh runs the CPU test suite against it)
ifeq ($(SAN),1)
CFLAGS  += -O1 -fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=undefined
OBJDIR  := $(HOST)/_build_san
LIBENC  := $(PKG)/libencoder_san.so
SANLINK := -fsanitize=address,undefined
else
LIBENC  := $(PKG)/libencoder.so
SANLINK :=
endif
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared
HOSTSRC := m1_bitvector.c m1_stream.c m1_vlc.c m1_blk.c m1_stages.c m1_decode_helpers.c m1_driver.c m1_stb_stub.c
HOSTOBJ := $(addprefix $(OBJDIR)/,$(HOSTSRC:.c=.o))
# stb_image v2.30 (public domain): compiled from the copy the reference vendors, or from STB_IMAGE_H
STB_IMAGE_H ?= /root/reference/include/stb_image.h
STBOBJ  := $(OBJDIR)/stb_image.o

.PHONY: all sharedlib jni cuda oracle clean
all: encoder

cuda: $(PKG)/libm1cu.so
$(PKG)/libm1cu.so: $(PKG)/csrc/m1cu_kernels.cu $(PKG)/csrc/m1cu_api.cu $(wildcard $(PKG)/csrc/*.h $(PKG)/csrc/*.cuh) include/m1cu.h
	$(NVCC) $(NVFLAGS) -o $@ $(PKG)/csrc/m1cu_kernels.cu $(PKG)/csrc/m1cu_api.cu

$(OBJDIR)/%.o: $(HOST)/%.c $(wildcard include/*.h) $(HOST)/m1_vlc_data.inc
	@mkdir -p $(OBJDIR)
	$(CC) $(CFLAGS) -c $< -o $@

ifneq ($(wildcard $(STB_IMAGE_H)),)
$(STBOBJ): $(STB_IMAGE_H)
	@mkdir -p $(OBJDIR)
	$(CC) -O2 -fPIC -w -x c -DSTB_IMAGE_IMPLEMENTATION -c $(STB_IMAGE_H) -o $@
endif
STBLINK := $(if $(wildcard $(STB_IMAGE_H))$(wildcard $(STBOBJ)),$(STBOBJ),)
ifeq ($(STBLINK),)
$(warning *** stb_image.h not found at STB_IMAGE_H=$(STB_IMAGE_H): libencoder.so is being built WITHOUT a JPEG decoder.)
$(warning *** mpeg_encode_procedure() will load no picture and return -1; pass STB_IMAGE_H=/path/to/stb_image.h (v2.30, public domain).)
endif

sharedlib: $(LIBENC)
$(LIBENC): $(HOSTOBJ) $(STBLINK) $(PKG)/libm1cu.so
	$(CC) -shared $(SANLINK) -o $@ $(HOSTOBJ) $(STBLINK) -L$(PKG) -lm1cu -Wl,-rpath,'$$ORIGIN' -lm -lpthread
ifneq ($(SAN),1)
	ln -sf $(PKG)/libencoder.so libencoder.so
endif

encoder: main.c $(PKG)/libencoder.so
	$(CC) $(CFLAGS) -o $@ main.c -L$(PKG) -lencoder -lm1cu -Wl,-rpath,'$$ORIGIN/$(PKG)' -lm

jni: encoder_jni.c $(PKG)/libencoder.so
	$(CC) $(CFLAGS) -I$(JAVA_HOME)/include -I$(JAVA_HOME)/include/linux -I$(JAVA_HOME)/include/darwin -shared \
	    -o libencoder_jni.so encoder_jni.c -L$(PKG) -lencoder -lm1cu -Wl,-rpath,'$$ORIGIN/$(PKG)' -lm

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(OBJDIR) encoder libencoder.so libencoder_jni.so $(PKG)/libencoder.so
