#!/bin/sh
# Host C under AddressSanitizer + UndefinedBehaviorSanitizer:
# builds libencoder_san.so (make sharedlib SAN=1) and oracle/libm1oracle_san.so (make -C oracle san) and runs the
# CPU parts of the host-library, JNI and oracle test files against them.  Needs no GPU.  Leak checking is off:
# the reference API hands malloc'd buffers to the caller by design (bitvector_new, convert_rgb_to_ycbcr, ...),
# and CPython itself never frees everything.  Output: profiles/r2_host_sanitizers.txt.
cd "$(dirname "$0")/.." || exit 1
set -e
SANCC=${SANCC:-/usr/bin/gcc}     # a gcc that ships libasan/libubsan (the image's $CC may not)
make -s sharedlib SAN=1 CC=$SANCC
make -s -C oracle san CC=$SANCC
ASAN=$(readlink -f "$($SANCC -print-file-name=libasan.so)"); UBSAN=$(readlink -f "$($SANCC -print-file-name=libubsan.so)")
OUT=profiles/r2_host_sanitizers.txt
{
  echo "# $(date -u +%FT%TZ)  gcc $($SANCC -dumpversion)  -fsanitize=address,undefined -fno-sanitize-recover=undefined"
  echo "# libencoder_san.so + libm1oracle_san.so, LD_PRELOAD=$ASAN:$UBSAN"
} > $OUT
set +e
LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
M1_HOSTLIB=$PWD/ec504_imageencoder_b200/libencoder_san.so M1_ORACLE_LIB=$PWD/oracle/libm1oracle_san.so M1_SANITIZER_RUN=1 \
  python -m pytest tests/test_host_library.py tests/test_oracle_golden.py tests/test_oracle_vs_ref.py tests/test_decoder.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -25 >> $OUT
rc=$?
grep -c "runtime error\|AddressSanitizer" $OUT | sed 's/^/# sanitizer reports in this log: /' >> $OUT
cat $OUT
exit $rc
