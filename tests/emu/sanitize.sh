#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  The CPU stand-in for compute-sanitizer: the unmodified .cu sources built for the launch emulator with
# AddressSanitizer (and, second pass, UndefinedBehaviorSanitizer), then a few frames of every pass (sanitize_frames.py).
#   tests/emu/sanitize.sh            -> builds tests/emu/_build_san/{asan,ubsan}/libkjb_emu.so and prints what the sanitizers report
set -e
cd "$(dirname "$0")"
CSRC=../../kajiya_b200/csrc
COMMON="-O1 -g -std=c++17 -fPIC -march=x86-64-v3 -ffp-contract=off -fno-fast-math -fwrapv -DKJB_EMU -include cuda_shim.h -pthread -fno-omit-frame-pointer -w"
for kind in asan ubsan; do
    if [ $kind = asan ]; then SAN="-fsanitize=address"; RT=$(gcc -print-file-name=libasan.so); else SAN="-fsanitize=undefined -fno-sanitize=float-divide-by-zero,vptr"; RT=$(gcc -print-file-name=libubsan.so); fi
    out=_build_san/$kind; mkdir -p $out
    for f in $CSRC/*.cu; do g++ $COMMON $SAN -x c++ -c $f -o $out/$(basename $f).o & done
    g++ $COMMON $SAN -c $CSRC/kjb_bvh.cpp -o $out/bvh.o & g++ $COMMON $SAN -c $CSRC/host/kjb_world.cpp -o $out/world.o & g++ $COMMON $SAN -c cuda_shim.cpp -o $out/shim.o &
    wait
    g++ -shared $SAN -o $out/libkjb_emu.so $out/*.o
    echo "== $kind"
    LD_PRELOAD=$RT ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=0 KJB_EMU_THREADS=1 python sanitize_frames.py $out/libkjb_emu.so
done
