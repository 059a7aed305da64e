#!/bin/bash
# TEST INFRASTRUCTURE -- the emulated kernels under AddressSanitizer with exact-size numpy buffers: any out-of-bounds read
# or write of an activation / parameter / gradient buffer by a kernel is reported with the kernel's source line.
# (ucontext stacks are not instrumented: --param asan-stack=0.)  Usage: tests/emu/run_asan.sh [pytest args]
set -e
cd "$(dirname "$0")/../.."
B=/tmp/medt_emu_asan
mkdir -p $B && python tests/emu/prep.py $B > /dev/null
for f in $B/*.cpp tests/emu/emu.cpp; do
  g++ -std=c++20 -O1 -g -fPIC -fsanitize=address --param asan-stack=0 -fno-omit-frame-pointer -Wno-unknown-pragmas \
      -I tests/emu/include -I $B -c $f -o $B/$(basename ${f%.cpp}).o &
done
wait
g++ -shared -fsanitize=address -o $B/libmedt_emu.so $B/*.o
EMU_TIGHT=1 EMU_LIB=$B/libmedt_emu.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  python -m pytest tests/test_emu_kernels.py -q -x "$@"
