#!/bin/bash
# TEST INFRASTRUCTURE -- race check of the kernels without a GPU: the emulated library under ThreadSanitizer with one
# TSan fiber per GPU thread (see emu.cpp).  Usage: tests/emu/run_tsan.sh [MT_DEBUG_FLAGS value]; prints the number of
# distinct race reports and leaves the full log in /tmp/medt_emu_tsan/log.txt.
set -e
cd "$(dirname "$0")/../.."
B=/tmp/medt_emu_tsan
mkdir -p $B && python tests/emu/prep.py $B > /dev/null
F="-std=c++20 -O1 -g -fPIC -fsanitize=thread -DEMU_TSAN -fno-omit-frame-pointer -Wno-unknown-pragmas -I tests/emu/include -I $B"
for f in $B/*.cpp tests/emu/race_harness.cpp; do
  g++ $F -c $f -o $B/$(basename ${f%.cpp}).o &
done
# the scheduler itself is NOT instrumented (its bookkeeping is outside the race model); it only calls the fiber API
g++ -std=c++20 -O1 -g -fPIC -DEMU_TSAN -I tests/emu/include -c tests/emu/emu.cpp -o $B/emu.o &
wait
g++ -fsanitize=thread -o $B/race_harness $B/*.o
MT_DEBUG_FLAGS=${1:-0} TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" $B/race_harness > $B/log.txt 2>&1 || true
grep -E "done|DONE|FAILED" $B/log.txt | tail -3
echo "race reports: $(grep -c 'WARNING: ThreadSanitizer: data race' $B/log.txt || true)"
