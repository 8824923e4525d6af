#!/bin/sh
# AddressSanitizer build of the host emulation (tests only): heap buffers stand in for device memory, so
# out-of-range global loads / stores of the kernels are caught on the CPU.  Usage:
#   sh tests/emu/build_asan.sh /tmp/asan
#   ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(g++ -print-file-name=libasan.so) \
#     WB_EMU_LIB=/tmp/asan/libworld_b200_emu_asan.so python -m pytest tests/test_emu_parity.py -q
# (the fuzzers under tests/fuzz/ honour WB_EMU_LIB as well)
set -e
OUT=${1:-/tmp/asan}
cd "$(dirname "$0")"
SRC=../../world_b200/csrc
mkdir -p "$OUT"
for f in wb_api wb_rng wb_cheaptrick wb_d4c wb_stonemask wb_synthesis wb_codec wb_fileio wb_matlab wb_f0common wb_dio wb_harvest wb_host wb_multi; do
  g++ -x c++ -std=c++17 -O1 -g -fPIC -DWB_EMU -ffp-contract=off -mfma -fsanitize=address -fno-omit-frame-pointer \
      -I../../include -c $SRC/$f.cu -o "$OUT/$f.o" &
done
wait
g++ -O1 -g -fPIC -DWB_EMU -fsanitize=address -c emu_globals.cpp -o "$OUT/emu_globals.o"
g++ -shared -fsanitize=address -o "$OUT/libworld_b200_emu_asan.so" "$OUT"/*.o -lm
echo "built $OUT/libworld_b200_emu_asan.so"
