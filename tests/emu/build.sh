#!/bin/sh
# Builds the single-thread host emulation of the kernel sources (tests only, never shipped).
set -e
cd "$(dirname "$0")"
SRC=../../world_b200/csrc
OUT=libworld_b200_emu.so
FILES="$SRC/wb_api.cu $SRC/wb_rng.cu $SRC/wb_cheaptrick.cu $SRC/wb_d4c.cu $SRC/wb_stonemask.cu $SRC/wb_synthesis.cu $SRC/wb_codec.cu $SRC/wb_fileio.cu $SRC/wb_matlab.cu $SRC/wb_f0common.cu $SRC/wb_dio.cu $SRC/wb_harvest.cu $SRC/wb_host.cu $SRC/wb_multi.cu"
OBJS=""
for f in $FILES; do
  o="_$(basename $f .cu).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC -name '*.cuh' -newer $o -o -name '*.h' -newer $o)" ]; then
    g++ -x c++ -std=c++17 -O2 -fPIC -DWB_EMU $WB_EMU_DEFS -ffp-contract=off -mfma -I../../include -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
g++ -O2 -fPIC -DWB_EMU -c emu_globals.cpp -o _emu_globals.o
g++ -shared -o $OUT $OBJS _emu_globals.o -lm
echo built $OUT
