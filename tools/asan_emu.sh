#!/bin/sh
# Bounds / UB check of the kernel SOURCES: compute-sanitizer is closed on this GPU pool, so the host emulator build
# (-DHFB200_EMU) is compiled with AddressSanitizer + UBSan and driven through NTT / Merkle ops and full segments.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
/usr/bin/g++ -x c++ -DHFB200_EMU -O1 -g -std=c++17 -fPIC -fopenmp -fsanitize=address,undefined -fno-omit-frame-pointer \
    -Wno-unknown-pragmas -I/usr/local/cuda/include -shared -o /tmp/libhfb200_emu_asan.so "$ROOT/hyperfridge-r0_b200/csrc/hfb200.cu"
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0 OMP_NUM_THREADS=4 python "$ROOT/tools/asan_probe.py"
