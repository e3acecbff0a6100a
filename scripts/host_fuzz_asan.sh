#!/bin/bash
# AddressSanitizer + UBSan over the host-side readers (HDF5 / inflate / zstd / streamvbyte): CPU only, no GPU needed.
#   bash scripts/host_fuzz_asan.sh [seed] [n_hdf5] [n_zstd]
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -x c++ nanodecoder_b200/csrc/fast5.cu -x c++ nanodecoder_b200/csrc/vbz.cu -o /tmp/libh5asan.so
# libstdc++ is preloaded next to libasan so that the __cxa_throw interceptor finds the real function under python
ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) /usr/lib/x86_64-linux-gnu/libstdc++.so.6" \
    python scripts/host_fuzz.py "${1:-1}" "${2:-3000}" "${3:-4000}"
# read assembly (difflib restatement, votes) and the .signal text parser, checked against difflib on the way
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -x c++ nanodecoder_b200/csrc/assembly.cu -o /tmp/libasm_asan.so
ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) /usr/lib/x86_64-linux-gnu/libstdc++.so.6" \
    python scripts/host_fuzz_assembly.py /tmp/libasm_asan.so "${1:-1}" 1500
