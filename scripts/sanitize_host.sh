#!/bin/bash
# AddressSanitizer + UBSan over the HOST code (no GPU needed):
#   1. the oracle (oracle/dccm_oracle.c) under its known-answer tests,
#   2. the product's table code (dennou-ccm_b200/csrc/dccm_tables.cpp: grids, Jones99 / bilinear / regular-grid
#      generators, row-band and separable variants, text / binary readers incl. malformed files) through ctypes.
# The device code has its own visit: scripts/gpu_sanitize.sh (compute-sanitizer).
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
cp oracle/libdccm_oracle.so /tmp/libdccm_oracle.keep
trap 'cp /tmp/libdccm_oracle.keep oracle/libdccm_oracle.so; touch oracle/libdccm_oracle.so' EXIT
gcc -O1 -g -march=x86-64-v3 -fopenmp -ffp-contract=off -fno-fast-math -fPIC -std=gnu11 $SAN -shared \
    -o oracle/libdccm_oracle.so oracle/dccm_oracle.c -lm
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python -m pytest tests/test_oracle_kat.py -x -q -p no:cacheprovider
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off $SAN -shared -o /tmp/libtables_asan.so scripts/asan_stub.cpp dennou-ccm_b200/csrc/dccm_tables.cpp
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 DCCM_ASAN_LIB=/tmp/libtables_asan.so python scripts/asan_tables.py
echo "host sanitizers: clean"
