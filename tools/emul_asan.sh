#!/bin/bash
# Runs the host build of the device work items (tests/cpp/cw_emul.cpp over csrc/cw_core.cuh) under
# AddressSanitizer + UBSan through the CPU test suite tests/test_cw_emul.py: pool layout, record
# offsets, list and table indices of the code the CUDA kernels execute, checked on the CPU.
# (compute-sanitizer is not available on the GPU pool.)
set -e
cd "$(dirname "$0")/.."
mkdir -p build
g++ -std=c++17 -O1 -g -ffp-contract=off -mfma -fsanitize=address,undefined -fno-omit-frame-pointer \
    -shared -fPIC -I include -I planner-rules-mcts_b200/csrc -x c++ tests/cpp/cw_emul.cpp -o build/libcw_emul.so
touch build/libcw_emul.so
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python -m pytest tests/test_cw_emul.py -x -q
rc=$?
rm -f build/libcw_emul.so   # the next test run rebuilds the plain library
exit $rc
