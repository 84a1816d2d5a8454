#!/bin/bash
# Build the CPU logic-simulator of the SIMT kernels (TEST INFRASTRUCTURE ONLY; see cuda_sim.h).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/_build"
g++ -O1 -g -std=c++17 -fPIC -shared -DIINS_CPUSIM -I"$HERE" -x c++ "$ROOT/iins_vae_b200/csrc/iins_runtime.cu" \
    -o "$HERE/_build/libiins_cpusim.so" -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable
echo "built $HERE/_build/libiins_cpusim.so"
