#!/usr/bin/env bash
# TEST INFRASTRUCTURE — compiles the C restatement (oracle/tnet_oracle.c) into
# oracle/_build/libtnet_oracle.so.  Loaded only by tests/, smoke() and bench.py's cpu_baseline leg.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/_build"
# -ffp-contract=off: no implicit FMA contraction, so float/double intermediates stay where the
# reference kernels have them; -fopenmp only parallelises the GEMM row loop.
gcc -O2 -std=c99 -fPIC -shared -ffp-contract=off -fopenmp -o "$HERE/_build/libtnet_oracle.so" "$HERE/tnet_oracle.c" -lm
echo "built $HERE/_build/libtnet_oracle.so"
