#!/bin/bash
# Builds oracle/quad/libelbo_quad.so (binary128 ELBO, TEST INFRASTRUCTURE ONLY).  Output is git-ignored (*.so).
set -e
cd "$(dirname "$0")"
gcc -O2 -fopenmp -fPIC -shared -o libelbo_quad.so elbo_quad.c -lquadmath -lm
echo built "$(pwd)/libelbo_quad.so"
