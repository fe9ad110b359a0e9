#!/bin/bash
# development builds of the library with measurement-only switches compiled in (never the product build):
#   scripts/build_dev.sh tune  -> build/dev/libnlml_tune.so   (-DNLML_GEN_TUNE: pin the run-time-rank kernel's configuration)
# use with NLML_HPE_LIB=build/dev/libnlml_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/dev
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o build/dev/libnlml_${name}.so nlml_hpe_b200/csrc/tucker_fit.cu nlml_hpe_b200/csrc/mlp_forward.cu
