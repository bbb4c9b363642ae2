#!/bin/bash
# build locally, stop on failure, then run a command on the GPU box: tools/gb.sh <timeout> '<command>'
set -o pipefail
python -m multimodal_siamese_cd_b200.build 2>&1 | grep -E "error|rror" | head -10
python -m multimodal_siamese_cd_b200.build > /dev/null 2>&1 || { echo "BUILD FAILED"; exit 1; }
t=$1; shift
gpurun --timeout $t -- "$@"
