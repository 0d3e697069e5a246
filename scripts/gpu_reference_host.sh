#!/bin/bash
# Run tests/test_gpu_reference_host.py on a B200: the reference's own host Python over the product shims.
# /root/reference does not exist on the GPU box, so a SCRATCH copy of the few host files the test imports is placed
# under oracle/_ref/pyref (git-ignored like the rest of oracle/_ref, shipped by gpurun), the test runs there with
# RSDF_REFERENCE_ROOT pointing at it, and the copy is removed again -- nothing of the reference enters the repo.
# Usage: scripts/gpu_reference_host.sh [extra pytest args]      (log -> gpurun_out/reference_host.txt)
set -u
cd "$(dirname "$0")/.."
REF=/root/reference
D=oracle/_ref/pyref
rm -rf "$D"; mkdir -p "$D"
trap 'rm -rf "$D"' EXIT
(cd "$REF" && cp --parents models/*.py systems/utils.py utils/__init__.py utils/misc.py lib/pbr/__init__.py \
    lib/pbr/light.py lib/pbr/utils/*.py "$OLDPWD/$D")
/usr/local/graft/bin/gpurun --timeout 1500 -- "mkdir -p gpurun_out; RSDF_REFERENCE_ROOT=\$PWD/$D timeout 1400 python -m pytest tests/test_gpu_reference_host.py -q -s -rs $* 2>&1 | tee gpurun_out/reference_host.txt | tail -60"
