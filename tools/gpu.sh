#!/bin/bash
# build the library (stale .so files travel to the GPU box otherwise), then run a script under gpurun in the background
# usage: tools/gpu.sh <tag> <script> [gpurun args...]
set -e
tag=$1; script=$2; shift 2
python /root/repo/atdn_vslam_b200/build.py > /dev/null
(gpurun "$@" --timeout 1500 -- "bash $script" > /root/repo/gpurun_out/${tag}_call.log 2>&1 &)
echo "started $tag"
