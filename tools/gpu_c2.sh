#!/bin/bash
# tests + bench config 2 + one full ncu capture of k_hide on config 2
TAG=$1
cd ${GRAFT_REPO_ROOT:-.}
bash tools/gpu_quick.sh $TAG 2
bash tools/gpu_prof.sh $TAG 2:1.0:k_hide:3
