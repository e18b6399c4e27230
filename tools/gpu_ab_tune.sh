#!/bin/bash
# like gpu_ab.sh with an AQH_TUNE value: bash tools/gpu_ab_tune.sh <tag> <tune> "<config>:<scale> ..." <lib1> <lib2> ...
TAG=$1; export AQH_TUNE=$2; shift 2
exec bash tools/gpu_ab.sh $TAG "$@"
