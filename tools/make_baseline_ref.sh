#!/bin/bash
# Recipe for baseline/_ref: the UNMODIFIED reference tree (facebookresearch/param), copied as the package
# `param_bench` — the name its own runners import themselves by (train/comms/pt/comms.py:15-36).
# baseline/_ref is git-ignored (it is not product source and never enters history) but NOT gpurun-ignored, so it
# travels to the GPU box, where /root/reference does not exist.  Nothing is edited; the two stub modules the
# reference's et_replay needs for imports it does not use on this path (pydot, intervaltree — absent from the
# image) are created in memory by param_b200/integration/refpath.py, not written here.
set -e
cd "$(dirname "$0")/.."
SRC=${1:-/root/reference}
[ -d "$SRC/train/comms/pt" ] || { echo "no reference tree at $SRC"; exit 1; }
rm -rf baseline/_ref
mkdir -p baseline/_ref
cp -r "$SRC" baseline/_ref/param_bench
rm -rf baseline/_ref/param_bench/.git
( cd baseline/_ref/param_bench && find . -type f | sort | xargs sha1sum ) | sha1sum | awk '{print $1}' > baseline/_ref/TREE_SHA1
echo "baseline/_ref/param_bench: $(find baseline/_ref/param_bench -type f | wc -l) files, tree sha1 $(cat baseline/_ref/TREE_SHA1)"
