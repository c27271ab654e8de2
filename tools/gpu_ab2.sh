#!/bin/bash
for lib in "" $@; do
  echo "== lib: ${lib:-default}"
  AN_LIB_PATH=${lib:+$PWD/anim-nerf_b200/$lib} timeout 120 python tools/profile_mlp.py 1048576 --train 2>&1 | tail -1
  AN_LIB_PATH=${lib:+$PWD/anim-nerf_b200/$lib} timeout 120 python tools/profile_mlp.py 1048576 2>&1 | tail -1
done
