#!/usr/bin/env bash
# config 4 (hourly Blended, 1 M HRUs on the 200 x 200 grid): store-stream depth / occupancy variants of the sweep
cp ravenhydroframework_b200/librvn_gpu.so /tmp/keep.so
for v in variants/*.so; do
  cp "$v" ravenhydroframework_b200/librvn_gpu.so
  echo "$(basename $v .so): $(timeout 400 python tools/prof_shapes.py blended_hourly_grid 1 50000 12 2>&1 | tail -1)"
done
cp /tmp/keep.so ravenhydroframework_b200/librvn_gpu.so
