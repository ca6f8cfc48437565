#!/usr/bin/env bash
# Development helper (GPU box): parity tests with one variant, then bench all variants/*.so
cp ravenhydroframework_b200/librvn_gpu.so /tmp/keep.so
cp "variants/$1.so" ravenhydroframework_b200/librvn_gpu.so
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -4
cp /tmp/keep.so ravenhydroframework_b200/librvn_gpu.so
tools/run_variants.sh "$2"
