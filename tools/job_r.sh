#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
