#!/usr/bin/env bash
# GPU parity of the new process algorithms + the complete golden suite
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
