#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py tests/test_reference_inputs.py -m gpu -q -x 2>&1 | tail -6
