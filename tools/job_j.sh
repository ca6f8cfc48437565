#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "assim or diffusive or res_ or muskingum" 2>&1 | tail -5
