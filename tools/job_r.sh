#!/usr/bin/env bash
# full GPU suite on the current tree
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
