#!/usr/bin/env bash
# depth of the sweep -> routing hand-off ring and priority of the routing stream on the headline workload
bash tools/run_variants.sh q3 --skip-other
