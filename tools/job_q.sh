#!/usr/bin/env bash
# footprint of the routing CTAs next to the sweep: tail kernel one warp per member, 32 / 64 / 128-thread routing CTAs
bash tools/run_variants.sh q2 --skip-other
